#!/usr/bin/env python
"""bench.py -- throughput of the batched KilobotsEnv.step hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2p|c3|c5|c1]

A "step" is one `KilobotsEnv.step` (10 physics sub-steps of 0.1 s, gym_kilobots/envs/kilobots_env.py:161-215)
for EVERY environment of the batch.  Default workload (BASELINE.json configs[1]): 4096 envs x 15
PhototaxisKilobots + 4 CornerQuads + circular gradient light (QuadAssemblyKilobotsEnv) per GPU.
Weak scaling: every rank owns its own 4096 environments; no data-path collective (NCCL only
all-reduces the timing and the episode statistics).

One JSON line on stdout (rank 0).  `value` = kilobot-steps/s with all inputs resident in HBM,
`e2e` = the same metric through KilobotsVecEnv.step with HOST numpy buffers (pinned staging,
H2D + D2H inside the timed region).  `--impl reference` times the CPU oracle (oracle/, a Box2D
restatement -- pybox2d itself is not installable here) on all host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gym_kilobots_b200 import _abi as abi  # noqa: E402
from gym_kilobots_b200 import scenarios as SC  # noqa: E402

METRIC = "kilobot-steps/s"
UNIT = "kilobot-steps/s"


def build_scenario(name, num_envs, env_offset=0):
    if name == "c2":
        return SC.c2_quad_assembly(num_envs, seed=0, env_offset=env_offset, degenerate=True)
    if name == "c2p":
        return SC.c2_quad_assembly(num_envs, seed=0, env_offset=env_offset, degenerate=False)
    if name == "c3":
        return SC.c3_shapes(num_envs, seed=0, env_offset=env_offset)
    if name == "c5":
        return SC.c5_small(num_envs, seed=0, env_offset=env_offset)
    if name == "c1":
        return SC.c1_single_env(num_envs, seed=0, env_offset=env_offset)
    raise SystemExit("unknown workload %s" % name)


DEFAULT_ENVS = {"c2": 4096, "c2p": 4096, "c3": 8192, "c5": 1 << 20, "c1": 1}
WORKLOAD_DESC = {
    "c2": "QuadAssemblyKilobotsEnv: 4096 envs x (15 PhototaxisKilobot + 4 CornerQuad 0.15m + CircularGradientLight r=0.2) per GPU, reference spawn",
    "c2p": "QuadAssembly scene, non-degenerate spawn, 4096 envs x 15 kilobots per GPU",
    "c3": "8192 envs x (50 PhototaxisKilobot + LForm/Triangle/Circle object) per GPU",
    "c5": "2^20 envs x (4 PhototaxisKilobot + Quad) per GPU",
    "c1": "single env: 10 PhototaxisKilobot + Quad + CircularGradientLight",
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes_per_env_step(N, M, P, C, A, L):
    """minimal persistent state read+written once per env-step, action in, obs out."""
    return 2.0 * (32 * (N + M) + 16 * N + 16 * P + 40 * C + 32) + 4 * A + (12 * (N + M) + 4 * L + 8)


def algorithmic_flops_per_env_step(N, M, pair_tests, points, pos_iters_per_island_step, substeps=10):
    """SURVEY.md 8(d) F_alg, per env-step (sums over the sub-steps are passed in as per-substep means)."""
    return substeps * (76 * N + 60 * M + 12 * pair_tests + points * (102 + 620 + 80 * pos_iters_per_island_step))


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per kb_step launch from the committed `ncu --set full` capture
    of this command (profiles/traffic.json, written from the .ncu-rep by tools/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def time_oracle(workload, sample_envs, steps, warmup, threads):
    """CPU oracle on the host cores: kilobot-steps/s on a bounded sample of the workload."""
    from oracle import kbo
    kbo.build()
    sc = build_scenario(workload, sample_envs)
    ob = kbo.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts, threads=threads)
    ob.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, sc.num_envs, steps + warmup)
    for t in range(warmup):
        ob.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        ob.step(acts[t])
    dt = time.perf_counter() - t0
    N = sc.scenes[0].num_kilobots
    return sc.num_envs * steps * N / dt, dt / steps, N


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = min(DEFAULT_ENVS[args.workload], args.ref_envs)
    value, sec_per_step, N = time_oracle(args.workload, sample, args.steps, args.warmup, threads)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "sample_envs": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d envs of the workload x %d env-steps, oracle/libkbo.so (Box2D restatement; "
                                   "pybox2d is not installable offline), one thread per core" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gym_kilobots_b200.envs import KilobotsVecEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E = args.envs or DEFAULT_ENVS[args.workload]
    sc = build_scenario(args.workload, E, env_offset=rank * E)
    env = KilobotsVecEnv(sc, device=local_rank)
    env.reset()
    b = env.batch
    total = args.steps + args.warmup
    acts_h = SC.random_actions(sc, E, total, seed=1 + rank)
    acts_d = torch.as_tensor(acts_h, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # clocks / throttle reasons are sampled from the first warm-up step to the end of the end-to-end leg (the
    # timed region of a default run is only ~25 ms, shorter than nvidia-smi's sampling period)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---------------- device-resident leg
    for t in range(args.warmup):
        env.step_device(acts_d[t])
    cnt0 = b.counters().astype(np.float64).sum(0)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (not inside the event pair)
        starts[i].record()
        env.step_device(acts_d[args.warmup + i])
        ends[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    elapsed = float(step_ms.sum()) * 1e-3
    cnt1 = b.counters().astype(np.float64).sum(0)
    status = b.status.cpu().numpy()

    # ---------------- end-to-end leg: public API, host buffers.  Same episode and the same step indices as the
    # device-resident leg (reset, W untimed steps through the host path, then the K timed ones), so that the two
    # numbers differ by the host<->device path only
    env.reset()
    for t in range(args.warmup):
        env.step(acts_h[t])
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        obs, rew, done, info = env.step(acts_h[args.warmup + i])
    torch.cuda.synchronize(dev)
    e2e_elapsed = time.perf_counter() - e0
    h2d, d2h = env.host_io_bytes()
    # keep the GPU under the same load (untimed) until the sampler has seen it for at least ~0.6 s
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:
        for _ in range(20):
            env.step_device(acts_d[args.warmup])
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([elapsed, e2e_elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, e2e_elapsed = float(t[0]), float(t[1])
        stats = torch.tensor([float((status != 0).sum()), float(E * args.steps)], dtype=torch.float64, device=dev)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)  # episode statistics: the only collective on this path
        failed_envs, env_steps_total = float(stats[0]), float(stats[1])
    else:
        failed_envs, env_steps_total = float((status != 0).sum()), float(E * args.steps)

    N, M = b.N, b.M
    env_steps_s = env_steps_total / elapsed
    value = env_steps_s * N
    e2e_value = env_steps_total * N / e2e_elapsed

    out = None
    if rank == 0:
        d = cnt1 - cnt0
        sub = max(d[abi.COUNTER_NAMES.index("substeps")], 1.0)
        C_mean = d[abi.COUNTER_NAMES.index("contacts")] / sub
        pts = d[abi.COUNTER_NAMES.index("points")] / sub
        lvls = d[abi.COUNTER_NAMES.index("levels")] / sub
        isl = max(d[abi.COUNTER_NAMES.index("islands")] / sub, 1e-9)
        pit = d[abi.COUNTER_NAMES.index("pos_iters")] / sub
        ptests = d[abi.COUNTER_NAMES.index("pair_tests")] / sub
        peaks, peak_src = measured_peaks()
        bytes_env = algorithmic_bytes_per_env_step(N, M, b.P, C_mean, b.A, b.L)
        launch_s = float(step_ms.mean()) * 1e-3
        achieved = bytes_env * E / launch_s / 1e9
        flops_env = algorithmic_flops_per_env_step(N, M, ptests, pts, pit / isl)
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        lc = b.launch_config()
        traffic = measured_traffic(args.workload) if E == DEFAULT_ENVS[args.workload] else None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                "kernel": "kb_step_kernel<%d>" % lc["lanes_per_env"], "launch": lc,
                "algorithmic_bytes_per_env_step": bytes_env, "launch_ms": launch_s * 1e3,
                "note": "state I/O is ~5 KB per env-step against ~70 k warp-instructions of sequential-impulse "
                        "solving: the kernel is bound by dependent-issue latency, not by HBM (DESIGN.md section 5)",
                "fp32": {"algorithmic_flops_per_env_step": flops_env,
                         "achieved_tflops": flops_env * E / launch_s / 1e12, "peak_tflops": fp32_peak,
                         "frac": flops_env * E / launch_s / 1e12 / fp32_peak,
                         "note": "non-tensor FP32 issue roofline (SURVEY 8d); no dense contraction on this path"}}
        cpu_val, cpu_sec, _ = time_oracle(args.workload, min(E, args.cpu_envs), max(2, min(args.steps, 5)), 1, 1)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(step_ms.mean()), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "envs_per_gpu": E, "kilobots_per_env": N,
                       "objects_per_env": M, "substeps_per_step": sc.scenes[0].steps_per_action,
                       "l2": "256 MiB device buffer zeroed between timed steps (L2 flush, outside the event pairs)",
                       "parallelism": "env-sharded x%d" % world},
            "env_steps_per_s": env_steps_s, "substeps_per_s": env_steps_s * sc.scenes[0].steps_per_action,
            "wall_s_timed_region": wall,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "KilobotsVecEnv.step(numpy) -> kb_step_host (pinned staging, H2D + kernel + D2H)"},
            "gpu_launches": int(args.steps),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "%d envs of the workload, oracle/libkbo.so single thread (Box2D restatement; "
                                       "pybox2d not installable offline)" % min(E, args.cpu_envs)},
            "sim_stats": {"contacts_per_substep": C_mean, "manifold_points_per_substep": pts,
                          "gs_levels_per_substep": lvls, "islands_per_substep": isl,
                          "pos_iters_per_island": pit / isl, "toi_events": d[abi.COUNTER_NAMES.index("toi_events")],
                          "envs_with_status_flags": failed_envs},
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(DEFAULT_ENVS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--cpu-envs", type=int, default=256, help="sample size of the cpu_baseline leg")
    ap.add_argument("--ref-envs", type=int, default=2048, help="sample size of --impl reference")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
