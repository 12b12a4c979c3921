"""The product on two GPUs (skipped on boxes with fewer): one process per GPU under torchrun, NCCL.
tools/allreduce_check.py runs KilobotsVecEnv on each rank's slice of 4096 global envs and checks that (1)
`all_reduce_episode_stats()` equals the sum of the ranks' per-env statistics and (2) the per-env observation hashes,
all-gathered, equal those of a single-GPU run of the same global env ids (1-vs-R determinism)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_all_reduce_and_determinism():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); profiles/allreduce_check_{2,8}gpu_r02.json hold the committed runs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tools", "allreduce_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["world_size"] == 2 and rep["backend"] == "nccl"
    assert rep["all_reduce_matches_sum_of_ranks"] and rep["per_env_obs_hashes_equal_single_gpu_run"]
