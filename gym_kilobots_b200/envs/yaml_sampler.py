"""YamlSceneSampler -- the random scene initialisation of YamlKilobotsEnv, vectorised over E environments and
evaluated ON THE DEVICE, so that auto-resetting finished environments never round-trips to the
host.  It restates, draw for draw, what `_configure_environment` does for ONE env in the reference:

  objects   'random' -> position U(world) * 0.7, orientation U(-pi, pi)      yaml_kilobots_env.py:194-198
  light     'random' -> U(world); 'object' -> on a circle of radius 1.2 * max(w, h) / 2 around a random object;
            momentum lights start with speed .01 in a random direction        :256-283
  kilobots  mean 'light' (the light, or a random component of a composite light per kilobot) / 'random'
            (U(world) * 0.9) / fixed; positions N(mean, std^2) clipped to the world bounds -+ 0.02   :327-354

The output are the `body_pose` [E, B, 3] / `light_state` [E, L] tensors `kb_reset` consumes (plumbing: plain torch
random ops, no simulation work).  The reference draws from numpy's unseeded global generator, so parity
is distributional: tests compare moments and supports with the E = 1 `YamlKilobotsEnv` facade.
Not reproduced: `shuffle(components)` of a composite light with init 'random' (:274-275) -- the component order
is kept.
"""
import math

import numpy as np


class YamlSceneSampler:
    def __init__(self, conf, num_envs, device="cuda", seed=0):
        import torch
        self.torch = torch
        self.conf = conf
        self.E = int(num_envs)
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        w, h = float(conf.width), float(conf.height)
        self.size = torch.tensor([w, h], dtype=torch.float64, device=self.device)
        self.lo = -self.size / 2
        self.hi = self.size / 2
        lights = self._light_leaves(getattr(conf, "light", None))
        self.lights = lights
        self.M = len(conf.objects)
        self.N = int(conf.kilobots.num)
        self.L = sum({"circular": 2, "momentum": 4, "linear": 1}[l.type] for l in lights)

    @staticmethod
    def _light_leaves(lc):
        if lc is None:
            return []
        if isinstance(lc, dict):   # components written as plain mappings in the YAML
            from types import SimpleNamespace
            lc = SimpleNamespace(type=lc.get("type", lc.get("obj_type")), init=lc.get("init"), radius=lc.get("radius"),
                                 components=lc.get("components"))
        if lc.type == "composite":
            out = []
            for c in lc.components:
                out += YamlSceneSampler._light_leaves(c)
            return out
        return [lc]

    def _rand(self, *shape):
        return self.torch.rand(*shape, dtype=self.torch.float64, device=self.device, generator=self.gen)

    def _randn(self, *shape):
        return self.torch.randn(*shape, dtype=self.torch.float64, device=self.device, generator=self.gen)

    def sample(self):
        """-> (body_pose [E, M + N, 3], light_state [E, L]) float64 tensors on the sampler's device."""
        torch, E = self.torch, self.E
        pose = torch.zeros((E, self.M + self.N, 3), dtype=torch.float64, device=self.device)
        # ---- objects (yaml_kilobots_env.py:191-198)
        extent = torch.zeros((E, max(self.M, 1)), dtype=torch.float64, device=self.device)
        for i, o in enumerate(self.conf.objects):
            if isinstance(o.init, str) and o.init == "random":
                pose[:, i, :2] = (self._rand(E, 2) * self.size + self.lo) * 0.7
                pose[:, i, 2] = self._rand(E) * 2 * math.pi - math.pi
            else:
                pose[:, i] = torch.tensor(np.asarray(o.init, dtype=np.float64), device=self.device)
            # Body.width / height as the light placement reads them (:258-259); a yaml 'circle' has radius = width
            extent[:, i] = max(float(o.width), float(o.height))
        # ---- lights (:256-283)
        light = torch.zeros((E, max(self.L, 1)), dtype=torch.float64, device=self.device)
        positions = []
        off = 0
        for lc in self.lights:
            if lc.type == "linear":
                light[:, off] = float(lc.init)
                off += 1
                continue
            if isinstance(lc.init, str) and lc.init == "random":
                p = self._rand(E, 2) * self.size + self.lo
            elif isinstance(lc.init, str) and lc.init == "object":
                which = torch.randint(0, self.M, (E,), device=self.device, generator=self.gen)
                idx = torch.arange(E, device=self.device)
                radius = 1.2 * extent[idx, which] / 2
                angle = self._rand(E) * 2 * math.pi - math.pi
                p = pose[idx, which, :2] + torch.stack([torch.cos(angle) * radius, torch.sin(angle) * radius], dim=1)
            else:
                p = torch.tensor(np.asarray(lc.init, dtype=np.float64), device=self.device).expand(E, 2).clone()
            light[:, off:off + 2] = p
            positions.append(p)
            if lc.type == "momentum":
                a = self._rand(E) * 2 * math.pi - math.pi
                light[:, off + 2] = torch.sin(a) * .01
                light[:, off + 3] = torch.cos(a) * .01
                off += 4
            else:
                off += 2
        # ---- kilobots (:327-354)
        mean = self.conf.kilobots.mean
        if isinstance(mean, str) and mean == "light" and not positions:
            mean = "random"
        if isinstance(mean, str) and mean == "light":
            if len(positions) == 1:
                mu = positions[0][:, None, :]
            else:
                stack = torch.stack(positions, dim=1)                       # [E, K, 2]
                pick = torch.randint(0, len(positions), (E, self.N), device=self.device, generator=self.gen)
                mu = torch.gather(stack, 1, pick[:, :, None].expand(E, self.N, 2))
        elif isinstance(mean, str) and mean == "random":
            mu = ((self._rand(E, 2) * self.size + self.lo) * 0.9)[:, None, :]
        else:
            mu = torch.tensor(np.asarray(mean, dtype=np.float64), device=self.device).expand(E, 1, 2)
        xy = self._randn(E, self.N, 2) * float(self.conf.kilobots.std) + mu
        xy = torch.minimum(torch.maximum(xy, self.lo + 0.02), self.hi - 0.02)
        pose[:, self.M:, :2] = xy
        return pose, light[:, :self.L]
