#!/usr/bin/env python
"""The TMA-pipelined rigid-water step at bench size against the LDG-staged kernel: same slab, same seed, bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from openmm_constv_b200 import _capi, integrator as hostapi, slab as slabmod

atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
groups = sys.argv[3].split(",") if len(sys.argv) > 3 else ["4"]
s, w_atoms, w_params, cons = slabmod.make_water_slab(atoms // 2, "single", seed=12345)
lib = _capi.load()


def run(env):
    os.environ.update(env)
    sysm = hostapi.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, 2 * s.zmax))
    sysm.addConstraints(*cons)
    it = hostapi.ConstVLangevinIntegrator(300.0, 1.0, 0.001)
    it.setRandomNumberSeed(99)
    ctx = hostapi.SlabContext(sysm, it, precision="single", flags=_capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE, padded=s.padded)
    ctx.load_slab(s)
    out = []
    for k in range(steps):
        it.step(1)
        if k in (0, steps - 1):
            torch.cuda.synchronize()
            out.append((ctx.posq.clone(), ctx.velm.clone(), ctx.force.clone()))
    print(env, "ok, step count", it._kernel.get_step_count(), flush=True)
    ctx.close()
    return out


ref = run({"CONSTV_SETTLE_WS": "0"})
for g in groups:
    got = run({"CONSTV_SETTLE_WS": "1", "CONSTV_WS_GROUPS": g})
    for (a, b) in zip(ref, got):
        for x, y, name in zip(a, b, ("posq", "velm", "force")):
            same = torch.equal(x.view(torch.int32) if x.dtype != torch.int64 else x, y.view(torch.int32) if y.dtype != torch.int64 else y)
            print("  groups", g, name, "bit-exact" if same else "DIFFERS", flush=True)
