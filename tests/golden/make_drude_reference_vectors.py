#!/usr/bin/env python
"""Generates tests/golden/drude_ref_host_*.npz from the REFERENCE'S OWN Drude kernel source.

Run in the build container (needs /root/reference):  python tests/golden/make_drude_reference_vectors.py
It builds oracle/_ref/libconstv_ref_host.so (the reference's vectorOps.cu + constVDrudeLangevin.cu +
constVLangevin.cu compiled unmodified by g++, see oracle/ref_drude_host.cpp), feeds it seeded slabs of
polarizable water in the launch order of CudaConstVKernels.cpp:307-344 (part 1, part 2, hard wall, image
rewrite) and stores inputs and outputs.  The oracle is NOT involved in producing these files; the host build
rounds every operation separately, so tests compare the oracle's uncontracted build against them
(tests/test_oracle_drude_vs_reference_source.py), and the GPU kernels within a stated tolerance.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from openmm_constv_b200 import slab as slabmod  # noqa: E402  (slab generator only: no compute)

T, FRICTION, TD, FRICTION_D, DT = 300.0, 5.0, 1.0, 20.0, 0.001
BOLTZ = 8.31446261815324e-3
CASES = [  # (precision, num_cells, num_real, max_drude_distance, steps)
    ("single", 2, 301, 0.02, 3),
    ("single", 4, 120, 0.0, 2),
    ("mixed", 2, 301, 0.02, 3),
    ("double", 2, 301, 0.02, 3),
]
MIXED = {"single": np.float32, "mixed": np.float64, "double": np.float64}


def scalars_of(max_dist):
    # CudaConstVKernels.cpp:261-270, computed here independently of the oracle
    vs = np.exp(-DT * FRICTION)
    fs = (1 - vs) / FRICTION / 4294967296.0
    ns = np.sqrt(2 * BOLTZ * T * FRICTION) * np.sqrt(0.5 * (1 - vs * vs) / FRICTION)
    vd = np.exp(-DT * FRICTION_D)
    fd = (1 - vd) / FRICTION_D / 4294967296.0
    nd = np.sqrt(2 * BOLTZ * TD * FRICTION_D) * np.sqrt(0.5 * (1 - vd * vd) / FRICTION_D)
    return np.array([vs, fs, ns, vd, fd, nd, max_dist, np.sqrt(BOLTZ * TD)], np.float64)


def main():
    import oracle
    from oracle import ref
    oracle.build_ref()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for ci, (precision, num_cells, num_real, max_dist, steps) in enumerate(CASES):
        s, normal, pairs = slabmod.make_drude_slab(num_real, precision, num_cells=num_cells, seed=200 + ci)
        scalars = scalars_of(max_dist)
        dt = MIXED[precision](DT)
        R = s.num_real
        noise = np.random.default_rng(300 + ci).standard_normal((steps, R + 3, 4)).astype(np.float32)
        rec = dict(precision=precision, num_cells=num_cells, num_atoms=s.num_atoms, steps=steps, zmax=s.zmax, dt=dt,
                   scalars=scalars, normal=normal, pairs=pairs, noise=noise, posq0=s.posq.copy(), velm0=s.velm.copy(),
                   force=s.force.copy(), masses=s.masses.copy(), box=np.asarray(s.box), random_index=3)
        if s.corr is not None:
            rec["corr0"] = s.corr.copy()
        inv = np.arange(s.padded, dtype=np.int32)
        delta = np.zeros_like(s.velm)
        for k in range(steps):
            ref.host_drude_step(precision, s.num_atoms, num_cells, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1), delta,
                                normal, pairs, scalars, dt, noise[k], 3, inv)
        rec.update(posq1=s.posq, velm1=s.velm)
        if s.corr is not None:
            rec["corr1"] = s.corr
        name = f"drude_ref_host_{precision}_c{num_cells}.npz"
        np.savez_compressed(os.path.join(out_dir, name), **rec)
        print("wrote", os.path.join(out_dir, name))


if __name__ == "__main__":
    main()
