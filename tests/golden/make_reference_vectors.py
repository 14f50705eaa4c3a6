#!/usr/bin/env python
"""Generates tests/golden/ref_host_*.npz from the REFERENCE'S OWN kernel source.

Run in the build container (needs /root/reference):  python tests/golden/make_reference_vectors.py
It builds oracle/_ref/libconstv_ref_host.so (the reference's constVLangevin.cu compiled unmodified
by g++, see oracle/ref_host.cpp), feeds it seeded slabs and stores inputs and outputs.  The oracle
is NOT involved in producing these files; tests/test_oracle_vs_reference_source.py checks it
against them.

The companion ref_cuda_*.npz fixtures come from the same source compiled by nvcc and run on a B200:
    gpurun -- python tests/golden/make_reference_vectors.py --cuda   (writes gpurun_out/golden/)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from openmm_constv_b200 import slab as slabmod  # noqa: E402  (slab generator only: no compute)

T, FRICTION, DT = 300.0, 1.0, 0.001
BOLTZ = 8.31446261815324e-3
CASES = [  # (precision, num_cells, num_real, permuted, steps)
    ("single", 2, 257, False, 3),
    ("single", 4, 96, True, 2),
    ("mixed", 2, 257, True, 3),
    ("double", 2, 257, False, 3),
    ("double", 4, 96, True, 2),
]
MIXED = {"single": np.float32, "mixed": np.float64, "double": np.float64}


def params_of(precision):
    # CudaConstVKernels.cpp:113-137, computed here independently of the oracle
    vscale = np.exp(-DT * FRICTION)
    fscale = (1 - vscale) / FRICTION
    noisescale = np.sqrt(BOLTZ * T * (1 - vscale * vscale))
    return np.array([vscale, fscale, noisescale], np.float64).astype(MIXED[precision])


def build_case(precision, num_cells, num_real, permuted, steps, seed):
    s = slabmod.make_slab(num_real, precision, num_cells=num_cells, seed=seed)
    atom_index = np.arange(s.padded, dtype=np.int32)
    if permuted:
        perm = np.random.default_rng(seed).permutation(s.num_atoms).astype(np.int32)
        n = s.num_atoms
        s.posq[:n] = s.posq[perm]
        if s.corr is not None:
            s.corr[:n] = s.corr[perm]
        s.velm[:n] = s.velm[perm]
        s.force[:, :n] = s.force[:, perm]
        atom_index[:n] = perm
    noise = np.random.default_rng(seed + 1).standard_normal((steps, s.padded, 4)).astype(np.float32)
    return s, atom_index, noise


def main():
    cuda = "--cuda" in sys.argv
    import oracle
    from oracle import ref
    if not cuda:
        oracle.build_ref()
    out_dir = os.path.join(ROOT, "gpurun_out", "golden") if cuda else os.path.dirname(os.path.abspath(__file__))
    os.makedirs(out_dir, exist_ok=True)
    for ci, (precision, num_cells, num_real, permuted, steps) in enumerate(CASES):
        s, atom_index, noise = build_case(precision, num_cells, num_real, permuted, steps, seed=100 + ci)
        prm = params_of(precision)
        dt = MIXED[precision](DT)
        rec = dict(precision=precision, num_cells=num_cells, num_atoms=s.num_atoms, steps=steps, zmax=s.zmax,
                   dt=dt, params=prm, atom_index=atom_index, noise=noise, posq0=s.posq.copy(), velm0=s.velm.copy(),
                   force=s.force.copy())
        if s.corr is not None:
            rec["corr0"] = s.corr.copy()
        if cuda:
            import torch
            dev = torch.device("cuda", 0)
            r = ref.CudaReference(precision, prm, dt, dev)
            t = {k: torch.as_tensor(v, device=dev) for k, v in
                 dict(posq=s.posq, velm=s.velm, force=s.force, ai=atom_index).items()}
            corr = torch.as_tensor(s.corr, device=dev) if s.corr is not None else None
            inv = torch.zeros_like(t["ai"])
            r.inverse_order(s.num_atoms, t["ai"], inv)
            inv[s.num_atoms:] = t["ai"][s.num_atoms:]
            delta = torch.zeros_like(t["velm"])
            for k in range(steps):
                xi = torch.as_tensor(noise[k], device=dev)
                r.step(s.num_atoms, num_cells, s.zmax, t["posq"], corr, t["velm"], t["force"], delta, xi, 0, inv)
            torch.cuda.synchronize()
            rec.update(posq1=t["posq"].cpu().numpy(), velm1=t["velm"].cpu().numpy())
            if corr is not None:
                rec["corr1"] = corr.cpu().numpy()
            name = f"ref_cuda_{precision}_c{num_cells}_{'perm' if permuted else 'ident'}.npz"
        else:
            inv = ref.host_inverse_order(atom_index)
            delta = np.zeros_like(s.velm)
            for k in range(steps):
                ref.host_step(precision, s.num_atoms, num_cells, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1),
                              delta, prm, dt, noise[k], 0, inv)
            rec.update(posq1=s.posq, velm1=s.velm)
            if s.corr is not None:
                rec["corr1"] = s.corr
            name = f"ref_host_{precision}_c{num_cells}_{'perm' if permuted else 'ident'}.npz"
        np.savez_compressed(os.path.join(out_dir, name), **rec)
        print("wrote", os.path.join(out_dir, name))


if __name__ == "__main__":
    main()
