"""Shared helpers for the parity tests (oracle drivers; no product code here)."""
from __future__ import annotations

import numpy as np

from openmm_constv_b200 import slab as slabmod

MIXED = slabmod.MIXED
REAL = slabmod.REAL


def identity_order(n_padded):
    return np.arange(n_padded, dtype=np.int32)


def oracle_params_for(oracle, precision, temperature, friction, dt):
    """Parameters as the kernel receives them (CudaConstVKernels.cpp:91-104): double for
    mixed/double, float for single; dt is stored as `mixed`."""
    p = oracle.params(temperature, friction, dt)
    m = MIXED[precision]
    return p.astype(m), m(dt)


def oracle_step(oracle, s, prm, dt, xi, inv=None, zshift=None, zero=True, random_index=0):
    """Advance slab ``s`` in place by one constraint-free step with the C oracle."""
    if inv is None:
        inv = identity_order(s.padded)
    if zshift is None:
        zshift = s.zmax
    delta = np.zeros((s.padded, 4), MIXED[s.precision])
    oracle.step(s.precision, s.num_atoms, s.num_cells, zshift, s.posq, s.corr, s.velm,
                s.force.reshape(-1), delta, prm, dt, xi, random_index, inv, zero, zero)
    return delta


def numpy_step(s, prm, dt, xi, inv=None, zshift=None, zero=True, random_index=0):
    """Same step with the independent numpy restatement."""
    from oracle import constv_numpy as onp
    if inv is None:
        inv = identity_order(s.padded)
    if zshift is None:
        zshift = s.zmax
    r, m = REAL[s.precision], MIXED[s.precision]
    delta = np.zeros((s.padded, 4), m)
    onp.part1(m, s.velm, s.force, delta, prm, dt, xi, random_index, s.num_atoms)
    onp.part2(r, m, s.posq, s.corr, delta, s.velm, dt, s.num_atoms)
    onp.mirror(r, m, s.num_real, s.num_cells, zshift, s.posq, s.corr, inv)
    if zero:
        onp.zero_images(s.num_atoms, s.num_real, s.velm, s.force, inv)
    return delta


def permute_slab(s, perm):
    """Return a copy of ``s`` whose slot ``k`` holds original atom ``perm[k]`` (OpenMM's
    atomIndex array); padding slots stay in place."""
    assert sorted(np.asarray(perm).tolist()) == list(range(s.num_atoms))
    return slabmod.permute_slab(s, perm)


def bits(a):
    """View float arrays as unsigned ints for bit-exact comparison (NaN-safe, -0 != +0)."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64)


def assert_bits_equal(a, b, what=""):
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.argwhere(ba != bb)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {ba.size} words differ; first at {i}: "
                             f"{a[i]!r} vs {b[i]!r}")
