"""Pins the oracle against the REFERENCE'S OWN kernel source.

oracle/_ref/libconstv_ref_host.so is /root/reference/platforms/cuda/src/kernels/constVLangevin.cu
compiled unmodified by g++ (oracle/ref_host.cpp emulates the CUDA execution model, ref_preamble.h
supplies OpenMM's typedef preamble).  Built without FMA contraction it must agree BIT FOR BIT with
the oracle's ORACLE_NO_CONTRACT build on every array, in the three precision models, with 2 and 4
cells, in identity and permuted slot order: that pins operand order, types, promotions, skips and
indexing of the restatement.  (Which operations nvcc fuses is pinned on the GPU by
tests/test_gpu_reference.py against the same source compiled by nvcc.)

The committed fixtures under tests/golden/ref_host_*.npz are outputs of that library
(tests/golden/make_reference_vectors.py), so the same check also runs where neither /root/reference
nor oracle/_ref exist.
"""
import glob
import os

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
T, FRICTION, DT = 300.0, 1.0, 0.001


@pytest.fixture(scope="module")
def ref(oracle):
    oracle.build_ref()
    from oracle import ref as r
    if not r.available("host"):
        pytest.skip("oracle/_ref/libconstv_ref_host.so is not built (needs /root/reference)")
    return r


def run_case(oracle, ref, precision, num_cells, num_real, permuted, steps, seed):
    s = slabmod.make_slab(num_real, precision, num_cells=num_cells, seed=seed)
    atom_index = H.identity_order(s.padded)
    if permuted:
        perm = np.random.default_rng(seed).permutation(s.num_atoms).astype(np.int32)
        s, atom_index = H.permute_slab(s, perm)
    a, b = s.copy(), s.copy()
    prm, dt = H.oracle_params_for(oracle, precision, T, FRICTION, DT)
    inv_ref = ref.host_inverse_order(atom_index)
    inv_orc = oracle.inverse_order(atom_index)
    assert np.array_equal(inv_ref, inv_orc)
    m = H.MIXED[precision]
    da, db = np.zeros((s.padded, 4), m), np.zeros((s.padded, 4), m)
    rng = np.random.default_rng(seed + 1)
    for k in range(steps):
        xi = rng.standard_normal((s.padded + 7, 4)).astype(np.float32)
        ref.host_step(precision, a.num_atoms, num_cells, a.zmax, a.posq, a.corr, a.velm, a.force.reshape(-1), da,
                      prm, dt, xi, 7, inv_ref)
        with oracle.uncontracted():
            oracle.step(precision, b.num_atoms, num_cells, b.zmax, b.posq, b.corr, b.velm, b.force.reshape(-1), db,
                        prm, dt, xi, 7, inv_orc, False, False)
        H.assert_bits_equal(a.posq, b.posq, f"posq step {k}")
        H.assert_bits_equal(a.velm, b.velm, f"velm step {k}")
        H.assert_bits_equal(da, db, f"posDelta step {k}")
        if a.corr is not None:
            H.assert_bits_equal(a.corr, b.corr, f"posqCorrection step {k}")
        assert np.array_equal(a.force, b.force)  # the reference never writes forces
    return a


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("permuted", [False, True])
def test_oracle_matches_reference_source_bit_for_bit(oracle, ref, precision, num_cells, permuted):
    run_case(oracle, ref, precision, num_cells, 3001, permuted, steps=4, seed=11 + num_cells)


def test_single_kernels_match_individually(oracle, ref):
    """part 1, part 2 and the image rewrite one at a time (the split the constraint hand-off uses)."""
    for precision in slabmod.PRECISIONS:
        s = slabmod.make_slab(777, precision, seed=3)
        a, b = s.copy(), s.copy()
        prm, dt = H.oracle_params_for(oracle, precision, T, 0.0, DT)  # friction 0: fscale = dt branch
        m = H.MIXED[precision]
        xi = np.random.default_rng(5).standard_normal((s.padded, 4)).astype(np.float32)
        da, db = np.zeros((s.padded, 4), m), np.zeros((s.padded, 4), m)
        inv = H.identity_order(s.padded)
        ref.host_part1(precision, a.num_atoms, a.velm, a.force.reshape(-1), da, prm, dt, xi, 0)
        with oracle.uncontracted():
            oracle.part1(precision, b.velm, b.force.reshape(-1), db, prm, dt, xi, 0, b.num_atoms)
        H.assert_bits_equal(a.velm, b.velm, "part1 velm")
        H.assert_bits_equal(da, db, "part1 posDelta")
        ref.host_part2(precision, a.num_atoms, a.posq, a.corr, da, a.velm, dt)
        with oracle.uncontracted():
            oracle.part2(precision, b.posq, b.corr, db, b.velm, dt, b.num_atoms)
        H.assert_bits_equal(a.posq, b.posq, "part2 posq")
        H.assert_bits_equal(a.velm, b.velm, "part2 velm")
        ref.host_mirror(precision, a.num_real, 2, a.zmax, a.posq, a.corr, inv)
        with oracle.uncontracted():
            oracle.mirror(precision, b.num_real, 2, b.zmax, b.posq, b.corr, inv)
        H.assert_bits_equal(a.posq, b.posq, "mirror posq")


def test_windowed_reference_step_equals_whole_step(oracle, ref):
    """bench.py times bounded windows of a large slab with constv_ref_host_step_range."""
    for precision in slabmod.PRECISIONS:
        s = slabmod.make_slab(1000, precision, num_cells=4, seed=8)
        a, b = s.copy(), s.copy()
        prm, dt = H.oracle_params_for(oracle, precision, T, FRICTION, DT)
        xi = np.random.default_rng(1).standard_normal((s.padded, 4)).astype(np.float32)
        inv = H.identity_order(s.padded)
        m = H.MIXED[precision]
        da, db = np.zeros((s.padded, 4), m), np.zeros((s.padded, 4), m)
        ref.host_step(precision, a.num_atoms, 4, a.zmax, a.posq, a.corr, a.velm, a.force.reshape(-1), da, prm, dt, xi, 0, inv)
        for lo, hi in ((0, 333), (333, 334), (334, 1000)):
            ref.host_step_range(precision, b.num_atoms, 4, b.zmax, b.posq, b.corr, b.velm, b.force.reshape(-1), db, prm,
                                dt, xi, 0, inv, lo, hi)
        H.assert_bits_equal(a.posq, b.posq, "posq")
        H.assert_bits_equal(a.velm, b.velm, "velm")
        H.assert_bits_equal(da, db, "posDelta")


def test_contracted_oracle_stays_within_rounding_of_the_reference_source(oracle, ref):
    """The default oracle fuses what nvcc fuses; against the uncontracted host build of the
    reference that is a last-bit difference only (north_star: 1e-6 relative per step)."""
    s = slabmod.make_slab(5000, "single", seed=21)
    a, b = s.copy(), s.copy()
    prm, dt = H.oracle_params_for(oracle, "single", T, FRICTION, DT)
    xi = np.random.default_rng(2).standard_normal((s.padded, 4)).astype(np.float32)
    inv = H.identity_order(s.padded)
    d = np.zeros((s.padded, 4), np.float32)
    ref.host_step("single", a.num_atoms, 2, a.zmax, a.posq, None, a.velm, a.force.reshape(-1), d, prm, dt, xi, 0, inv)
    oracle.step("single", b.num_atoms, 2, b.zmax, b.posq, None, b.velm, b.force.reshape(-1), d.copy(), prm, dt, xi, 0,
                inv, False, False)
    n = s.num_atoms
    # a velocity is a sum of three terms of thermal size (~0.5 nm/ps) that may cancel, so the bound
    # is relative to that size: one float ulp of the terms, not of the result
    v_thermal = np.sqrt(np.mean(s.velm[: s.num_real, :3][s.velm[: s.num_real, 3] != 0] ** 2))
    assert np.max(np.abs(a.velm[:n, :3] - b.velm[:n, :3])) < 1e-6 * v_thermal
    assert np.max(np.abs(a.posq[:n, :3] - b.posq[:n, :3])) < 1e-6 * np.max(np.abs(s.posq[:n, :3]))
    assert not np.array_equal(a.velm, b.velm)  # and the two roundings really are different


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_host_*.npz"))))
def test_committed_reference_vectors(oracle, path):
    """Golden vectors generated from the reference's kernel source (host build): the oracle's
    uncontracted build must reproduce them from the stored inputs."""
    g = np.load(path)
    precision = str(g["precision"])
    num_cells, num_atoms, steps = int(g["num_cells"]), int(g["num_atoms"]), int(g["steps"])
    zmax, dt = float(g["zmax"]), H.MIXED[precision](g["dt"])
    prm = g["params"].astype(H.MIXED[precision])
    posq, velm, force = g["posq0"].copy(), g["velm0"].copy(), g["force"].copy()
    corr = g["corr0"].copy() if precision == "mixed" else None
    inv = oracle.inverse_order(g["atom_index"])
    delta = np.zeros_like(velm)
    for k in range(steps):
        with oracle.uncontracted():
            oracle.step(precision, num_atoms, num_cells, zmax, posq, corr, velm, force.reshape(-1), delta, prm, dt,
                        g["noise"][k], 0, inv, False, False)
    H.assert_bits_equal(posq, g["posq1"], "posq")
    H.assert_bits_equal(velm, g["velm1"], "velm")
    if corr is not None:
        H.assert_bits_equal(corr, g["corr1"], "posqCorrection")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_cuda_*.npz"))))
def test_committed_reference_gpu_vectors(oracle, path):
    """Golden vectors recorded from the reference's kernel source compiled by nvcc and run on a B200
    (make_reference_vectors.py --cuda): the oracle's default build, which fuses what nvcc fuses,
    must reproduce them bit for bit."""
    g = np.load(path)
    precision = str(g["precision"])
    num_cells, num_atoms, steps = int(g["num_cells"]), int(g["num_atoms"]), int(g["steps"])
    zmax, dt = float(g["zmax"]), H.MIXED[precision](g["dt"])
    prm = g["params"].astype(H.MIXED[precision])
    posq, velm, force = g["posq0"].copy(), g["velm0"].copy(), g["force"].copy()
    corr = g["corr0"].copy() if precision == "mixed" else None
    inv = oracle.inverse_order(g["atom_index"])
    delta = np.zeros_like(velm)
    for k in range(steps):
        oracle.step(precision, num_atoms, num_cells, zmax, posq, corr, velm, force.reshape(-1), delta, prm, dt,
                    g["noise"][k], 0, inv, False, False)
    H.assert_bits_equal(posq, g["posq1"], "posq")
    H.assert_bits_equal(velm, g["velm1"], "velm")
    if corr is not None:
        H.assert_bits_equal(corr, g["corr1"], "posqCorrection")
