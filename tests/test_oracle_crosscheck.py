"""Two independent restatements of the reference kernels (C: oracle/constv_oracle.c, numpy:
oracle/constv_numpy.py) must agree bit for bit.  The reference holds no golden vectors for this
path, so this cross-check plus the property tests are what pins the oracle.
Also covers the cases the reference's own tests never reach: image particles, the neutral-atom
skip, numCells > 2, both mirror modes, reordered slots, and the validation messages."""
import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests.helpers import (assert_bits_equal, numpy_step, oracle_params_for, oracle_step,
                           permute_slab)

T, GAMMA, DT = 300.0, 1.0, 0.001


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
def test_c_oracle_equals_numpy_restatement(oracle, precision, num_cells):
    s0 = slabmod.make_slab(600, precision, num_cells=num_cells, seed=3)
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    a, b = s0.copy(), s0.copy()
    for step in range(3):
        xi = oracle.fill_noise(s0.padded, seed=11, step_index=step)
        f = slabmod.new_forces(s0, step)
        a.force[:] = f
        b.force[:] = f
        da = oracle_step(oracle, a, prm, dt, xi)
        db = numpy_step(b, prm, dt, xi)
        assert_bits_equal(da, db, "posDelta")
        assert_bits_equal(a.posq, b.posq, "posq")
        assert_bits_equal(a.velm, b.velm, "velm")
        if precision == "mixed":
            assert_bits_equal(a.corr, b.corr, "posqCorrection")
        assert np.array_equal(a.force, b.force)
        ke_a = oracle.kinetic_energy(precision, a.velm, a.force.reshape(-1), 0.5 * DT, a.num_atoms)
        from oracle import constv_numpy as onp
        ke_b = onp.kinetic_energy(slabmod.MIXED[precision], b.velm, b.force, 0.5 * DT, b.num_atoms)
        assert ke_a == pytest.approx(ke_b, rel=1e-13)


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_reordered_slots_give_the_same_atoms(oracle, precision):
    """invAtomOrder indirection (constVLangevin.cu:142,153,155,170-177): permuting the slots and
    stepping must equal stepping and then permuting."""
    s0 = slabmod.make_slab(257, precision, seed=9)
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    perm = np.random.default_rng(1).permutation(s0.num_atoms).astype(np.int32)
    sp, atom_index = permute_slab(s0, perm)
    inv = oracle.inverse_order(atom_index)
    assert np.array_equal(inv[atom_index], np.arange(s0.padded))
    xi = oracle.fill_noise(s0.padded, seed=5, step_index=0)
    xip = oracle.fill_noise(s0.padded, seed=5, step_index=0, atom_index=atom_index)
    oracle_step(oracle, s0, prm, dt, xi)
    oracle_step(oracle, sp, prm, dt, xip, inv=inv)
    n = s0.num_atoms
    assert_bits_equal(sp.posq[:n], s0.posq[perm], "posq")
    assert_bits_equal(sp.velm[:n], s0.velm[perm], "velm")
    assert np.array_equal(sp.force[:, :n], s0.force[:, perm])


def test_mirror_reference_mode_and_exact_negation(oracle):
    s = slabmod.make_slab(500, "single", seed=2)
    R = s.num_real
    prm, dt = oracle_params_for(oracle, "single", T, GAMMA, DT)
    xi = oracle.fill_noise(s.padded, 1, 0)
    ref, neg = s.copy(), s.copy()
    oracle_step(oracle, ref, prm, dt, xi)                 # z_img = -z + 2*zmax (reference)
    oracle_step(oracle, neg, prm, dt, xi, zshift=0.0)     # z_img = -z         (north_star)
    charged = s.posq[:R, 3] != 0
    assert charged.any() and (~charged).any()
    # real atoms identical in both modes
    assert_bits_equal(ref.posq[:R], neg.posq[:R])
    # exact negation: bit pattern of z differs from the real atom's only in the sign bit
    zr = neg.posq[:R, 2][charged]
    zi = neg.posq[R:2 * R, 2][charged]
    assert np.array_equal(zi.view(np.uint32) ^ np.uint32(0x80000000), zr.view(np.uint32))
    # reference mode: one rounding of the double expression
    expect = (-ref.posq[:R, 2].astype(np.float64) + 2.0 * s.zmax).astype(np.float32)
    assert_bits_equal(ref.posq[R:2 * R, 2][charged], expect[charged])
    # x, y copied; image keeps its own charge
    for t in (ref, neg):
        assert_bits_equal(t.posq[R:2 * R, :2][charged], t.posq[:R, :2][charged])
        assert_bits_equal(t.posq[R:2 * R, 3], s.posq[R:2 * R, 3])
    # neutral atoms' images are NOT rewritten (constVLangevin.cu:143)
    assert_bits_equal(ref.posq[R:2 * R][~charged], s.posq[R:2 * R][~charged])
    # both modes agree modulo the periodic box 2*zmax
    d = ref.posq[R:2 * R, 2][charged].astype(np.float64) - zi.astype(np.float64)
    assert np.allclose(d, 2 * s.zmax, atol=1e-6)


def test_image_zeroing_and_massless_invariants(oracle):
    s = slabmod.make_slab(300, "single", seed=4)
    R, N = s.num_real, s.num_atoms
    s.velm[R:N, :3] = 7.0           # garbage in image velocities: must be wiped
    assert np.any(s.force[:, R:N] != 0)
    before = s.copy()
    prm, dt = oracle_params_for(oracle, "single", T, GAMMA, DT)
    oracle_step(oracle, s, prm, dt, oracle.fill_noise(s.padded, 1, 0))
    assert not s.velm[R:N].any()
    assert not s.force[:, R:N].any()
    assert np.array_equal(s.force[:, :R], before.force[:, :R])   # real forces untouched
    massless = before.velm[:R, 3] == 0
    assert massless.any()
    assert_bits_equal(s.velm[:R][massless], before.velm[:R][massless])
    assert_bits_equal(s.posq[:R][massless], before.posq[:R][massless])
    # without zeroing the reference leaves image velm/force alone
    t = before.copy()
    oracle_step(oracle, t, prm, dt, oracle.fill_noise(s.padded, 1, 0), zero=False)
    assert_bits_equal(t.velm[R:N], before.velm[R:N])
    assert np.array_equal(t.force, before.force)


def test_validation_messages(oracle):
    m = np.array([1.0, 2.0, 0.0, 0.0])
    assert oracle.validate(4, 2, m, 8.0, -1.0) == 4.0           # zmax derived from the box
    assert oracle.validate(4, 2, m, 8.0, 4.0) == 4.0
    with pytest.raises(ValueError, match="Not even number of cells"):
        oracle.validate(3, 3, m[:3], 8.0, -1.0)
    with pytest.raises(ValueError, match="not multiple of number of cells"):
        oracle.validate(3, 2, m[:3], 8.0, -1.0)
    with pytest.raises(ValueError, match="Image Particle has nonzero mass"):
        oracle.validate(4, 2, np.array([1.0, 2.0, 0.0, 1.0]), 8.0, -1.0)
    with pytest.raises(ValueError, match="does not match with the provided zmax"):
        oracle.validate(4, 2, m, 8.0, 3.9)


def test_parameters(oracle):
    from oracle import constv_numpy as onp
    for T_, g, dt in [(300.0, 1.0, 0.001), (0.0, 0.1, 0.01), (100.0, 0.0, 0.002), (372.4, 91.0, 0.004)]:
        p = oracle.params(T_, g, dt)
        assert np.array_equal(p, onp.params(T_, g, dt, oracle.BOLTZ))
    assert oracle.params(100.0, 0.0, 0.002)[1] == 0.002      # friction == 0 -> fscale = dt
    assert oracle.params(100.0, 0.0, 0.002)[2] == 0.0


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_single_pass_range_step_equals_three_kernel_step(oracle, precision):
    """oracle_step_range (the fused single-pass loop timed as the CPU baseline) is the same
    function as the reference's three-kernel sequence, also when applied window by window."""
    s0 = slabmod.make_slab(1000, precision, num_cells=4, seed=21)
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    xi = oracle.fill_noise(s0.padded, 8, 0)
    a, b = s0.copy(), s0.copy()
    oracle_step(oracle, a, prm, dt, xi)
    for lo, hi in [(0, 100), (100, 101), (101, 1000)]:
        oracle.step_range(precision, b.num_atoms, b.num_cells, b.zmax, b.posq, b.corr, b.velm,
                          b.force.reshape(-1), prm, dt, xi, 0, lo, hi)
    assert_bits_equal(a.posq, b.posq)
    assert_bits_equal(a.velm, b.velm)
    if precision == "mixed":
        assert_bits_equal(a.corr, b.corr)
    assert np.array_equal(a.force, b.force)
