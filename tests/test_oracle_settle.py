"""The SETTLE restatement (oracle/constv_oracle_settle_impl.h): PARITY UNPINNED against OpenMM, which owns
the reference's constraint solver and is absent.  Pinned instead on
 * an independent solution of the same constraint problem: iterative SHAKE in float64 (oracle/settle_numpy.py),
 * the invariants of the algorithm: the three distances restored, the centre of mass untouched, corrections
   along the OLD bond vectors only (Miyamoto & Kollman 1992, eqs. 1-5),
 * OpenMM's cluster detection rule restated twice (C ABI host code vs numpy; the C ABI half needs no GPU).
"""
import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from oracle import settle_numpy
from tests import helpers as H

T, GAMMA, DT = 300.0, 1.0, 0.002


def positions64(s):
    p = s.posq[:, :3].astype(np.float64)
    if s.corr is not None:
        p = p + s.corr[:, :3]
    return p


def unconstrained_step(oracle, s, seed=3):
    """part 1 on a copy: returns posDelta (the unconstrained displacements) and the velm part 1 wrote."""
    prm, dt = H.oracle_params_for(oracle, s.precision, T, GAMMA, DT)
    xi = np.random.default_rng(seed).standard_normal((s.padded, 4)).astype(np.float32)
    delta = np.zeros((s.padded, 4), H.MIXED[s.precision])
    velm = s.velm.copy()
    oracle.part1(s.precision, velm, s.force.reshape(-1), delta, prm, dt, xi, 0, s.num_atoms)
    return delta, velm, prm, dt, xi


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_settle_matches_converged_shake(oracle, precision):
    s, atoms, params, _ = slabmod.make_water_slab(3000, precision, seed=21)
    delta, velm, *_ = unconstrained_step(oracle, s)
    old = positions64(s)
    unc = old + delta[:, :3].astype(np.float64)
    masses = 1.0 / velm[atoms, 3].astype(np.float64)
    want = settle_numpy.shake_molecules(old[atoms], unc[atoms], masses, params[:, 0].astype(np.float64),
                                        params[:, 1].astype(np.float64))
    oracle.settle_positions(precision, s.posq, s.corr, delta, velm, atoms, params)
    got = old[atoms] + delta[atoms, :3].astype(np.float64)
    # the float64 SHAKE solution is exact to ~1e-15.  SETTLE in float arithmetic carries ~1e-7 nm; in double
    # it is limited by the height of the canonical triangle, sqrt(dAB^2 - (dBC/2)^2), which the formulation
    # restated here evaluates from the FLOAT parameters in float arithmetic (~6e-8 relative)
    tol = 3e-7 if precision == "single" else 2e-8
    assert np.max(np.abs(got - want)) < tol
    # the constraint moved things: the unconstrained bonds were off by ~1e-3 nm
    d = np.linalg.norm(unc[atoms[:, 0]] - unc[atoms[:, 1]], axis=1)
    assert np.max(np.abs(d - params[:, 0])) > 1e-4


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_settle_invariants(oracle, precision):
    s, atoms, params, _ = slabmod.make_water_slab(3000, precision, seed=22)
    delta, velm, *_ = unconstrained_step(oracle, s, seed=4)
    old = positions64(s)
    unc = old + delta[:, :3].astype(np.float64)
    untouched = np.ones(s.padded, bool)
    untouched[atoms.reshape(-1)] = False
    before = delta.copy()
    oracle.settle_positions(precision, s.posq, s.corr, delta, velm, atoms, params)
    new = old + delta[:, :3].astype(np.float64)
    # all three inherit the float evaluation of the squared parameters (see above)
    for (i, j, d, rel) in ((0, 1, params[:, 0], 2e-7), (0, 2, params[:, 0], 2e-7), (1, 2, params[:, 1], 2e-7)):
        r = np.linalg.norm(new[atoms[:, i]] - new[atoms[:, j]], axis=1)
        assert np.max(np.abs(r / d - 1)) < (2e-6 if precision == "single" else rel)
    m = 1.0 / velm[atoms, 3].astype(np.float64)
    com_unc = np.einsum("na,nak->nk", m, unc[atoms]) / m.sum(1)[:, None]
    com_new = np.einsum("na,nak->nk", m, new[atoms]) / m.sum(1)[:, None]
    assert np.max(np.abs(com_unc - com_new)) < (2e-7 if precision == "single" else 1e-13)
    # the constraint forces act along the OLD bonds: the correction of each atom lies in the old molecular plane
    normal = np.cross(old[atoms[:, 1]] - old[atoms[:, 0]], old[atoms[:, 2]] - old[atoms[:, 0]])
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    for a in range(3):
        corr = new[atoms[:, a]] - unc[atoms[:, a]]
        assert np.max(np.abs(np.einsum("nk,nk->n", corr, normal))) < (2e-7 if precision == "single" else 1e-13)
    # atoms outside molecules are not touched
    H.assert_bits_equal(delta[untouched], before[untouched], "posDelta of unconstrained atoms")


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
def test_settle_step_composition_and_rigid_trajectory(oracle, precision, num_cells):
    """oracle.settle_step = part 1, SETTLE, part 2, image rewrite, zeroing in the reference's order; over a
    trajectory the molecules stay rigid and the velocities stay on the constraint surface."""
    s, atoms, params, _ = slabmod.make_water_slab(999, precision, num_cells=num_cells, seed=23)
    a, b = s.copy(), s.copy()
    prm, dt = H.oracle_params_for(oracle, precision, T, GAMMA, DT)
    inv = H.identity_order(s.padded)
    m = H.MIXED[precision]
    rng = np.random.default_rng(5)
    for k in range(6):
        p_prev = positions64(a)
        xi = rng.standard_normal((s.padded, 4)).astype(np.float32)
        da, db = np.zeros((s.padded, 4), m), np.zeros((s.padded, 4), m)
        oracle.settle_step(precision, a.num_atoms, num_cells, a.zmax, a.posq, a.corr, a.velm, a.force.reshape(-1), da, prm, dt,
                           xi, 0, inv, atoms, params)
        oracle.part1(precision, b.velm, b.force.reshape(-1), db, prm, dt, xi, 0, b.num_atoms)
        oracle.settle_positions(precision, b.posq, b.corr, db, b.velm, atoms, params)
        oracle.part2(precision, b.posq, b.corr, db, b.velm, dt, b.num_atoms)
        oracle.mirror(precision, b.num_real, num_cells, b.zmax, b.posq, b.corr, inv)
        oracle.zero_images(precision, b.num_atoms, b.num_real, b.velm, b.force.reshape(-1), inv)
        H.assert_bits_equal(a.posq, b.posq, "posq")
        H.assert_bits_equal(a.velm, b.velm, "velm")
    p = positions64(a)
    rel = 5e-6 if precision == "single" else 2e-7
    for (i, j, d) in ((0, 1, params[:, 0]), (0, 2, params[:, 0]), (1, 2, params[:, 1])):
        r = np.linalg.norm(p[atoms[:, i]] - p[atoms[:, j]], axis=1)
        assert np.max(np.abs(r / d - 1)) < rel
    # v = delta/dt of a rigid displacement: the relative velocity of two atoms of a molecule is perpendicular
    # to their mid-step bond, (r_new - r_old).(r_new + r_old) = |r_new|^2 - |r_old|^2 = 0
    v = a.velm[:, :3].astype(np.float64)
    bond = 0.5 * (p[atoms[:, 1]] - p[atoms[:, 0]] + p_prev[atoms[:, 1]] - p_prev[atoms[:, 0]])
    vrel = v[atoms[:, 1]] - v[atoms[:, 0]]
    radial = np.abs(np.einsum("nk,nk->n", bond, vrel)) / np.linalg.norm(bond, axis=1)
    assert np.max(radial) < (2e-3 if precision == "single" else 1e-4) * np.median(np.linalg.norm(vrel, axis=1))
    # images follow (charged atoms)
    R = a.num_real
    charged = a.posq[:R, 3] != 0
    assert np.array_equal(a.posq[R:2 * R, 0][charged], a.posq[:R, 0][charged])


def test_cluster_detection_rule():
    """numpy restatement of OpenMM's rule on the water slab's constraint list, plus the cases it must refuse."""
    s, atoms, params, (p1, p2, dist) = slabmod.make_water_slab(500, "single", seed=24)
    got_atoms, got_params, uncovered = settle_numpy.find_settle_clusters(s.num_atoms, p1, p2, dist)
    assert uncovered == 0
    assert np.array_equal(got_atoms, atoms) and np.array_equal(got_params, params)
    # shuffled constraint order and swapped endpoints give the same clusters
    order = np.random.default_rng(1).permutation(len(p1))
    a2, pr2, unc2 = settle_numpy.find_settle_clusters(s.num_atoms, p2[order], p1[order], dist[order])
    assert unc2 == 0 and np.array_equal(a2, atoms) and np.array_equal(pr2, params)
    # a molecule listed H, O, H: the apex is found by the equal distances, not by position in the list
    a3, pr3, _ = settle_numpy.find_settle_clusters(3, [0, 1, 0], [1, 2, 2], [0.1, 0.1, 0.1633])
    assert a3.tolist() == [[1, 0, 2]] and np.allclose(pr3, [[0.1, 0.1633]])
    # a chain (no closing bond) and a 4-ring are left to the general solver
    _, _, unc = settle_numpy.find_settle_clusters(4, [0, 1], [1, 2], [0.1, 0.1])
    assert unc == 2
    _, _, unc = settle_numpy.find_settle_clusters(4, [0, 1, 2, 3], [1, 2, 3, 0], [0.1] * 4)
    assert unc == 4
    with pytest.raises(ValueError, match="must be the same"):
        settle_numpy.find_settle_clusters(3, [0, 1, 0], [1, 2, 2], [0.1, 0.11, 0.12])


def test_capi_cluster_detection_matches_the_numpy_rule():
    """constv_settle_find_clusters (pure host code in the C-ABI library, no GPU) against the numpy restatement."""
    import ctypes as C
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    vp = C.c_void_p

    def find(n, p1, p2, dist):
        p1, p2 = np.ascontiguousarray(p1, np.int32), np.ascontiguousarray(p2, np.int32)
        dist = np.ascontiguousarray(dist, np.float64)
        atoms = np.zeros((max(len(p1) // 3, 1), 3), np.int32)
        params = np.zeros((max(len(p1) // 3, 1), 2), np.float32)
        found, other = C.c_int32(0), C.c_int32(0)
        rc = lib.constv_settle_find_clusters(n, len(p1), p1.ctypes.data_as(vp), p2.ctypes.data_as(vp), dist.ctypes.data_as(vp),
                                             atoms.ctypes.data_as(vp), params.ctypes.data_as(vp), C.byref(found), C.byref(other))
        return rc, atoms[:found.value], params[:found.value], other.value

    s, atoms, params, (p1, p2, dist) = slabmod.make_water_slab(700, "single", seed=25)
    order = np.random.default_rng(2).permutation(len(p1))
    for a, b, d in ((p1, p2, dist), (p2[order], p1[order], dist[order])):
        rc, got_atoms, got_params, other = find(s.num_atoms, a, b, d)
        want_atoms, want_params, want_other = settle_numpy.find_settle_clusters(s.num_atoms, a, b, d)
        assert rc == 0 and other == want_other == 0
        assert np.array_equal(got_atoms, want_atoms) and np.array_equal(got_params, want_params)
        assert np.array_equal(got_atoms, atoms)
    # extra constraints: a lone bond elsewhere is left over; a bond ON a water atom breaks that triangle
    rc, got_atoms, _, other = find(s.num_atoms, np.append(p1, 0), np.append(p2, 1), np.append(dist, 0.3))
    assert rc == 0 and other == 1 and len(got_atoms) == len(atoms)
    rc, got_atoms, _, other = find(s.num_atoms, np.append(p1, 0), np.append(p2, atoms[0, 0]), np.append(dist, 0.3))
    assert rc == 0 and other == 4 and len(got_atoms) == len(atoms) - 1
    rc, got_atoms, got_params, other = find(3, [0, 1, 0], [1, 2, 2], [0.1, 0.1, 0.1633])
    assert rc == 0 and got_atoms.tolist() == [[1, 0, 2]] and np.allclose(got_params, [[0.1, 0.1633]])
    rc, *_ = find(3, [0, 1, 0], [1, 2, 2], [0.1, 0.11, 0.12])
    assert rc == _capi.ERR_SYSTEM and b"must be the same" in lib.constv_last_error()
    rc, *_ = find(3, [0], [3], [0.1])
    assert rc == _capi.ERR_INVALID


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_velocity_constraints_and_rigid_kinetic_energy(oracle, precision):
    """oracle.kinetic_energy_rigid: the half-step-shifted velocities of the molecule atoms are projected onto
    the constraints (what OpenMM's computeKineticEnergy does for a constrained System; PARITY UNPINNED).  The
    projection is pinned by what defines it: relative velocities perpendicular to the bonds, each molecule's
    linear and angular momentum unchanged (constraint forces are internal and central), idempotence, and the
    kinetic energy never growing."""
    s, atoms, params, _ = slabmod.make_water_slab(3000, precision, seed=26)
    shift = 0.0005
    ke, v = oracle.kinetic_energy_rigid(precision, s.posq, s.corr, s.velm, s.force.reshape(-1), shift, atoms, s.num_atoms, True)
    ke_free = oracle.kinetic_energy(precision, s.velm, s.force.reshape(-1), shift, s.num_atoms)
    assert 0.5 * ke_free < ke < ke_free
    p = positions64(s)
    m = 1.0 / s.velm[atoms, 3].astype(np.float64)
    mixed = H.MIXED[precision]
    f = s.force[:, : s.num_atoms].T.astype(mixed)
    w = s.velm[: s.num_atoms, 3:4]
    v_free = (s.velm[: s.num_atoms, :3] + (mixed(shift / 4294967296.0) * f) * w).astype(np.float64)
    for i, j in ((0, 1), (0, 2), (1, 2)):
        d = p[atoms[:, i]] - p[atoms[:, j]]
        dv = v[atoms[:, i]] - v[atoms[:, j]]
        cosine = np.abs(np.einsum("nk,nk->n", d, dv)) / (np.linalg.norm(d, axis=1) * np.linalg.norm(dv, axis=1))
        assert np.max(cosine) < 1e-12
    lin_before = np.einsum("na,nak->nk", m, v_free[atoms])
    lin_after = np.einsum("na,nak->nk", m, v[atoms])
    assert np.max(np.abs(lin_before - lin_after)) < 2e-5 * np.max(np.abs(lin_before))  # v_free is recomputed in numpy: float rounding
    ang_before = np.einsum("na,nak->nk", m, np.cross(p[atoms], v_free[atoms]))
    ang_after = np.einsum("na,nak->nk", m, np.cross(p[atoms], v[atoms]))
    assert np.max(np.abs(ang_before - ang_after)) < 2e-5 * np.max(np.abs(ang_before))
    untouched = np.ones(s.num_atoms, bool)
    untouched[atoms.reshape(-1)] = False
    massive = s.velm[: s.num_atoms, 3] != 0
    assert np.allclose(v[untouched & massive], v_free[untouched & massive], rtol=1e-6, atol=0)
    # idempotence: velocities already on the constraint surface are left alone
    t = s.copy()
    t.velm[: s.num_atoms, :3] = v.astype(mixed)
    t.force[:] = 0
    ke2, v2 = oracle.kinetic_energy_rigid(precision, t.posq, t.corr, t.velm, t.force.reshape(-1), 0.0, atoms, t.num_atoms, True)
    scale = np.max(np.abs(v))
    assert np.max(np.abs(v2 - t.velm[: s.num_atoms, :3].astype(np.float64))) < (1e-6 if precision == "single" else 1e-12) * scale
    # after a constrained step (v = delta/dt of a rigid displacement) nothing is left to project at the mid-step
    # geometry; at the end-of-step geometry the projection removes only the chord/arc difference
    prm, dt = H.oracle_params_for(oracle, precision, T, GAMMA, DT)
    a = s.copy()
    xi = np.random.default_rng(6).standard_normal((s.padded, 4)).astype(np.float32)
    oracle.settle_step(precision, a.num_atoms, 2, a.zmax, a.posq, a.corr, a.velm, a.force.reshape(-1),
                       np.zeros((s.padded, 4), mixed), prm, dt, xi, 0, H.identity_order(s.padded), atoms, params)
    a.force[:] = 0
    ke_a = oracle.kinetic_energy_rigid(precision, a.posq, a.corr, a.velm, a.force.reshape(-1), 0.0, atoms, a.num_atoms)
    ke_a_free = oracle.kinetic_energy(precision, a.velm, a.force.reshape(-1), 0.0, a.num_atoms)
    assert 0.97 * ke_a_free < ke_a <= ke_a_free * (1 + 1e-12)


def test_free_rigid_rotors_conserve_kinetic_energy(oracle):
    """Dynamics, not just geometry: friction 0 (no noise), no forces -- every water molecule is a free rigid
    rotor.  The constrained step (v = constrained displacement / dt) must neither inject nor drain energy: the
    kinetic energy stays put over 1500 steps, and so does each molecule's angular momentum about its centre of
    mass (a solver that mishandled the rotation about the molecular axes would show up here)."""
    precision = "double"
    s, atoms, params, _ = slabmod.make_water_slab(600, precision, seed=27)
    s.force[:] = 0
    R = s.num_real
    m = s.masses[:R]
    prm, dt = H.oracle_params_for(oracle, precision, 300.0, 0.0, 0.001)
    inv = H.identity_order(s.padded)
    delta = np.zeros((s.padded, 4), np.float64)
    xi = np.zeros((s.padded, 4), np.float32)
    mm = m[atoms]

    def observables(t):
        v = t.velm[:R, :3]
        p = positions64(t)[:R]
        com = np.einsum("na,nak->nk", mm, p[atoms]) / mm.sum(1)[:, None]
        vcom = np.einsum("na,nak->nk", mm, v[atoms]) / mm.sum(1)[:, None]
        ang = np.einsum("na,nak->nk", mm, np.cross(p[atoms] - com[:, None], v[atoms] - vcom[:, None]))
        return 0.5 * np.sum(m[:, None] * v ** 2), ang

    oracle.settle_step(precision, s.num_atoms, 2, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1), delta, prm, dt, xi, 0, inv,
                       atoms, params)  # the first step projects the Maxwell-Boltzmann velocities onto rigid motion
    ke0, ang0 = observables(s)
    for k in range(1500):
        oracle.settle_step(precision, s.num_atoms, 2, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1), delta, prm, dt, xi, 0,
                           inv, atoms, params)
    ke1, ang1 = observables(s)
    assert abs(ke1 / ke0 - 1) < 1e-6
    # the angular momentum vector of a free asymmetric top is conserved in the laboratory frame; v is the chord
    # velocity (mid-step) while positions are end-of-step, hence the O((omega dt)^2) tolerance
    assert np.max(np.linalg.norm(ang1 - ang0, axis=1)) < 2e-3 * np.max(np.linalg.norm(ang0, axis=1))
