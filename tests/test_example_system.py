"""BASELINE.json configs[0-1] on the REAL topology of the reference's example (example/nacl_1m_spce), without
OpenMM: the fixture tests/golden/nacl_1m_spce_system.npz is what `forcefield.createSystem(topology,
constraints=AllBonds)` makes of nacl_1m_eq.pdb + spce.xml as far as the integrator is concerned (derived by
the committed script tests/golden/make_example_system.py), the image particles are added by the example's own
recipe (nacl_constv.py:60-95, slab.example_system), and the step is the one the plugin runs for this System:
part 1, the rigid-water constraints (SETTLE), part 2, image rewrite.

CPU: the fixture against the example's facts and, where /root/reference exists, against a fresh derivation;
the oracle trajectory keeps every molecule rigid and every image on its partner.
GPU (-m gpu): 1000-step trajectories through constv_step_fused_settle, in-kernel Philox noise and the
example's applied field efield*q*z as the injected force, bit for bit against the oracle; a friction-0 run.
The forces of the example's force field (PME, Lennard-Jones) are OpenMM's and out of scope: injected.
"""
import os

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "nacl_1m_spce_system.npz")
T, GAMMA, DT = 300.0, 1.0, 0.001            # nacl_constv.py:12-16
EFIELD = 1.0 * 96.48533212331002 / 4.2183   # constV = 1 V over zmax, in kJ/mol/nm/e (nacl_constv.py:29-30)


def field_forces(s):
    """CustomExternalForce('efield*q*z') on the real atoms (nacl_constv.py:69-72,90-91): F_z = -efield*q."""
    f = np.zeros((3, s.padded), np.int64)
    f[2, :s.num_real] = np.rint(-EFIELD * s.posq[:s.num_real, 3].astype(np.float64) * 4294967296.0).astype(np.int64)
    return f


def test_fixture_is_the_example():
    d = np.load(FIXTURE)
    mass, q = d["mass"], d["charge"]
    assert len(mass) == 7494
    assert np.sum(mass == 22.98977) == 38 and np.sum(mass == 35.45) == 38            # SOD, CLA
    assert np.sum(mass == 15.9994) == 2046 and np.sum(mass == 1.008) == 4092          # SPC/E water
    assert np.sum(mass == 0.0) == 1280 and np.all(q[mass == 0.0] == 0.0)              # wall carbons: massless, neutral
    assert abs(q.sum()) < 1e-9 and set(np.unique(q)) == {-1.0, -0.8476, 0.0, 0.4238, 1.0}
    assert np.allclose(d["box_nm"], [3.9352, 4.26, 2 * 4.2183]) and float(d["zmax_nm"]) == 4.2183
    z = d["pos_mA"][:, 2] / 1e4
    assert set(np.unique(z[mass == 0.0])) == {0.0, 4.2183}                             # the two electrodes
    assert z[mass > 0].min() > 0.0 and z[mass > 0].max() < 4.2183
    # constraints=AllBonds + rigidWater: two O-H at 0.1 nm and one H-H at 2*0.1*sin(109.47 deg / 2) per water
    dist = d["constraint_distance"]
    assert len(dist) == 3 * 2046 and np.sum(dist == 0.1) == 2 * 2046
    hh = np.unique(dist[dist != 0.1])
    assert len(hh) == 1 and abs(hh[0] - 2 * 0.1 * np.sin(0.5 * 1.91061193216)) < 1e-12
    if os.path.exists("/root/reference/example/nacl_1m_spce/nacl_1m_eq.pdb"):  # the fixture is today's derivation
        import importlib.util
        spec = importlib.util.spec_from_file_location("mes", os.path.join(ROOT, "tests", "golden", "make_example_system.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        fresh = mod.build()
        for key in ("pos_mA", "mass", "charge", "constraint_p1", "constraint_p2", "constraint_distance", "box_nm"):
            assert np.array_equal(fresh[key], d[key]), key


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_images_follow_the_examples_recipe(precision):
    s, atoms, params, cons = slabmod.example_system(FIXTURE, precision)
    R = s.num_real
    assert (s.num_atoms, s.num_cells, s.zmax) == (14988, 2, 4.2183)
    assert np.all(s.masses[R:] == 0) and np.all(s.velm[R:, 3] == 0) and np.all(s.velm[:, :3] == 0)
    q = s.posq[:R, 3]
    assert np.array_equal(s.posq[R:2 * R, 3], -q)                                       # nacl_constv.py:87
    assert np.array_equal(s.posq[R:2 * R, :2], s.posq[:R, :2])
    p = s.posq[:, 2].astype(np.float64) + (s.corr[:, 2] if s.corr is not None else 0)
    tol = 1e-6 if precision == "single" else 1e-12
    assert np.allclose(p[R:2 * R][q != 0], -p[:R][q != 0], atol=tol, rtol=0)           # :80
    assert np.allclose(p[R:2 * R][q == 0], -p[:R][q == 0] + 0.001, atol=tol, rtol=0)   # :85
    assert len(atoms) == 2046 and np.all(s.masses[atoms[:, 0]] == 15.9994) and np.all(s.masses[atoms[:, 1:]] == 1.008)
    # the library finds the same rigid molecules from the constraint list (what the plugin's initialize() does)
    import ctypes as C
    from openmm_constv_b200 import _capi
    p1, p2, dist = cons
    found_atoms = np.zeros((len(p1) // 3, 3), np.int32)
    found_params = np.zeros((len(p1) // 3, 2), np.float32)
    found, other = C.c_int32(0), C.c_int32(0)
    vp = C.c_void_p
    _capi.check(_capi.load().constv_settle_find_clusters(
        s.num_atoms, len(p1), np.ascontiguousarray(p1).ctypes.data_as(vp), np.ascontiguousarray(p2).ctypes.data_as(vp),
        np.ascontiguousarray(dist).ctypes.data_as(vp), found_atoms.ctypes.data_as(vp), found_params.ctypes.data_as(vp),
        C.byref(found), C.byref(other)))
    assert (found.value, other.value) == (2046, 0)
    assert np.array_equal(found_atoms, atoms) and np.array_equal(found_params, params)


def bond_errors(s, atoms, params):
    p = s.posq[:, :3].astype(np.float64)
    if s.corr is not None:
        p = p + s.corr[:, :3]
    out = []
    for (i, j, d) in ((0, 1, params[:, 0]), (0, 2, params[:, 0]), (1, 2, params[:, 1])):
        out.append(np.max(np.abs(np.linalg.norm(p[atoms[:, i]] - p[atoms[:, j]], axis=1) / d - 1)))
    return max(out)


@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_oracle_trajectory_on_the_example(oracle, precision):
    """The PDB's waters are not exactly on the SPC/E geometry (printed to 0.001 A, and a few percent off the
    constraint lengths); the first constrained step puts every molecule on it and it stays there; images stay
    on their partners; massless wall atoms and images never move or gain a velocity."""
    s, atoms, params, cons = slabmod.example_system(FIXTURE, precision)
    s.force[:] = field_forces(s)
    assert bond_errors(s, atoms, params) > 1e-3
    walls = s.masses[:s.num_real] == 0
    wall_pos = s.posq[:s.num_real][walls].copy()
    prm, dt = H.oracle_params_for(oracle, precision, T, GAMMA, DT)
    inv = H.identity_order(s.padded)
    delta = np.zeros((s.padded, 4), H.MIXED[precision])
    R = s.num_real
    for k in range(60):
        xi = oracle.fill_noise(s.padded, 11, k)
        s.force[:] = field_forces(s)
        oracle.settle_step(precision, s.num_atoms, 2, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1), delta, prm, dt, xi, 0,
                           inv, atoms, params)
        assert bond_errors(s, atoms, params) < (3e-5 if precision == "single" else 3e-7)
    assert np.array_equal(s.posq[:R][walls], wall_pos) and np.all(s.velm[R:] == 0) and np.all(s.velm[:R][walls] == 0)
    charged = s.posq[:R, 3] != 0
    z = s.posq[:R, 2].astype(np.float64) + (s.corr[:R, 2] if s.corr is not None else 0)
    zi = s.posq[R:2 * R, 2].astype(np.float64) + (s.corr[R:2 * R, 2] if s.corr is not None else 0)
    if precision == "single":  # (float)(-(double)z + 2 zmax): one rounding
        want = (-s.posq[:R, 2].astype(np.float64) + 2.0 * s.zmax).astype(np.float32)
        assert np.array_equal(s.posq[R:2 * R, 2][charged].view(np.uint32), want[charged].view(np.uint32))
    else:
        assert np.allclose(zi[charged], -z[charged] + 2 * s.zmax, atol=1e-12, rtol=0)
    assert np.all(s.force[:, R:] == 0)


# ---- GPU -------------------------------------------------------------------------------------------------------------

def _context(s, cons, friction, seed):
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from openmm_constv_b200 import _capi, integrator as hostapi
    _capi.load()
    sysm = hostapi.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, s.num_cells * s.zmax))
    sysm.addConstraints(*cons)
    it = hostapi.ConstVLangevinIntegrator(T, friction, DT, s.num_cells)
    it.setRandomNumberSeed(seed)
    ctx = hostapi.SlabContext(sysm, it, precision=s.precision, padded=s.padded)
    ctx.load_slab(s)
    assert it._kernel.fused_constraints and it._kernel.num_rigid_molecules == 2046
    return ctx, it


def _download(ctx, s):
    out = s.copy()
    out.posq, out.velm, out.force = ctx.posq.cpu().numpy(), ctx.velm.cpu().numpy(), ctx.force.cpu().numpy()
    if ctx.corr is not None:
        out.corr = ctx.corr.cpu().numpy()
    return out


def _assert_same(a, b):
    H.assert_bits_equal(a.posq, b.posq, "posq")
    H.assert_bits_equal(a.velm, b.velm, "velm")
    if a.corr is not None:
        H.assert_bits_equal(a.corr, b.corr, "posqCorrection")
    assert np.array_equal(a.force, b.force), "force"


@pytest.mark.gpu
@pytest.mark.parametrize("precision,friction", [("mixed", GAMMA), ("single", GAMMA), ("mixed", 0.0), ("single", 0.0)])
def test_gpu_1000_step_trajectory_on_the_example(oracle, precision, friction):
    """configs[1]: "same nacl_1m_spce slab on 1 B200 with the new integrator, trajectory-checked".  1000 steps
    of dt = 1 fs at 300 K through the plugin's path for this System (constv_step_fused_settle: the constraints are
    solved inside the step kernel), Philox noise, the applied field re-injected every step (the step zeroes image
    forces only).  The oracle is fed the Gaussians the library publishes for (seed, step): every array equal bit
    for bit every 100 steps and at the end.  friction = 0: no noise, no damping -- the trajectory must also
    conserve KE + sum(efield q z) to the integrator's accuracy (north_star's friction-0 check)."""
    import torch
    from openmm_constv_b200._capi import check
    s0, atoms, params, cons = slabmod.example_system(FIXTURE, precision)
    f = field_forces(s0)
    s0.force[:] = f
    ctx, it = _context(s0, cons, friction, seed=2026)
    lib, slab = it._kernel._lib, it._kernel.handle
    g = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    f_dev = torch.as_tensor(f, device="cuda")
    ref = s0.copy()
    prm, dt = H.oracle_params_for(oracle, precision, T, friction, DT)
    inv = H.identity_order(s0.padded)
    delta = np.zeros((s0.padded, 4), H.MIXED[precision])
    R = s0.num_real
    q = s0.posq[:R, 3].astype(np.float64)
    m = s0.masses[:R]

    def energy(s):
        z = s.posq[:R, 2].astype(np.float64) + (s.corr[:R, 2] if s.corr is not None else 0)
        v = s.velm[:R, :3].astype(np.float64)
        return 0.5 * np.sum(m[:, None] * v * v), np.sum(EFIELD * q * z)

    energies = []
    for k in range(1000):
        check(lib.constv_generate_noise(slab, k, g.data_ptr(), ctx.stream_ptr()))
        xi = g.cpu().numpy()
        ctx.force.copy_(f_dev)
        it.step(1)
        ref.force[:] = f
        oracle.settle_step(precision, ref.num_atoms, 2, ref.zmax, ref.posq, ref.corr, ref.velm, ref.force.reshape(-1), delta,
                           prm, dt, xi, 0, inv, atoms, params)
        if k % 100 == 99:
            _assert_same(_download(ctx, s0), ref)
            energies.append(energy(ref))
    assert it._kernel.get_step_count() == 1000
    assert bond_errors(ref, atoms, params) < (3e-5 if precision == "single" else 3e-7)
    if friction == 0.0:
        # After the first step has put the waters on the SPC/E geometry the motion is conservative: the ions and
        # waters fall in the field (KE grows to ~1800 kJ/mol, the field energy drops by as much) while KE + U stays
        # within the leapfrog's O(dt) bookkeeping error (velocities live half a step behind the positions).
        total = [ke + u for ke, u in energies]
        assert max(ke for ke, _ in energies) > 500.0
        assert max(total) - min(total) < 0.02 * max(ke for ke, _ in energies), energies
    else:
        ke = it.computeKineticEnergy()
        assert ke == pytest.approx(oracle.kinetic_energy_rigid(precision, ref.posq, ref.corr, ref.velm, ref.force.reshape(-1),
                                                               0.5 * DT, atoms, ref.num_atoms), rel=1e-12)
        dof = 3 * np.sum(m > 0) - 3 * 2046
        temp = 2 * ke / (dof * slabmod.BOLTZ)
        assert 200.0 < temp < 1500.0  # thermostatted from v = 0 plus the heat of the first step's geometry snap
    ctx.close()
