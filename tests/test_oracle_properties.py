"""The analytic properties the reference's own CUDA test asserts
(/root/reference/platforms/cuda/tests/TestCudaConstVLangevinIntegrator.cpp), re-expressed without
OpenMM and applied to the oracle.  These are what pins the oracle's integrator arithmetic
(SURVEY.md 8c): testSingleBond (:52-94), testTemperature (:96-129),
testConstrainedMasslessParticles' v == 0 check (:206-209).

Unlike the reference's test systems (which have no image particles and would be rejected by
initialize) every system here is a valid slab: R real atoms + R massless images."""
import numpy as np
import pytest

FORCE_SCALE = 4294967296.0


class TinySlab:
    """R real atoms + R images in OpenMM layout, forces from a python callback."""

    def __init__(self, oracle, precision, pos, mass, charge, zmax=50.0):
        from openmm_constv_b200.slab import MIXED, REAL, padded_atoms
        self.o, self.precision = oracle, precision
        R = len(mass)
        self.R, self.N, self.P = R, 2 * R, padded_atoms(2 * R)
        r, m = REAL[precision], MIXED[precision]
        self.posq = np.zeros((self.P, 4), r)
        self.posq[:R, :3] = pos
        self.posq[:R, 3] = charge
        self.posq[R:2 * R, :3] = pos * [1, 1, -1]
        self.posq[R:2 * R, 3] = -np.asarray(charge)
        self.corr = np.zeros((self.P, 4), r) if precision == "mixed" else None
        self.velm = np.zeros((self.P, 4), m)
        self.velm[:R, 3] = [0.0 if x == 0 else 1.0 / x for x in mass]
        self.force = np.zeros((3, self.P), np.int64)
        self.delta = np.zeros((self.P, 4), m)
        self.inv = np.arange(self.P, dtype=np.int32)
        self.zmax = zmax
        self.time = 0.0
        self.step_count = 0

    def set_forces(self, f):
        self.force[:, : self.R] = np.rint(np.asarray(f).T * FORCE_SCALE).astype(np.int64)

    def positions(self):
        p = self.posq[: self.R, :3].astype(np.float64)
        if self.corr is not None:
            p += self.corr[: self.R, :3]
        return p

    def step(self, T, gamma, dt, force_fn, seed=1):
        from openmm_constv_b200.slab import MIXED
        m = MIXED[self.precision]
        prm = self.o.params(T, gamma, dt).astype(m)
        self.set_forces(force_fn(self.positions()))
        xi = self.o.fill_noise(self.P, seed, self.step_count)
        self.o.step(self.precision, self.N, 2, self.zmax, self.posq, self.corr, self.velm,
                    self.force.reshape(-1), self.delta, prm, m(dt), xi, 0, self.inv, True, True)
        self.time += dt
        self.step_count += 1

    def kinetic_energy(self, dt, force_fn):
        self.set_forces(force_fn(self.positions()))
        return self.o.kinetic_energy(self.precision, self.velm, self.force.reshape(-1), 0.5 * dt,
                                     self.N)


def bond_force(k, r0):
    def f(p):
        d = p[1] - p[0]
        r = np.linalg.norm(d)
        g = k * (r - r0) * d / r
        return np.array([g, -g])
    return f


def bond_energy(k, r0, p):
    return 0.5 * k * (np.linalg.norm(p[1] - p[0]) - r0) ** 2


@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
def test_single_bond_damped_oscillator(oracle, precision):
    """TestCudaConstVLangevinIntegrator.cpp:52-80: T=0, friction 0.1, dt 0.01, two mass-2 particles
    on a k=1, r0=1.5 bond started at separation 2; compare with the analytic damped oscillator
    to 0.02 for 1000 steps."""
    pos = np.array([[-1.0, 0, 10.0], [1.0, 0, 10.0]])
    s = TinySlab(oracle, precision, pos, [2.0, 2.0], [0.0, 0.0])
    f = bond_force(1.0, 1.5)
    freq = np.sqrt(1 - 0.05 * 0.05)
    for _ in range(1000):
        t = s.time
        dist = 1.5 + 0.5 * np.exp(-0.05 * t) * np.cos(freq * t)
        speed = -0.5 * np.exp(-0.05 * t) * (0.05 * np.cos(freq * t) + freq * np.sin(freq * t))
        p = s.positions()
        assert abs(p[0, 0] + 0.5 * dist) < 0.02 and abs(p[1, 0] - 0.5 * dist) < 0.02
        assert abs(s.velm[0, 0] + 0.5 * speed) < 0.02 and abs(s.velm[1, 0] - 0.5 * speed) < 0.02
        s.step(0.0, 0.1, 0.01, f)
    # neutral atoms: images never rewritten, image velocities stay exactly zero
    assert not s.velm[2:4].any()


@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
def test_single_bond_energy_conservation_without_friction(oracle, precision):
    """:84-93: friction 0 -> total energy (half-step-shifted KE + PE) conserved within 1 %."""
    pos = np.array([[-1.0, 0, 10.0], [1.0, 0, 10.0]])
    s = TinySlab(oracle, precision, pos, [2.0, 2.0], [0.0, 0.0])
    f = bond_force(1.0, 1.5)
    e0 = s.kinetic_energy(0.01, f) + bond_energy(1.0, 1.5, s.positions())
    for _ in range(1000):
        e = s.kinetic_energy(0.01, f) + bond_energy(1.0, 1.5, s.positions())
        assert e == pytest.approx(e0, rel=0.01)
        s.step(0.0, 0.0, 0.01, f)


def test_equipartition_temperature(oracle):
    """:96-129: <KE> = 1.5 N kT.  8 charged particles (so the mirror runs every step) in
    independent harmonic wells; friction 2, dt 0.01, T = 100."""
    n, T = 8, 100.0
    rng = np.random.default_rng(0)
    centre = np.column_stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(5, 9, n)])
    s = TinySlab(oracle, "double", centre.copy(), [2.0] * n, [1.0, -1.0] * (n // 2))

    def f(p):
        return -5.0 * (p - centre)

    for _ in range(3000):
        s.step(T, 2.0, 0.01, f, seed=99)
    ke, nsamp = 0.0, 6000
    for _ in range(nsamp):
        ke += s.kinetic_energy(0.01, f)
        s.step(T, 2.0, 0.01, f, seed=99)
    ke /= nsamp
    expected = 0.5 * n * 3 * oracle.BOLTZ * T
    # friction 2/ps -> velocity autocorrelation ~0.5 ps = 50 steps: ~120 independent samples of
    # a chi^2 with 24 dof -> relative sigma ~ sqrt(2/24/120) = 2.6 %; 4 sigma bound
    assert ke == pytest.approx(expected, rel=0.11)
    # images mirrored exactly through the reference expression
    R = s.R
    expect = -s.posq[:R, 2] + 2.0 * s.zmax
    assert np.array_equal(s.posq[R:2 * R, 2], expect)


def test_massless_particles_keep_zero_velocity(oracle):
    """:196-209: a massless particle's velocity stays exactly 0.0 and it does not move."""
    pos = np.array([[-1.0, 0, 3.0], [1.0, 0, 3.0]])
    s = TinySlab(oracle, "single", pos, [0.0, 1.0], [0.5, -0.5])
    p0 = s.posq[0].copy()
    for _ in range(5):
        s.step(300.0, 2.0, 0.01, lambda p: np.ones_like(p) * 100.0)
    assert s.velm[0, 0] == 0.0 and not s.velm[0].any()
    assert np.array_equal(s.posq[0], p0)
    assert s.velm[1, :3].any()
