"""The oracle's ConstVDrudeLangevinIntegrator restatement against the REFERENCE'S OWN Drude kernel
source compiled for the host (oracle/ref_drude_host.cpp: vectorOps.cu + constVDrudeLangevin.cu,
unmodified): bit for bit on every array with the uncontracted oracle build, three precision
models, with and without the hard wall."""
import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H

T, FRICTION, TD, FRICTION_D, DT = 300.0, 5.0, 1.0, 20.0, 0.001


@pytest.fixture(scope="module")
def ref(oracle):
    oracle.build_ref()
    from oracle import ref as r
    if not r.available("host"):
        pytest.skip("oracle/_ref/libconstv_ref_host.so is not built (needs /root/reference)")
    return r


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("max_dist", [0.0, 0.02])
def test_drude_oracle_matches_reference_source_bit_for_bit(oracle, ref, precision, num_cells, max_dist):
    s, normal, pairs = slabmod.make_drude_slab(2001, precision, num_cells=num_cells, seed=40 + num_cells)
    a, b = s.copy(), s.copy()
    scalars = oracle.drude_params(T, FRICTION, TD, FRICTION_D, max_dist, DT)
    m = H.MIXED[precision]
    dt = m(DT)
    inv = H.identity_order(s.padded)
    da, db = np.zeros((s.padded, 4), m), np.zeros((s.padded, 4), m)
    rng = np.random.default_rng(3)
    bounced = 0
    for k in range(4):
        xi = rng.standard_normal((s.padded + 2 * len(pairs) + 11, 4)).astype(np.float32)
        before = a.velm.copy()
        ref.host_drude_step(precision, a.num_atoms, num_cells, a.zmax, a.posq, a.corr, a.velm, a.force.reshape(-1), da,
                            normal, pairs, scalars, dt, xi, 11, inv)
        with oracle.uncontracted():
            oracle.drude_step(precision, b.num_atoms, num_cells, b.zmax, b.posq, b.corr, b.velm, b.force.reshape(-1), db,
                              normal, pairs, scalars.astype(m), dt, xi, 11, inv, False, False)
        H.assert_bits_equal(a.velm, b.velm, f"velm step {k}")
        H.assert_bits_equal(a.posq, b.posq, f"posq step {k}")
        H.assert_bits_equal(da, db, f"posDelta step {k}")
        if a.corr is not None:
            H.assert_bits_equal(a.corr, b.corr, f"posqCorrection step {k}")
        assert not np.array_equal(before, a.velm)
        d = (a.posq[pairs[:, 0], :3].astype(np.float64) - a.posq[pairs[:, 1], :3]).astype(np.float64)
        bounced += int(np.sum(np.linalg.norm(d, axis=1) > 0.02))
    assert np.all(np.isfinite(a.velm)) and np.all(np.isfinite(a.posq))
    if max_dist > 0:  # the wall is really exercised, and holds to rounding after each step
        d = a.posq[pairs[:, 0], :3].astype(np.float64) - a.posq[pairs[:, 1], :3]
        assert np.max(np.linalg.norm(d, axis=1)) < 0.02 * (1 + 1e-3)
    else:
        assert bounced > 0


def test_hardwall_with_massless_parent_matches_reference_source(oracle, ref):
    """constVDrudeLangevin.cu:141-165: only the Drude particle moves when its parent is massless
    (part 1 would produce NaN for such a pair, so the kernel is driven alone)."""
    import ctypes as C
    for precision in slabmod.PRECISIONS:
        s, normal, pairs = slabmod.make_drude_slab(801, precision, seed=7)
        s.velm[pairs[:, 1], 3] = 0  # massless parents
        s.velm[pairs[:, 1], :3] = 0
        a, b = s.copy(), s.copy()
        sfx = ref.SUFFIX[precision]
        m = H.MIXED[precision]
        h = ref.host()
        _, dt = ref.host_arrays(precision, np.zeros(3), DT)
        corr_a = a.corr if a.corr is not None else np.zeros((1, 4), H.REAL[precision])
        getattr(h, "constv_ref_drude_host_configure_" + sfx)(C.c_int(a.num_real), C.c_int(a.padded), C.c_int(len(normal)),
                                                              C.c_int(len(pairs)))
        getattr(h, "constv_ref_drude_host_hardwall_" + sfx)(ref._p(a.posq), ref._p(corr_a), ref._p(a.velm), ref._p(pairs),
                                                             ref._p(dt), C.c_double(0.01), C.c_double(0.09))
        with oracle.uncontracted():
            oracle.drude_hardwall(precision, b.posq, b.corr, b.velm, pairs, m(DT), 0.01, 0.09)
        H.assert_bits_equal(a.posq, b.posq, "posq")
        H.assert_bits_equal(a.velm, b.velm, "velm")
        assert not np.array_equal(a.posq, s.posq)
        assert np.array_equal(a.posq[pairs[:, 1]], s.posq[pairs[:, 1]])  # parents did not move


def test_drude_parameters(oracle):
    """CudaConstVKernels.cpp:262-270."""
    p = oracle.drude_params(300.0, 5.0, 1.0, 20.0, 0.02, 0.001)
    kB = 8.31446261815324e-3
    assert p[0] == np.exp(-0.001 * 5.0) and p[3] == np.exp(-0.001 * 20.0)
    assert p[1] == (1 - p[0]) / 5.0 / 4294967296.0
    assert p[2] == np.sqrt(2 * kB * 300.0 * 5.0) * np.sqrt(0.5 * (1 - p[0] * p[0]) / 5.0)
    assert p[6] == 0.02 and p[7] == np.sqrt(kB * 1.0)


import glob  # noqa: E402
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "drude_ref_host_*.npz"))))
def test_drude_oracle_matches_committed_reference_vectors(oracle, path):
    """Fixtures recorded from the reference's Drude kernel source compiled for the host
    (tests/golden/make_drude_reference_vectors.py): the same check where /root/reference is absent."""
    g = np.load(path)
    precision, num_cells, num_atoms, steps = str(g["precision"]), int(g["num_cells"]), int(g["num_atoms"]), int(g["steps"])
    m = H.MIXED[precision]
    posq, velm, force = g["posq0"].copy(), g["velm0"].copy(), g["force"].copy()
    corr = g["corr0"].copy() if precision == "mixed" else None
    inv = H.identity_order(posq.shape[0])
    delta = np.zeros((posq.shape[0], 4), m)
    with oracle.uncontracted():
        for k in range(steps):
            oracle.drude_step(precision, num_atoms, num_cells, float(g["zmax"]), posq, corr, velm, force.reshape(-1), delta,
                              g["normal"], g["pairs"], g["scalars"].astype(m), m(g["dt"]), g["noise"][k], int(g["random_index"]),
                              inv, False, False)
    H.assert_bits_equal(posq, g["posq1"], "posq")
    H.assert_bits_equal(velm, g["velm1"], "velm")
    if corr is not None:
        H.assert_bits_equal(corr, g["corr1"], "posqCorrection")
    assert not np.array_equal(g["velm0"], g["velm1"])
