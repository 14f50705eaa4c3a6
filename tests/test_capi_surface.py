"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares (and nothing else), its host-only entry points behave like the reference's host
code, and the product has no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from openmm_constv_b200 import _capi, build
    build.build_capi()
    _capi.load()
    return _capi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "constv_b200.h")).read()
    return sorted(set(re.findall(r"CONSTV_API\s+[\w\s\*]+?\b(constv_\w+)\s*\(", text)))


def test_library_exports_exactly_the_declared_symbols(capi):
    declared = _declared_symbols()
    assert declared == sorted(capi.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "constv_" in l)
    assert exported == declared
    lib = capi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.constv_version() == 121


def test_header_cites_the_reference_for_every_entry_point():
    text = open(os.path.join(ROOT, "include", "constv_b200.h")).read()
    assert text.count("Kernels.cpp:") >= 12 and "Langevin.cu:" in text


def test_library_is_sm100a_only(capi):
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_validate_system_matches_reference_messages(capi):
    m = np.array([1.0, 2.0, 0.0, 0.0])
    assert capi.validate_system(m, 2, 8.0, -1.0) == 4.0
    assert capi.validate_system(m, 2, 8.0, 4.0) == 4.0
    cases = [
        (m[:3], 3, 8.0, -1.0, "Not even number of cells"),
        (m[:3], 2, 8.0, -1.0, "Number of particles is not multiple of number of cells"),
        (np.array([1.0, 2.0, 0.0, 1.0]), 2, 8.0, -1.0, "Image Particle has nonzero mass"),
        (m, 2, 8.0, 3.9, "Unit cell dimension does not match with the provided zmax value"),
    ]
    for masses, cells, box, cell, msg in cases:
        with pytest.raises(capi.ConstVError) as ei:
            capi.validate_system(masses, cells, box, cell)
        assert ei.value.code == capi.ERR_SYSTEM and str(ei.value) == msg


def test_validate_system_agrees_with_oracle(capi, oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 9))
        cells = int(rng.integers(1, 5))
        masses = rng.choice([0.0, 1.0], n)
        box = float(rng.choice([8.0, 7.9]))
        cell = float(rng.choice([-1.0, 4.0, 2.0]))
        try:
            expect = oracle.validate(n, cells, masses, box, cell)
            err = None
        except ValueError as e:
            expect, err = None, str(e)
        try:
            got = capi.validate_system(masses, cells, box, cell)
            assert err is None and got == expect
        except capi.ConstVError as e:
            assert str(e) == err


def test_no_cpu_fallback(capi):
    """Without a CUDA device the compute entry points fail with CONSTV_ERR_CUDA -- they never
    compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = capi.Config()
    cfg.struct_size = C.sizeof(capi.Config)
    cfg.device, cfg.num_atoms, cfg.padded_num_atoms, cfg.num_cells = 0, 64, 64, 2
    cfg.precision, cfg.mirror_mode, cfg.flags, cfg.zmax = 0, 0, capi.FLAGS_DEFAULT, 4.0
    h = C.c_void_p(None)
    rc = capi.load().constv_create(C.byref(cfg), C.byref(h))
    assert rc == capi.ERR_CUDA and not h.value
    from openmm_constv_b200.integrator import ConstVLangevinIntegrator, SlabContext, SlabSystem
    sys_ = SlabSystem()
    sys_.addParticles([1.0, 0.0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SlabContext(sys_, ConstVLangevinIntegrator(300, 1, 0.001))


def test_create_rejects_bad_configs(capi):
    lib = capi.load()
    h = C.c_void_p(None)

    def cfg(**kw):
        c = capi.Config()
        c.struct_size = C.sizeof(capi.Config)
        c.device, c.num_atoms, c.padded_num_atoms, c.num_cells = 0, 64, 64, 2
        c.precision, c.mirror_mode, c.flags, c.zmax = 0, 0, capi.FLAGS_DEFAULT, 4.0
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    for bad, code in [(cfg(struct_size=8), capi.ERR_INVALID), (cfg(num_cells=3), capi.ERR_SYSTEM),
                      (cfg(num_atoms=63), capi.ERR_SYSTEM), (cfg(padded_num_atoms=32), capi.ERR_INVALID),
                      (cfg(precision=7), capi.ERR_INVALID), (cfg(zmax=-1.0), capi.ERR_INVALID),
                      (cfg(mirror_mode=1, num_cells=4), capi.ERR_INVALID)]:
        assert lib.constv_create(C.byref(bad), C.byref(h)) == code
        assert lib.constv_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under openmm_constv_b200/ may reference it."""
    pkg = os.path.join(ROOT, "openmm_constv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "libconstv_oracle" not in text, f
                assert "constv_oracle" not in text, f


def test_integrator_surface_matches_reference():
    """Same constructor signature, accessors and defaults as constvplugin.i:65-80 /
    ConstVLangevinIntegrator.cpp:44-52, and the unbound-step error (cpp:78-79)."""
    from openmm_constv_b200.integrator import ConstVLangevinIntegrator, OpenMMException
    it = ConstVLangevinIntegrator(372.4, 1.234, 0.0018)
    assert (it.getTemperature(), it.getFriction(), it.getStepSize()) == (372.4, 1.234, 0.0018)
    assert it.getNumCells() == 2 and it.getCellSize() == -1
    assert it.getConstraintTolerance() == 1e-5 and it.getRandomNumberSeed() == 0
    it2 = ConstVLangevinIntegrator(300, 1, 0.001, 4, 2.5)
    assert it2.getNumCells() == 4 and it2.getCellSize() == 2.5
    for name in ("NumCells", "CellSize", "Temperature", "Friction", "RandomNumberSeed"):
        assert hasattr(it, "get" + name) and hasattr(it, "set" + name)
    assert it.getKernelNames() == ["IntegrateConstVLangevinStep"]
    with pytest.raises(OpenMMException, match="This Integrator is not bound to a context!"):
        it.step(1)
