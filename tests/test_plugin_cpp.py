"""The OpenMM-facing C++ plugin (openmm_constv_b200/plugin): builds against the OpenMM shim and runs
its C++ tests, which are modelled on the reference's own
(serialization/tests/TestSerializeConstVIntegrator.cpp, platforms/cuda/tests/
TestCudaConstVLangevinIntegrator.cpp, run once per precision as tests/CMakeLists.txt:22-24 does)."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "openmm_constv_b200", "plugin")


@pytest.fixture(scope="module")
def plugin_built():
    from openmm_constv_b200 import build
    build.build_all()
    return PLUGIN


def run(binary, *args, timeout=600):
    res = subprocess.run([os.path.join(PLUGIN, "bin", binary), *args], capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, f"{binary} {' '.join(args)} failed:\n{res.stdout}\n{res.stderr}"
    assert "Done" in res.stdout
    return res.stdout


def test_serialization_round_trip(plugin_built):
    run("TestSerializeConstVIntegrator")


def test_plugin_surface_without_gpu(plugin_built):
    run("TestConstVPluginSurface")


def test_plugin_exports_the_reference_symbols(plugin_built):
    """CudaConstVKernelFactory.cpp:38-61 and ConstVSerializationProxyRegistration.cpp:63-65.  In a child
    process: loading the shim RTLD_GLOBAL brings the system libcudart into this process's global symbol
    scope, which a later `import torch` (with its own bundled CUDA runtime) does not survive."""
    code = f"""
import ctypes, os
P = {PLUGIN!r}
ctypes.CDLL(os.path.join(P, "lib", "libOpenMMShim.so"), mode=ctypes.RTLD_GLOBAL)
api = ctypes.CDLL(os.path.join(P, "lib", "libOpenMMConstV.so"), mode=ctypes.RTLD_GLOBAL)
assert hasattr(api, "registerConstVSerializationProxies")
plugin = ctypes.CDLL(os.path.join(P, "lib", "plugins", "libConstVPluginCUDA.so"))
for name in ("registerPlatforms", "registerKernelFactories", "registerConstVCudaKernelFactories"):
    assert hasattr(plugin, name), name
print("ok")
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


def test_constvplugin_python_module():
    """python/constvplugin.i:65-80: module name, class, constructor defaults, accessors."""
    sys.path.insert(0, os.path.join(PLUGIN, "python"))
    try:
        import constvplugin
    finally:
        sys.path.pop(0)
    integ = constvplugin.ConstVLangevinIntegrator(300.0, 1.0, 0.001)
    assert integ.getNumCells() == 2 and integ.getCellSize() == -1
    assert integ.getRandomNumberSeed() == 0 and integ.getConstraintTolerance() == 1e-5
    for setter, getter, value in (("setNumCells", "getNumCells", 4), ("setCellSize", "getCellSize", 4.2183),
                                  ("setTemperature", "getTemperature", 310.0), ("setFriction", "getFriction", 0.5),
                                  ("setRandomNumberSeed", "getRandomNumberSeed", 42)):
        getattr(integ, setter)(value)
        got = getattr(integ, getter)()
        got = getattr(got, "_value", got)  # Quantity when an OpenMM unit package exists
        assert got == value
    with pytest.raises(constvplugin.OpenMMException, match="not bound to a context"):
        integ.step(1)
    text = open(os.path.join(PLUGIN, "python", "constvplugin.i")).read()
    assert "%module constvplugin" in text and "class ConstVLangevinIntegrator : public Integrator" in text


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
def test_cuda_integrator_through_the_plugin(plugin_built, precision):
    out = run("TestCudaConstVLangevinIntegrator", precision, timeout=900)
    assert f"Done ({precision})" in out


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
def test_cuda_drude_integrator_through_the_plugin(plugin_built, precision):
    """ConstVDrudeLangevinIntegrator -> IntegrateConstVDrudeLangevinStepKernel -> constv_drude_* (the reference
    has no test for this integrator; properties of its algorithm, see the C++ file's header)."""
    out = run("TestCudaConstVDrudeLangevinIntegrator", precision, timeout=900)
    assert f"Done ({precision})" in out


def test_constvplugin_python_module_drude():
    """python/constvplugin.i:82-100: ConstVDrudeLangevinIntegrator(temperature, friction, drudeTemperature,
    drudeFriction, stepSize, numCells=2, zmax=-1) and its accessors."""
    sys.path.insert(0, os.path.join(PLUGIN, "python"))
    try:
        import constvplugin
    finally:
        sys.path.pop(0)
    integ = constvplugin.ConstVDrudeLangevinIntegrator(300.0, 5.0, 1.0, 20.0, 0.001)
    assert integ.getNumCells() == 2 and integ.getCellSize() == -1 and integ.getMaxDrudeDistance() == 0
    for setter, getter, value in (("setDrudeTemperature", "getDrudeTemperature", 2.0), ("setDrudeFriction", "getDrudeFriction", 10.0),
                                  ("setMaxDrudeDistance", "getMaxDrudeDistance", 0.02), ("setTemperature", "getTemperature", 310.0),
                                  ("setFriction", "getFriction", 0.5), ("setNumCells", "getNumCells", 4)):
        getattr(integ, setter)(value)
        got = getattr(integ, getter)()
        got = getattr(got, "_value", got)
        assert got == value
    with pytest.raises(constvplugin.OpenMMException, match="Distance cannot be negative"):
        integ.setMaxDrudeDistance(-1)
    with pytest.raises(constvplugin.OpenMMException, match="not bound to a context"):
        integ.step(1)
    assert integ.getKernelNames() == ["IntegrateConstVDrudeLangevinStep"]
    text = open(os.path.join(PLUGIN, "python", "constvplugin.i")).read()
    assert "class ConstVDrudeLangevinIntegrator : public Integrator" in text


def _surfaces():
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_swig_surface", os.path.join(ROOT, "tests", "golden", "make_swig_surface.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ours = mod.swig_surface(open(os.path.join(PLUGIN, "python", "constvplugin.i")).read())
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_swig_surface.json")))
    return ours, ref


def test_swig_interface_matches_the_reference_surface():
    """constvplugin.i declares the reference's module surface (python/constvplugin.i: module name, vector
    templates, the two classes with their constructors, defaults, accessors and the SEVEN unit-decorated
    getters) and nothing that changes the behaviour of a reference call.  The only additions are
    get/setRandomNumberSeed on the Drude class, which the reference's C++ class has but its .i forgot."""
    ours, ref = _surfaces()
    assert ours["module"] == ref["module"] == "constvplugin"
    assert ours["templates"] == ref["templates"]
    assert ours["unit_getters"] == ref["unit_getters"]  # in particular: getCellSize is NOT decorated
    assert set(ours["classes"]) == set(ref["classes"])
    for name, rc in ref["classes"].items():
        oc = ours["classes"][name]
        assert oc["base"] == rc["base"]
        assert [(t, d) for t, _, d in oc["constructor"]] == [(t, d) for t, _, d in rc["constructor"]]
        for method, sig in rc["methods"].items():
            assert oc["methods"][method] == sig, (name, method)
        extra = set(oc["methods"]) - set(rc["methods"])
        assert extra <= {"getRandomNumberSeed", "setRandomNumberSeed"}, extra
    if os.path.exists("/root/reference/python/constvplugin.i"):  # the fixture is what the reference says today
        import importlib.util
        spec = importlib.util.spec_from_file_location("mss", os.path.join(ROOT, "tests", "golden", "make_swig_surface.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.swig_surface(open("/root/reference/python/constvplugin.i").read()) == ref


def test_ctypes_module_has_the_swig_surface():
    """The ctypes constvplugin module (what runs without SWIG/OpenMM) offers every class, constructor default
    and method the .i declares, decorates exactly the getters the .i decorates, and nothing else."""
    ours, _ = _surfaces()
    sys.path.insert(0, os.path.join(PLUGIN, "python"))
    try:
        import constvplugin
    finally:
        sys.path.pop(0)
    import inspect
    for name, info in ours["classes"].items():
        cls = getattr(constvplugin, name)
        params = list(inspect.signature(cls.__init__).parameters.values())[1:]
        assert len(params) == len(info["constructor"])
        for p, (typ, _, default) in zip(params, info["constructor"]):
            if default is None:
                assert p.default is inspect.Parameter.empty
            else:
                assert p.default == (int(default) if typ == "int" else float(default))
        for method in info["methods"]:
            assert callable(getattr(cls, method)), (name, method)
        decorated = {m for m in info["methods"] if f"{name}::{m}" in ours["unit_getters"]}
        # a getter is decorated in the ctypes module iff the class overrides it (constvplugin.py wraps the host getters)
        overridden = {m for m in info["methods"] if m.startswith("get") and m in cls.__dict__}
        assert overridden == decorated, (name, overridden, decorated)
