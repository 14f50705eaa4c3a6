"""Builds the SWIG `constvplugin` module against a real OpenMM install.

    swig -python -c++ -o ConstVPluginWrapper.cpp -I$OPENMM_DIR/include constvplugin.i
    OPENMM_DIR=... CONSTV_PLUGIN_DIR=<openmm_constv_b200/plugin> python setup.py build_ext --inplace

(the two steps the reference's python/CMakeLists.txt drives).  Not runnable in the build container:
SWIG and OpenMM are absent there; see constvplugin.py for the module the tests use.
"""
import os
from setuptools import Extension, setup

openmm_dir = os.environ.get("OPENMM_DIR", "/usr/local/openmm")
plugin_dir = os.environ.get("CONSTV_PLUGIN_DIR", os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

extension = Extension(
    name="_constvplugin",
    sources=["ConstVPluginWrapper.cpp"],
    libraries=["OpenMM", "OpenMMConstV"],
    include_dirs=[os.path.join(openmm_dir, "include"), os.path.join(plugin_dir, "openmmapi", "include")],
    library_dirs=[os.path.join(openmm_dir, "lib"), os.path.join(plugin_dir, "lib")],
    runtime_library_dirs=[os.path.join(openmm_dir, "lib"), os.path.join(plugin_dir, "lib")],
)

setup(name="constvplugin", version="1.0", py_modules=["constvplugin"], ext_modules=[extension])
