"""The product path never touches the oracle: nothing under openmm_constv_b200/ (Python, CUDA, the C++ plugin) or
include/ imports, includes, links or executes anything under oracle/; only tests/, __graft_entry__.smoke() and bench.py's
CPU / reference legs may.  A textual check, so that a stray import is caught without a GPU."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = [os.path.join(ROOT, "openmm_constv_b200"), os.path.join(ROOT, "include")]
SOURCE_EXT = (".py", ".cu", ".cuh", ".inc", ".cpp", ".h", ".hpp", ".i", ".sh", "Makefile")

IMPORTS = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|#\s*include\s*[\"<][^\">]*oracle[^\">]*[\">]|libconstv_oracle|oracle/_ref|constv_oracle", re.M)


def product_sources():
    for top in PRODUCT:
        for dirpath, dirnames, filenames in os.walk(top):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "bin", "lib", "build")]
            for name in filenames:
                if name.endswith(SOURCE_EXT):
                    yield os.path.join(dirpath, name)


def test_no_product_source_reaches_into_the_oracle():
    offenders = []
    for path in product_sources():
        with open(path, errors="replace") as fh:
            text = fh.read()
        for m in IMPORTS.finditer(text):
            offenders.append(f"{os.path.relpath(path, ROOT)}: {m.group(0).strip()}")
    assert not offenders, offenders


def test_the_native_library_is_required():
    """_capi.load() raises when the CUDA library is missing instead of falling back."""
    from openmm_constv_b200 import _capi
    saved_path, saved_lib = _capi.LIB_PATH, _capi._lib
    try:
        _capi._lib = None
        _capi.LIB_PATH = os.path.join(ROOT, "openmm_constv_b200", "no_such_library.so")
        try:
            _capi.load()
        except ImportError as exc:
            assert "no CPU fallback" in str(exc).replace("There is no", "no")
        else:
            raise AssertionError("load() returned without the library")
    finally:
        _capi.LIB_PATH, _capi._lib = saved_path, saved_lib
