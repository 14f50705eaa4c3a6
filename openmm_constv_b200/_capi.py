"""ctypes binding of the C ABI in ``include/constv_b200.h`` (``libconstv_b200.so``).

This is the reference-side binding a Python host uses (INTEGRATION.md shows the C++ one).  The
library is the product's only compute path: loading fails loudly when it has not been built, and
every call raises :class:`ConstVError` on a non-zero return code.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libconstv_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_SYSTEM = 0, 1, 2, 3, 4
PRECISION = {"single": 0, "mixed": 1, "double": 2}
MIRROR_REFERENCE, MIRROR_NEGATE = 0, 1
FLAG_ZERO_IMAGE_VELM = 1 << 0
FLAG_ZERO_IMAGE_FORCE = 1 << 1
FLAG_CACHE_IMAGE_CHARGE = 1 << 2
FLAG_PDL = 1 << 3
FLAG_KEEP_IN_L2 = 1 << 4
FORCE_FIXED64, FORCE_FLOAT32 = 0, 1
FLAGS_DEFAULT = FLAG_ZERO_IMAGE_VELM | FLAG_ZERO_IMAGE_FORCE | FLAG_PDL

# every symbol include/constv_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "constv_validate_system", "constv_create", "constv_destroy", "constv_bind_arrays",
    "constv_alloc_arrays", "constv_get_arrays", "constv_set_parameters", "constv_get_parameters",
    "constv_update_inverse_order", "constv_refresh_image_charges", "constv_step_part1",
    "constv_step_part2_mirror", "constv_step_fused", "constv_step_fused_batched",
    "constv_kinetic_energy", "constv_kinetic_energy_device", "constv_generate_noise",
    "constv_get_step_count", "constv_set_step_count", "constv_get_time", "constv_set_time",
    "constv_upload_state", "constv_download_state", "constv_step_host", "constv_step_host_submit",
    "constv_step_host_collect", "constv_step_host_submit_state", "constv_last_error",
    "constv_version", "constv_launch_count", "constv_last_kernel", "constv_set_launch_shape", "constv_set_kernel_variant",
    "constv_set_chains", "constv_get_chains", "constv_step_fused_chain", "constv_step_fused_multi",
    "constv_drude_configure", "constv_drude_random_count", "constv_drude_set_parameters",
    "constv_drude_get_parameters", "constv_drude_step_part1", "constv_drude_step_part2",
    "constv_drude_step_fused", "constv_drude_hardwall",
    "constv_settle_find_clusters", "constv_settle_configure", "constv_step_fused_settle", "constv_settle_positions",
    "constv_step_fused_settle_chain", "constv_drude_step_fused_chain", "constv_step_fused_settle_batched",
    "constv_mirror", "constv_generate_philox_words",
]


class ConstVError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("num_atoms", C.c_int32),
        ("padded_num_atoms", C.c_int32), ("num_cells", C.c_int32), ("precision", C.c_int32),
        ("mirror_mode", C.c_int32), ("flags", C.c_uint32), ("zmax", C.c_double),
        ("seed", C.c_uint64), ("global_atom_offset", C.c_uint64), ("boltz", C.c_double),
    ]


_lib = None


def load() -> C.CDLL:
    """Load the native library; raise if it is missing (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m openmm_constv_b200.build` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "constv_validate_system": [i32, i32, vp, dbl, dbl, C.POINTER(dbl)],
        "constv_create": [C.POINTER(Config), C.POINTER(vp)],
        "constv_destroy": [vp],
        "constv_bind_arrays": [vp, vp, vp, vp, vp, vp, vp],
        "constv_alloc_arrays": [vp],
        "constv_get_arrays": [vp] + [C.POINTER(vp)] * 5,
        "constv_set_parameters": [vp, dbl, dbl, dbl],
        "constv_get_parameters": [vp, C.POINTER(dbl * 4)],
        "constv_update_inverse_order": [vp, vp],
        "constv_refresh_image_charges": [vp, vp],
        "constv_step_part1": [vp, vp, u32, vp],
        "constv_step_part2_mirror": [vp, vp],
        "constv_step_fused": [vp, vp, u32, vp],
        "constv_step_fused_batched": [C.POINTER(vp), i32, vp],
        "constv_kinetic_energy": [vp, dbl, C.POINTER(dbl), vp],
        "constv_kinetic_energy_device": [vp, dbl, vp, vp],
        "constv_generate_noise": [vp, u64, vp, vp],
        "constv_get_step_count": [vp, C.POINTER(u64)],
        "constv_set_step_count": [vp, u64],
        "constv_get_time": [vp, C.POINTER(dbl)],
        "constv_set_time": [vp, dbl],
        "constv_upload_state": [vp, vp, vp, vp, vp, vp],
        "constv_download_state": [vp, vp, vp, vp, vp, vp],
        "constv_step_host": [vp, vp, vp, vp, C.POINTER(dbl), vp],
        "constv_step_host_submit": [vp, vp, i32, C.POINTER(u64), vp],
        "constv_step_host_collect": [vp, u64, C.POINTER(dbl)],
        "constv_step_host_submit_state": [vp, vp, i32, vp, vp, C.POINTER(u64), vp],
        "constv_set_launch_shape": [vp, i32, i32],
        "constv_set_kernel_variant": [vp, i32, i32, i32, i32],
        "constv_set_chains": [vp, i32],
        "constv_get_chains": [vp, C.POINTER(i32)],
        "constv_step_fused_chain": [vp, i32, vp, u32, vp],
        "constv_step_fused_multi": [vp, i32, vp],
        "constv_drude_configure": [vp, i32, vp],
        "constv_drude_random_count": [vp, C.POINTER(i32)],
        "constv_drude_set_parameters": [vp, dbl, dbl, dbl, dbl, dbl, dbl],
        "constv_drude_get_parameters": [vp, C.POINTER(dbl * 9)],
        "constv_drude_step_part1": [vp, vp, u32, vp],
        "constv_drude_step_part2": [vp, vp],
        "constv_drude_step_fused": [vp, vp, u32, vp],
        "constv_drude_hardwall": [vp, vp],
        "constv_settle_find_clusters": [i32, i32, vp, vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)],
        "constv_settle_configure": [vp, i32, vp, vp],
        "constv_step_fused_settle": [vp, vp, u32, vp],
        "constv_settle_positions": [vp, vp],
        "constv_step_fused_settle_chain": [vp, i32, vp, u32, vp],
        "constv_drude_step_fused_chain": [vp, i32, vp, u32, vp],
        "constv_step_fused_settle_batched": [C.POINTER(vp), i32, vp],
        "constv_mirror": [vp, vp],
        "constv_generate_philox_words": [vp, u64, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.constv_last_error.restype = C.c_char_p
    lib.constv_last_error.argtypes = []
    lib.constv_version.restype = C.c_int
    lib.constv_launch_count.restype = C.c_uint64
    lib.constv_last_kernel.restype = C.c_char_p
    lib.constv_last_kernel.argtypes = []
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise ConstVError(rc, load().constv_last_error().decode())


def launch_count() -> int:
    return int(load().constv_launch_count())


def last_kernel() -> str:
    """Kernel behind this thread's most recent launch where the library had a choice ('' otherwise)."""
    return load().constv_last_kernel().decode()


def validate_system(masses, num_cells: int, box_z: float, cell_size: float) -> float:
    """Host-only system checks; raises ConstVError(ERR_SYSTEM) with the reference's message."""
    import numpy as np
    m = np.ascontiguousarray(masses, np.float64)
    z = C.c_double(0.0)
    check(load().constv_validate_system(len(m), num_cells, m.ctypes.data_as(C.c_void_p), box_z,
                                        cell_size, C.byref(z)))
    return z.value
