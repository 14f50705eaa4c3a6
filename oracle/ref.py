"""ctypes front-end of oracle/_ref/: the REFERENCE's own kernel source
(/root/reference/platforms/cuda/src/kernels/constVLangevin.cu) compiled unmodified

  * by g++ for the host with the CUDA execution model emulated (``libconstv_ref_host.so``), and
  * by nvcc for sm_100a behind plain C launchers (``libconstv_ref_cuda.so``; ``_fastmath`` is the
    same with --use_fast_math).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's reference legs).  The libraries are built in
the container that has /root/reference (``oracle.build_ref()``) and travel with the tree.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SUFFIX = {"single": "single", "mixed": "mixed", "double": "double"}
REAL = {"single": np.float32, "mixed": np.float32, "double": np.float64}
MIXED = {"single": np.float32, "mixed": np.float64, "double": np.float64}

_host = None
_cuda = {}


def available(kind: str = "host") -> bool:
    name = {"host": "libconstv_ref_host.so", "cuda": "libconstv_ref_cuda.so",
            "cuda_fastmath": "libconstv_ref_cuda_fastmath.so"}[kind]
    return os.path.exists(os.path.join(REF_DIR, name))


def host() -> C.CDLL:
    global _host
    if _host is None:
        _host = C.CDLL(os.path.join(REF_DIR, "libconstv_ref_host.so"))
        _host.constv_ref_host_num_threads.restype = C.c_int
    return _host


def cuda(fastmath: bool = False) -> C.CDLL:
    if fastmath not in _cuda:
        name = "libconstv_ref_cuda_fastmath.so" if fastmath else "libconstv_ref_cuda.so"
        _cuda[fastmath] = C.CDLL(os.path.join(REF_DIR, name))
    return _cuda[fastmath]


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def host_arrays(precision, prm, step_size):
    """The two little device arrays the kernels read (CudaConstVKernels.cpp:71,122-137 and
    integration.getStepSize()): params[3] and dt = (0, stepSize), both `mixed`."""
    m = MIXED[precision]
    return np.ascontiguousarray(prm, m), np.array([0, step_size], m)


def host_step(precision, num_atoms, num_cells, zmax, posq, corr, velm, force, pos_delta, prm, step_size,
              random4, random_index, inv_order):
    """One constraint-free step of the reference (part 1, part 2, image rewrite) on the host."""
    P = velm.shape[0]
    params, dt = host_arrays(precision, prm, step_size)
    if corr is None:  # the kernels take the pointer in every mode (unused unless mixed)
        corr = np.zeros((1, 4), REAL[precision])
    fn = getattr(host(), "constv_ref_host_step_" + SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(P), C.c_int(num_cells), C.c_double(zmax), _p(posq), _p(corr), _p(velm),
       _p(force), _p(pos_delta), _p(params), _p(dt), _p(random4), C.c_uint(random_index), _p(inv_order))


def host_step_range(precision, num_atoms, num_cells, zmax, posq, corr, velm, force, pos_delta, prm, step_size,
                    random4, random_index, inv_order, lo, hi):
    """The step restricted to real atoms [lo, hi) and their images (identity order)."""
    P = velm.shape[0]
    params, dt = host_arrays(precision, prm, step_size)
    if corr is None:
        corr = np.zeros((1, 4), REAL[precision])
    fn = getattr(host(), "constv_ref_host_step_range_" + SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(P), C.c_int(num_cells), C.c_double(zmax), _p(posq), _p(corr), _p(velm),
       _p(force), _p(pos_delta), _p(params), _p(dt), _p(random4), C.c_uint(random_index), _p(inv_order),
       C.c_int(lo), C.c_int(hi))


def host_part1(precision, num_atoms, velm, force, pos_delta, prm, step_size, random4, random_index):
    params, dt = host_arrays(precision, prm, step_size)
    fn = getattr(host(), "constv_ref_host_part1_" + SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(velm.shape[0]), _p(velm), _p(force), _p(pos_delta), _p(params), _p(dt),
       _p(random4), C.c_uint(random_index))


def host_part2(precision, num_atoms, posq, corr, pos_delta, velm, step_size):
    _, dt = host_arrays(precision, np.zeros(3), step_size)
    if corr is None:
        corr = np.zeros((1, 4), REAL[precision])
    fn = getattr(host(), "constv_ref_host_part2_" + SUFFIX[precision])
    fn(C.c_int(num_atoms), _p(posq), _p(corr), _p(pos_delta), _p(velm), _p(dt))


def host_mirror(precision, num_real, num_cells, zmax, posq, corr, inv_order):
    if corr is None:
        corr = np.zeros((1, 4), REAL[precision])
    fn = getattr(host(), "constv_ref_host_mirror_" + SUFFIX[precision])
    fn(C.c_int(num_real), C.c_int(num_cells), C.c_double(zmax), _p(posq), _p(corr), _p(inv_order))


def host_inverse_order(atom_index):
    atom_index = np.ascontiguousarray(atom_index, np.int32)
    inv = np.empty_like(atom_index)
    host().constv_ref_host_inverse_order_single(C.c_int(len(atom_index)), _p(atom_index), _p(inv))
    return inv


def host_drude_step(precision, num_atoms, num_cells, zmax, posq, corr, velm, force, pos_delta, normal, pairs,
                    scalars, step_size, random4, random_index, inv_order):
    """One constraint-free step of the reference's Drude integrator on the host, in its launch order
    (CudaConstVKernels.cpp:317-349): Drude part 1, part 2, hard wall (if maxDrudeDistance > 0),
    image rewrite.  scalars = oracle.drude_params(...) (doubles; narrowed to `mixed` by the launcher
    as CudaConstVKernels.cpp:284-312 does)."""
    h = host()
    sfx = SUFFIX[precision]
    P = velm.shape[0]
    num_real = num_atoms // num_cells
    _, dt = host_arrays(precision, np.zeros(3), step_size)
    if corr is None:
        corr = np.zeros((1, 4), REAL[precision])
    sc = [C.c_double(float(x)) for x in scalars]
    getattr(h, "constv_ref_drude_host_configure_" + sfx)(C.c_int(num_real), C.c_int(P), C.c_int(len(normal)),
                                                          C.c_int(len(pairs)))
    getattr(h, "constv_ref_drude_host_part1_" + sfx)(_p(velm), _p(force), _p(pos_delta), _p(normal), _p(pairs), _p(dt),
                                                      *sc[:6], _p(random4), C.c_uint(random_index))
    getattr(h, "constv_ref_drude_host_part2_" + sfx)(_p(posq), _p(corr), _p(pos_delta), _p(velm), _p(dt))
    if scalars[6] > 0:
        getattr(h, "constv_ref_drude_host_hardwall_" + sfx)(_p(posq), _p(corr), _p(velm), _p(pairs), _p(dt), sc[6], sc[7])
    host_mirror(precision, num_real, num_cells, zmax, posq, corr, inv_order)


def host_num_threads() -> int:
    return host().constv_ref_host_num_threads()


def host_set_num_threads(n: int) -> None:
    host().constv_ref_host_set_num_threads(C.c_int(n))


class CudaReference:
    """The reference's kernels on the GPU: holds the two small parameter arrays on the device and
    launches part 1 / part 2 / image rewrite on torch tensors laid out as OpenMM's arrays."""

    def __init__(self, precision, prm, step_size, device, fastmath=False, max_blocks=0):
        import torch
        self.torch = torch
        self.lib = cuda(fastmath)
        self.precision = precision
        self.max_blocks = int(max_blocks)
        params, dt = host_arrays(precision, prm, step_size)
        self.params = torch.as_tensor(params, device=device)
        self.dt = torch.as_tensor(dt, device=device)
        self.dummy_corr = torch.zeros((1, 4), dtype=torch.float32 if precision != "double" else torch.float64,
                                      device=device)
        sfx = SUFFIX[precision]
        self._part1 = getattr(self.lib, "constv_ref_cuda_part1_" + sfx)
        self._part2 = getattr(self.lib, "constv_ref_cuda_part2_" + sfx)
        self._mirror = getattr(self.lib, "constv_ref_cuda_mirror_" + sfx)
        self._inverse = getattr(self.lib, "constv_ref_cuda_inverse_order_" + sfx)
        vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint
        self._part1.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, u32, i32, vp]
        self._part2.argtypes = [i32, vp, vp, vp, vp, vp, i32, vp]
        self._mirror.argtypes = [i32, i32, C.c_double, vp, vp, vp, i32, vp]
        self._inverse.argtypes = [i32, vp, vp, i32, vp]

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise RuntimeError(f"reference kernel launch failed: cudaError {rc}")

    def part1(self, num_atoms, velm, force, pos_delta, random4, random_index, stream=0):
        self._check(self._part1(num_atoms, velm.shape[0], velm.data_ptr(), force.data_ptr(), pos_delta.data_ptr(),
                                self.params.data_ptr(), self.dt.data_ptr(), random4.data_ptr(), random_index,
                                self.max_blocks, stream))

    def part2(self, num_atoms, posq, corr, pos_delta, velm, stream=0):
        corr = self.dummy_corr if corr is None else corr
        self._check(self._part2(num_atoms, posq.data_ptr(), corr.data_ptr(), pos_delta.data_ptr(), velm.data_ptr(),
                                self.dt.data_ptr(), self.max_blocks, stream))

    def mirror(self, num_real, num_cells, zmax, posq, corr, inv_order, stream=0):
        corr = self.dummy_corr if corr is None else corr
        self._check(self._mirror(num_real, num_cells, zmax, posq.data_ptr(), corr.data_ptr(), inv_order.data_ptr(),
                                 self.max_blocks, stream))

    def inverse_order(self, num_atoms, atom_index, inv_order, stream=0):
        self._check(self._inverse(num_atoms, atom_index.data_ptr(), inv_order.data_ptr(), self.max_blocks, stream))

    def step(self, num_atoms, num_cells, zmax, posq, corr, velm, force, pos_delta, random4, random_index, inv_order,
             stream=0):
        """CudaConstVKernels.cpp:146-166 without constraints: three launches."""
        self.part1(num_atoms, velm, force, pos_delta, random4, random_index, stream)
        self.part2(num_atoms, posq, corr, pos_delta, velm, stream)
        self.mirror(num_atoms // num_cells, num_cells, zmax, posq, corr, inv_order, stream)


class CudaDrudeReference(CudaReference):
    """The reference's Drude kernels on the GPU (oracle/ref_drude_cuda.cu) plus the shared image
    rewrite: the launch sequence of CudaConstVKernels.cpp:317-349 without constraints."""

    def __init__(self, precision, scalars, step_size, device, normal, pairs, num_real, padded, max_blocks=0):
        super().__init__(precision, np.zeros(3), step_size, device, max_blocks=max_blocks)
        torch = self.torch
        self.scalars = [float(x) for x in scalars]
        self.normal = torch.as_tensor(np.ascontiguousarray(normal, np.int32), device=device)
        self.pairs = torch.as_tensor(np.ascontiguousarray(pairs, np.int32).reshape(-1, 2), device=device)
        self.num_real, self.num_normal, self.num_pairs = int(num_real), len(normal), len(pairs)
        sfx = SUFFIX[precision]
        vp, i32, u32, dbl = C.c_void_p, C.c_int, C.c_uint, C.c_double
        self._configure = getattr(self.lib, "constv_ref_drude_cuda_configure_" + sfx)
        self._dpart1 = getattr(self.lib, "constv_ref_drude_cuda_part1_" + sfx)
        self._dpart2 = getattr(self.lib, "constv_ref_drude_cuda_part2_" + sfx)
        self._hardwall = getattr(self.lib, "constv_ref_drude_cuda_hardwall_" + sfx)
        self._configure.argtypes = [i32, i32, i32, i32]
        self._dpart1.argtypes = [i32, vp, vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, dbl, vp, u32, i32, vp]
        self._dpart2.argtypes = [i32, vp, vp, vp, vp, vp, i32, vp]
        self._hardwall.argtypes = [i32, vp, vp, vp, vp, vp, dbl, dbl, i32, vp]
        self._check(self._configure(self.num_real, int(padded), self.num_normal, self.num_pairs))

    def drude_step(self, num_atoms, num_cells, zmax, posq, corr, velm, force, pos_delta, random4, random_index,
                   inv_order, stream=0):
        c = self.dummy_corr if corr is None else corr
        s = self.scalars
        self._check(self._dpart1(self.num_real, velm.data_ptr(), force.data_ptr(), pos_delta.data_ptr(),
                                 self.normal.data_ptr(), self.pairs.data_ptr(), self.dt.data_ptr(), *s[:6],
                                 random4.data_ptr(), random_index, self.max_blocks, stream))
        self._check(self._dpart2(self.num_real, posq.data_ptr(), c.data_ptr(), pos_delta.data_ptr(), velm.data_ptr(),
                                 self.dt.data_ptr(), self.max_blocks, stream))
        if s[6] > 0 and self.num_pairs > 0:
            self._check(self._hardwall(self.num_pairs, posq.data_ptr(), c.data_ptr(), velm.data_ptr(),
                                       self.pairs.data_ptr(), self.dt.data_ptr(), s[6], s[7], self.max_blocks, stream))
        self.mirror(num_atoms // num_cells, num_cells, zmax, posq, corr, inv_order, stream)
