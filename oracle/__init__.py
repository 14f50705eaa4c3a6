"""CPU oracle for the ConstVLangevinIntegrator step path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/libconstv_oracle.so`` (built from ``constv_oracle.c`` by
``make -C oracle``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
``openmm_constv_b200`` never does.

The functions restate /root/reference/platforms/cuda/src/kernels/constVLangevin.cu and the host
code around it (citations are in ``constv_oracle.c``).  Parity status: integrator arithmetic pinned
by the reference tests' analytic properties; image handling "parity unpinned" (no reference test
touches an image particle).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libconstv_oracle.so")
_LIB_PATH_NOFMA = os.path.join(_HERE, "libconstv_oracle_nofma.so")

BOLTZ = 8.31446261815324e-3  # kJ/mol/K, OpenMM >= 7.4 SimTKOpenMMRealType.h [recalled, SURVEY 8c]

PRECISIONS = ("single", "mixed", "double")
_SUFFIX = {"single": "f32", "mixed": "mixed", "double": "f64"}
REAL = {"single": np.float32, "mixed": np.float32, "double": np.float64}
MIXED = {"single": np.float32, "mixed": np.float64, "double": np.float64}

ERRORS = {
    1: "Not even number of cells",
    2: "Number of particles is not multiple of number of cells",
    3: "Image Particle has nonzero mass",
    4: "Unit cell dimension does not match with the provided zmax value",
}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    if force or any(
        not os.path.exists(t) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(t)
            for f in ("constv_oracle.c", "constv_oracle_impl.h", "constv_oracle_drude_impl.h",
                      "constv_oracle_settle_impl.h", "Makefile"))
        for t in (_LIB_PATH, _LIB_PATH_NOFMA)
    ):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


def build_ref(force: bool = False, reference_root: str = "/root/reference"):
    """Compile the reference's own kernel source into oracle/_ref/ (host emulation + nvcc sm_100a),
    see oracle/Makefile `ref`.  Only possible where the reference tree exists (the build
    container); elsewhere the prebuilt files that travelled with the tree are used.  Returns the
    _ref directory, or None if neither the sources nor prebuilt libraries are there."""
    ref_dir = os.path.join(_HERE, "_ref")
    src = os.path.join(reference_root, "platforms", "cuda", "src", "kernels", "constVLangevin.cu")
    if os.path.exists(src):
        cmd = ["make", "-C", _HERE, "ref", "REF_ROOT=" + reference_root] + (["-B"] if force else [])
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + res.stdout + res.stderr)
    return ref_dir if os.path.exists(os.path.join(ref_dir, "libconstv_ref_host.so")) else None


_libs = {}
_variant = "fma"


def _load(path):
    so = C.CDLL(path)
    so.oracle_kinetic_energy_f32.restype = C.c_double
    so.oracle_kinetic_energy_mixed.restype = C.c_double
    so.oracle_kinetic_energy_f64.restype = C.c_double
    for sfx in ("f32", "mixed", "f64"):
        getattr(so, "oracle_kinetic_energy_rigid_" + sfx).restype = C.c_double
    so.oracle_num_threads.restype = C.c_int
    return so


def lib() -> C.CDLL:
    if _variant not in _libs:
        build()
        _libs[_variant] = _load(_LIB_PATH if _variant == "fma" else _LIB_PATH_NOFMA)
    return _libs[_variant]


class uncontracted:
    """Context manager: inside it every oracle function runs the ORACLE_NO_CONTRACT build (each
    operation rounded separately), which is what a host compilation of the reference's kernel
    source without FMA contraction computes (oracle/ref_host.cpp)."""

    def __enter__(self):
        global _variant
        self._saved = _variant
        _variant = "nofma"
        return self

    def __exit__(self, *exc):
        global _variant
        _variant = self._saved
        return False


def _p(a, dtype=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    if dtype is not None:
        assert a.dtype == np.dtype(dtype), (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def _mixed_scalar(precision, v):
    return C.c_float(v) if MIXED[precision] is np.float32 else C.c_double(v)


def params(temperature, friction, step_size, boltz=BOLTZ):
    out = np.zeros(3, np.float64)
    lib().oracle_params(C.c_double(temperature), C.c_double(friction), C.c_double(step_size),
                        C.c_double(boltz), _p(out))
    return out


def validate(num_atoms, num_cells, masses, box_z, cell_size):
    """Returns zmax or raises ValueError with the reference's message."""
    masses = np.ascontiguousarray(masses, np.float64)
    z = C.c_double(0)
    rc = lib().oracle_validate(C.c_int(num_atoms), C.c_int(num_cells), _p(masses),
                               C.c_double(box_z), C.c_double(cell_size), C.byref(z))
    if rc:
        raise ValueError(ERRORS[rc])
    return z.value


def inverse_order(atom_index):
    atom_index = np.ascontiguousarray(atom_index, np.int32)
    inv = np.empty_like(atom_index)
    lib().oracle_inverse_order(C.c_int(len(atom_index)), _p(atom_index), _p(inv))
    return inv


def part1(precision, velm, force, pos_delta, prm, step_size, random4, random_index=0,
          num_atoms=None):
    P = velm.shape[0]
    n = P if num_atoms is None else num_atoms
    m = MIXED[precision]
    prm = np.ascontiguousarray(prm, m)
    fn = getattr(lib(), "oracle_part1_" + _SUFFIX[precision])
    fn(C.c_int(n), C.c_int(P), _p(velm, m), _p(force, np.int64), _p(pos_delta, m), _p(prm),
       _mixed_scalar(precision, step_size), _p(random4, np.float32), C.c_uint(random_index))


def part2(precision, posq, posq_corr, pos_delta, velm, step_size, num_atoms=None):
    n = velm.shape[0] if num_atoms is None else num_atoms
    r, m = REAL[precision], MIXED[precision]
    fn = getattr(lib(), "oracle_part2_" + _SUFFIX[precision])
    fn(C.c_int(n), _p(posq, r), _p(posq_corr, r) if posq_corr is not None else None,
       _p(pos_delta, m), _p(velm, m), _mixed_scalar(precision, step_size))


def mirror(precision, num_real, num_cells, zshift_per_cell, posq, posq_corr, inv_order):
    r = REAL[precision]
    fn = getattr(lib(), "oracle_mirror_" + _SUFFIX[precision])
    fn(C.c_int(num_real), C.c_int(num_cells), C.c_double(zshift_per_cell), _p(posq, r),
       _p(posq_corr, r) if posq_corr is not None else None, _p(inv_order, np.int32))


def zero_images(precision, num_atoms, num_real, velm, force, inv_order, zero_velm=True,
                zero_force=True):
    P = velm.shape[0]
    fn = getattr(lib(), "oracle_zero_images_" + _SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(num_real), C.c_int(P), _p(velm, MIXED[precision]),
       _p(force, np.int64), _p(inv_order, np.int32), C.c_int(zero_velm), C.c_int(zero_force))


def kinetic_energy(precision, velm, force, time_shift, num_atoms=None):
    P = velm.shape[0]
    n = P if num_atoms is None else num_atoms
    fn = getattr(lib(), "oracle_kinetic_energy_" + _SUFFIX[precision])
    return fn(C.c_int(n), C.c_int(P), _p(velm, MIXED[precision]), _p(force, np.int64),
              C.c_double(time_shift))


def step(precision, num_atoms, num_cells, zshift_per_cell, posq, posq_corr, velm, force,
         pos_delta, prm, step_size, random4, random_index, inv_order, zero_velm=True,
         zero_force=True):
    """One constraint-free step in the reference's launch order, all arrays updated in place."""
    P = velm.shape[0]
    r, m = REAL[precision], MIXED[precision]
    prm = np.ascontiguousarray(prm, m)
    fn = getattr(lib(), "oracle_step_" + _SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(P), C.c_int(num_cells), C.c_double(zshift_per_cell),
       _p(posq, r), _p(posq_corr, r) if posq_corr is not None else None, _p(velm, m),
       _p(force, np.int64), _p(pos_delta, m), _p(prm), _mixed_scalar(precision, step_size),
       _p(random4, np.float32), C.c_uint(random_index), _p(inv_order, np.int32),
       C.c_int(zero_velm), C.c_int(zero_force))


def step_range(precision, num_atoms, num_cells, zshift_per_cell, posq, posq_corr, velm, force, prm,
               step_size, random4, random_index, lo, hi, zero_velm=True, zero_force=True):
    """The step restricted to real atoms [lo, hi) and their images, identity order, single pass
    (the loop a CPU port of the fused step would run; used for the timed CPU baseline)."""
    P = velm.shape[0]
    r, m = REAL[precision], MIXED[precision]
    prm = np.ascontiguousarray(prm, m)
    fn = getattr(lib(), "oracle_step_range_" + _SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(P), C.c_int(num_cells), C.c_double(zshift_per_cell),
       _p(posq, r), _p(posq_corr, r) if posq_corr is not None else None, _p(velm, m),
       _p(force, np.int64), _p(prm), _mixed_scalar(precision, step_size),
       _p(random4, np.float32), C.c_uint(random_index), C.c_int(zero_velm), C.c_int(zero_force),
       C.c_int(lo), C.c_int(hi))


# ---- ConstVDrudeLangevinIntegrator (constVDrudeLangevin.cu, CudaConstVKernels.cpp:252-349) ----------

def drude_params(temperature, friction, drude_temperature, drude_friction, max_drude_distance, step_size,
                 boltz=BOLTZ):
    """[vscale, fscale/2^32, noisescale, vscaleDrude, fscaleDrude/2^32, noisescaleDrude,
    maxDrudeDistance, hardwallscaleDrude] in double."""
    out = np.zeros(8, np.float64)
    lib().oracle_drude_params(C.c_double(temperature), C.c_double(friction), C.c_double(drude_temperature),
                              C.c_double(drude_friction), C.c_double(max_drude_distance), C.c_double(step_size),
                              C.c_double(boltz), _p(out))
    return out


def drude_part1(precision, velm, force, pos_delta, normal, pairs, scalars, step_size, random4, random_index=0):
    m = MIXED[precision]
    s = np.ascontiguousarray(scalars, m)
    fn = getattr(lib(), "oracle_drude_part1_" + _SUFFIX[precision])
    fn(C.c_int(len(normal)), C.c_int(len(pairs)), C.c_int(velm.shape[0]), _p(velm, m), _p(force, np.int64),
       _p(pos_delta, m), _p(normal, np.int32), _p(pairs, np.int32), _mixed_scalar(precision, step_size), _p(s),
       _p(random4, np.float32), C.c_uint(random_index))


def drude_hardwall(precision, posq, posq_corr, velm, pairs, step_size, max_drude_distance, hardwall_scale):
    r, m = REAL[precision], MIXED[precision]
    fn = getattr(lib(), "oracle_drude_hardwall_" + _SUFFIX[precision])
    fn(C.c_int(len(pairs)), _p(posq, r), _p(posq_corr, r) if posq_corr is not None else None, _p(velm, m),
       _p(pairs, np.int32), _mixed_scalar(precision, step_size), _mixed_scalar(precision, max_drude_distance),
       _mixed_scalar(precision, hardwall_scale))


def drude_step(precision, num_atoms, num_cells, zshift_per_cell, posq, posq_corr, velm, force, pos_delta,
               normal, pairs, scalars, step_size, random4, random_index, inv_order, zero_velm=True,
               zero_force=True):
    """One constraint-free Drude step in the reference's launch order (+ the explicit image zeroing)."""
    r, m = REAL[precision], MIXED[precision]
    s = np.ascontiguousarray(scalars, m)
    fn = getattr(lib(), "oracle_drude_step_" + _SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(velm.shape[0]), C.c_int(num_cells), C.c_double(zshift_per_cell), _p(posq, r),
       _p(posq_corr, r) if posq_corr is not None else None, _p(velm, m), _p(force, np.int64), _p(pos_delta, m),
       C.c_int(len(normal)), _p(normal, np.int32), C.c_int(len(pairs)), _p(pairs, np.int32), _p(s),
       _mixed_scalar(precision, step_size), _p(random4, np.float32), C.c_uint(random_index),
       _p(inv_order, np.int32), C.c_int(zero_velm), C.c_int(zero_force))


# ---- rigid 3-site molecules: the constraint hand-off (SETTLE, constv_oracle_settle_impl.h) --------------

def settle_positions(precision, posq, posq_corr, pos_delta, velm, cluster_atoms, cluster_params):
    """SETTLE on posDelta for every (apex, b, c) molecule; params = (d(apex-b), d(b-c)) float32."""
    r, m = REAL[precision], MIXED[precision]
    atoms = np.ascontiguousarray(cluster_atoms, np.int32).reshape(-1, 3)
    prm = np.ascontiguousarray(cluster_params, np.float32).reshape(-1, 2)
    fn = getattr(lib(), "oracle_settle_positions_" + _SUFFIX[precision])
    fn(C.c_int(len(atoms)), _p(posq, r), _p(posq_corr, r) if posq_corr is not None else None, _p(pos_delta, m),
       _p(velm, m), _p(atoms), _p(prm))


def settle_step(precision, num_atoms, num_cells, zshift_per_cell, posq, posq_corr, velm, force, pos_delta, prm,
                step_size, random4, random_index, inv_order, cluster_atoms, cluster_params, zero_velm=True,
                zero_force=True):
    """part 1 -> SETTLE -> part 2 -> image rewrite -> image zeroing."""
    r, m = REAL[precision], MIXED[precision]
    atoms = np.ascontiguousarray(cluster_atoms, np.int32).reshape(-1, 3)
    cp = np.ascontiguousarray(cluster_params, np.float32).reshape(-1, 2)
    prm = np.ascontiguousarray(prm, m)
    fn = getattr(lib(), "oracle_settle_step_" + _SUFFIX[precision])
    fn(C.c_int(num_atoms), C.c_int(velm.shape[0]), C.c_int(num_cells), C.c_double(zshift_per_cell), _p(posq, r),
       _p(posq_corr, r) if posq_corr is not None else None, _p(velm, m), _p(force, np.int64), _p(pos_delta, m),
       _p(prm), _mixed_scalar(precision, step_size), _p(random4, np.float32), C.c_uint(random_index),
       _p(inv_order, np.int32), C.c_int(zero_velm), C.c_int(zero_force), C.c_int(len(atoms)), _p(atoms), _p(cp))


def kinetic_energy_rigid(precision, posq, posq_corr, velm, force, time_shift, cluster_atoms, num_atoms=None,
                         return_velocities=False):
    """KE with the shifted velocities projected onto the rigid molecules' constraints (what OpenMM's
    computeKineticEnergy reports for a constrained System)."""
    r, m = REAL[precision], MIXED[precision]
    n = velm.shape[0] if num_atoms is None else num_atoms
    atoms = np.ascontiguousarray(cluster_atoms, np.int32).reshape(-1, 3)
    out = np.zeros((n, 3), np.float64) if return_velocities else None
    fn = getattr(lib(), "oracle_kinetic_energy_rigid_" + _SUFFIX[precision])
    ke = fn(C.c_int(n), C.c_int(velm.shape[0]), _p(posq, r), _p(posq_corr, r) if posq_corr is not None else None,
            _p(velm, m), _p(force, np.int64), C.c_double(time_shift), C.c_int(len(atoms)), _p(atoms),
            _p(out) if out is not None else None)
    return (ke, out) if return_velocities else ke


def philox4x32_10(counter, key):
    counter = np.ascontiguousarray(counter, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32_10(_p(counter), _p(key), _p(out))
    return out


def gaussian4(seed, atom, step_index):
    out = np.zeros(4, np.float32)
    lib().oracle_gaussian4(C.c_uint64(seed), C.c_uint64(atom), C.c_uint64(step_index), _p(out))
    return out


def fill_noise(num_atoms, seed, step_index, atom_index=None, global_atom_offset=0):
    out = np.zeros((num_atoms, 4), np.float32)
    ai = None if atom_index is None else np.ascontiguousarray(atom_index, np.int32)
    lib().oracle_fill_noise(C.c_int(num_atoms), _p(ai), C.c_uint64(seed),
                            C.c_uint64(global_atom_offset), C.c_uint64(step_index), _p(out))
    return out


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(C.c_int(n))
