"""Python host mirror of the reference's plugin interface for the ConstV Langevin step path.

Same names, argument meaning and error behaviour as the reference (scychon/openmm_constV):

  ConstVLangevinIntegrator            openmmapi/include/openmm/ConstVLangevinIntegrator.h:46-176,
                                      openmmapi/src/ConstVLangevinIntegrator.cpp:44-85,
                                      python/constvplugin.i:65-80
  IntegrateConstVLangevinStepKernel   openmmapi/include/openmm/ConstVKernels.h:49-77 (interface),
                                      platforms/cuda/src/CudaConstVKernels.cpp:53-177 (CUDA impl)
  SlabSystem / SlabContext            the small part of OpenMM's System / Context / CudaContext the
                                      path touches, standing in for OpenMM, which is
                                      not available here.  Forces are injected: the context's
                                      ``force_fn`` plays the role of context.calcForcesAndEnergy.

  ConstVDrudeLangevinIntegrator       openmmapi/include/openmm/ConstVDrudeLangevinIntegrator.h,
                                      openmmapi/src/ConstVDrudeLangevinIntegrator.cpp:44-129,
                                      python/constvplugin.i:82-100
  IntegrateConstVDrudeLangevinStepKernel  ConstVKernels.h:82-108 (interface),
                                      platforms/cuda/src/CudaConstVKernels.cpp:179-357 (CUDA impl)
  DrudeForce                          the part of OpenMM's DrudeForce the kernel reads
                                      (getNumParticles / getParticleParameters, Kernels.cpp:219-224)

All compute goes through the C ABI (``_capi``); torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import ConstVError, check

_TORCH_REAL = {"single": "float32", "mixed": "float32", "double": "float64"}
_TORCH_MIXED = {"single": "float32", "mixed": "float64", "double": "float64"}


class OpenMMException(Exception):
    """Stands in for OpenMM::OpenMMException (same messages as the reference)."""


def _raise_openmm(err: ConstVError):
    if err.code == _capi.ERR_SYSTEM:
        raise OpenMMException(str(err)) from None
    raise err


class _ConstraintList:
    """(particle1, particle2, distance) triples held as arrays; behaves like the list of tuples it replaces."""

    def __init__(self, before, p1, p2, dist):
        if isinstance(before, _ConstraintList):
            p1, p2, dist = np.concatenate([before.p1, p1]), np.concatenate([before.p2, p2]), np.concatenate([before.dist, dist])
        elif len(before):
            b = list(before)
            p1 = np.concatenate([np.asarray([x[0] for x in b], np.int32), p1])
            p2 = np.concatenate([np.asarray([x[1] for x in b], np.int32), p2])
            dist = np.concatenate([np.asarray([x[2] for x in b], np.float64), dist])
        self.p1, self.p2, self.dist = p1, p2, dist

    def __len__(self):
        return len(self.p1)

    def __getitem__(self, i):
        return (int(self.p1[i]), int(self.p2[i]), float(self.dist[i]))

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def append(self, item):
        self.p1 = np.append(self.p1, np.int32(item[0]))
        self.p2 = np.append(self.p2, np.int32(item[1]))
        self.dist = np.append(self.dist, np.float64(item[2]))

    def arrays(self):
        return self.p1, self.p2, self.dist


class SlabSystem:
    """The slice of OpenMM::System the integrator reads: particle masses and the periodic box
    (CudaConstVKernels.cpp:79-89)."""

    def __init__(self):
        self._masses: list[float] = []
        self._box = ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))
        self._forces: list = []
        self._constraints: list[tuple] = []

    def addConstraint(self, particle1: int, particle2: int, distance: float) -> int:
        self._constraints.append((int(particle1), int(particle2), float(distance)))
        return len(self._constraints) - 1

    def addConstraints(self, particle1, particle2, distance) -> None:
        """Many constraints at once (arrays).  A million-constraint water slab is listed in milliseconds instead of
        through a Python loop; getConstraintParameters(i) answers the same either way."""
        p1 = np.ascontiguousarray(particle1, np.int32).reshape(-1)
        p2 = np.ascontiguousarray(particle2, np.int32).reshape(-1)
        d = np.ascontiguousarray(distance, np.float64).reshape(-1)
        if not (len(p1) == len(p2) == len(d)):
            raise ValueError("addConstraints: arrays of different lengths")
        self._constraints = _ConstraintList(self._constraints, p1, p2, d)

    def getNumConstraints(self) -> int:
        return len(self._constraints)

    def getConstraintParameters(self, index: int):
        return self._constraints[index]

    def addForce(self, force) -> int:
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self) -> int:
        return len(self._forces)

    def getForce(self, index: int):
        return self._forces[index]

    def addParticle(self, mass: float) -> int:
        self._masses.append(float(mass))
        return len(self._masses) - 1

    def addParticles(self, masses) -> None:
        self._masses.extend(np.asarray(masses, np.float64).tolist())

    def getParticleMasses(self) -> np.ndarray:
        return np.asarray(self._masses, np.float64)

    def getNumParticles(self) -> int:
        return len(self._masses)

    def getParticleMass(self, index: int) -> float:
        return self._masses[index]

    def setParticleMass(self, index: int, mass: float) -> None:
        self._masses[index] = float(mass)

    def setDefaultPeriodicBoxVectors(self, a, b, c) -> None:
        self._box = (tuple(a), tuple(b), tuple(c))

    def getDefaultPeriodicBoxVectors(self):
        return self._box


class DrudeForce:
    """The slice of OpenMM::DrudeForce the Drude kernel reads (CudaConstVKernels.cpp:219-224):
    per entry the Drude particle ``particle``, its parent ``particle1`` and the parameters the
    integrator ignores."""

    def __init__(self):
        self._particles: list[tuple] = []

    def addParticle(self, particle, particle1, particle2=-1, particle3=-1, particle4=-1, charge=0.0,
                    polarizability=0.0, aniso12=0.0, aniso34=0.0) -> int:
        self._particles.append((int(particle), int(particle1), int(particle2), int(particle3), int(particle4),
                                float(charge), float(polarizability), float(aniso12), float(aniso34)))
        return len(self._particles) - 1

    def getNumParticles(self) -> int:
        return len(self._particles)

    def getParticleParameters(self, index: int):
        return self._particles[index]


class ConstVLangevinIntegrator:
    """Langevin dynamics with exact image charges between slab electrodes.

    ``ConstVLangevinIntegrator(temperature, frictionCoeff, stepSize, numCells=2, zmax=-1)``:
    kelvin, 1/ps, ps; ``numCells`` counts real + image cells; ``zmax`` < 0 derives the cell size
    from the periodic box (ConstVLangevinIntegrator.h:57).  Constraint tolerance defaults to 1e-5
    and the random seed to 0 (ConstVLangevinIntegrator.cpp:50-51).
    """

    def __init__(self, temperature, frictionCoeff, stepSize, numCells=2, zmax=-1):
        self._context = None
        self._kernel = None
        self.setTemperature(temperature)
        self.setFriction(frictionCoeff)
        self.setStepSize(stepSize)
        self.setNumCells(numCells)
        self.setCellSize(zmax)
        self.setConstraintTolerance(1e-5)
        self.setRandomNumberSeed(0)

    # -- accessors (ConstVLangevinIntegrator.h:63-144 and OpenMM::Integrator) ------------------
    def getNumCells(self): return self._num_cells
    def setNumCells(self, cells): self._num_cells = int(cells)
    def getCellSize(self): return self._cell_size
    def setCellSize(self, zmax): self._cell_size = float(zmax)
    def getTemperature(self): return self._temperature
    def setTemperature(self, temp): self._temperature = float(temp)
    def getFriction(self): return self._friction
    def setFriction(self, coeff): self._friction = float(coeff)
    def getRandomNumberSeed(self): return self._seed
    def setRandomNumberSeed(self, seed): self._seed = int(seed)
    def getStepSize(self): return self._step_size
    def setStepSize(self, size): self._step_size = float(size)
    def getConstraintTolerance(self): return self._constraint_tol
    def setConstraintTolerance(self, tol): self._constraint_tol = float(tol)

    # -- Integrator protocol (ConstVLangevinIntegrator.cpp:54-85) ------------------------------
    def _initialize(self, context: "SlabContext") -> None:
        if self._context is not None and self._context is not context:
            raise OpenMMException("This Integrator is already bound to a context")
        self._context = context
        self._kernel = IntegrateConstVLangevinStepKernel(IntegrateConstVLangevinStepKernel.Name(),
                                                         context)
        self._kernel.initialize(context.getSystem(), self)

    def _cleanup(self) -> None:
        if self._kernel is not None:
            self._kernel.close()
        self._kernel = None
        self._context = None

    def getKernelNames(self):
        return [IntegrateConstVLangevinStepKernel.Name()]

    def computeKineticEnergy(self) -> float:
        if self._context is None:
            raise OpenMMException("This Integrator is not bound to a context!")
        return self._kernel.computeKineticEnergy(self._context, self)

    def step(self, steps: int) -> None:
        if self._context is None:
            raise OpenMMException("This Integrator is not bound to a context!")
        steps = int(steps)
        if steps > 1 and self._kernel.can_execute_many(self._context):
            # nothing changes between the steps (no force callback, no constraint hand-off, Philox noise): the
            # reference's loop body (ConstVLangevinIntegrator.cpp:80-83) is n x execute, issued in one call
            self._kernel.execute_many(self._context, self, steps)
            return
        for _ in range(steps):
            self._context.updateContextState()
            self._context.calcForcesAndEnergy(True, False)
            self._kernel.execute(self._context, self)


class IntegrateConstVLangevinStepKernel:
    """B200 implementation of the kernel interface ConstVKernels.h:49-77, the analogue of
    CudaIntegrateConstVLangevinStepKernel (CudaConstVKernels.h:46-79): ``initialize``,
    ``execute``, ``computeKineticEnergy``; owns the C-ABI slab handle."""

    @staticmethod
    def Name() -> str:
        return "IntegrateConstVLangevinStep"

    def __init__(self, name: str, context: "SlabContext"):
        if name != self.Name():
            raise OpenMMException(f"Tried to create kernel with illegal kernel name '{name}'")
        self._lib = _capi.load()
        self._ctx = context
        self._slab = C.c_void_p(None)

    def initialize(self, system: SlabSystem, integrator: ConstVLangevinIntegrator) -> None:
        ctx = self._ctx
        n = system.getNumParticles()
        masses = system.getParticleMasses()
        box_z = system.getDefaultPeriodicBoxVectors()[2][2]
        try:  # CudaConstVKernels.cpp:76-96
            zmax = _capi.validate_system(masses, integrator.getNumCells(), box_z,
                                         integrator.getCellSize())
        except ConstVError as err:
            _raise_openmm(err)
        seed = integrator.getRandomNumberSeed()
        if seed == 0:  # ConstVLangevinIntegrator.h:138-140: a unique seed per Context
            seed = int.from_bytes(os.urandom(8), "little") | 1
        cfg = _capi.Config()
        cfg.struct_size = C.sizeof(_capi.Config)
        cfg.device = ctx.device_index
        cfg.num_atoms = n
        cfg.padded_num_atoms = ctx.padded
        cfg.num_cells = integrator.getNumCells()
        cfg.precision = _capi.PRECISION[ctx.precision]
        cfg.mirror_mode = ctx.mirror_mode
        cfg.flags = ctx.flags
        cfg.zmax = zmax
        cfg.seed = seed & 0xFFFFFFFFFFFFFFFF
        cfg.global_atom_offset = ctx.global_atom_offset
        cfg.boltz = 0.0
        check(self._lib.constv_create(C.byref(cfg), C.byref(self._slab)))
        self.zmax = zmax
        self.seed = cfg.seed
        self.rebind()
        self._configure_rigid_molecules(system)

    def _configure_rigid_molecules(self, system: SlabSystem) -> None:
        """The constraint hand-off (CudaConstVKernels.cpp:153): when EVERY constraint of the System belongs
        to a rigid 3-site molecule (the reference's example: SPC/E water, constraints=AllBonds) the solver
        runs inside the step kernel; otherwise the split path hands posDelta to the caller's solver."""
        self.num_rigid_molecules = 0
        self.fused_constraints = False
        nc = system.getNumConstraints() if hasattr(system, "getNumConstraints") else 0
        if nc == 0:
            return
        held = getattr(system, "_constraints", None)
        if isinstance(held, _ConstraintList):
            p1, p2, dist = (np.ascontiguousarray(a) for a in held.arrays())
        else:
            cons = [system.getConstraintParameters(i) for i in range(nc)]
            p1 = np.ascontiguousarray([c[0] for c in cons], np.int32)
            p2 = np.ascontiguousarray([c[1] for c in cons], np.int32)
            dist = np.ascontiguousarray([c[2] for c in cons], np.float64)
        atoms = np.zeros((max(nc // 3, 1), 3), np.int32)
        params = np.zeros((max(nc // 3, 1), 2), np.float32)
        found, other = C.c_int32(0), C.c_int32(0)
        vp = C.c_void_p
        try:
            check(self._lib.constv_settle_find_clusters(system.getNumParticles(), nc, p1.ctypes.data_as(vp),
                                                        p2.ctypes.data_as(vp), dist.ctypes.data_as(vp),
                                                        atoms.ctypes.data_as(vp), params.ctypes.data_as(vp),
                                                        C.byref(found), C.byref(other)))
        except ConstVError as err:
            _raise_openmm(err)
        self.num_rigid_molecules = found.value
        if other.value == 0 and found.value > 0:
            check(self._lib.constv_settle_configure(self._slab, found.value, atoms.ctypes.data_as(vp),
                                                    params.ctypes.data_as(vp)))
            self.fused_constraints = True

    def rebind(self) -> None:
        ctx = self._ctx
        check(self._lib.constv_bind_arrays(
            self._slab, ctx.posq.data_ptr(), ctx.corr.data_ptr() if ctx.corr is not None else None,
            ctx.velm.data_ptr(), ctx.force.data_ptr(), ctx.pos_delta.data_ptr(),
            ctx.atom_index.data_ptr() if ctx.atom_index is not None else None))
        if ctx.atom_index is not None:
            check(self._lib.constv_update_inverse_order(self._slab, ctx.stream_ptr()))

    def execute(self, context: "SlabContext", integrator: ConstVLangevinIntegrator) -> None:
        lib, slab, st = self._lib, self._slab, context.stream_ptr()
        # the raw numbers (kelvin, 1/ps): subclasses may decorate the getters with units
        check(lib.constv_set_parameters(slab, integrator._temperature, integrator._friction,
                                        integrator.getStepSize()))
        if context.atoms_were_reordered:  # CudaConstVKernels.cpp:139-142
            check(lib.constv_update_inverse_order(slab, st))
            context.atoms_were_reordered = False
        rnd, idx = context.prepare_random_numbers()
        if context.apply_constraints is None and getattr(self, "fused_constraints", False):
            # part 1, the constraints (rigid 3-site molecules), part 2 and the image rewrite in one kernel
            check(lib.constv_step_fused_settle(slab, rnd, idx, st))
        elif context.apply_constraints is None:
            if context.getSystem().getNumConstraints() > 0:
                raise OpenMMException("the System has constraints the step kernel cannot solve itself: "
                                      "give the context an apply_constraints callback")
            check(lib.constv_step_fused(slab, rnd, idx, st))
        else:  # CudaConstVKernels.cpp:146-166 with the constraint hand-off at :153
            check(lib.constv_step_part1(slab, rnd, idx, st))
            context.apply_constraints(context, integrator.getConstraintTolerance())
            check(lib.constv_step_part2_mirror(slab, st))
        context.time += integrator.getStepSize()
        context.step_count += 1

    def can_execute_many(self, context: "SlabContext") -> bool:
        return (type(self) is IntegrateConstVLangevinStepKernel and context.force_fn is None
                and context.apply_constraints is None and context.random_pool is None
                and not getattr(self, "fused_constraints", False) and context.getSystem().getNumConstraints() == 0)

    def execute_many(self, context: "SlabContext", integrator: ConstVLangevinIntegrator, steps: int) -> None:
        """``steps`` x execute in one C-ABI call (constv_step_fused_multi; on a chained slab the library forks
        onto its own streams)."""
        lib, slab, st = self._lib, self._slab, context.stream_ptr()
        check(lib.constv_set_parameters(slab, integrator._temperature, integrator._friction, integrator.getStepSize()))
        if context.atoms_were_reordered:
            check(lib.constv_update_inverse_order(slab, st))
            context.atoms_were_reordered = False
        check(lib.constv_step_fused_multi(slab, steps, st))
        context.time += steps * integrator.getStepSize()
        context.step_count += steps

    def computeKineticEnergy(self, context: "SlabContext", integrator: ConstVLangevinIntegrator) -> float:
        ke = C.c_double(0.0)
        check(self._lib.constv_kinetic_energy(self._slab, 0.5 * integrator.getStepSize(),
                                              C.byref(ke), context.stream_ptr()))
        return ke.value

    # -- extras of the B200 path -----------------------------------------------------------------
    @property
    def handle(self) -> C.c_void_p:
        return self._slab

    def kinetic_energy_device(self, time_shift: float, out_tensor) -> None:
        check(self._lib.constv_kinetic_energy_device(self._slab, time_shift, out_tensor.data_ptr(),
                                                     self._ctx.stream_ptr()))

    def get_step_count(self) -> int:
        v = C.c_uint64(0)
        check(self._lib.constv_get_step_count(self._slab, C.byref(v)))
        return v.value

    def set_step_count(self, count: int) -> None:
        check(self._lib.constv_set_step_count(self._slab, count))

    def settle_positions(self) -> None:
        """The solver alone on pos_delta (for an apply_constraints callback of the split path)."""
        check(self._lib.constv_settle_positions(self._slab, self._ctx.stream_ptr()))

    def set_launch_shape(self, pairs_per_thread: int = 0, ctas_per_sm: int = 0) -> None:
        check(self._lib.constv_set_launch_shape(self._slab, pairs_per_thread, ctas_per_sm))

    def set_kernel_variant(self, variant: int = 0, block_size: int = 0, stages: int = 0,
                           ctas_per_sm: int = 0) -> None:
        check(self._lib.constv_set_kernel_variant(self._slab, variant, block_size, stages, ctas_per_sm))

    def close(self) -> None:
        if self._slab:
            self._lib.constv_destroy(self._slab)
            self._slab = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConstVDrudeLangevinIntegrator(ConstVLangevinIntegrator):
    """Dual-thermostat Langevin dynamics for Drude oscillators with exact image charges.

    ``ConstVDrudeLangevinIntegrator(temperature, frictionCoeff, drudeTemperature, drudeFrictionCoeff,
    stepSize, numCells=2, zmax=-1)`` (ConstVDrudeLangevinIntegrator.cpp:44-55): ordinary particles and
    each pair's centre of mass are thermalised at (temperature, frictionCoeff), each pair's relative
    motion at (drudeTemperature, drudeFrictionCoeff); maxDrudeDistance (default 0 = off) is the hard
    wall on the Drude displacement.  The System must hold exactly one DrudeForce.
    """

    def __init__(self, temperature, frictionCoeff, drudeTemperature, drudeFrictionCoeff, stepSize,
                 numCells=2, zmax=-1):
        super().__init__(temperature, frictionCoeff, stepSize, numCells, zmax)
        self.setDrudeTemperature(drudeTemperature)
        self.setDrudeFriction(drudeFrictionCoeff)
        self.setMaxDrudeDistance(0)

    def getDrudeTemperature(self): return self._drude_temperature
    def setDrudeTemperature(self, temp): self._drude_temperature = float(temp)
    def getDrudeFriction(self): return self._drude_friction
    def setDrudeFriction(self, coeff): self._drude_friction = float(coeff)
    def getMaxDrudeDistance(self): return self._max_drude_distance

    def setMaxDrudeDistance(self, distance):
        if distance < 0:  # ConstVDrudeLangevinIntegrator.cpp:61-65
            raise OpenMMException("setMaxDrudeDistance: Distance cannot be negative")
        self._max_drude_distance = float(distance)

    def _initialize(self, context: "SlabContext") -> None:
        if self._context is not None and self._context is not context:
            raise OpenMMException("This Integrator is already bound to a context")
        system = context.getSystem()
        force = None  # ConstVDrudeLangevinIntegrator.cpp:70-80
        for i in range(system.getNumForces()):
            if isinstance(system.getForce(i), DrudeForce):
                if force is not None:
                    raise OpenMMException("The System contains multiple DrudeForces")
                force = system.getForce(i)
        if force is None:
            raise OpenMMException("The System does not contain a DrudeForce")
        self._context = context
        self._kernel = IntegrateConstVDrudeLangevinStepKernel(IntegrateConstVDrudeLangevinStepKernel.Name(), context)
        self._kernel.initialize(system, self, force)

    def getKernelNames(self):
        return [IntegrateConstVDrudeLangevinStepKernel.Name()]


class IntegrateConstVDrudeLangevinStepKernel(IntegrateConstVLangevinStepKernel):
    """B200 implementation of ConstVKernels.h:82-108, the analogue of
    CudaIntegrateConstVDrudeLangevinStepKernel (CudaConstVKernels.cpp:179-357)."""

    @staticmethod
    def Name() -> str:
        return "IntegrateConstVDrudeLangevinStep"

    def _configure_rigid_molecules(self, system: SlabSystem) -> None:
        # the Drude step always hands constraints over (CudaConstVKernels.cpp:315)
        self.num_rigid_molecules = 0
        self.fused_constraints = False

    def initialize(self, system: SlabSystem, integrator: ConstVDrudeLangevinIntegrator, force: DrudeForce = None) -> None:
        super().initialize(system, integrator)
        pairs = np.zeros((max(force.getNumParticles(), 1), 2), np.int32)
        for i in range(force.getNumParticles()):  # CudaConstVKernels.cpp:219-225
            prm = force.getParticleParameters(i)
            pairs[i] = prm[0], prm[1]
        check(self._lib.constv_drude_configure(self._slab, force.getNumParticles(), pairs.ctypes.data_as(C.c_void_p)))
        n = C.c_int32(0)
        check(self._lib.constv_drude_random_count(self._slab, C.byref(n)))
        self.random_count = n.value
        self.num_pairs = force.getNumParticles()

    def execute(self, context: "SlabContext", integrator: ConstVDrudeLangevinIntegrator) -> None:
        lib, slab, st = self._lib, self._slab, context.stream_ptr()
        check(lib.constv_drude_set_parameters(slab, integrator._temperature, integrator._friction,
                                              integrator._drude_temperature, integrator._drude_friction,
                                              integrator._max_drude_distance, integrator.getStepSize()))
        if context.atoms_were_reordered:  # CudaConstVKernels.cpp:303-306
            check(lib.constv_update_inverse_order(slab, st))
            context.atoms_were_reordered = False
        rnd, idx = context.prepare_random_numbers(self.random_count)  # :307
        if context.apply_constraints is None:
            if context.getSystem().getNumConstraints() > 0:
                raise OpenMMException("the System has constraints: the Drude step hands them to the caller's solver, "
                                      "give the context an apply_constraints callback")
            check(lib.constv_drude_step_fused(slab, rnd, idx, st))
        else:  # :308-344 with the constraint hand-off at :315
            check(lib.constv_drude_step_part1(slab, rnd, idx, st))
            context.apply_constraints(context, integrator.getConstraintTolerance())
            check(lib.constv_drude_step_part2(slab, st))
        context.time += integrator.getStepSize()
        context.step_count += 1


class SlabContext:
    """Device-resident state in OpenMM's CUDA layouts plus the Context calls the integrator makes
    (updateContextState, calcForcesAndEnergy, getState-like accessors).

    ``force_fn(context)`` must fill ``context.force`` ((3,P) int64 fixed point) on the device each
    step; None keeps the current buffer (injected forces).  ``apply_constraints(context, tol)``,
    when given, is called between part 1 and part 2 on ``context.pos_delta``.
    """

    def __init__(self, system: SlabSystem, integrator: ConstVLangevinIntegrator, precision="single",
                 device=0, mirror_mode=_capi.MIRROR_REFERENCE, flags=_capi.FLAGS_DEFAULT,
                 force_fn=None, apply_constraints=None, random_pool=None, global_atom_offset=0,
                 padded=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("SlabContext needs a CUDA device: the ConstV step path has no CPU fallback")
        _capi.load()
        self.torch = torch
        self._system = system
        self._integrator = integrator
        self.precision = precision
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.mirror_mode = mirror_mode
        self.flags = flags
        self.force_fn = force_fn
        self.apply_constraints = apply_constraints
        self.random_pool = random_pool       # optional (M,4) float32 device tensor of injected noise
        self.random_index = 0
        self.global_atom_offset = int(global_atom_offset)
        n = system.getNumParticles()
        self.num_atoms = n
        self.padded = int(padded) if padded is not None else (n + 31) // 32 * 32
        rt = getattr(torch, _TORCH_REAL[precision])
        mt = getattr(torch, _TORCH_MIXED[precision])
        P = self.padded
        self.posq = torch.zeros((P, 4), dtype=rt, device=self.device)
        self.corr = torch.zeros((P, 4), dtype=rt, device=self.device) if precision == "mixed" else None
        self.velm = torch.zeros((P, 4), dtype=mt, device=self.device)
        self.force = torch.zeros((3, P), dtype=torch.int64, device=self.device)
        self.pos_delta = torch.zeros((P, 4), dtype=mt, device=self.device)
        self.atom_index = None
        self.atoms_were_reordered = False
        self.time = 0.0
        self.step_count = 0
        masses = system.getParticleMasses()
        inv = np.zeros(P, np.float64)
        inv[:n][masses != 0] = 1.0 / masses[masses != 0]
        self.velm[:, 3] = torch.as_tensor(inv, dtype=mt, device=self.device)
        integrator._initialize(self)

    # -- what the integrator calls ---------------------------------------------------------------
    def getSystem(self): return self._system
    def getIntegrator(self): return self._integrator
    def updateContextState(self): pass

    def calcForcesAndEnergy(self, include_forces=True, include_energy=False):
        if self.force_fn is not None:
            self.force_fn(self)

    def stream_ptr(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def prepare_random_numbers(self, count=None):
        """Mirrors integration.prepareRandomNumbers / getRandom (CudaConstVKernels.cpp:146,148)
        when a pool of injected Gaussians is attached; (None, 0) selects the Philox stream."""
        if self.random_pool is None:
            return None, 0
        count = self.padded if count is None else int(count)
        idx = self.random_index
        if idx + count > self.random_pool.shape[0]:
            raise RuntimeError("injected random pool exhausted")
        self.random_index += count
        return self.random_pool.data_ptr(), idx

    # -- state I/O -------------------------------------------------------------------------------
    def load_slab(self, slab) -> None:
        """Upload a ``slab.Slab`` (host arrays in the same layouts)."""
        t = self.torch
        assert slab.padded == self.padded and slab.precision == self.precision
        self.posq.copy_(t.from_numpy(slab.posq))
        if self.corr is not None:
            self.corr.copy_(t.from_numpy(slab.corr))
        self.velm.copy_(t.from_numpy(slab.velm))
        self.force.copy_(t.from_numpy(slab.force))

    def setPositions(self, positions, charges=None):
        t = self.torch
        p = np.asarray(positions, np.float64)
        n = len(p)
        lo = p.astype(self.posq.cpu().numpy().dtype)
        self.posq[:n, :3] = t.as_tensor(lo, device=self.device)
        if self.corr is not None:
            self.corr[:n, :3] = t.as_tensor((p - lo.astype(np.float64)).astype(np.float32), device=self.device)
        if charges is not None:
            self.posq[:n, 3] = t.as_tensor(np.asarray(charges), dtype=self.posq.dtype, device=self.device)

    def setVelocities(self, velocities):
        v = np.asarray(velocities, np.float64)
        self.velm[: len(v), :3] = self.torch.as_tensor(v, dtype=self.velm.dtype, device=self.device)

    def setForces(self, forces_kj_mol_nm):
        """Inject forces (N,3) in kJ/mol/nm as OpenMM's fixed-point buffer."""
        f = np.rint(np.asarray(forces_kj_mol_nm, np.float64).T * 4294967296.0).astype(np.int64)
        self.force[:, : f.shape[1]] = self.torch.as_tensor(f, device=self.device)

    def getPositions(self):
        p = self.posq[: self.num_atoms, :3].double()
        if self.corr is not None:
            p = p + self.corr[: self.num_atoms, :3].double()
        return p.cpu().numpy()

    def getVelocities(self):
        return self.velm[: self.num_atoms, :3].double().cpu().numpy()

    def getKineticEnergy(self):
        return self._integrator.computeKineticEnergy()

    def getTime(self): return self.time
    def getStepCount(self): return self.step_count

    def set_atom_order(self, atom_index) -> None:
        """Install a slot permutation (what cu.reorderAtoms() leaves behind): ``atom_index[slot]``
        is the original atom held by ``slot``.  The caller permutes the state arrays itself."""
        t = self.torch
        ai = np.arange(self.padded, dtype=np.int32)
        ai[: len(atom_index)] = atom_index
        self.atom_index = t.as_tensor(ai, device=self.device)
        self.atoms_were_reordered = True
        self._integrator._kernel.rebind()

    def reinitialize(self):
        self._integrator._cleanup()
        self.time = 0.0
        self.step_count = 0
        self.random_index = 0
        self._integrator._initialize(self)

    def close(self):
        self._integrator._cleanup()
