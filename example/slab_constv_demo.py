#!/usr/bin/env python
"""The flow of the reference's example (example/nacl_1m_spce/nacl_constv.py) on the B200 step path, for a host
without OpenMM: NaCl in rigid SPC/E water between two wall electrodes with image charges, constant applied
field, ConstVLangevinIntegrator(300 K, 1/ps, 1 fs).

What the reference's script gets from OpenMM (force field, PME, constraints, reporters) is replaced here by what
this repository ships: a synthetic slab of the example's composition (`slab.make_water_slab`), the applied field
`efield*q*z` of nacl_constv.py:69-72 as the only force (evaluated on the device every step, as a `force_fn`
callback = context.calcForcesAndEnergy), the rigid-water constraints solved inside the step kernel, and a
StateDataReporter-like print.  `from constvplugin import ConstVLangevinIntegrator` is the reference's own import.
With no interatomic forces the ions (1 % of the atoms) drift freely in the field at q*E/(m*friction) ~ 1 nm/ps for
1 V, which the kinetic temperature counts: ~310 K at 1 V, 300 K at --volts 0.

    python example/slab_constv_demo.py [--real-atoms 7494] [--steps 2000] [--volts 1.0] [--precision mixed]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "openmm_constv_b200", "plugin", "python"))

from constvplugin import ConstVLangevinIntegrator, SlabContext, SlabSystem  # noqa: E402
from openmm_constv_b200 import slab as slabmod  # noqa: E402

FARADAY = 96.48533212  # kJ/mol per volt per elementary charge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--real-atoms", type=int, default=7494)   # the example's slab (nacl_1m_eq.pdb: 7494 real atoms)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--report", type=int, default=500)
    ap.add_argument("--volts", type=float, default=1.0)
    ap.add_argument("--precision", default="mixed", choices=["single", "mixed", "double"])
    args = ap.parse_args()
    import torch

    slab, molecules, params, constraints = slabmod.make_water_slab(args.real_atoms, args.precision, seed=2024)
    R = slab.num_real
    system = SlabSystem()
    system.addParticles(slab.masses)
    system.setDefaultPeriodicBoxVectors((slab.box[0], 0, 0), (0, slab.box[1], 0), (0, 0, 2 * slab.zmax))
    system.addConstraints(*constraints)                      # constraints=AllBonds on rigid water

    integrator = ConstVLangevinIntegrator(300.0, 1.0, 0.001)  # kelvin, 1/ps, ps (nacl_constv.py:12-16,36)
    integrator.setRandomNumberSeed(7)

    efield = args.volts * FARADAY / slab.zmax                 # kJ/mol/nm per elementary charge (nacl_constv.py:29-30)

    def applied_field(ctx):
        # F_z = -d/dz (efield*q*z) = -efield*q on every real atom, as OpenMM's fixed-point force buffer
        q = ctx.posq[:R, 3].double()
        ctx.force.zero_()
        ctx.force[2, :R] = torch.round(-efield * q * 4294967296.0).to(torch.int64)

    context = SlabContext(system, integrator, precision=args.precision, force_fn=applied_field, padded=slab.padded)
    context.load_slab(slab)
    kernel = integrator._kernel
    print(f"{slab.num_atoms} particles ({R} real + {slab.num_atoms - R} image), {len(molecules)} rigid water molecules, "
          f"constraints solved inside the step kernel: {kernel.fused_constraints}")
    print(f"{'step':>8} {'time/ps':>9} {'T_kin/K':>9} {'max |r_OH-0.1|/nm':>18} {'max |z_img + z - 2 zmax|':>25}")
    n_dof = 3 * int(np.sum(slab.masses[:R] > 0)) - 3 * len(molecules)
    done = 0
    while done < args.steps:
        n = min(args.report, args.steps - done)
        integrator.step(n)
        done += n
        pos = context.getPositions()
        ke = context.getKineticEnergy()   # half-step-shifted velocities, projected onto the constraints
        temp = 2 * ke / (n_dof * slabmod.BOLTZ)
        oh = np.linalg.norm(pos[molecules[:, 1]] - pos[molecules[:, 0]], axis=1)
        charged = slab.posq[:R, 3] != 0
        mirror = np.max(np.abs(pos[R:2 * R, 2][charged] + pos[:R, 2][charged] - 2 * slab.zmax))
        print(f"{context.getStepCount():8d} {context.getTime():9.3f} {temp:9.1f} {np.max(np.abs(oh - 0.1)):18.2e} {mirror:25.2e}")
    context.close()


if __name__ == "__main__":
    main()
