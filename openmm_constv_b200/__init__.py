"""B200-native ConstVLangevinIntegrator step path (drop-in for scychon/openmm_constV's
platforms/cuda ConstVLangevin kernels).  See DESIGN.md.

The compute path is the C-ABI library built from ``csrc/`` (``include/constv_b200.h``); it has no
CPU fallback: every entry point raises if ``libconstv_b200.so`` is missing.
"""
__version__ = "0.1.1"
