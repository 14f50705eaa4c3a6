"""Multi-GPU range partition of one slab (new in the B200 path: the reference is
single-device, CudaConstVKernelFactory.cpp:64).

Real atoms are split into contiguous ranges, one per rank; every image lives on its partner's GPU,
so a step needs NO inter-GPU traffic.  The Philox stream is keyed by the GLOBAL original atom index
(``global_atom_offset`` = first real atom of the shard), so a partitioned run reproduces the
single-GPU trajectory bit for bit.  The only collective is a scalar sum for the kinetic energy,
issued when the energy is asked for -- never per step.
"""
from __future__ import annotations


def shard_range(num_real: int, rank: int, world: int, granule: int = 32) -> tuple[int, int]:
    """Real-atom range [lo, hi) of ``rank``: balanced, contiguous, boundaries on multiples of
    ``granule`` atoms (keeps every shard's arrays 128-byte aligned when carved out of one slab)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    units = (num_real + granule - 1) // granule
    lo = min(num_real, units * rank // world * granule)
    hi = min(num_real, units * (rank + 1) // world * granule)
    return lo, hi


def all_ranges(num_real: int, world: int, granule: int = 32) -> list[tuple[int, int]]:
    return [shard_range(num_real, r, world, granule) for r in range(world)]


def molecule_ranges(num_real: int, world: int, groups, granule: int = 32) -> list[tuple[int, int]]:
    """Ranges for a slab whose step couples atoms (rigid molecules, Drude pairs): the balanced boundaries of
    ``all_ranges`` moved down to the nearest point that cuts no group.  ``groups`` = (n, k) array of the slots
    that must stay together (cluster atoms, or (Drude particle, parent) pairs).  A group is never split, so the
    partitioned step is still free of inter-GPU traffic and reproduces the single-GPU trajectory bit for bit."""
    import numpy as np
    g = np.asarray(groups).reshape(len(groups), -1)
    first, last = (g.min(axis=1), g.max(axis=1)) if len(g) else (np.zeros(0, int), np.zeros(0, int))
    # cut[b] = number of groups that a boundary placed before slot b would split
    cut = np.zeros(num_real + 2, np.int64)
    np.add.at(cut, first + 1, 1)
    np.add.at(cut, last + 1, -1)
    cut = np.cumsum(cut)
    bounds = [0]
    for r in range(1, world):
        b = shard_range(num_real, r, world, granule)[0]
        while b > bounds[-1] and cut[b] != 0:
            b -= 1
        bounds.append(max(b, bounds[-1]))
    bounds.append(num_real)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def slice_groups(groups, lo: int, hi: int):
    """The groups (cluster atoms / Drude pairs) of the shard [lo, hi) in the shard's own slot numbering, and the
    mask that selects them from the full list."""
    import numpy as np
    g = np.asarray(groups).reshape(len(groups), -1)
    mask = (g.min(axis=1) >= lo) & (g.max(axis=1) < hi) if len(g) else np.zeros(0, bool)
    return (g[mask] - lo).astype(np.int32), mask


def allreduce_scalar_sum(value_tensor):
    """Sum a 1-element tensor over the process group in place (NCCL on GPUs over NVLink/NVSwitch,
    gloo in the CPU tests); a no-op without an initialised group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(value_tensor, op=dist.ReduceOp.SUM)
    return value_tensor


def allreduce_scalar_max(value_tensor):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(value_tensor, op=dist.ReduceOp.MAX)
    return value_tensor


def temperature_from_kinetic_energy(ke_total: float, num_massive_atoms: int, boltz: float,
                                    num_constraints: int = 0) -> float:
    """T = 2 KE / (dof k_B) after the reduction (dof = 3 per massive atom minus constraints)."""
    dof = 3 * num_massive_atoms - num_constraints
    return 2.0 * ke_total / (dof * boltz) if dof > 0 else 0.0
