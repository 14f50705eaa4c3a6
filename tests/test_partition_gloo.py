"""The N > 1 path on CPU: world_size-2 (and 3) gloo process groups exercise the range partition,
the global-index keyed noise stream and the scalar kinetic-energy all-reduce, with the oracle
standing in for the GPU step on each rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openmm_constv_b200 import partition, slab as slabmod


def test_ranges_cover_the_slab_exactly_once():
    for R in (1, 31, 32, 33, 1000, 500_000, 8_000_001):
        for world in (1, 2, 3, 4, 8):
            rs = partition.all_ranges(R, world)
            assert rs[0][0] == 0 and rs[-1][1] == R
            for (a, b), (c, d) in zip(rs, rs[1:]):
                assert b == c and a <= b
            assert all(lo % 32 == 0 for lo, _ in rs)
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 32 or R < 32 * world


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        full = slabmod.make_slab(5003, "single", seed=77)
        lo, hi = partition.shard_range(full.num_real, rank, world)
        shard = slabmod.slice_slab(full, lo, hi)
        assert shard.global_atom_offset == lo
        prm = oracle.params(300.0, 1.0, 0.001).astype(np.float32)
        inv = np.arange(shard.padded, dtype=np.int32)
        delta = np.zeros((shard.padded, 4), np.float32)
        for k in range(3):
            xi = oracle.fill_noise(shard.padded, 5, k, global_atom_offset=shard.global_atom_offset)
            oracle.step("single", shard.num_atoms, 2, shard.zmax, shard.posq, None, shard.velm,
                        shard.force.reshape(-1), delta, prm, np.float32(0.001), xi, 0, inv, True, True)
        ke = torch.tensor([oracle.kinetic_energy("single", shard.velm, shard.force.reshape(-1), 0.0005,
                                                 shard.num_atoms)], dtype=torch.float64)
        partition.allreduce_scalar_sum(ke)          # the one collective of the partition
        tmax = partition.allreduce_scalar_max(torch.tensor([float(rank)], dtype=torch.float64))
        np.savez(os.path.join(out, f"rank{rank}.npz"), lo=lo, hi=hi, posq=shard.posq, velm=shard.velm,
                 force=shard.force, ke=ke.numpy(), tmax=tmax.numpy(), n=shard.num_atoms)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_run_reproduces_the_single_rank_trajectory(tmp_path, oracle, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = slabmod.make_slab(5003, "single", seed=77)
    prm = oracle.params(300.0, 1.0, 0.001).astype(np.float32)
    inv = np.arange(full.padded, dtype=np.int32)
    delta = np.zeros((full.padded, 4), np.float32)
    for k in range(3):
        xi = oracle.fill_noise(full.padded, 5, k)
        oracle.step("single", full.num_atoms, 2, full.zmax, full.posq, None, full.velm,
                    full.force.reshape(-1), delta, prm, np.float32(0.001), xi, 0, inv, True, True)
    ke_full = oracle.kinetic_energy("single", full.velm, full.force.reshape(-1), 0.0005, full.num_atoms)
    R = full.num_real
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        r = hi - lo
        # real block and image block of the shard against the same atoms of the unpartitioned run
        for name, ref in (("posq", full.posq), ("velm", full.velm)):
            got = z[name]
            assert np.array_equal(got[:r].view(np.uint32), ref[lo:hi].view(np.uint32)), (name, rank)
            assert np.array_equal(got[r:2 * r].view(np.uint32), ref[R + lo:R + hi].view(np.uint32)), (name, rank)
        assert np.array_equal(z["force"][:, :r], full.force[:, lo:hi])
        assert float(z["ke"][0]) == pytest.approx(ke_full, rel=1e-12)
        assert float(z["tmax"][0]) == world - 1


def test_molecule_ranges_never_split_a_group():
    for seed, world in ((1, 2), (2, 3), (3, 8)):
        s, atoms, params, _ = slabmod.make_water_slab(4001, "single", seed=seed)
        rs = partition.molecule_ranges(s.num_real, world, atoms)
        assert rs[0][0] == 0 and rs[-1][1] == s.num_real
        for (a, b), (c, d) in zip(rs, rs[1:]):
            assert b == c and a <= b
        owner = np.searchsorted([hi for _, hi in rs], atoms, side="right")
        assert np.all(owner[:, 0] == owner[:, 1]) and np.all(owner[:, 0] == owner[:, 2])
        total = 0
        for lo, hi in rs:
            local, mask = partition.slice_groups(atoms, lo, hi)
            total += len(local)
            assert np.all(local >= 0) and np.all(local < hi - lo)
        assert total == len(atoms)
        balanced = partition.all_ranges(s.num_real, world)
        assert all(abs(a[0] - b[0]) <= 3 for a, b in zip(rs, balanced))  # moved by less than one molecule
    # scattered pairs (Drude particle far from its parent) can force a boundary far down, never a split
    pairs = np.array([[10, 1500], [1700, 1701]], np.int32)
    assert partition.molecule_ranges(2000, 2, pairs) == [(0, 10), (10, 2000)]
    assert partition.molecule_ranges(2000, 2, np.array([[10, 900]], np.int32)) == partition.all_ranges(2000, 2)


def _water_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        full, atoms, params, _ = slabmod.make_water_slab(5003, "single", seed=78)
        lo, hi = partition.molecule_ranges(full.num_real, world, atoms)[rank]
        shard = slabmod.slice_slab(full, lo, hi)
        local, mask = partition.slice_groups(atoms, lo, hi)
        prm = oracle.params(300.0, 1.0, 0.002).astype(np.float32)
        inv = np.arange(shard.padded, dtype=np.int32)
        delta = np.zeros((shard.padded, 4), np.float32)
        for k in range(3):
            xi = oracle.fill_noise(shard.padded, 5, k, global_atom_offset=shard.global_atom_offset)
            oracle.settle_step("single", shard.num_atoms, 2, shard.zmax, shard.posq, None, shard.velm, shard.force.reshape(-1),
                               delta, prm, np.float32(0.002), xi, 0, inv, local, params[mask])
        np.savez(os.path.join(out, f"water{rank}.npz"), lo=lo, hi=hi, posq=shard.posq, velm=shard.velm)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_rigid_water_reproduces_the_single_rank_trajectory(tmp_path, oracle, world):
    """The step that solves the rigid molecules' constraints itself shards the same way, as long as no molecule
    is cut (partition.molecule_ranges): still no data-path collective, still the same bits."""
    mp.spawn(_water_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full, atoms, params, _ = slabmod.make_water_slab(5003, "single", seed=78)
    prm = oracle.params(300.0, 1.0, 0.002).astype(np.float32)
    inv = np.arange(full.padded, dtype=np.int32)
    delta = np.zeros((full.padded, 4), np.float32)
    for k in range(3):
        xi = oracle.fill_noise(full.padded, 5, k)
        oracle.settle_step("single", full.num_atoms, 2, full.zmax, full.posq, None, full.velm, full.force.reshape(-1), delta, prm,
                           np.float32(0.002), xi, 0, inv, atoms, params)
    R = full.num_real
    for rank in range(world):
        z = np.load(tmp_path / f"water{rank}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        r = hi - lo
        for name, ref in (("posq", full.posq), ("velm", full.velm)):
            got = z[name]
            assert np.array_equal(got[:r].view(np.uint32), ref[lo:hi].view(np.uint32)), (name, rank)
            assert np.array_equal(got[r:2 * r].view(np.uint32), ref[R + lo:R + hi].view(np.uint32)), (name, rank)
