"""The multi-GPU workloads are ONE seeded global slab made of lateral tiles (slab.slab_tile): a rank builds only
its own tile, and that tile must be exactly the partition's slice of the unpartitioned slab -- arrays, masses
and global atom offset -- for the plain, rigid-water and Drude slabs.  bench.py's N > 1 arm, its 16M-atom leg and
its partition-parity leg all rest on this."""
import numpy as np
import pytest

from openmm_constv_b200 import partition, slab as slabmod


@pytest.mark.parametrize("precision", ["single", "mixed"])
@pytest.mark.parametrize("tiles", [1, 2, 4, 8])
def test_tile_is_the_partition_slice_of_the_global_slab(precision, tiles):
    R = 640 * tiles
    parts = [slabmod.slab_tile(R, tiles, t, precision, seed=99) for t in range(tiles)]
    full = slabmod.concat_slabs(parts)
    assert full.num_real == R and full.num_atoms == 2 * R and full.global_atom_offset == 0
    for t, tile in enumerate(parts):
        lo, hi = partition.shard_range(R, t, tiles)
        assert (lo, hi) == (tile.global_atom_offset, tile.global_atom_offset + tile.num_real)
        cut = slabmod.slice_slab(full, lo, hi)
        assert cut.global_atom_offset == lo
        for name in ("posq", "velm", "force", "masses"):
            assert np.array_equal(getattr(cut, name), getattr(tile, name)), name
        if precision == "mixed":
            assert np.array_equal(cut.corr, tile.corr)
    # tiles sit side by side in x and have the same composition
    xs = [(p.posq[:p.num_real, 0].min(), p.posq[:p.num_real, 0].max()) for p in parts]
    for (a0, a1), (b0, b1) in zip(xs, xs[1:]):
        assert a1 <= b0 + 1e-6
    assert len({tuple(sorted(p.pair_classes().items())) for p in parts}) == 1
    assert abs(full.box[0] - tiles * parts[0].box[0]) < 1e-9


def test_tile_count_must_cut_on_32_atom_boundaries():
    with pytest.raises(ValueError):
        slabmod.slab_tile(1000, 3, 0)
    with pytest.raises(ValueError):
        slabmod.slab_tile(100, 2, 0)  # 50 real atoms per tile is not a multiple of 32
    assert slabmod.slab_tile(100, 1, 0).num_real == 100  # a single tile is any size


def test_water_and_drude_tiles_keep_their_groups_local():
    s, atoms, params, cons = slabmod.slab_tile(1280, 2, 1, "single", seed=5, maker=slabmod.make_water_slab)
    assert s.global_atom_offset == 640 and atoms.max() < s.num_real and len(cons[0]) == 3 * len(atoms)
    d, normal, pairs = slabmod.slab_tile(1280, 2, 1, "single", seed=5, maker=slabmod.make_drude_slab)
    assert d.global_atom_offset == 640 and pairs.max() < d.num_real


def test_bench_workload_config_is_the_same_object_in_both_arms():
    """The B200 arm and --impl reference must print an identical `config`."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n in (1, 2, 4, 8):
        c = bench.workload_config(1_000_000, n, "single")
        assert c["total_atoms"] == 1_000_000 * n and c["atoms_per_gpu"] == 1_000_000
        assert c == bench.workload_config(1_000_000, n, "single")
        c16 = bench.workload_config(1_000_000, n, "single", True)
        assert c16["total_atoms"] == 16_000_000 and c16["atoms_per_gpu"] == 16_000_000 // n
    assert bench.auto_replicas(56_000_000) == 6 and bench.auto_replicas(900_000_000) == 1
