"""The integer stream under the Gaussians, pinned bit for bit ON THE DEVICE: the raw
Philox4x32-10 words the step kernels draw for (seed; global atom, step), read back through the debug entry
constv_generate_philox_words, against

  * the three Random123 known-answer vectors (Salmon et al., SC'11; kat_vectors, philox4x32 10 rounds), each
    reached through the entry's own (seed, global_atom_offset, step) arguments,
  * the C oracle's Philox (itself pinned on the same vectors, tests/test_oracle_philox.py) for every atom of a
    slab, for reordered slots, a shard's global_atom_offset, and steps >= 2^32,
  * a vectorised numpy restatement for all atoms at once.
"""
import ctypes as C

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod

pytestmark = pytest.mark.gpu

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def philox_numpy(atom, step, seed):
    """Philox4x32-10 for arrays of 64-bit atom indices, one step and one seed (uint64 arithmetic)."""
    atom = np.asarray(atom, np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    c = [atom & m32, atom >> np.uint64(32), np.full_like(atom, step & 0xFFFFFFFF), np.full_like(atom, step >> 32)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack(c, axis=1).astype(np.uint32)


def make_context(s, seed, offset=0):
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from openmm_constv_b200 import _capi, integrator as hostapi
    _capi.load()
    sysm = hostapi.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, s.num_cells * s.zmax))
    it = hostapi.ConstVLangevinIntegrator(300.0, 1.0, 0.001, s.num_cells)
    it.setRandomNumberSeed(seed)
    ctx = hostapi.SlabContext(sysm, it, precision=s.precision, global_atom_offset=offset, padded=s.padded)
    ctx.load_slab(s)
    return ctx, it


def words(ctx, it, step, n):
    import torch
    from openmm_constv_b200 import _capi
    out = torch.zeros((ctx.padded, 4), dtype=torch.int32, device=ctx.device)
    _capi.check(_capi.load().constv_generate_philox_words(it._kernel.handle, C.c_uint64(step), out.data_ptr(), ctx.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint32)[:n]


@pytest.mark.parametrize("ctr,key,expect", KAT)
def test_device_philox_known_answers(ctr, key, expect):
    """counter = (atom lo, atom hi, step lo, step hi), key = (seed lo, seed hi): slot 0 of a slab whose
    global_atom_offset is the vector's atom."""
    s = slabmod.make_slab(16, "single", seed=3)
    atom = ctr[0] | (ctr[1] << 32)
    step = ctr[2] | (ctr[3] << 32)
    seed = key[0] | (key[1] << 32)
    if seed == 0:
        # seed 0 means "pick one" at the integrator level (ConstVLangevinIntegrator.h:138-140): go through the C ABI config
        from openmm_constv_b200 import _capi
        import torch
        lib = _capi.load()
        cfg = _capi.Config()
        cfg.struct_size = C.sizeof(_capi.Config)
        cfg.device, cfg.num_atoms, cfg.padded_num_atoms, cfg.num_cells = 0, s.num_atoms, s.padded, 2
        cfg.precision, cfg.mirror_mode, cfg.flags, cfg.zmax = 0, 0, _capi.FLAGS_DEFAULT, s.zmax
        cfg.seed, cfg.global_atom_offset, cfg.boltz = 0, atom, 0.0
        h = C.c_void_p(None)
        _capi.check(lib.constv_create(C.byref(cfg), C.byref(h)))
        out = torch.zeros((s.padded, 4), dtype=torch.int32, device="cuda:0")
        _capi.check(lib.constv_generate_philox_words(h, C.c_uint64(step), out.data_ptr(), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint32)[0]
        lib.constv_destroy(h)
    else:
        ctx, it = make_context(s, seed, offset=atom)
        got = words(ctx, it, step, 1)[0]
        ctx.close()
    assert [int(x) for x in got] == expect


@pytest.mark.parametrize("step", [0, 1, 12345, 2**32 - 1, 2**32, 2**33 + 5, 2**64 - 1])
def test_device_philox_equals_oracle_for_every_atom(oracle, step):
    s = slabmod.make_slab(5000, "single", seed=4)
    seed, offset = 0xDEADBEEFCAFE, 0
    ctx, it = make_context(s, seed, offset)
    got = words(ctx, it, step, s.num_atoms)
    ctx.close()
    want = philox_numpy(np.arange(s.num_atoms, dtype=np.uint64) + np.uint64(offset), step, seed)
    assert np.array_equal(got, want)
    for a in (0, 1, 31, 32, 4999, 5000, 9999):  # and the C oracle, atom by atom
        o = oracle.philox4x32_10([a & 0xFFFFFFFF, a >> 32, step & 0xFFFFFFFF, step >> 32], [seed & 0xFFFFFFFF, seed >> 32])
        assert np.array_equal(got[a], o), a


@pytest.mark.parametrize("offset", [1, 300, 2**31, 2**32 - 7, 2**32 + 17, 2**40 + 3])
def test_device_philox_global_atom_offset(oracle, offset):
    """A shard of the multi-GPU range partition draws the words of its GLOBAL atom indices, across the 2^32
    boundary of the low counter word too."""
    s = slabmod.make_slab(700, "single", seed=5)
    seed, step = 99, 2**32 + 3
    ctx, it = make_context(s, seed, offset)
    got = words(ctx, it, step, s.num_atoms)
    ctx.close()
    atoms = np.arange(s.num_atoms, dtype=np.uint64) + np.uint64(offset)
    assert np.array_equal(got, philox_numpy(atoms, step, seed))
    for i in (0, 6, 7, 8, 699, 1399):
        a = int(atoms[i])
        o = oracle.philox4x32_10([a & 0xFFFFFFFF, a >> 32, step & 0xFFFFFFFF, step >> 32], [seed & 0xFFFFFFFF, seed >> 32])
        assert np.array_equal(got[i], o), i


def test_device_philox_follows_the_original_atom_index():
    """Reordered slots: slot s draws the words of atom_index[s], not of s."""
    s = slabmod.make_slab(1000, "single", seed=6)
    perm = np.random.default_rng(1).permutation(s.num_atoms).astype(np.int32)
    t, atom_index = slabmod.permute_slab(s, perm)
    ctx, it = make_context(t, 7)
    ctx.set_atom_order(perm)
    got = words(ctx, it, 11, s.num_atoms)
    ctx.close()
    assert np.array_equal(got, philox_numpy(perm.astype(np.uint64), 11, 7))


def test_gaussians_are_the_published_function_of_the_words(oracle):
    """constv_generate_noise == Box-Muller of exactly these words (4e-6: device SFU log/sin/cos against the
    oracle's double evaluation of the same definition, DESIGN.md "RNG")."""
    import torch
    from openmm_constv_b200 import _capi
    s = slabmod.make_slab(3000, "single", seed=8)
    ctx, it = make_context(s, 5150)
    w = words(ctx, it, 9, s.num_atoms).astype(np.float64)
    noise = torch.zeros((s.padded, 4), dtype=torch.float32, device=ctx.device)
    _capi.check(_capi.load().constv_generate_noise(it._kernel.handle, 9, noise.data_ptr(), ctx.stream_ptr()))
    xi = noise.cpu().numpy()[: s.num_atoms]
    ctx.close()
    u = (np.floor(w / 512.0) + 0.5) * 2.0 ** -23
    two_pi, pi = float(np.float32(6.2831854820251464844)), float(np.float32(3.1415927410125732422))
    r1, r2 = np.sqrt(-2 * np.log(u[:, 0])), np.sqrt(-2 * np.log(u[:, 2]))
    a1, a2 = u[:, 1] * two_pi - pi, u[:, 3] * two_pi - pi
    want = np.stack([r1 * np.cos(a1), r1 * np.sin(a1), r2 * np.cos(a2), r2 * np.sin(a2)], axis=1)
    assert np.max(np.abs(xi - want)) < 4e-6
    assert np.max(np.abs(xi - oracle.fill_noise(s.num_atoms, 5150, 9))) < 4e-6
