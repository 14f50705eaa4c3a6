"""Parity of the ConstVDrudeLangevinIntegrator CUDA path, through the C ABI (constv_drude_*):

* against the CPU oracle (oracle/constv_oracle_drude_impl.h, pinned bit for bit on the reference's
  constVDrudeLangevin.cu by tests/test_oracle_drude_vs_reference_source.py), and
* against the REFERENCE'S OWN Drude kernels compiled unmodified for sm_100a and launched on the same B200
  in the order of CudaConstVKernels.cpp:307-344 (oracle/ref_drude_cuda.cu).

With injected Gaussians (the reference's pool addressing, constVDrudeLangevin.cu:18,30) every array must
agree BIT FOR BIT: fused step, split (constraint hand-off) path, reordered slots, hard wall on and off,
three precision models.
"""
import ctypes as C

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H
from tests.test_gpu_parity import assert_slab_bits_equal, download, gpu  # noqa: F401

pytestmark = pytest.mark.gpu

T, FRICTION, TD, FRICTION_D, DT = 300.0, 5.0, 1.0, 20.0, 0.001


def make_drude_context(gpu, s, pairs, max_dist=0.0, seed=77, flags=None, pool=None, **kw):
    from openmm_constv_b200 import _capi
    import torch
    sysm = gpu.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, s.num_cells * s.zmax))
    force = gpu.DrudeForce()
    for p, p1 in pairs:
        force.addParticle(int(p), int(p1), -1, -1, -1, -1.8476, 0.001, 1.0, 1.0)
    sysm.addForce(force)
    it = gpu.ConstVDrudeLangevinIntegrator(T, FRICTION, TD, FRICTION_D, DT, s.num_cells)
    it.setMaxDrudeDistance(max_dist)
    it.setRandomNumberSeed(seed)
    ctx = gpu.SlabContext(sysm, it, precision=s.precision,
                          flags=_capi.FLAGS_DEFAULT if flags is None else flags,
                          random_pool=None if pool is None else torch.as_tensor(pool).cuda(),
                          global_atom_offset=s.global_atom_offset, padded=s.padded, **kw)
    ctx.load_slab(s)
    return ctx, it


def oracle_drude_step(oracle, s, normal, pairs, scalars, xi, random_index=0, inv=None, zero=True):
    m = H.MIXED[s.precision]
    delta = np.zeros((s.padded, 4), m)
    oracle.drude_step(s.precision, s.num_atoms, s.num_cells, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1), delta,
                      normal, pairs, scalars.astype(m), m(DT), xi, random_index,
                      H.identity_order(s.padded) if inv is None else inv, zero, zero)
    return delta


def mixed_scalars(oracle, precision, max_dist):
    """The eight coefficients as the kernels receive them (CudaConstVKernels.cpp:274-301)."""
    return oracle.drude_params(T, FRICTION, TD, FRICTION_D, max_dist, DT).astype(H.MIXED[precision])


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("max_dist", [0.0, 0.02])
@pytest.mark.parametrize("num_real", [41, 6007])
@pytest.mark.parametrize("variant", [0, 1, 5])
def test_drude_fused_step_bit_exact(gpu, oracle, precision, num_cells, max_dist, num_real, variant):
    """variant 0 = the staged kernel whose warps draw chunks of work items (two slots per thread), 1 = the gather kernel,
    5 = the staged kernel with one work item per thread."""
    import torch
    s0, normal, pairs = slabmod.make_drude_slab(num_real, precision, num_cells=num_cells, seed=11)
    nsteps, R = 4, s0.num_real
    pool = np.random.default_rng(5).standard_normal((nsteps * R + 7, 4)).astype(np.float32)
    ctx, it = make_drude_context(gpu, s0, pairs, max_dist, pool=pool)
    it._kernel.set_kernel_variant(variant)
    assert it._kernel.random_count == len(normal) + 2 * len(pairs) == R
    ref = s0.copy()
    scalars = mixed_scalars(oracle, precision, max_dist)
    for k in range(nsteps):
        f = slabmod.new_forces(s0, k)
        ctx.force.copy_(torch.from_numpy(f))
        ref.force[:] = f
        it.step(1)
        oracle_drude_step(oracle, ref, normal, pairs, scalars, pool, random_index=k * R)
        assert_slab_bits_equal(download(ctx, s0), ref)
    assert ctx.getStepCount() == nsteps and it._kernel.get_step_count() == nsteps
    if max_dist > 0 and len(pairs):
        d = ref.posq[pairs[:, 0], :3].astype(np.float64) - ref.posq[pairs[:, 1], :3]
        assert np.max(np.linalg.norm(d, axis=1)) < max_dist * (1 + 1e-3)
    ctx.close()


@pytest.mark.parametrize("num_real", [253, 255, 256, 257, 259, 511, 513, 767, 769, 1023, 1025, 1279])
def test_drude_sizes_around_the_tile_boundaries(gpu, oracle, num_real):
    """The balanced kernel stages 256 slots per CTA (two per thread): real ranges that end just before, on and just after a
    tile boundary, with the 4-site molecules starting at every residue of 4 (the ion count varies with the size); hard
    wall on, the reference's pool addressing."""
    import torch
    s0, normal, pairs = slabmod.make_drude_slab(num_real, "single", seed=100 + num_real)
    nsteps, R = 3, s0.num_real
    pool = np.random.default_rng(num_real).standard_normal((nsteps * R + 7, 4)).astype(np.float32)
    ctx, it = make_drude_context(gpu, s0, pairs, 0.02, pool=pool)
    ref = s0.copy()
    scalars = mixed_scalars(oracle, "single", 0.02)
    for k in range(nsteps):
        f = slabmod.new_forces(s0, k)
        ctx.force.copy_(torch.from_numpy(f))
        ref.force[:] = f
        it.step(1)
        oracle_drude_step(oracle, ref, normal, pairs, scalars, pool, random_index=k * R)
        assert_slab_bits_equal(download(ctx, s0), ref)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("shift", [1, 37, 300])
@pytest.mark.parametrize("variant", [0, 5])
def test_drude_pairs_across_tiles(gpu, oracle, precision, shift, variant):
    """Pairs whose parent sits in another tile (128 or 256 slots) than the Drude particle (here: every Drude particle
    paired with the parent `shift` molecules further on; wall off, the geometry is meaningless): the owner
    handles such a parent in global memory and the parent's own thread stays silent."""
    s0, normal, pairs = slabmod.make_drude_slab(5003, precision, seed=19)
    pairs = np.stack([pairs[:, 0], np.roll(pairs[:, 1], -shift)], axis=1).astype(np.int32)
    pool = np.random.default_rng(10).standard_normal((2 * s0.num_real, 4)).astype(np.float32)
    ctx, it = make_drude_context(gpu, s0, pairs, 0.0, pool=pool)
    it._kernel.set_kernel_variant(variant)
    ref = s0.copy()
    scalars = mixed_scalars(oracle, precision, 0.0)
    tile = 256 if variant == 0 else 128
    far = np.sum(pairs[:, 0] // tile != pairs[:, 1] // tile)
    assert far > (0 if shift == 1 else len(pairs) // 2)
    for k in range(2):
        it.step(1)
        oracle_drude_step(oracle, ref, normal, pairs, scalars, pool, random_index=k * s0.num_real)
        assert_slab_bits_equal(download(ctx, s0), ref)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("max_dist", [0.0, 0.02])
def test_drude_split_path_equals_oracle(gpu, oracle, precision, max_dist):
    """part 1 | constraint hand-off | part 2 + hard wall + image rewrite (one kernel)."""
    s0, normal, pairs = slabmod.make_drude_slab(3001, precision, num_cells=4, seed=12)
    pool = np.random.default_rng(6).standard_normal((s0.num_real, 4)).astype(np.float32)
    seen = {}
    ctx, it = make_drude_context(gpu, s0, pairs, max_dist, pool=pool,
                                 apply_constraints=lambda c, tol: seen.update(tol=tol, delta=c.pos_delta.cpu().numpy()))
    it.step(1)
    ref = s0.copy()
    delta = oracle_drude_step(oracle, ref, normal, pairs, mixed_scalars(oracle, precision, max_dist), pool)
    assert seen["tol"] == 1e-5
    R = s0.num_real
    written = np.zeros(s0.padded, bool)
    written[:R] = s0.velm[:R, 3] != 0
    written[pairs.reshape(-1)] = True
    H.assert_bits_equal(seen["delta"][written], delta[written], "posDelta")
    assert_slab_bits_equal(download(ctx, s0), ref)
    assert it._kernel.get_step_count() == 1
    ctx.close()


def role_preserving_permutation(s, rng):
    """OpenMM reorders atoms by swapping whole identical molecules, so a slot keeps its role: shuffle
    the 4-site molecules among themselves (as blocks) and the image particles freely among the image
    slots."""
    R, N = s.num_real, s.num_atoms
    n_ion, n_mol = s.meta["n_ion"], s.meta["n_pairs"]
    perm = np.arange(N, dtype=np.int32)
    mol_order = rng.permutation(n_mol)
    for k in range(4):
        perm[n_ion + k: n_ion + 4 * n_mol: 4] = n_ion + 4 * mol_order + k
    perm[R:N] = R + rng.permutation(N - R)
    return perm


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("split,variant", [(False, 0), (False, 1), (False, 5), (True, 0)])
@pytest.mark.parametrize("side_array,num_cells", [(False, 2), (True, 2), (True, 4)])
def test_drude_reordered_slots(gpu, oracle, precision, split, variant, side_array, num_cells):
    """variant 0 | 5 = the staged kernels, variant 1 = thread per slot.  side_array: with the image-charge snapshot the
    balanced kernel reads the images' slots and charges from the per-slot tables instead of gathering them."""
    from openmm_constv_b200 import _capi
    s0, normal, pairs = slabmod.make_drude_slab(2501, precision, num_cells=num_cells, seed=13)
    perm = role_preserving_permutation(s0, np.random.default_rng(4))
    sp, atom_index = H.permute_slab(s0, perm)
    inv = np.empty_like(atom_index)
    inv[atom_index] = np.arange(len(atom_index), dtype=np.int32)
    pool = np.random.default_rng(8).standard_normal((2 * s0.num_real, 4)).astype(np.float32)
    ctx, it = make_drude_context(gpu, sp, pairs, 0.02, pool=pool,
                                 flags=_capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side_array else 0),
                                 apply_constraints=(lambda c, t: None) if split else None)
    ctx.set_atom_order(perm)
    it._kernel.set_kernel_variant(variant)
    ref = sp.copy()
    scalars = mixed_scalars(oracle, precision, 0.02)
    for k in range(2):
        it.step(1)
        oracle_drude_step(oracle, ref, normal, pairs, scalars, pool, random_index=k * s0.num_real, inv=inv)
        assert_slab_bits_equal(download(ctx, sp), ref)
    ctx.close()


@pytest.fixture(scope="module")
def refmod():
    from oracle import ref
    if not ref.available("cuda"):
        pytest.skip("oracle/_ref/libconstv_ref_cuda.so is not built (needs /root/reference at build time)")
    if not hasattr(ref.cuda(), "constv_ref_drude_cuda_part1_single"):
        pytest.skip("the reference build has no Drude kernels")
    return ref


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("max_dist", [0.0, 0.02])
@pytest.mark.parametrize("split", [False, True])
def test_drude_kernels_equal_reference_kernels(gpu, oracle, refmod, precision, max_dist, split):
    """Image zeroing off = the reference's behaviour; its four launches against our one (or two)."""
    import torch
    s0, normal, pairs = slabmod.make_drude_slab(20011, precision, num_cells=2, seed=14)
    R, nsteps = s0.num_real, 3
    pools = [np.random.default_rng(70 + k).standard_normal((R, 4)).astype(np.float32) for k in range(nsteps)]
    dev = torch.device("cuda", 0)
    scalars = oracle.drude_params(T, FRICTION, TD, FRICTION_D, max_dist, DT)
    r = refmod.CudaDrudeReference(precision, scalars, H.MIXED[precision](DT), dev, normal, pairs, R, s0.padded,
                                  max_blocks=6 * 148)
    posq, velm, force = (torch.as_tensor(a, device=dev) for a in (s0.posq, s0.velm, s0.force))
    corr = torch.as_tensor(s0.corr, device=dev) if s0.corr is not None else None
    inv = torch.arange(s0.padded, dtype=torch.int32, device=dev)
    delta = torch.zeros_like(velm)
    for xi in pools:
        r.drude_step(s0.num_atoms, s0.num_cells, s0.zmax, posq, corr, velm, force, delta, torch.as_tensor(xi, device=dev), 0, inv)
    torch.cuda.synchronize()
    want = s0.copy()
    want.posq, want.velm, want.force = posq.cpu().numpy(), velm.cpu().numpy(), force.cpu().numpy()
    if corr is not None:
        want.corr = corr.cpu().numpy()

    ctx, it = make_drude_context(gpu, s0, pairs, max_dist, flags=0, pool=np.concatenate(pools),
                                 apply_constraints=(lambda c, t: None) if split else None)
    it.step(nsteps)
    assert_slab_bits_equal(download(ctx, s0), want)
    ctx.close()
    # ... and the CPU oracle agrees with the reference kernels as nvcc contracts them
    orc = s0.copy()
    for xi in pools:
        oracle_drude_step(oracle, orc, normal, pairs, scalars.astype(H.MIXED[precision]), xi, zero=False)
    assert_slab_bits_equal(orc, want)


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_drude_hardwall_kernel_alone(gpu, oracle, precision):
    """constv_drude_hardwall = applyHardWallConstraints (constVDrudeLangevin.cu:108-211), including the
    massless-parent branch (:141-165) that a full step cannot reach (part 1 divides by the parent's mass)."""
    s0, normal, pairs = slabmod.make_drude_slab(4001, precision, seed=15)
    half = pairs[: len(pairs) // 2, 1]
    s0.velm[half, 3] = 0
    s0.velm[half, :3] = 0
    s0.masses[half] = 0
    ctx, it = make_drude_context(gpu, s0, pairs, 0.01)
    lib, slab = it._kernel._lib, it._kernel.handle
    from openmm_constv_b200._capi import check
    check(lib.constv_drude_set_parameters(slab, T, FRICTION, TD, FRICTION_D, 0.01, DT))
    check(lib.constv_drude_hardwall(slab, ctx.stream_ptr()))
    ref = s0.copy()
    sc = mixed_scalars(oracle, precision, 0.01)
    oracle.drude_hardwall(precision, ref.posq, ref.corr, ref.velm, pairs, H.MIXED[precision](DT), sc[6], sc[7])
    got = download(ctx, s0)
    assert_slab_bits_equal(got, ref)
    assert not np.array_equal(got.posq, s0.posq)
    assert np.array_equal(got.posq[half], s0.posq[half])  # massless parents stay put
    ctx.close()


@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_drude_philox_stream(gpu, oracle, precision):
    """random4 == NULL: an ordinary particle draws the Philox block of its own atom index, a pair the
    Drude particle's block for the centre of mass and the parent's for the relative motion.  The oracle
    is driven with exactly those Gaussians (as the device generated them) laid out in the reference's pool
    addressing, once for the ordinary particles and once for the pairs (the two regions of the
    reference's pool overlap, constVDrudeLangevin.cu:18 vs :30)."""
    import torch
    s0, normal, pairs = slabmod.make_drude_slab(5003, precision, seed=16)
    R = s0.num_real
    ctx, it = make_drude_context(gpu, s0, pairs, 0.02, seed=1234)
    lib, slab = it._kernel._lib, it._kernel.handle
    from openmm_constv_b200._capi import check
    g = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    check(lib.constv_generate_noise(slab, 0, g.data_ptr(), ctx.stream_ptr()))
    g = g.cpu().numpy()
    it.step(1)
    got = download(ctx, s0)
    scalars = mixed_scalars(oracle, precision, 0.02)
    # ordinary particles
    a = s0.copy()
    oracle_drude_step(oracle, a, normal, pairs, scalars, g)
    # pairs
    b = s0.copy()
    pool = np.zeros((R + 1, 4), np.float32)
    pool[len(normal) + 2 * np.arange(len(pairs))] = g[pairs[:, 0]]
    pool[len(normal) + 2 * np.arange(len(pairs)) + 1] = g[pairs[:, 1]]
    oracle_drude_step(oracle, b, normal, pairs, scalars, pool)
    for arr in ("posq", "velm") + (("corr",) if s0.corr is not None else ()):
        ga, aa, ba = getattr(got, arr), getattr(a, arr), getattr(b, arr)
        H.assert_bits_equal(ga[normal], aa[normal], arr + " (ordinary particles)")
        H.assert_bits_equal(ga[R + normal], aa[R + normal], arr + " (images of ordinary particles)")
        pp = pairs.reshape(-1)
        H.assert_bits_equal(ga[pp], ba[pp], arr + " (pairs)")
        H.assert_bits_equal(ga[R + pp], ba[R + pp], arr + " (images of pairs)")
    # a second context with the same seed reproduces it; another seed does not
    ctx2, it2 = make_drude_context(gpu, s0, pairs, 0.02, seed=1234)
    it2.step(1)
    assert_slab_bits_equal(download(ctx2, s0), got)
    ctx3, it3 = make_drude_context(gpu, s0, pairs, 0.02, seed=99)
    it3.step(1)
    assert not np.array_equal(download(ctx3, s0).velm, got.velm)
    for c in (ctx, ctx2, ctx3):
        c.close()


@pytest.mark.parametrize("precision", ["single", "mixed"])
@pytest.mark.parametrize("chains", [2, 7])
def test_drude_chains(gpu, precision, chains):
    """The fused Drude step as chains: the single launch's bits, in turn on one stream or one stream per chain."""
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0, normal, pairs = slabmod.make_drude_slab(20011, precision, seed=21)
    nsteps = 4
    ctx_a, it_a = make_drude_context(gpu, s0, pairs, 0.02, seed=6)
    it_a.step(nsteps)
    want = download(ctx_a, s0)
    ctx_a.close()
    ctx_b, it_b = make_drude_context(gpu, s0, pairs, 0.02, seed=6)
    _capi.check(lib.constv_set_chains(it_b._kernel.handle, chains))
    it_b.step(nsteps)
    assert it_b._kernel.get_step_count() == nsteps
    assert_slab_bits_equal(download(ctx_b, s0), want)
    ctx_b.close()
    ctx_c, it_c = make_drude_context(gpu, s0, pairs, 0.02, seed=6)
    h = it_c._kernel.handle
    _capi.check(lib.constv_set_chains(h, chains))
    _capi.check(lib.constv_drude_set_parameters(h, T, FRICTION, TD, FRICTION_D, 0.02, DT))
    streams = [torch.cuda.Stream() for _ in range(chains)]
    torch.cuda.synchronize()
    for q in reversed(range(chains)):
        for _ in range(nsteps):
            _capi.check(lib.constv_drude_step_fused_chain(h, q, None, 0, streams[q].cuda_stream))
    torch.cuda.synchronize()
    assert it_c._kernel.get_step_count() == nsteps
    assert_slab_bits_equal(download(ctx_c, s0), want)
    assert lib.constv_drude_step_fused_chain(h, -1, None, 0, None) == _capi.ERR_INVALID
    ctx_c.close()


def test_drude_relative_motion_thermalises_at_the_drude_temperature(gpu):
    """Free pairs (no forces), Philox noise: after many relaxation times the pair's relative velocity
    carries kB*T_drude per degree of freedom and its centre of mass kB*T (the dual thermostat of
    constVDrudeLangevin.cu:31-63)."""
    s0, normal, pairs = slabmod.make_drude_slab(40000, "single", seed=17)
    s0.force[:] = 0
    ctx, it = make_drude_context(gpu, s0, pairs, 0.0, seed=5)
    it.setFriction(50.0)
    it.setDrudeFriction(50.0)
    it.step(400)  # 20 relaxation times
    v = ctx.velm.cpu().numpy().astype(np.float64)
    w1, w2 = v[pairs[:, 0], 3], v[pairs[:, 1], 3]
    m1, m2 = 1 / w1, 1 / w2
    vrel = v[pairs[:, 1], :3] - v[pairs[:, 0], :3]
    vcm = (m1[:, None] * v[pairs[:, 0], :3] + m2[:, None] * v[pairs[:, 1], :3]) / (m1 + m2)[:, None]
    mu = m1 * m2 / (m1 + m2)
    kB = 8.31446261815324e-3
    t_rel = np.mean(mu[:, None] * vrel ** 2) / kB
    t_cm = np.mean((m1 + m2)[:, None] * vcm ** 2) / kB
    assert abs(t_rel - TD) < 0.03 * TD
    assert abs(t_cm - T) < 0.03 * T
    ctx.close()


def test_drude_error_paths(gpu):
    from openmm_constv_b200 import _capi
    s0, normal, pairs = slabmod.make_drude_slab(301, "single", seed=18)
    sysm = gpu.SlabSystem()
    sysm.addParticles(s0.masses)
    sysm.setDefaultPeriodicBoxVectors((s0.box[0], 0, 0), (0, s0.box[1], 0), (0, 0, s0.num_cells * s0.zmax))
    it = gpu.ConstVDrudeLangevinIntegrator(T, FRICTION, TD, FRICTION_D, DT)
    with pytest.raises(gpu.OpenMMException, match="The System does not contain a DrudeForce"):
        gpu.SlabContext(sysm, it, padded=s0.padded)
    sysm.addForce(gpu.DrudeForce())
    sysm.addForce(gpu.DrudeForce())
    with pytest.raises(gpu.OpenMMException, match="The System contains multiple DrudeForces"):
        gpu.SlabContext(sysm, gpu.ConstVDrudeLangevinIntegrator(T, FRICTION, TD, FRICTION_D, DT), padded=s0.padded)
    with pytest.raises(gpu.OpenMMException, match="Distance cannot be negative"):
        it.setMaxDrudeDistance(-1.0)
    with pytest.raises(gpu.OpenMMException, match="not bound to a context"):
        gpu.ConstVDrudeLangevinIntegrator(T, FRICTION, TD, FRICTION_D, DT).step(1)
    # C ABI: the Drude calls need a configured pair list and Drude coefficients; a Drude slab refuses the
    # plain Langevin step, and a particle cannot sit in two pairs
    ctx, it2 = make_drude_context(gpu, s0, pairs)
    lib, slab = it2._kernel._lib, it2._kernel.handle
    assert lib.constv_drude_step_fused(slab, None, 0, None) == _capi.ERR_STATE  # parameters not set yet
    it2.step(1)
    assert lib.constv_step_fused(slab, None, 0, None) == _capi.ERR_STATE
    assert b"ConstVDrudeLangevinIntegrator" in lib.constv_last_error()
    bad = np.array([[pairs[0, 0], pairs[0, 1]], [pairs[0, 0], pairs[1, 1]]], np.int32)
    assert lib.constv_drude_configure(slab, 2, bad.ctypes.data_as(C.c_void_p)) == _capi.ERR_INVALID
    bad = np.array([[s0.num_real, 0]], np.int32)  # an image particle
    assert lib.constv_drude_configure(slab, 1, bad.ctypes.data_as(C.c_void_p)) == _capi.ERR_INVALID
    it2.step(1)  # the failed configure calls left the slab usable
    ctx.close()
    # a Drude System with constraints and no solver to hand them to
    sysm = gpu.SlabSystem()
    sysm.addParticles(s0.masses)
    sysm.setDefaultPeriodicBoxVectors((s0.box[0], 0, 0), (0, s0.box[1], 0), (0, 0, s0.num_cells * s0.zmax))
    force = gpu.DrudeForce()
    for p, p1 in pairs:
        force.addParticle(int(p), int(p1))
    sysm.addForce(force)
    sysm.addConstraint(int(normal[0]), int(normal[1]), 0.1)
    it3 = gpu.ConstVDrudeLangevinIntegrator(T, FRICTION, TD, FRICTION_D, DT)
    ctx3 = gpu.SlabContext(sysm, it3, padded=s0.padded)
    ctx3.load_slab(s0)
    with pytest.raises(gpu.OpenMMException, match="apply_constraints"):
        it3.step(1)
    ctx3.close()


import glob  # noqa: E402
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "drude_ref_host_*.npz"))))
def test_drude_committed_reference_host_vectors(gpu, path):
    """Fixtures recorded from the reference's Drude kernel source compiled for the HOST, where every
    operation is rounded separately; the GPU contracts multiply-adds as nvcc does for the reference's own
    kernels, so agreement is to rounding: 1e-6 of the thermal velocity / of the coordinate range per step
    (north_star's fp32 tolerance), 1e-11 where velocities and positions are carried in double."""
    import torch
    g = np.load(path)
    precision, num_cells, num_atoms, steps = str(g["precision"]), int(g["num_cells"]), int(g["num_atoms"]), int(g["steps"])
    s = slabmod.make_drude_slab(num_atoms // num_cells, precision, num_cells=num_cells, seed=1)[0]
    s.posq, s.velm, s.force, s.masses = g["posq0"].copy(), g["velm0"].copy(), g["force"].copy(), g["masses"].copy()
    if precision == "mixed":
        s.corr = g["corr0"].copy()
    s.zmax = float(g["zmax"])
    R = num_atoms // num_cells
    pool = np.zeros((steps * R + 8, 4), np.float32)
    off = int(g["random_index"])
    for k in range(steps):  # the context advances its pool index by R per step, starting at 0
        pool[k * R: k * R + R] = g["noise"][k][off: off + R]
    sc = g["scalars"]
    ctx, it = make_drude_context(gpu, s, g["pairs"], float(sc[6]), flags=0, pool=pool)
    ctx.velm.copy_(torch.as_tensor(s.velm))
    it.step(steps)
    n = num_atoms
    got_v, got_p = ctx.velm.cpu().numpy().astype(np.float64), ctx.posq.cpu().numpy().astype(np.float64)
    want_v, want_p = g["velm1"].astype(np.float64), g["posq1"].astype(np.float64)
    if precision == "mixed":
        got_p[:, :3] += ctx.corr.cpu().numpy()[:, :3]
        want_p[:, :3] += g["corr1"][:, :3]
    massive = g["velm0"][:R, 3] != 0
    v_thermal = np.sqrt(np.mean(g["velm0"][:R, :3][massive].astype(np.float64) ** 2))
    tol = 1e-6 if precision == "single" else 1e-11
    assert np.max(np.abs(got_v[:n, :3] - want_v[:n, :3])) < steps * tol * v_thermal * 20
    assert np.max(np.abs(got_p[:n, :3] - want_p[:n, :3])) < steps * tol * np.max(np.abs(want_p[:n, :3]))
    ctx.close()
