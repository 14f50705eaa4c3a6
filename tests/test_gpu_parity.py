"""Parity of the CUDA path against the CPU oracle, through the C ABI (via the Python host mirror,
which only forwards to include/constv_b200.h entry points).

Bars (BASELINE.json north_star): image positions bit-exact; positions and velocities within fp32
relative 1e-6 per step and over 1000-step friction-0 trajectories.  The CUDA path is written to the
oracle's canonical operation order, so with injected noise the tests below demand BIT-EXACT
equality of every array; the Philox Gaussians are compared within 4e-6 absolute (device float
log/sincospi against the oracle's double evaluation), the underlying stream being exact.
"""
import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests.helpers import (assert_bits_equal, identity_order, oracle_params_for, oracle_step,
                           permute_slab)

pytestmark = pytest.mark.gpu

T, GAMMA, DT = 300.0, 1.0, 0.001


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from openmm_constv_b200 import _capi, integrator
    _capi.load()  # raises if libconstv_b200.so is missing: no fallback
    return integrator


def make_context(gpu, s, T_=T, gamma=GAMMA, dt=DT, seed=77, mirror_mode=0, flags=None, pool=None,
                 **kw):
    from openmm_constv_b200 import _capi
    import torch
    sysm = gpu.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, s.num_cells * s.zmax))
    it = gpu.ConstVLangevinIntegrator(T_, gamma, dt, s.num_cells)
    it.setRandomNumberSeed(seed)
    ctx = gpu.SlabContext(sysm, it, precision=s.precision, mirror_mode=mirror_mode,
                          flags=_capi.FLAGS_DEFAULT if flags is None else flags,
                          random_pool=None if pool is None else torch.as_tensor(pool).cuda(),
                          global_atom_offset=s.global_atom_offset, padded=s.padded, **kw)
    ctx.load_slab(s)
    return ctx, it


def download(ctx, s):
    out = s.copy()
    out.posq = ctx.posq.cpu().numpy()
    out.velm = ctx.velm.cpu().numpy()
    out.force = ctx.force.cpu().numpy()
    if ctx.corr is not None:
        out.corr = ctx.corr.cpu().numpy()
    return out


def assert_slab_bits_equal(a, b):
    assert_bits_equal(a.posq, b.posq, "posq")
    assert_bits_equal(a.velm, b.velm, "velm")
    if a.corr is not None:
        assert_bits_equal(a.corr, b.corr, "posqCorrection")
    assert np.array_equal(a.force, b.force), "force"


# ---- injected noise: everything bit-exact ---------------------------------------------------------

@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells,mirror_mode", [(2, 0), (2, 1), (4, 0)])
@pytest.mark.parametrize("num_real", [1, 33, 4999])
def test_fused_step_bit_exact(gpu, oracle, precision, num_cells, mirror_mode, num_real):
    import torch
    s0 = slabmod.make_slab(num_real, precision, num_cells=num_cells, seed=5)
    nsteps = 4
    pool = np.concatenate([oracle.fill_noise(s0.padded, 21, k) for k in range(nsteps)])
    ctx, it = make_context(gpu, s0, mirror_mode=mirror_mode, pool=pool)
    ref = s0.copy()
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    for k in range(nsteps):
        f = slabmod.new_forces(s0, k)
        ctx.force.copy_(torch.from_numpy(f))
        ref.force[:] = f
        it.step(1)
        oracle_step(oracle, ref, prm, dt, pool, zshift=(0.0 if mirror_mode else s0.zmax),
                    random_index=k * s0.padded)
        assert_slab_bits_equal(download(ctx, s0), ref)
    assert ctx.getStepCount() == nsteps and it._kernel.get_step_count() == nsteps
    assert ctx.getTime() == pytest.approx(nsteps * DT)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("ilp,ctas", [(1, 1), (2, 3), (4, 16)])
def test_launch_shapes_give_identical_results(gpu, oracle, precision, ilp, ctas):
    s0 = slabmod.make_slab(70001, precision, seed=8)
    pool = oracle.fill_noise(s0.padded, 3, 0)
    ctx, it = make_context(gpu, s0, pool=pool)
    it._kernel.set_launch_shape(ilp, ctas)
    it.step(1)
    ref = s0.copy()
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    oracle_step(oracle, ref, prm, dt, pool)
    assert_slab_bits_equal(download(ctx, s0), ref)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_split_path_equals_fused_and_oracle(gpu, oracle, precision):
    """part 1 | constraint hand-off (a no-op callback here) | part 2 + mirror."""
    s0 = slabmod.make_slab(3001, precision, num_cells=4, seed=6)
    pool = oracle.fill_noise(s0.padded, 9, 0)
    seen = {}

    def constraints(ctx, tol):
        seen["tol"] = tol
        seen["delta"] = ctx.pos_delta.cpu().numpy()

    ctx, it = make_context(gpu, s0, pool=pool, apply_constraints=constraints)
    it.step(1)
    ref = s0.copy()
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    delta = oracle_step(oracle, ref, prm, dt, pool)
    assert seen["tol"] == 1e-5
    massive = s0.velm[:, 3] != 0
    assert_bits_equal(seen["delta"][massive], delta[massive], "posDelta")
    assert_slab_bits_equal(download(ctx, s0), ref)
    assert it._kernel.get_step_count() == 1
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("split", [False, True])
def test_reordered_slots(gpu, oracle, precision, split):
    """Behind OpenMM the slots are permuted (atomIndex); invAtomOrder is rebuilt on the device."""
    s0 = slabmod.make_slab(2501, precision, num_cells=2, seed=10)
    perm = np.random.default_rng(3).permutation(s0.num_atoms).astype(np.int32)
    sp, atom_index = permute_slab(s0, perm)
    pool = oracle.fill_noise(s0.padded, 4, 0, atom_index=atom_index)
    ctx, it = make_context(gpu, sp, pool=pool, apply_constraints=(lambda c, t: None) if split else None)
    ctx.set_atom_order(perm)
    it.step(1)
    ref = sp.copy()
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    oracle_step(oracle, ref, prm, dt, pool, inv=oracle.inverse_order(atom_index))
    assert_slab_bits_equal(download(ctx, sp), ref)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("reordered", [False, True])
@pytest.mark.parametrize("side_array", [False, True])
def test_mirror_alone(gpu, oracle, precision, num_cells, reordered, side_array):
    """constv_mirror = the updateImageParticlePositions launch alone (CudaConstVKernels.cpp:165-166), what a host with
    virtual sites calls after computeVirtualSites: someone moves real atoms (here: a scatter of new positions written
    straight into posq, as the virtual-site kernel would), the images follow bit for bit, neutral atoms' images stay
    put, and neither velocities, forces nor the step counter change."""
    import torch
    from openmm_constv_b200 import _capi
    s0 = slabmod.make_slab(3001, precision, num_cells=num_cells, seed=15)
    perm = np.random.default_rng(4).permutation(s0.num_atoms).astype(np.int32) if reordered else None
    sp, atom_index = permute_slab(s0, perm) if reordered else (s0, identity_order(s0.padded))
    flags = _capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side_array else 0)
    ctx, it = make_context(gpu, sp, flags=flags)
    if reordered:
        ctx.set_atom_order(perm)
    it.step(2)
    before = download(ctx, sp)
    # move every third real atom (by ORIGINAL index) somewhere else
    R = s0.num_real
    inv = oracle.inverse_order(atom_index)
    movers = inv[np.arange(0, R, 3)]
    rng = np.random.default_rng(8)
    new_xyz = (rng.random((len(movers), 3)) * [sp.box[0], sp.box[1], sp.zmax]).astype(before.posq.dtype)
    moved = before.copy()
    moved.posq[movers, :3] = new_xyz
    if moved.corr is not None:
        moved.corr[movers, :3] = (rng.random((len(movers), 3)) * 1e-8).astype(moved.corr.dtype)
        ctx.corr.copy_(torch.from_numpy(moved.corr))
    ctx.posq.copy_(torch.from_numpy(moved.posq))
    _capi.check(_capi.load().constv_mirror(it._kernel.handle, ctx.stream_ptr()))
    got = download(ctx, sp)
    oracle.mirror(precision, R, num_cells, sp.zmax, moved.posq, moved.corr, inv)
    assert_bits_equal(got.posq, moved.posq, "posq after the image rewrite")
    if moved.corr is not None:
        assert_bits_equal(got.corr, moved.corr, "posqCorrection after the image rewrite")
    assert_bits_equal(got.velm, before.velm, "velm untouched")
    assert np.array_equal(got.force, before.force)
    assert it._kernel.get_step_count() == 2
    assert not np.array_equal(got.posq, before.posq)
    ctx.close()


def test_trajectory_1000_steps_friction_zero(gpu, oracle):
    """north_star: reference-matching 1000-step trajectory with friction 0 (fp32).  Static injected
    forces; every 100 steps all arrays must still be bit-identical to the oracle's."""
    s0 = slabmod.make_slab(2000, "single", seed=12, force_sigma=50.0, image_force_garbage=False)
    from openmm_constv_b200 import _capi
    flags = _capi.FLAG_ZERO_IMAGE_VELM | _capi.FLAG_PDL      # keep the force buffer constant
    ctx, it = make_context(gpu, s0, T_=300.0, gamma=0.0, dt=0.001, flags=flags)
    ref = s0.copy()
    prm, dt = oracle_params_for(oracle, "single", 300.0, 0.0, 0.001)
    assert prm[2] == 0.0
    zero_noise = np.zeros((s0.padded, 4), np.float32)
    inv = identity_order(s0.padded)
    delta = np.zeros((s0.padded, 4), np.float32)
    for k in range(1000):
        it.step(1)
        oracle.step("single", ref.num_atoms, 2, ref.zmax, ref.posq, None, ref.velm,
                    ref.force.reshape(-1), delta, prm, dt, zero_noise, 0, inv, True, False)
        if (k + 1) % 100 == 0:
            got = download(ctx, s0)
            assert_slab_bits_equal(got, ref)
            rel = np.abs(got.posq[:, :3] - ref.posq[:, :3]).max() / np.abs(ref.posq[:, :3]).max()
            assert rel <= 1e-6
    ctx.close()


# ---- Philox stream ---------------------------------------------------------------------------------

def test_philox_gaussians_match_oracle(gpu, oracle):
    import ctypes as C
    import torch
    from openmm_constv_b200 import _capi
    s0 = slabmod.make_slab(40000, "single", seed=1)
    ctx, it = make_context(gpu, s0, seed=0xDEADBEEFCAFE)
    out = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    for step in (0, 1, 2**33 + 5):
        _capi.check(_capi.load().constv_generate_noise(it._kernel.handle, step, out.data_ptr(), ctx.stream_ptr()))
        got = out.cpu().numpy()[: s0.num_atoms]
        want = oracle.fill_noise(s0.num_atoms, 0xDEADBEEFCAFE, step)
        assert np.max(np.abs(got - want)) < 4e-6
    assert abs(got.mean()) < 0.02 and abs(got.std() - 1) < 0.02
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_philox_step_consumes_exactly_the_published_stream(gpu, oracle, precision):
    """A Philox-driven fused step equals the oracle fed with the noise constv_generate_noise
    publishes for the same (seed, step) -- bit for bit -- and advances the device step counter;
    the global atom offset shifts the stream (multi-GPU range partition)."""
    import torch
    from openmm_constv_b200 import _capi
    full = slabmod.make_slab(6000, precision, seed=14)
    shard = slabmod.slice_slab(full, 2000, 4500)
    ctx, it = make_context(gpu, shard, seed=99)
    ref = shard.copy()
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    noise = torch.zeros((shard.padded, 4), dtype=torch.float32, device="cuda")
    for k in range(3):
        _capi.check(_capi.load().constv_generate_noise(it._kernel.handle, k, noise.data_ptr(), ctx.stream_ptr()))
        xi = noise.cpu().numpy()
        want = oracle.fill_noise(shard.num_atoms, 99, k, global_atom_offset=2000)
        assert np.max(np.abs(xi[: shard.num_atoms] - want)) < 4e-6
        it.step(1)
        oracle_step(oracle, ref, prm, dt, xi)
        assert_slab_bits_equal(download(ctx, shard), ref)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_real,num_cells", [(1, 2), (3, 2), (6, 2), (1027, 2), (50001, 2), (20002, 4)])
@pytest.mark.parametrize("block,stages", [(256, 4), (512, 2), (128, 6)])
def test_tma_pipelined_kernel_bit_exact(gpu, oracle, precision, num_real, num_cells, block, stages):
    """The TMA-ring kernel (cp.async.bulk + mbarrier pipeline) against the oracle, including slabs
    smaller than one quad, ragged ends (R % 4 != 0) and ranges shorter than a chunk."""
    import torch
    from openmm_constv_b200 import _capi
    s0 = slabmod.make_slab(num_real, precision, num_cells=num_cells, seed=40)
    ctx, it = make_context(gpu, s0, seed=7, flags=_capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE)
    it._kernel.set_kernel_variant(2, block, stages, 0)
    ref = s0.copy()
    prm, dt = oracle_params_for(oracle, precision, T, GAMMA, DT)
    noise = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    for k in range(3):
        _capi.check(_capi.load().constv_generate_noise(it._kernel.handle, k, noise.data_ptr(), ctx.stream_ptr()))
        f = slabmod.new_forces(s0, k)
        ctx.force.copy_(torch.from_numpy(f))
        ref.force[:] = f
        it.step(1)
        oracle_step(oracle, ref, prm, dt, noise.cpu().numpy())
        assert_slab_bits_equal(download(ctx, s0), ref)
    assert it._kernel.get_step_count() == 3
    ctx.close()


def test_tma_kernel_refuses_ineligible_slabs(gpu):
    from openmm_constv_b200 import _capi
    s0 = slabmod.make_slab(1000, "single", seed=41)
    ctx, it = make_context(gpu, s0)      # no image-charge side array
    it._kernel.set_kernel_variant(2, 256, 4, 0)
    with pytest.raises(_capi.ConstVError, match="TMA kernel needs"):
        it.step(1)
    ctx.close()


def test_seed_determinism(gpu):
    """TestCudaConstVLangevinIntegrator.cpp:212-268: same seed -> same trajectory, different seed
    -> different; seed 0 -> a fresh seed per context."""
    s0 = slabmod.make_slab(1000, "single", seed=2)

    def run(seed):
        ctx, it = make_context(gpu, s0, seed=seed)
        it.step(10)
        p = ctx.posq.cpu().numpy()
        ctx.close()
        return p

    a, b, c = run(5), run(5), run(10)
    assert_bits_equal(a, b)
    assert not np.array_equal(a, c)
    assert not np.array_equal(run(0), run(0))


def test_cuda_graph_replay_advances_the_stream(gpu):
    """The step counter lives on the device, so replaying a captured graph keeps advancing the
    Philox stream: 2 replays of a 4-step graph == 8 eager steps."""
    import torch
    s0 = slabmod.make_slab(5000, "single", seed=3)
    ctx_a, it_a = make_context(gpu, s0, seed=42)
    it_a.step(8)
    want = download(ctx_a, s0)
    ctx_a.close()
    ctx_b, it_b = make_context(gpu, s0, seed=42)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        it_b.step(0)
        g = torch.cuda.CUDAGraph()
        # parameters must be set before capture; warm nothing else
        from openmm_constv_b200 import _capi
        _capi.check(_capi.load().constv_set_parameters(it_b._kernel.handle, T, GAMMA, DT))
        with torch.cuda.graph(g, stream=stream):
            for _ in range(4):
                _capi.check(_capi.load().constv_step_fused(it_b._kernel.handle, None, 0, ctx_b.stream_ptr()))
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    assert it_b._kernel.get_step_count() == 8
    assert_slab_bits_equal(download(ctx_b, s0), want)
    ctx_b.close()


# ---- chains: independent real-atom ranges with their own step counters --------------------------------

@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("chains", [2, 5, 16])
def test_chains_are_bit_identical_to_the_single_launch(gpu, precision, chains):
    """constv_set_chains cuts the slab into contiguous ranges (one step counter each); whether the ranges
    are launched in turn on one stream or each on a stream of its own, every array equals the
    unchained step's bit for bit (the Philox key is (seed; atom, step), not the launch)."""
    import ctypes as C
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(70001, precision, seed=12)
    nsteps = 5
    ctx_a, it_a = make_context(gpu, s0, seed=9)
    it_a.step(nsteps)
    want = download(ctx_a, s0)
    ctx_a.close()

    # every chain in turn on the caller's stream
    ctx_b, it_b = make_context(gpu, s0, seed=9)
    _capi.check(lib.constv_set_chains(it_b._kernel.handle, chains))
    got = C.c_int32(0)
    _capi.check(lib.constv_get_chains(it_b._kernel.handle, C.byref(got)))
    assert got.value == chains
    launches = _capi.launch_count()
    it_b.step(nsteps)
    assert _capi.launch_count() - launches == nsteps * chains
    assert it_b._kernel.get_step_count() == nsteps
    assert_slab_bits_equal(download(ctx_b, s0), want)
    ctx_b.close()

    # one stream per chain, all steps of a chain back to back: chains run ahead of each other
    ctx_c, it_c = make_context(gpu, s0, seed=9)
    h = it_c._kernel.handle
    _capi.check(lib.constv_set_chains(h, chains))
    _capi.check(lib.constv_set_parameters(h, T, GAMMA, DT))
    streams = [torch.cuda.Stream() for _ in range(chains)]
    torch.cuda.synchronize()
    for q in reversed(range(chains)):
        for _ in range(nsteps):
            _capi.check(lib.constv_step_fused_chain(h, q, None, 0, streams[q].cuda_stream))
    torch.cuda.synchronize()
    assert it_c._kernel.get_step_count() == nsteps
    assert_slab_bits_equal(download(ctx_c, s0), want)
    t = C.c_double(0.0)
    _capi.check(lib.constv_get_time(h, C.byref(t)))
    assert t.value == pytest.approx(nsteps * DT)
    ctx_c.close()


def test_chains_with_the_image_charge_side_array(gpu):
    """CONSTV_FLAG_CACHE_IMAGE_CHARGE: the snapshot of the images' charges is taken lazily by the first step.
    When that first step is a chain on a stream of its own, no other chain may rewrite posq[image] before the
    snapshot is complete (a race would freeze garbage charges into the side array)."""
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(200003, "single", seed=15)
    flags = _capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE
    ctx_a, it_a = make_context(gpu, s0, seed=31, flags=flags)
    it_a.step(3)
    want = download(ctx_a, s0)
    ctx_a.close()
    for trial in range(3):
        ctx, it = make_context(gpu, s0, seed=31, flags=flags)
        h = it._kernel.handle
        _capi.check(lib.constv_set_chains(h, 8))
        _capi.check(lib.constv_set_parameters(h, T, GAMMA, DT))
        streams = [torch.cuda.Stream() for _ in range(8)]
        torch.cuda.synchronize()
        for k in range(3):
            for q in reversed(range(8)):
                _capi.check(lib.constv_step_fused_chain(h, q, None, 0, streams[q].cuda_stream))
        torch.cuda.synchronize()
        assert_slab_bits_equal(download(ctx, s0), want)
        ctx.close()


@pytest.mark.parametrize("chains", [1, 3])
def test_step_fused_multi(gpu, chains):
    """constv_step_fused_multi(n) == n x constv_step_fused, eagerly and replayed from a CUDA graph; with chains
    the library forks onto its own streams and joins back into the caller's."""
    import ctypes as C
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(40001, "single", seed=14)
    ctx_a, it_a = make_context(gpu, s0, seed=21)
    it_a.step(12)
    want = download(ctx_a, s0)
    ctx_a.close()
    ctx_b, it_b = make_context(gpu, s0, seed=21)
    h = it_b._kernel.handle
    _capi.check(lib.constv_set_chains(h, chains))
    _capi.check(lib.constv_set_parameters(h, T, GAMMA, DT))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        _capi.check(lib.constv_step_fused_multi(h, 4, stream.cuda_stream))   # eager (creates the side streams)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            _capi.check(lib.constv_step_fused_multi(h, 4, stream.cuda_stream))
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    assert it_b._kernel.get_step_count() == 12
    assert_slab_bits_equal(download(ctx_b, s0), want)
    t = C.c_double(0.0)
    _capi.check(lib.constv_get_time(h, C.byref(t)))
    assert t.value == pytest.approx(8 * DT)  # host-side time advances when a call is issued, not when a graph replays
    assert lib.constv_step_fused_multi(h, -1, None) == _capi.ERR_INVALID
    ctx_b.close()


def test_chains_guard_rails(gpu):
    import ctypes as C
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(1000, "single", seed=13)  # 4 tiles of 256 pairs
    ctx, it = make_context(gpu, s0, seed=9, apply_constraints=None)
    h = it._kernel.handle
    assert lib.constv_set_chains(h, 0) == _capi.ERR_INVALID
    assert lib.constv_set_chains(h, 17) == _capi.ERR_INVALID
    _capi.check(lib.constv_set_chains(h, 16))  # clamped to the number of tiles
    got = C.c_int32(0)
    _capi.check(lib.constv_get_chains(h, C.byref(got)))
    assert got.value == 4
    _capi.check(lib.constv_set_parameters(h, T, GAMMA, DT))
    assert lib.constv_step_fused_chain(h, 4, None, 0, None) == _capi.ERR_INVALID
    _capi.check(lib.constv_step_fused_chain(h, 1, None, 0, None))
    count = C.c_uint64(0)
    assert lib.constv_get_step_count(h, C.byref(count)) == _capi.ERR_STATE  # chain 1 is ahead
    assert b"out of step" in lib.constv_last_error()
    assert lib.constv_set_chains(h, 1) == _capi.ERR_STATE
    for q in (0, 2, 3):
        _capi.check(lib.constv_step_fused_chain(h, q, None, 0, None))
    _capi.check(lib.constv_get_step_count(h, C.byref(count)))
    assert count.value == 1
    assert lib.constv_step_part1(h, None, 0, None) == _capi.ERR_STATE  # the split path is unchained
    _capi.check(lib.constv_set_chains(h, 1))
    _capi.check(lib.constv_get_step_count(h, C.byref(count)))
    assert count.value == 1  # the counter survives re-chaining
    it.step(1)
    assert it._kernel.get_step_count() == 2
    ctx.close()


# ---- kinetic energy ----------------------------------------------------------------------------------

@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_real", [7, 100003])
def test_kinetic_energy(gpu, oracle, precision, num_real):
    s0 = slabmod.make_slab(num_real, precision, seed=15)
    ctx, it = make_context(gpu, s0)
    want = oracle.kinetic_energy(precision, s0.velm, s0.force.reshape(-1), 0.5 * DT, s0.num_atoms)
    got = it.computeKineticEnergy()
    assert got == pytest.approx(want, rel=1e-12)
    assert ctx.getKineticEnergy() == got     # deterministic reduction order
    # equipartition sanity on the Maxwell-Boltzmann slab: without the half-step force shift the
    # kinetic energy is 1.5 kT per massive atom
    import ctypes as C
    from openmm_constv_b200 import _capi
    ke0 = C.c_double(0)
    _capi.check(_capi.load().constv_kinetic_energy(it._kernel.handle, 0.0, C.byref(ke0), ctx.stream_ptr()))
    assert ke0.value == pytest.approx(oracle.kinetic_energy(precision, s0.velm, s0.force.reshape(-1), 0.0, s0.num_atoms), rel=1e-12)
    dof = 3 * int(np.sum(s0.velm[: s0.num_real, 3] != 0))
    if num_real > 1000:
        assert ke0.value == pytest.approx(0.5 * dof * slabmod.BOLTZ * 300.0, rel=0.02)
    ctx.close()


# ---- ensemble launch -----------------------------------------------------------------------------------

@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_batched_ensemble_equals_individual_steps(gpu, oracle, precision):
    import ctypes as C
    from openmm_constv_b200 import _capi
    sizes = [1000, 777, 2048, 1, 5000]
    slabs = [slabmod.make_slab(r, precision, seed=30 + i) for i, r in enumerate(sizes)]
    temps = [250.0 + 25 * i for i in range(len(sizes))]
    singles, batch = [], []
    for s, t in zip(slabs, temps):
        singles.append(make_context(gpu, s, T_=t, seed=1234))
        batch.append(make_context(gpu, s, T_=t, seed=1234))
    lib = _capi.load()
    handles = (C.c_void_p * len(batch))(*[it._kernel.handle for _, it in batch])
    for _, it in batch:
        _capi.check(lib.constv_set_parameters(it._kernel.handle, it.getTemperature(), GAMMA, DT))
    for k in range(3):
        for ctx, it in singles:
            it.step(1)
        _capi.check(lib.constv_step_fused_batched(handles, len(batch), batch[0][0].stream_ptr()))
    for (ca, ia), (cb, ib), s in zip(singles, batch, slabs):
        assert_slab_bits_equal(download(cb, s), download(ca, s))
        assert ib._kernel.get_step_count() == 3
        ca.close()
        cb.close()


# ---- host-buffer entry point ---------------------------------------------------------------------------

def test_step_host_roundtrip(gpu, oracle):
    import ctypes as C
    import torch
    from openmm_constv_b200 import _capi
    s0 = slabmod.make_slab(3000, "single", seed=17)
    ctx, it = make_context(gpu, s0, seed=5)
    lib = _capi.load()
    _capi.check(lib.constv_set_parameters(it._kernel.handle, T, GAMMA, DT))
    f = slabmod.new_forces(s0, 3)
    host_force = torch.from_numpy(np.ascontiguousarray(f[:, : s0.num_real])).pin_memory()
    host_posq = torch.zeros((s0.num_atoms, 4), dtype=torch.float32).pin_memory()
    host_velm = torch.zeros((s0.num_atoms, 4), dtype=torch.float32).pin_memory()
    noise = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    _capi.check(lib.constv_generate_noise(it._kernel.handle, 0, noise.data_ptr(), ctx.stream_ptr()))
    ke = C.c_double(0.0)
    _capi.check(lib.constv_step_host(it._kernel.handle, host_force.data_ptr(), host_posq.data_ptr(),
                                     host_velm.data_ptr(), C.byref(ke), ctx.stream_ptr()))
    ref = s0.copy()
    ref.force[:, : s0.num_real] = f[:, : s0.num_real]
    prm, dt = oracle_params_for(oracle, "single", T, GAMMA, DT)
    oracle_step(oracle, ref, prm, dt, noise.cpu().numpy())
    assert_bits_equal(host_posq.numpy(), ref.posq[: s0.num_atoms])
    assert_bits_equal(host_velm.numpy(), ref.velm[: s0.num_atoms])
    want_ke = oracle.kinetic_energy("single", ref.velm, ref.force.reshape(-1), 0.5 * DT, s0.num_atoms)
    assert ke.value == pytest.approx(want_ke, rel=1e-12)
    ctx.close()


@pytest.mark.parametrize("fmt", ["fixed64", "float32"])
def test_pipelined_host_stepping_equals_synchronous(gpu, oracle, fmt):
    """constv_step_host_submit/_collect (upload of step k+1 overlapping step k) must give exactly what
    the synchronous constv_step_host gives; float32 host forces are converted on the device as
    OpenMM's force kernels fill the buffer: (long long)(f * 2^32), truncating."""
    import ctypes as C
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(5003, "single", seed=23)
    R, nsteps = s0.num_real, 7
    rng = np.random.default_rng(4)
    f32 = [(rng.standard_normal((3, R)) * 1000.0).astype(np.float32) for _ in range(nsteps)]
    fixed = [np.trunc(f.astype(np.float64) * 4294967296.0).astype(np.int64) for f in f32]  # float*2^32 is exact
    assert all(np.array_equal(np.trunc(f * np.float32(4294967296.0)).astype(np.int64), g) for f, g in zip(f32, fixed))
    # synchronous path with int64 forces
    ctx_a, it_a = make_context(gpu, s0, seed=5)
    _capi.check(lib.constv_set_parameters(it_a._kernel.handle, T, GAMMA, DT))
    ke_a = []
    for g in fixed:
        hf = torch.from_numpy(np.ascontiguousarray(g)).pin_memory()
        ke = C.c_double(0.0)
        _capi.check(lib.constv_step_host(it_a._kernel.handle, hf.data_ptr(), None, None, C.byref(ke), ctx_a.stream_ptr()))
        ke_a.append(ke.value)
    # pipelined path, two tickets in flight
    ctx_b, it_b = make_context(gpu, s0, seed=5)
    _capi.check(lib.constv_set_parameters(it_b._kernel.handle, T, GAMMA, DT))
    src = fixed if fmt == "fixed64" else f32
    code = _capi.FORCE_FIXED64 if fmt == "fixed64" else _capi.FORCE_FLOAT32
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in src]
    ke_b, tickets = [], []
    for k in range(nsteps):
        t = C.c_uint64(0)
        _capi.check(lib.constv_step_host_submit(it_b._kernel.handle, pinned[k].data_ptr(), code, C.byref(t), ctx_b.stream_ptr()))
        tickets.append(t.value)
        if k >= 1:
            ke = C.c_double(0.0)
            _capi.check(lib.constv_step_host_collect(it_b._kernel.handle, tickets[k - 1], C.byref(ke)))
            ke_b.append(ke.value)
    ke = C.c_double(0.0)
    _capi.check(lib.constv_step_host_collect(it_b._kernel.handle, tickets[-1], C.byref(ke)))
    ke_b.append(ke.value)
    assert tickets == list(range(nsteps))
    assert ke_a == ke_b
    torch.cuda.synchronize()
    assert_slab_bits_equal(download(ctx_b, s0), download(ctx_a, s0))
    # discipline: a third outstanding ticket and out-of-order collection are refused
    t = C.c_uint64(0)
    for _ in range(2):
        _capi.check(lib.constv_step_host_submit(it_b._kernel.handle, pinned[0].data_ptr(), code, C.byref(t), ctx_b.stream_ptr()))
    assert lib.constv_step_host_submit(it_b._kernel.handle, pinned[0].data_ptr(), code, C.byref(t), ctx_b.stream_ptr()) == _capi.ERR_STATE
    assert lib.constv_step_host_collect(it_b._kernel.handle, t.value, None) == _capi.ERR_STATE
    _capi.check(lib.constv_step_host_collect(it_b._kernel.handle, t.value - 1, None))
    _capi.check(lib.constv_step_host_collect(it_b._kernel.handle, t.value, None))
    ctx_a.close()
    ctx_b.close()


@pytest.mark.parametrize("which", ["posq", "posq_velm", "velm"])
def test_pipelined_state_readback_equals_synchronous(gpu, oracle, which):
    """constv_step_host_submit_state: every ticket's posq / velm, read back on the second copy stream while the
    next ticket's forces upload, are the arrays the synchronous constv_step_host returns for that step -- i.e.
    the next step's kernels really wait for the read-back of the arrays they overwrite."""
    import ctypes as C
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(200003, "single", seed=29)  # large enough that a read-back outlasts the next step's kernels
    R, N, nsteps = s0.num_real, s0.num_atoms, 6
    rng = np.random.default_rng(9)
    f32 = [(rng.standard_normal((3, R)) * 1000.0).astype(np.float32) for _ in range(nsteps)]
    fixed = [np.trunc(f.astype(np.float64) * 4294967296.0).astype(np.int64) for f in f32]
    ctx_a, it_a = make_context(gpu, s0, seed=5)
    _capi.check(lib.constv_set_parameters(it_a._kernel.handle, T, GAMMA, DT))
    want = []
    for g in fixed:
        hf = torch.from_numpy(np.ascontiguousarray(g)).pin_memory()
        hp = torch.zeros((N, 4), dtype=torch.float32).pin_memory()
        hv = torch.zeros((N, 4), dtype=torch.float32).pin_memory()
        ke = C.c_double(0.0)
        _capi.check(lib.constv_step_host(it_a._kernel.handle, hf.data_ptr(), hp.data_ptr(), hv.data_ptr(), C.byref(ke), ctx_a.stream_ptr()))
        want.append((hp.numpy().copy(), hv.numpy().copy(), ke.value))
    ctx_b, it_b = make_context(gpu, s0, seed=5)
    _capi.check(lib.constv_set_parameters(it_b._kernel.handle, T, GAMMA, DT))
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in f32]
    out_p = [torch.zeros((N, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    out_v = [torch.zeros((N, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    h = it_b._kernel.handle

    def check(k, ke):
        if "posq" in which:
            assert_bits_equal(out_p[k % 2].numpy(), want[k][0])
        if "velm" in which:
            assert_bits_equal(out_v[k % 2].numpy(), want[k][1])
        assert ke == want[k][2]

    prev = None
    for k in range(nsteps):
        t = C.c_uint64(0)
        _capi.check(lib.constv_step_host_submit_state(h, pinned[k].data_ptr(), _capi.FORCE_FLOAT32,
                                                      out_p[k % 2].data_ptr() if "posq" in which else None,
                                                      out_v[k % 2].data_ptr() if "velm" in which else None,
                                                      C.byref(t), ctx_b.stream_ptr()))
        if prev is not None:
            ke = C.c_double(0.0)
            _capi.check(lib.constv_step_host_collect(h, prev, C.byref(ke)))
            check(k - 1, ke.value)
        prev = t.value
    ke = C.c_double(0.0)
    _capi.check(lib.constv_step_host_collect(h, prev, C.byref(ke)))
    check(nsteps - 1, ke.value)
    # a plain ticket behind a state ticket still waits for the read-back before it rewrites the arrays
    t = C.c_uint64(0)
    _capi.check(lib.constv_step_host_submit_state(h, pinned[0].data_ptr(), _capi.FORCE_FLOAT32, out_p[0].data_ptr(), None, C.byref(t), ctx_b.stream_ptr()))
    t2 = C.c_uint64(0)
    _capi.check(lib.constv_step_host_submit(h, pinned[1].data_ptr(), _capi.FORCE_FLOAT32, C.byref(t2), ctx_b.stream_ptr()))
    _capi.check(lib.constv_step_host_collect(h, t.value, None))
    snapshot = out_p[0].numpy().copy()
    _capi.check(lib.constv_step_host_collect(h, t2.value, None))
    hf = torch.from_numpy(np.ascontiguousarray(fixed[0])).pin_memory()
    hp = torch.zeros((N, 4), dtype=torch.float32).pin_memory()
    _capi.check(lib.constv_step_host(it_a._kernel.handle, hf.data_ptr(), hp.data_ptr(), None, None, ctx_a.stream_ptr()))
    assert_bits_equal(snapshot, hp.numpy())
    ctx_a.close()
    ctx_b.close()


# ---- full-size, size-independent properties ------------------------------------------------------------

def test_full_size_slab_properties_and_oracle(gpu, oracle):
    """BASELINE config: 1M-atom slab (500k real + 500k image).  Checked against the oracle (the C
    oracle handles 1M atoms in milliseconds) and through properties evaluated on the device:
    image z == float(-double(z) + 2*zmax) bit for bit, x/y/charge carried, image velm and force
    zero, massless atoms untouched, neutral atoms' images untouched."""
    import torch
    s0 = slabmod.make_slab(500_000, "single", seed=12345)
    R = s0.num_real
    ctx, it = make_context(gpu, s0, seed=2024)
    before_posq = ctx.posq.clone()
    before_velm = ctx.velm.clone()
    noise = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    from openmm_constv_b200 import _capi
    _capi.check(_capi.load().constv_generate_noise(it._kernel.handle, 0, noise.data_ptr(), ctx.stream_ptr()))
    it.step(1)
    posq, velm, force = ctx.posq, ctx.velm, ctx.force
    charged = before_posq[:R, 3] != 0
    massless = before_velm[:R, 3] == 0
    expect_z = (-posq[:R, 2].double() + 2.0 * s0.zmax).float()
    assert torch.equal(posq[R:2 * R, 2][charged].view(torch.int32), expect_z[charged].view(torch.int32))
    assert torch.equal(posq[R:2 * R, :2][charged], posq[:R, :2][charged])
    assert torch.equal(posq[R:2 * R, 3], before_posq[R:2 * R, 3])
    assert torch.equal(posq[R:2 * R][~charged], before_posq[R:2 * R][~charged])
    assert not velm[R:2 * R].any() and not force[:, R:2 * R].any()
    assert torch.equal(posq[:R][massless], before_posq[:R][massless])
    assert torch.equal(velm[:R][massless], before_velm[:R][massless])
    assert not torch.equal(posq[:R][~massless], before_posq[:R][~massless])
    ref = s0.copy()
    prm, dt = oracle_params_for(oracle, "single", T, GAMMA, DT)
    oracle_step(oracle, ref, prm, dt, noise.cpu().numpy())
    assert_slab_bits_equal(download(ctx, s0), ref)
    ctx.close()


# ---- error behaviour and the reference tests' physics -----------------------------------------------------

def test_error_messages_through_the_plugin_interface(gpu):
    sysm = gpu.SlabSystem()
    sysm.addParticles([1.0, 1.0, 0.0, 0.0])
    sysm.setDefaultPeriodicBoxVectors((3, 0, 0), (0, 3, 0), (0, 0, 8.0))
    for args, masses, msg in [
        ((300, 1, 0.001, 3), None, "Not even number of cells"),
        ((300, 1, 0.001, 4), [1.0, 0, 0, 0, 0, 0], "Number of particles is not multiple of number of cells"),
        ((300, 1, 0.001, 2), [1.0, 1.0, 0.0, 2.0], "Image Particle has nonzero mass"),
        ((300, 1, 0.001, 2, 3.9), None, "Unit cell dimension does not match with the provided zmax value"),
    ]:
        sy = sysm
        if masses is not None:
            sy = gpu.SlabSystem()
            sy.addParticles(masses)
            sy.setDefaultPeriodicBoxVectors((3, 0, 0), (0, 3, 0), (0, 0, 8.0))
        with pytest.raises(gpu.OpenMMException, match=msg):
            gpu.SlabContext(sy, gpu.ConstVLangevinIntegrator(*args))
    it = gpu.ConstVLangevinIntegrator(300, 1, 0.001)
    c1 = gpu.SlabContext(sysm, it)
    with pytest.raises(gpu.OpenMMException, match="already bound to a context"):
        gpu.SlabContext(sysm, it)
    c1.close()


def test_single_bond_damped_oscillator_on_gpu(gpu):
    """TestCudaConstVLangevinIntegrator.cpp:52-94 with the bond force evaluated on the device
    by a torch expression standing in for context.calcForcesAndEnergy."""
    import torch
    sysm = gpu.SlabSystem()
    sysm.addParticles([2.0, 2.0, 0.0, 0.0])
    sysm.setDefaultPeriodicBoxVectors((10, 0, 0), (0, 10, 0), (0, 0, 100.0))

    def bond(ctx):
        p = ctx.posq[:2, :3].double()
        d = p[1] - p[0]
        r = d.norm()
        g = 1.0 * (r - 1.5) * d / r
        f = torch.stack([g, -g])                      # (2,3)
        ctx.force[:, :2] = torch.round(f.T * 4294967296.0).to(torch.int64)

    for precision in ("single", "mixed", "double"):
        it = gpu.ConstVLangevinIntegrator(0, 0.1, 0.01)
        ctx = gpu.SlabContext(sysm, it, precision=precision, force_fn=bond)
        ctx.setPositions([[-1, 0, 10], [1, 0, 10], [-1, 0, -10], [1, 0, -10]], charges=[0, 0, 0, 0])
        freq = np.sqrt(1 - 0.05 * 0.05)
        for _ in range(1000):
            t = ctx.getTime()
            dist = 1.5 + 0.5 * np.exp(-0.05 * t) * np.cos(freq * t)
            if _ % 50 == 0:
                p = ctx.getPositions()
                assert abs(p[0, 0] + 0.5 * dist) < 0.02 and abs(p[1, 0] - 0.5 * dist) < 0.02
            it.step(1)
        # friction 0: energy conservation within 1 %
        it.setFriction(0.0)
        ctx.setPositions([[-1, 0, 10], [1, 0, 10]])
        ctx.setVelocities([[0, 0, 0], [0, 0, 0]])

        def energy():
            bond(ctx)
            p = ctx.getPositions()
            return it.computeKineticEnergy() + 0.5 * (np.linalg.norm(p[1] - p[0]) - 1.5) ** 2

        e0 = energy()
        for _ in range(200):
            it.step(5)
            assert energy() == pytest.approx(e0, rel=0.01)
        ctx.close()


def test_massless_particles_keep_zero_velocity(gpu):
    """TestCudaConstVLangevinIntegrator.cpp:196-209."""
    sysm = gpu.SlabSystem()
    sysm.addParticles([0.0, 1.0, 0.0, 0.0])
    sysm.setDefaultPeriodicBoxVectors((10, 0, 0), (0, 10, 0), (0, 0, 20.0))
    it = gpu.ConstVLangevinIntegrator(300.0, 2.0, 0.01)
    ctx = gpu.SlabContext(sysm, it)
    ctx.setPositions([[-1, 0, 3], [1, 0, 3], [-1, 0, -3], [1, 0, -3]], charges=[0.5, -0.5, -0.5, 0.5])
    ctx.setForces([[100, 100, 100]] * 4)
    p0 = ctx.getPositions()
    it.step(5)
    v = ctx.getVelocities()
    assert v[0, 0] == 0.0 and not v[0].any() and v[1].any()
    assert np.array_equal(ctx.getPositions()[0], p0[0])
    ctx.close()
