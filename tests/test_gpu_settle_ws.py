"""The rigid-water step as a warp-specialised TMA pipeline (csrc/constv_settle_ws_kernels.cuh: bulk asynchronous
copies in and out of a shared-memory ring, producer / consumer groups / storer): same bits as the CPU oracle and
as the staged kernel (kernel variants 0 and 3; the pipeline is variant 4 and is not the default: it is slower), for

  * every precision, with the image-charge side array (without it the library falls back to the LDG-staged kernel),
  * runs that start on every alignment of the 4-slot load window and the 2-element force-plane window
    (1..9 ions before the water, odd and even numbers of real atoms), tiny slabs, several tiles per CTA,
  * tiles that are NOT uniformly massive / charged (massless or neutral atoms inside the ion run): the per-slot
    store path of the storer warp,
  * flags that switch the image zeroing off, chains, CUDA-graph replay.

Every test feeds the oracle the Gaussians the library publishes for (seed, step), so the comparison is bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H

pytestmark = pytest.mark.gpu

T, GAMMA, DT = 300.0, 1.0, 0.002


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from openmm_constv_b200 import _capi, integrator
    _capi.load()
    return integrator


WS = 4  # constv_set_kernel_variant: the TMA-pipelined rigid-water kernel (0 = auto = the staged kernel; 3 = the same, staged through registers)
WS_PIPELINE = 4
BULK = 6  # the staged kernel's shape with the pipeline's bulk stage-in (constv_settle_step_bulk_kernel)


@pytest.fixture(params=[WS, BULK], ids=["pipeline", "bulk_stage_in"], autouse=True)
def kernel_variant(request):
    """Every test of this module runs once per kernel: both read the same 4-slot-aligned load windows."""
    global WS
    before, WS = WS, request.param
    yield request.param
    WS = before


def make_context(gpu, s, cons, seed=77, flags=None, variant=None):
    from openmm_constv_b200 import _capi
    variant = WS if variant is None else variant
    sysm = gpu.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, s.num_cells * s.zmax))
    sysm.addConstraints(*cons)
    it = gpu.ConstVLangevinIntegrator(T, GAMMA, DT, s.num_cells)
    it.setRandomNumberSeed(seed)
    ctx = gpu.SlabContext(sysm, it, precision=s.precision, flags=_capi.FLAGS_DEFAULT if flags is None else flags,
                          global_atom_offset=s.global_atom_offset, padded=s.padded)
    ctx.load_slab(s)
    it._kernel.set_kernel_variant(variant)
    assert it._kernel.fused_constraints
    return ctx, it


def download(ctx, s):
    out = s.copy()
    out.posq, out.velm, out.force = ctx.posq.cpu().numpy(), ctx.velm.cpu().numpy(), ctx.force.cpu().numpy()
    if ctx.corr is not None:
        out.corr = ctx.corr.cpu().numpy()
    return out


def assert_same(a, b):
    H.assert_bits_equal(a.posq, b.posq, "posq")
    H.assert_bits_equal(a.velm, b.velm, "velm")
    if a.corr is not None:
        H.assert_bits_equal(a.corr, b.corr, "posqCorrection")
    assert np.array_equal(a.force, b.force), "force"


def run_against_oracle(gpu, oracle, s0, atoms, params, cons, nsteps=3, flags=None, seed=77, zero=True):
    """nsteps Philox steps through the WS kernel; the oracle is driven with the published Gaussians."""
    import torch
    from openmm_constv_b200 import _capi
    ctx, it = make_context(gpu, s0, cons, seed=seed, flags=flags)
    lib, slab = it._kernel._lib, it._kernel.handle
    g = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    ref = s0.copy()
    prm, dt = H.oracle_params_for(oracle, s0.precision, T, GAMMA, DT)
    delta = np.zeros((s0.padded, 4), H.MIXED[s0.precision])
    inv = H.identity_order(s0.padded)
    launches = _capi.launch_count()
    for k in range(nsteps):
        _capi.check(lib.constv_generate_noise(slab, k, g.data_ptr(), ctx.stream_ptr()))
        xi = g.cpu().numpy()
        f = slabmod.new_forces(s0, k)
        ctx.force.copy_(torch.from_numpy(f))
        ref.force[:] = f
        it.step(1)
        oracle.settle_step(s0.precision, ref.num_atoms, s0.num_cells, ref.zmax, ref.posq, ref.corr, ref.velm, ref.force.reshape(-1),
                           delta, prm, dt, xi, 0, inv, atoms, params, *(() if zero else (False, False)))
        assert_same(download(ctx, s0), ref)
    assert it._kernel.get_step_count() == nsteps
    # which kernel ran: the variant under test needs the image-charge side array and padded % 4 == 0, else the staged kernel
    eligible = bool((_capi.FLAGS_DEFAULT if flags is None else flags) & _capi.FLAG_CACHE_IMAGE_CHARGE) and s0.padded % 4 == 0 and s0.num_cells == 2
    want = {WS_PIPELINE: "constv_settle_step_ws_kernel", BULK: "constv_settle_step_bulk_kernel"}[WS] if eligible else "constv_settle_step_tiled_kernel"
    assert _capi.last_kernel() == want
    got = download(ctx, s0)
    ctx.close()
    return got


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("side_array", [False, True])
@pytest.mark.parametrize("num_real", [7, 37, 6007, 60001, 400003])
def test_ws_step_bit_exact(gpu, oracle, precision, side_array, num_real):
    """side_array False: the pipeline needs the image-charge side array, so the library falls back to the LDG-staged
    kernel -- same bits either way.  400003 real atoms: every CTA walks several tiles (the stages are reused)."""
    from openmm_constv_b200 import _capi
    if num_real > 100000 and precision != "single":
        pytest.skip("the large case runs in single precision only")
    s0, atoms, params, cons = slabmod.make_water_slab(num_real, precision, seed=41)
    flags = _capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side_array else 0)
    run_against_oracle(gpu, oracle, s0, atoms, params, cons, flags=flags, nsteps=3 if num_real < 100000 else 2)


@pytest.mark.parametrize("n_ion", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_ws_every_alignment_of_the_load_and_store_windows(gpu, oracle, n_ion):
    """The water run starts at slot n_ion and the wall run wherever the water ends: every residue of the 4-slot load
    window and of the 2-element force-plane store window, with R odd and even."""
    from openmm_constv_b200 import _capi
    for R in (1000 + n_ion, 1301 + n_ion):
        s0, atoms, params, cons = slabmod.make_water_slab(R, "single", seed=42 + n_ion, ion_fraction=n_ion / R)
        assert s0.meta["n_ion"] in (n_ion, n_ion + 1, n_ion + 2)  # the remainder of the water count goes to the ions
        for side in (False, True):
            flags = _capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side else 0)
            run_against_oracle(gpu, oracle, s0, atoms, params, cons, nsteps=2, flags=flags)


@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_ws_tiles_that_are_not_uniform(gpu, oracle, precision):
    """Massless and neutral atoms scattered through the ion run, a massless charged atom, a massive neutral one: the
    tile's records cannot go out as one bulk copy and the storer warp stores slot by slot."""
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(3000, precision, seed=43, ion_fraction=0.2)
    n_ion = s0.meta["n_ion"]
    rng = np.random.default_rng(5)
    R = s0.num_real
    for i in rng.choice(n_ion, n_ion // 3, replace=False):      # massless (walls among the ions)
        s0.velm[i, 3] = 0
        s0.velm[i, :3] = 0
        s0.masses[i] = 0
    for i in rng.choice(n_ion, n_ion // 3, replace=False):      # neutral: no image rewrite (constVLangevin.cu:143)
        s0.posq[i, 3] = 0
        s0.posq[R + i, 3] = 0
    for side in (False, True):
        flags = _capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side else 0)
        run_against_oracle(gpu, oracle, s0, atoms, params, cons, flags=flags)


def test_ws_without_image_zeroing(gpu, oracle):
    """No zeroing flags: the reference's behaviour (no zeroing kernel exists there): image velm / forces are left alone."""
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(5003, "single", seed=44)
    got = run_against_oracle(gpu, oracle, s0, atoms, params, cons, flags=_capi.FLAG_CACHE_IMAGE_CHARGE, zero=False)
    assert np.any(got.force[:, s0.num_real:s0.num_atoms] != 0)


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_ws_equals_the_ldg_staged_kernel(gpu, precision):
    """Same slab, same seed, 6 steps: variant 0 (the TMA pipeline) and variant 3 (the LDG-staged kernel) leave the same bits;
    launch counts say which kernel ran (one launch per step either way)."""
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(50021, precision, seed=45)
    out = []
    for variant in (WS, 3):
        ctx, it = make_context(gpu, s0, cons, seed=9, variant=variant, flags=_capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE)
        n0 = _capi.launch_count()
        it.step(6)
        # the first step takes the image-charge snapshot: one more launch
        assert _capi.launch_count() - n0 == 6 + 1
        out.append(download(ctx, s0))
        ctx.close()
    assert_same(out[0], out[1])


@pytest.mark.parametrize("chains", [2, 3])
def test_ws_chains_and_graph_replay(gpu, chains):
    """Chains = shares of the tile space with a step counter each; a captured graph of 4 steps replayed twice == 8 steps."""
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0, atoms, params, cons = slabmod.make_water_slab(30011, "single", seed=46)
    ctx, it = make_context(gpu, s0, cons, seed=3)
    it.step(8)
    want = download(ctx, s0)
    ctx.close()
    ctx, it = make_context(gpu, s0, cons, seed=3)
    _capi.check(lib.constv_set_chains(it._kernel.handle, chains))
    it.step(8)
    assert_same(download(ctx, s0), want)
    ctx.close()
    ctx, it = make_context(gpu, s0, cons, seed=3)
    h = it._kernel.handle
    _capi.check(lib.constv_set_parameters(h, T, GAMMA, DT))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        _capi.check(lib.constv_step_fused_settle(h, None, 0, stream.cuda_stream))  # outside capture: takes the snapshot
        stream.synchronize()
        _capi.check(lib.constv_set_step_count(h, 0))
        ctx.load_slab(s0)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(4):
                _capi.check(lib.constv_step_fused_settle(h, None, 0, stream.cuda_stream))
        graph.replay()
        graph.replay()
        stream.synchronize()
    torch.cuda.synchronize()
    assert_same(download(ctx, s0), want)
    ctx.close()


def _molecule_preserving_permutation(s, rng):
    """perm[new slot] = old slot: whole water molecules trade places, images freely (what OpenMM's re-sort does)."""
    R, N = s.num_real, s.num_atoms
    n_ion, n_mol = s.meta["n_ion"], s.meta["n_molecules"]
    perm = np.arange(N, dtype=np.int32)
    order = rng.permutation(n_mol)
    for k in range(3):
        perm[n_ion + k: n_ion + 3 * n_mol: 3] = n_ion + 3 * order + k
    perm[R:N] = R + rng.permutation(N - R)
    return perm


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("num_real", [1003, 40009])
def test_bulk_stage_in_on_reordered_slots(gpu, precision, num_cells, num_real):
    """Reordered slots (whole molecules swapped, images anywhere), Philox noise keyed by the original index, the per-slot image
    tables: kernel variant 6 (bulk stage-in, three more bulk copies for atomIndex and the tables) leaves the bits of
    variant 7 (the cp.async-staged kernel, itself pinned on the oracle in tests/test_gpu_settle.py), with a re-sort
    between steps."""
    import torch
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(num_real, precision, num_cells=num_cells, seed=51)
    perm = _molecule_preserving_permutation(s0, np.random.default_rng(11))
    sp, atom_index = H.permute_slab(s0, perm)
    N, P = s0.num_atoms, s0.padded
    out = []
    for variant, kernel in ((BULK, "constv_settle_step_bulk_kernel"), (7, "constv_settle_step_tiled_kernel")):
        ctx, it = make_context(gpu, sp, cons, seed=13, variant=variant, flags=_capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE)
        ctx.set_atom_order(perm)
        lib, slab = it._kernel._lib, it._kernel.handle
        order = np.arange(P, dtype=np.int32)
        order[:N] = atom_index[:N]
        for k in range(4):
            if k == 2:  # slots trade places again, in place, as OpenMM does it
                sigma = np.arange(P)
                sigma[:N] = _molecule_preserving_permutation(s0, np.random.default_rng(12))
                sig = torch.as_tensor(sigma, device="cuda")
                for arr in (ctx.posq, ctx.velm) + ((ctx.corr,) if ctx.corr is not None else ()):
                    arr.copy_(arr[sig].clone())
                ctx.force.copy_(ctx.force[:, sig].clone())
                order = order[sigma]
                ctx.atom_index.copy_(torch.as_tensor(order, device="cuda"))
                _capi.check(lib.constv_update_inverse_order(slab, ctx.stream_ptr()))
            ctx.force.copy_(torch.from_numpy(slabmod.new_forces(sp, k)))
            it.step(1)
            if k >= 1:  # the first step takes the image-charge snapshot and builds the tables behind the staged kernel
                assert _capi.last_kernel() == kernel
        out.append(download(ctx, sp))
        ctx.close()
    assert_same(out[0], out[1])


def test_automatic_choice_of_the_stage_in(gpu):
    """Variant 0: per-thread asynchronous copies for one launch of fewer than two waves of CTAs, the bulk stage-in for a
    chain's launch; same bits."""
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0, atoms, params, cons = slabmod.make_water_slab(30011, "single", seed=47)
    out = []
    for chains, kernel in ((1, "constv_settle_step_tiled_kernel"), (3, "constv_settle_step_bulk_kernel")):
        ctx, it = make_context(gpu, s0, cons, seed=3, variant=0, flags=_capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE)
        if chains > 1:
            _capi.check(lib.constv_set_chains(it._kernel.handle, chains))
        it.step(3)
        assert _capi.last_kernel() == kernel
        out.append(download(ctx, s0))
        ctx.close()
    assert_same(out[0], out[1])
