"""Parity of the rigid-molecule step (constv_step_fused_settle / constv_settle_positions) against the CPU
oracle (oracle/constv_oracle_settle_impl.h), through the C ABI.  The solver is OpenMM's, not the reference's
(CudaConstVKernels.cpp:153 hands posDelta to integration.applyConstraints): PARITY UNPINNED against OpenMM;
the oracle is pinned on converged SHAKE and on the algorithm's invariants (tests/test_oracle_settle.py).
Here: the CUDA kernels reproduce the oracle BIT FOR BIT -- fused step, split path, reordered slots, Philox
noise, three precision models -- and molecules stay rigid over a trajectory."""
import ctypes as C

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H
from tests.test_gpu_parity import assert_slab_bits_equal, download, gpu  # noqa: F401

pytestmark = pytest.mark.gpu

T, GAMMA, DT = 300.0, 1.0, 0.002


def make_water_context(gpu, s, constraints, seed=77, flags=None, pool=None, **kw):
    from openmm_constv_b200 import _capi
    import torch
    sysm = gpu.SlabSystem()
    sysm.addParticles(s.masses)
    sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, s.num_cells * s.zmax))
    sysm.addConstraints(*constraints)
    it = gpu.ConstVLangevinIntegrator(T, GAMMA, DT, s.num_cells)
    it.setRandomNumberSeed(seed)
    ctx = gpu.SlabContext(sysm, it, precision=s.precision, flags=_capi.FLAGS_DEFAULT if flags is None else flags,
                          random_pool=None if pool is None else torch.as_tensor(pool).cuda(),
                          global_atom_offset=s.global_atom_offset, padded=s.padded, **kw)
    ctx.load_slab(s)
    return ctx, it


def oracle_settle_step(oracle, s, atoms, params, xi, random_index=0, inv=None):
    prm, dt = H.oracle_params_for(oracle, s.precision, T, GAMMA, DT)
    delta = np.zeros((s.padded, 4), H.MIXED[s.precision])
    oracle.settle_step(s.precision, s.num_atoms, s.num_cells, s.zmax, s.posq, s.corr, s.velm, s.force.reshape(-1), delta, prm,
                       dt, xi, random_index, H.identity_order(s.padded) if inv is None else inv, atoms, params)
    return delta


def bond_errors(s, atoms, params):
    p = s.posq[:, :3].astype(np.float64)
    if s.corr is not None:
        p = p + s.corr[:, :3]
    out = []
    for (i, j, d) in ((0, 1, params[:, 0]), (0, 2, params[:, 0]), (1, 2, params[:, 1])):
        out.append(np.max(np.abs(np.linalg.norm(p[atoms[:, i]] - p[atoms[:, j]], axis=1) / d - 1)))
    return max(out)


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("num_real", [37, 6007])
@pytest.mark.parametrize("variant", [0, 1, 3])
def test_fused_settle_step_bit_exact(gpu, oracle, precision, num_cells, num_real, variant):
    """variant 0 = the kernel staged through shared memory (tiles of 128 work items, coalesced both ways; the stage filled
    by asynchronous copies), variant 3 = the same with the stage filled through registers,
    variant 1 = the gather kernel (a thread reads its three records itself)."""
    import torch
    s0, atoms, params, cons = slabmod.make_water_slab(num_real, precision, num_cells=num_cells, seed=31)
    nsteps = 4
    pool = np.random.default_rng(2).standard_normal((nsteps * s0.padded, 4)).astype(np.float32)
    ctx, it = make_water_context(gpu, s0, cons, pool=pool)
    it._kernel.set_kernel_variant(variant)
    assert it._kernel.fused_constraints and it._kernel.num_rigid_molecules == len(atoms)
    ref = s0.copy()
    launches = __import__("openmm_constv_b200")._capi.launch_count()
    for k in range(nsteps):
        f = slabmod.new_forces(s0, k)
        ctx.force.copy_(torch.from_numpy(f))
        ref.force[:] = f
        it.step(1)
        oracle_settle_step(oracle, ref, atoms, params, pool, random_index=k * s0.padded)
        assert_slab_bits_equal(download(ctx, s0), ref)
    assert __import__("openmm_constv_b200")._capi.launch_count() - launches == nsteps  # ONE kernel per step
    assert it._kernel.get_step_count() == nsteps
    assert bond_errors(ref, atoms, params) < (5e-6 if precision == "single" else 2e-7)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_split_path_with_the_library_solver_equals_fused(gpu, oracle, precision):
    """part 1 | constv_settle_positions on pos_delta | part 2 + image rewrite: three launches, same bits."""
    s0, atoms, params, cons = slabmod.make_water_slab(3001, precision, num_cells=2, seed=32)
    pool = np.random.default_rng(3).standard_normal((s0.padded, 4)).astype(np.float32)
    seen = {}

    def constraints(ctx, tol):
        seen["before"] = ctx.pos_delta.cpu().numpy()
        ctx.getIntegrator()._kernel.settle_positions()
        seen["after"] = ctx.pos_delta.cpu().numpy()

    ctx, it = make_water_context(gpu, s0, cons, pool=pool, apply_constraints=constraints)
    it.step(1)
    ref = s0.copy()
    delta = oracle_settle_step(oracle, ref, atoms, params, pool)
    assert_slab_bits_equal(download(ctx, s0), ref)
    massive = s0.velm[:, 3] != 0
    H.assert_bits_equal(seen["after"][massive], delta[massive], "constrained posDelta")
    assert not np.array_equal(seen["before"][atoms.reshape(-1)], seen["after"][atoms.reshape(-1)])
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_fused_settle_scattered_molecules(gpu, oracle, precision):
    """Molecules whose atoms are NOT laid out (apex, apex+1, apex+2) -- here neighbouring molecules trade the
    slots of their second hydrogen -- take the item-list path of the kernel (two dependent gathers) instead of
    the arithmetic runs; same bits."""
    import torch
    s0, atoms, params, _ = slabmod.make_water_slab(4001, precision, num_cells=2, seed=36)
    R, n = s0.num_real, len(atoms) // 2 * 2
    a_h2, b_h2 = atoms[0:n:2, 2].copy(), atoms[1:n:2, 2].copy()
    for cell in range(s0.num_cells):  # trade the records (and those of their images)
        x, y = a_h2 + cell * R, b_h2 + cell * R
        for arr in (s0.posq, s0.velm) + ((s0.corr,) if s0.corr is not None else ()):
            arr[x], arr[y] = arr[y].copy(), arr[x].copy()
        s0.force[:, x], s0.force[:, y] = s0.force[:, y].copy(), s0.force[:, x].copy()
    atoms = atoms.copy()
    atoms[0:n:2, 2], atoms[1:n:2, 2] = b_h2, a_h2
    p1 = np.concatenate([atoms[:, 0], atoms[:, 0], atoms[:, 1]])
    p2 = np.concatenate([atoms[:, 1], atoms[:, 2], atoms[:, 2]])
    dist = np.concatenate([params[:, 0], params[:, 0], params[:, 1]]).astype(np.float64)
    # (apex, b, c) as the detection rule orders them (b < c): SETTLE is not bitwise symmetric in b <-> c
    from oracle import settle_numpy
    atoms, params, other = settle_numpy.find_settle_clusters(s0.num_atoms, p1, p2, dist)
    assert other == 0 and np.any(atoms[:, 2] != atoms[:, 0] + 2)
    pool = np.random.default_rng(9).standard_normal((3 * s0.padded, 4)).astype(np.float32)
    ctx, it = make_water_context(gpu, s0, (p1, p2, dist), pool=pool)
    assert it._kernel.fused_constraints and it._kernel.num_rigid_molecules == len(atoms)
    ref = s0.copy()
    for k in range(3):
        it.step(1)
        oracle_settle_step(oracle, ref, atoms, params, pool, random_index=k * s0.padded)
        assert_slab_bits_equal(download(ctx, s0), ref)
    assert bond_errors(ref, atoms, params) < (5e-6 if precision == "single" else 2e-7)
    ctx.close()


def molecule_preserving_permutation(s, rng):
    R, N = s.num_real, s.num_atoms
    n_ion, n_mol = s.meta["n_ion"], s.meta["n_molecules"]
    perm = np.arange(N, dtype=np.int32)
    order = rng.permutation(n_mol)
    for k in range(3):
        perm[n_ion + k: n_ion + 3 * n_mol: 3] = n_ion + 3 * order + k
    perm[R:N] = R + rng.permutation(N - R)
    return perm


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("variant,num_cells", [(0, 2), (1, 2), (0, 4)])
@pytest.mark.parametrize("side_array", [False, True])
def test_fused_settle_reordered_slots(gpu, oracle, precision, variant, num_cells, side_array):
    """OpenMM swaps whole identical molecules: slots keep their roles, images move freely.  variant 0 = the
    staged kernel, variant 1 = the gather kernel.  side_array: with the image-charge snapshot the staged kernel reads the
    images' slots and charges from the per-slot tables (constv_image_by_slot_kernel) instead of gathering them."""
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(2501, precision, num_cells=num_cells, seed=33)
    perm = molecule_preserving_permutation(s0, np.random.default_rng(6))
    sp, atom_index = H.permute_slab(s0, perm)
    inv = np.empty_like(atom_index)
    inv[atom_index] = np.arange(len(atom_index), dtype=np.int32)
    pool = np.random.default_rng(4).standard_normal((2 * s0.padded, 4)).astype(np.float32)
    ctx, it = make_water_context(gpu, sp, cons, pool=pool,
                                 flags=_capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side_array else 0))
    ctx.set_atom_order(perm)
    it._kernel.set_kernel_variant(variant)
    ref = sp.copy()
    for k in range(2):
        it.step(1)
        oracle_settle_step(oracle, ref, atoms, params, pool, random_index=k * s0.padded, inv=inv)
        assert_slab_bits_equal(download(ctx, sp), ref)
    ctx.close()


@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_fused_settle_resort_between_steps(gpu, oracle, precision):
    """A second re-sort while stepping, as OpenMM does it: the state arrays and the atomIndex array change IN PLACE and the
    host calls constv_update_inverse_order.  The image-charge snapshot (indexed by atom) stays valid; the per-slot image
    tables of the staged kernel must follow the new order."""
    import torch
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(3001, precision, seed=35)
    N, P = s0.num_atoms, s0.padded
    perm1 = molecule_preserving_permutation(s0, np.random.default_rng(8))
    sp, atom_index = H.permute_slab(s0, perm1)
    pool = np.random.default_rng(5).standard_normal((4 * P, 4)).astype(np.float32)
    ctx, it = make_water_context(gpu, sp, cons, pool=pool, flags=_capi.FLAGS_DEFAULT | _capi.FLAG_CACHE_IMAGE_CHARGE)
    ctx.set_atom_order(perm1)
    lib, slab = it._kernel._lib, it._kernel.handle
    ref = sp.copy()

    def inverse(ai):
        inv = np.arange(P, dtype=np.int32)
        inv[ai[:N]] = np.arange(N, dtype=np.int32)
        return inv

    order = np.arange(P, dtype=np.int32)
    order[:N] = atom_index[:N]
    for k in range(4):
        if k == 2:  # slots trade places again (whole molecules, images freely): sigma[new slot] = old slot
            sigma = np.arange(P)
            sigma[:N] = molecule_preserving_permutation(s0, np.random.default_rng(9))
            sig = torch.as_tensor(sigma, device="cuda")
            for arr in (ctx.posq, ctx.velm) + ((ctx.corr,) if ctx.corr is not None else ()):
                arr.copy_(arr[sig].clone())
            ctx.force.copy_(ctx.force[:, sig].clone())
            order = order[sigma]
            ctx.atom_index.copy_(torch.as_tensor(order, device="cuda"))
            _capi.check(lib.constv_update_inverse_order(slab, ctx.stream_ptr()))
            ref.posq, ref.velm = ref.posq[sigma].copy(), ref.velm[sigma].copy()
            if ref.corr is not None:
                ref.corr = ref.corr[sigma].copy()
            ref.force = np.ascontiguousarray(ref.force[:, sigma])
        it.step(1)
        oracle_settle_step(oracle, ref, atoms, params, pool, random_index=k * P, inv=inverse(order))
        assert_slab_bits_equal(download(ctx, sp), ref)
    ctx.close()


@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_fused_settle_philox_and_rigid_trajectory(gpu, oracle, precision):
    import torch
    from openmm_constv_b200._capi import check
    s0, atoms, params, cons = slabmod.make_water_slab(20011, precision, seed=34)
    ctx, it = make_water_context(gpu, s0, cons, seed=4321)
    lib, slab = it._kernel._lib, it._kernel.handle
    g = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    check(lib.constv_generate_noise(slab, 0, g.data_ptr(), ctx.stream_ptr()))
    it.step(1)
    ref = s0.copy()
    oracle_settle_step(oracle, ref, atoms, params, g.cpu().numpy())
    assert_slab_bits_equal(download(ctx, s0), ref)
    # 300 more steps under the same (unphysically large, constant) forces: the molecules stay rigid
    it.step(300)
    got = download(ctx, s0)
    assert np.all(np.isfinite(got.posq)) and np.all(np.isfinite(got.velm))
    assert bond_errors(got, atoms, params) < (1e-4 if precision == "single" else 2e-7)  # rounding random-walks in float
    # same seed, same trajectory
    ctx2, it2 = make_water_context(gpu, s0, cons, seed=4321)
    it2.step(301)
    assert_slab_bits_equal(download(ctx2, s0), got)
    ctx.close()
    ctx2.close()


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("chains", [2, 5])
def test_fused_settle_chains(gpu, precision, chains):
    """The rigid-water step as chains (shares of the staged kernel's tiles, one step counter each): in turn on
    one stream, or each on a stream of its own and running ahead of the others -- the single launch's bits."""
    import torch
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0, atoms, params, cons = slabmod.make_water_slab(20011, precision, seed=37)
    nsteps = 4
    ctx_a, it_a = make_water_context(gpu, s0, cons, seed=5)
    it_a.step(nsteps)
    want = download(ctx_a, s0)
    ctx_a.close()
    ctx_b, it_b = make_water_context(gpu, s0, cons, seed=5)
    _capi.check(lib.constv_set_chains(it_b._kernel.handle, chains))
    it_b.step(nsteps)
    assert it_b._kernel.get_step_count() == nsteps
    assert_slab_bits_equal(download(ctx_b, s0), want)
    ctx_b.close()
    ctx_c, it_c = make_water_context(gpu, s0, cons, seed=5)
    h = it_c._kernel.handle
    _capi.check(lib.constv_set_chains(h, chains))
    _capi.check(lib.constv_set_parameters(h, T, GAMMA, DT))
    streams = [torch.cuda.Stream() for _ in range(chains)]
    torch.cuda.synchronize()
    for q in reversed(range(chains)):
        for _ in range(nsteps):
            _capi.check(lib.constv_step_fused_settle_chain(h, q, None, 0, streams[q].cuda_stream))
    torch.cuda.synchronize()
    assert it_c._kernel.get_step_count() == nsteps
    assert_slab_bits_equal(download(ctx_c, s0), want)
    assert lib.constv_step_fused_settle_chain(h, chains, None, 0, None) == _capi.ERR_INVALID
    it_c._kernel.set_kernel_variant(1)  # the gather kernel has no chains
    assert lib.constv_step_fused_settle_chain(h, 0, None, 0, None) == _capi.ERR_STATE
    ctx_c.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_real", [37, 100003])
def test_kinetic_energy_of_rigid_water(gpu, oracle, precision, num_real):
    """computeKineticEnergy on a System whose constraints are solved in the step: the half-step-shifted
    velocities are projected onto the rigid molecules' constraints first (as OpenMM's computeKineticEnergy does
    for a constrained System).  Against the oracle to 1e-12 (the per-molecule arithmetic is the same sequence of
    double operations; only the order of the final sum differs), before and after a step, identity and
    reordered slots."""
    s0, atoms, params, cons = slabmod.make_water_slab(num_real, precision, seed=38)
    ctx, it = make_water_context(gpu, s0, cons, seed=8)
    want = oracle.kinetic_energy_rigid(precision, s0.posq, s0.corr, s0.velm, s0.force.reshape(-1), 0.5 * DT, atoms, s0.num_atoms)
    free = oracle.kinetic_energy(precision, s0.velm, s0.force.reshape(-1), 0.5 * DT, s0.num_atoms)
    got = it.computeKineticEnergy()
    assert got == pytest.approx(want, rel=1e-12)
    if len(atoms) > 10:
        assert got < 0.9 * free  # Maxwell-Boltzmann velocities per atom: a third of the molecules' 9 dof is projected out
    assert ctx.getKineticEnergy() == got  # deterministic reduction order
    it.step(3)
    after = download(ctx, s0)
    want = oracle.kinetic_energy_rigid(precision, after.posq, after.corr, after.velm, after.force.reshape(-1), 0.5 * DT, atoms,
                                       s0.num_atoms)
    assert it.computeKineticEnergy() == pytest.approx(want, rel=1e-12)
    ctx.close()
    perm = molecule_preserving_permutation(s0, np.random.default_rng(7))
    sp, _ = H.permute_slab(s0, perm)
    ctx, it = make_water_context(gpu, sp, cons, seed=8)
    ctx.set_atom_order(perm)
    want = oracle.kinetic_energy_rigid(precision, sp.posq, sp.corr, sp.velm, sp.force.reshape(-1), 0.5 * DT, atoms, sp.num_atoms)
    assert it.computeKineticEnergy() == pytest.approx(want, rel=1e-12)
    ctx.close()


@pytest.mark.parametrize("precision", ["single", "mixed"])
def test_example_sized_rigid_water_trajectory(gpu, oracle, precision):
    """BASELINE configs[0-1] shape: the example's slab (7494 real + 7494 image particles, rigid SPC/E water,
    T = 300 K, friction 1/ps, dt = 1 fs; example/nacl_1m_spce/nacl_constv.py:12-16,56) with the applied field
    efield*q*z as the force, 200 steps with the in-kernel Philox noise, against the oracle driven with the
    Gaussians the library publishes for each step: every array bit for bit all along the trajectory."""
    import torch
    from openmm_constv_b200._capi import check
    s0, atoms, params, cons = slabmod.make_water_slab(7494, precision, seed=2024)
    R = s0.num_real
    efield = 1.0 * 96.48533212 / s0.zmax
    fz = np.rint(-efield * s0.posq[:R, 3].astype(np.float64) * 4294967296.0).astype(np.int64)
    s0.force[:] = 0
    s0.force[2, :R] = fz
    ctx, it = make_water_context(gpu, s0, cons, seed=99)
    it.setStepSize(0.001)
    lib, slab = it._kernel._lib, it._kernel.handle
    g = torch.zeros((s0.padded, 4), dtype=torch.float32, device="cuda")
    ref = s0.copy()
    prm, dt = H.oracle_params_for(oracle, precision, T, GAMMA, 0.001)
    inv = H.identity_order(s0.padded)
    delta = np.zeros((s0.padded, 4), H.MIXED[precision])
    for k in range(200):
        check(lib.constv_generate_noise(slab, k, g.data_ptr(), ctx.stream_ptr()))
        xi = g.cpu().numpy()
        ctx.force.zero_()
        ctx.force[2, :R] = torch.as_tensor(fz, device="cuda")  # the step zeroes image forces only; keep the field on
        it.step(1)
        ref.force[:] = 0
        ref.force[2, :R] = fz
        oracle.settle_step(precision, ref.num_atoms, 2, ref.zmax, ref.posq, ref.corr, ref.velm, ref.force.reshape(-1), delta,
                           prm, dt, xi, 0, inv, atoms, params)
        if k % 50 == 49:
            assert_slab_bits_equal(download(ctx, s0), ref)
    assert bond_errors(ref, atoms, params) < (2e-5 if precision == "single" else 2e-7)
    ke = it.computeKineticEnergy()
    assert ke == pytest.approx(oracle.kinetic_energy_rigid(precision, ref.posq, ref.corr, ref.velm, ref.force.reshape(-1),
                                                           0.0005, atoms, ref.num_atoms), rel=1e-12)
    ctx.close()


@pytest.mark.parametrize("precision", ["single", "mixed"])
@pytest.mark.parametrize("side_array", [False, True])
def test_batched_rigid_water_ensemble(gpu, precision, side_array):
    """constv_step_fused_settle_batched: 11 rigid-water slabs of different sizes and seeds in one call (two
    launches: groups of 8) == each slab stepped on its own.  side_array: with the image-charge snapshot the batch runs
    the bulk-staged kernel (constv_settle_step_bulk_batched_kernel), else the cp.async-staged one."""
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    flags = _capi.FLAGS_DEFAULT | (_capi.FLAG_CACHE_IMAGE_CHARGE if side_array else 0)
    sizes = [37, 500, 2049, 7494, 3000, 129, 1000, 6000, 4097, 999, 2500]
    solo, batch = [], []
    for k, n in enumerate(sizes):
        s0, atoms, params, cons = slabmod.make_water_slab(n, precision, seed=60 + k)
        ctx, it = make_water_context(gpu, s0, cons, seed=100 + k, flags=flags)
        it._kernel.set_kernel_variant(7)  # each slab on its own: the cp.async-staged kernel, pinned on the oracle above
        it.step(3)
        solo.append(download(ctx, s0))
        ctx.close()
        batch.append((s0,) + make_water_context(gpu, s0, cons, seed=100 + k, flags=flags))
    handles = (C.c_void_p * len(batch))(*[it._kernel.handle for _, _, it in batch])
    for _, _, it in batch:
        _capi.check(lib.constv_set_parameters(it._kernel.handle, T, GAMMA, DT))
    launches = _capi.launch_count()
    for _ in range(3):
        _capi.check(lib.constv_step_fused_settle_batched(handles, len(batch), None))
    # two launches per step (groups of 8), plus one image-charge snapshot per slab before the first step
    assert _capi.launch_count() - launches == 3 * 2 + (len(batch) if side_array else 0)
    assert _capi.last_kernel() == ("constv_settle_step_bulk_batched_kernel" if side_array else "constv_settle_step_tiled_batched_kernel")
    for want, (s0, ctx, it) in zip(solo, batch):
        assert it._kernel.get_step_count() == 3
        assert_slab_bits_equal(download(ctx, s0), want)
        ctx.close()


def test_settle_guard_rails(gpu):
    from openmm_constv_b200 import _capi
    s0, atoms, params, cons = slabmod.make_water_slab(301, "single", seed=35)
    ctx, it = make_water_context(gpu, s0, cons)
    lib, slab = it._kernel._lib, it._kernel.handle
    vp = C.c_void_p
    bad = np.array([[atoms[0, 0], atoms[0, 1], atoms[0, 2]], [atoms[0, 2], atoms[1, 1], atoms[1, 2]]], np.int32)
    prm = np.ascontiguousarray(params[:2])
    assert lib.constv_settle_configure(slab, 2, bad.ctypes.data_as(vp), prm.ctypes.data_as(vp)) == _capi.ERR_INVALID
    assert b"one molecule only" in lib.constv_last_error()
    bad = np.array([[s0.num_real, 0, 1]], np.int32)  # an image particle
    assert lib.constv_settle_configure(slab, 1, bad.ctypes.data_as(vp), prm.ctypes.data_as(vp)) == _capi.ERR_INVALID
    flat = np.array([[0.1, 0.25]], np.float32)  # d(b-c) >= 2 d(apex-b): no triangle
    assert lib.constv_settle_configure(slab, 1, atoms[:1].ctypes.data_as(vp), flat.ctypes.data_as(vp)) == _capi.ERR_INVALID
    it.step(1)  # the failed calls left the slab usable
    ctx.close()
    # constraints the step kernel cannot solve (a lone bond) need the caller's solver
    p1, p2, dist = cons
    sysm = gpu.SlabSystem()
    sysm.addParticles(s0.masses)
    sysm.setDefaultPeriodicBoxVectors((s0.box[0], 0, 0), (0, s0.box[1], 0), (0, 0, s0.num_cells * s0.zmax))
    sysm.addConstraints(p1, p2, dist)
    sysm.addConstraint(0, 1, 0.3)
    it = gpu.ConstVLangevinIntegrator(T, GAMMA, DT)
    ctx = gpu.SlabContext(sysm, it, padded=s0.padded)
    ctx.load_slab(s0)
    assert not it._kernel.fused_constraints and it._kernel.num_rigid_molecules == len(atoms)
    with pytest.raises(gpu.OpenMMException, match="apply_constraints"):
        it.step(1)
    ctx.close()
