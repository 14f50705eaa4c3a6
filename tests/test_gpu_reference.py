"""The product kernels against the REFERENCE'S OWN CUDA kernels on the same B200.

oracle/_ref/libconstv_ref_cuda.so is /root/reference/platforms/cuda/src/kernels/constVLangevin.cu
compiled unmodified by nvcc for sm_100a (oracle/ref_cuda.cu): the code OpenMM would build at run
time on this GPU, launched as CudaConstVKernels.cpp:146-166 launches it (part 1, part 2, image
rewrite).  With the same injected Gaussians the fused kernel, the split path and the reordered-slot
path must reproduce its arrays BIT FOR BIT (image zeroing off = the reference's behaviour), and so
must the CPU oracle: this pins the FMA contraction the oracle's canonical order assumes.

With --use_fast_math (OpenMM's default optimisation flag, recalled) the reference's own results
move by an ulp or two in single precision; the test states how far.
"""
import glob
import os

import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H
from tests.test_gpu_parity import assert_slab_bits_equal, download, gpu, make_context  # noqa: F401

pytestmark = pytest.mark.gpu

T, GAMMA, DT = 300.0, 1.0, 0.001
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def refmod():
    from oracle import ref
    if not ref.available("cuda"):
        pytest.skip("oracle/_ref/libconstv_ref_cuda.so is not built (needs /root/reference at build time)")
    return ref


def run_reference(refmod, oracle, s, atom_index, pools, fastmath=False, gamma=GAMMA):
    """Advance a copy of slab `s` with the reference's kernels on the GPU; returns (slab, posDelta)."""
    import torch
    dev = torch.device("cuda", 0)
    prm, dt = H.oracle_params_for(oracle, s.precision, T, gamma, DT)
    r = refmod.CudaReference(s.precision, prm, dt, dev, fastmath=fastmath, max_blocks=6 * 148)
    posq = torch.as_tensor(s.posq, device=dev)
    velm = torch.as_tensor(s.velm, device=dev)
    force = torch.as_tensor(s.force, device=dev)
    corr = torch.as_tensor(s.corr, device=dev) if s.corr is not None else None
    ai = torch.as_tensor(atom_index, device=dev)
    inv = ai.clone()
    r.inverse_order(s.num_atoms, ai, inv)
    delta = torch.zeros_like(velm)
    for xi in pools:
        r.step(s.num_atoms, s.num_cells, s.zmax, posq, corr, velm, force, delta, torch.as_tensor(xi, device=dev), 0, inv)
    torch.cuda.synchronize()
    out = s.copy()
    out.posq, out.velm, out.force = posq.cpu().numpy(), velm.cpu().numpy(), force.cpu().numpy()
    if corr is not None:
        out.corr = corr.cpu().numpy()
    return out, delta.cpu().numpy()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_cells", [2, 4])
@pytest.mark.parametrize("num_real", [33, 20011])
def test_fused_kernel_equals_reference_kernels(gpu, oracle, refmod, precision, num_cells, num_real):
    s0 = slabmod.make_slab(num_real, precision, num_cells=num_cells, seed=31)
    nsteps = 3
    pools = [np.random.default_rng(50 + k).standard_normal((s0.padded, 4)).astype(np.float32) for k in range(nsteps)]
    want, _ = run_reference(refmod, oracle, s0, H.identity_order(s0.padded), pools)
    ctx, it = make_context(gpu, s0, flags=0, pool=np.concatenate(pools))  # no zeroing: reference behaviour
    it.step(nsteps)
    assert_slab_bits_equal(download(ctx, s0), want)
    ctx.close()
    # ... and the CPU oracle agrees with the reference kernels as nvcc contracts them
    orc = s0.copy()
    prm, dt = H.oracle_params_for(oracle, precision, T, GAMMA, DT)
    for xi in pools:
        H.oracle_step(oracle, orc, prm, dt, xi, zero=False)
    assert_slab_bits_equal(orc, want)


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_split_path_equals_reference_kernels(gpu, oracle, refmod, precision):
    s0 = slabmod.make_slab(4097, precision, seed=32)
    pool = np.random.default_rng(7).standard_normal((s0.padded, 4)).astype(np.float32)
    want, want_delta = run_reference(refmod, oracle, s0, H.identity_order(s0.padded), [pool], gamma=0.0)
    seen = {}
    ctx, it = make_context(gpu, s0, gamma=0.0, flags=0, pool=pool,
                           apply_constraints=lambda c, tol: seen.update(delta=c.pos_delta.cpu().numpy()))
    it.step(1)
    massive = s0.velm[:, 3] != 0
    H.assert_bits_equal(seen["delta"][massive], want_delta[massive], "posDelta handed to the constraint solver")
    assert_slab_bits_equal(download(ctx, s0), want)
    ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
def test_reordered_slots_equal_reference_kernels(gpu, oracle, refmod, precision):
    s0 = slabmod.make_slab(2501, precision, num_cells=4, seed=33)
    perm = np.random.default_rng(3).permutation(s0.num_atoms).astype(np.int32)
    sp, atom_index = H.permute_slab(s0, perm)
    pools = [np.random.default_rng(60 + k).standard_normal((s0.padded, 4)).astype(np.float32) for k in range(2)]
    want, _ = run_reference(refmod, oracle, sp, atom_index, pools)
    ctx, it = make_context(gpu, sp, flags=0, pool=np.concatenate(pools))
    ctx.set_atom_order(perm)
    it.step(2)
    assert_slab_bits_equal(download(ctx, sp), want)
    ctx.close()


def test_fast_math_build_of_the_reference_stays_within_tolerance(gpu, oracle, refmod):
    """If OpenMM compiles its kernels with --use_fast_math the reference's single-precision square
    root and divisions are approximate; its own results then differ from its IEEE build (and from
    ours) by rounding only: < 1e-6 of the thermal velocity / of the coordinate per step."""
    if not refmod.available("cuda_fastmath"):
        pytest.skip("fast-math reference build absent")
    s0 = slabmod.make_slab(20011, "single", seed=34)
    pool = np.random.default_rng(9).standard_normal((s0.padded, 4)).astype(np.float32)
    a, _ = run_reference(refmod, oracle, s0, H.identity_order(s0.padded), [pool])
    b, _ = run_reference(refmod, oracle, s0, H.identity_order(s0.padded), [pool], fastmath=True)
    n, R = s0.num_atoms, s0.num_real
    massive = s0.velm[:R, 3] != 0
    v_thermal = np.sqrt(np.mean(s0.velm[:R, :3][massive] ** 2))
    assert np.max(np.abs(a.velm[:n, :3] - b.velm[:n, :3])) < 1e-6 * v_thermal
    assert np.max(np.abs(a.posq[:n, :3] - b.posq[:n, :3])) < 1e-6 * np.max(np.abs(s0.posq[:n, :3]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_cuda_*.npz"))))
def test_committed_reference_gpu_vectors(gpu, path):
    """Fixtures recorded from the reference's kernels on a B200 (make_reference_vectors.py --cuda)."""
    import torch
    g = np.load(path)
    precision, num_cells, num_atoms, steps = str(g["precision"]), int(g["num_cells"]), int(g["num_atoms"]), int(g["steps"])
    s = slabmod.make_slab(num_atoms // num_cells, precision, num_cells=num_cells, seed=1)
    s.posq, s.velm, s.force = g["posq0"].copy(), g["velm0"].copy(), g["force"].copy()
    if precision == "mixed":
        s.corr = g["corr0"].copy()
    s.zmax = float(g["zmax"])
    s.masses = np.zeros(num_atoms)  # validation only needs the images massless; velm carries 1/m
    atom_index = g["atom_index"]
    orig_w = np.zeros(num_atoms)
    orig_w[atom_index[:num_atoms]] = s.velm[:num_atoms, 3]
    s.masses = np.where(orig_w != 0, 1.0 / np.where(orig_w != 0, orig_w, 1.0), 0.0)
    ctx, it = make_context(gpu, s, flags=0, pool=g["noise"].reshape(-1, 4))
    ctx.velm.copy_(torch.as_tensor(s.velm))  # exact stored inverse masses, not 1/(1/w)
    if not np.array_equal(atom_index, np.arange(len(atom_index))):
        ctx.set_atom_order(atom_index[:num_atoms])
    it.step(steps)
    H.assert_bits_equal(ctx.posq.cpu().numpy(), g["posq1"], "posq")
    H.assert_bits_equal(ctx.velm.cpu().numpy(), g["velm1"], "velm")
    if precision == "mixed":
        H.assert_bits_equal(ctx.corr.cpu().numpy(), g["corr1"], "posqCorrection")
    ctx.close()
