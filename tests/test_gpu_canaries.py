"""Out-of-bounds canaries (compute-sanitizer is not available on the GPU pool): every device array the step
kernels touch is a slice of a larger buffer filled with a sentinel; after stepping through each kernel family --
fused, chains, reordered slots, the staged and gather rigid-water kernels, the staged and thread-per-slot Drude
kernels, the split paths, ragged sizes -- the sentinel regions must be untouched, bit for bit."""
import numpy as np
import pytest

from openmm_constv_b200 import slab as slabmod
from tests import helpers as H
from tests.test_gpu_parity import gpu, make_context  # noqa: F401
from tests.test_gpu_drude import make_drude_context
from tests.test_gpu_settle import make_water_context, molecule_preserving_permutation

pytestmark = pytest.mark.gpu

GUARD = 2048  # elements (rows of 4 for posq/velm, int64 words for the force planes) on either side
SENT32, SENT64 = 0x7FC0DEAD, 0x7FF8DEADBEEFDEAD


class Guarded:
    """Re-homes a SlabContext's arrays inside sentinel-filled buffers and re-binds them."""

    def __init__(self, ctx, it):
        import torch
        self.torch, self.ctx = torch, ctx
        self.big = {}
        for name in ("posq", "corr", "velm", "pos_delta"):
            t = getattr(ctx, name)
            if t is None:
                continue
            itype = torch.int32 if t.element_size() == 4 else torch.int64
            sent = SENT32 if t.element_size() == 4 else SENT64
            big = torch.full((t.shape[0] + 2 * GUARD, 4), sent, dtype=itype, device=t.device)
            view = big[GUARD:GUARD + t.shape[0]].view(t.dtype)
            view.copy_(t)
            self.big[name] = (big, sent, t.shape[0])
            setattr(ctx, name, view)
        f = ctx.force
        big = torch.full((f.numel() + 2 * GUARD,), SENT64, dtype=torch.int64, device=f.device)
        view = big[GUARD:GUARD + f.numel()].view(f.shape)
        view.copy_(f)
        self.big["force"] = (big, SENT64, None)
        ctx.force = view
        it._kernel.rebind()

    def check(self):
        self.torch.cuda.synchronize()
        for name, (big, sent, rows) in self.big.items():
            flat = big.reshape(-1)
            inner = flat.numel() - 2 * GUARD * (4 if rows is not None else 1)
            g = GUARD * (4 if rows is not None else 1)
            assert bool((flat[:g] == sent).all()), f"{name}: wrote below the array"
            assert bool((flat[g + inner:] == sent).all()), f"{name}: wrote above the array"


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_real,num_cells", [(1, 2), (255, 2), (4097, 2), (1000, 4)])
def test_langevin_kernels_stay_inside_their_arrays(gpu, precision, num_real, num_cells):
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0 = slabmod.make_slab(num_real, precision, num_cells=num_cells, seed=50)
    # fused, then chains
    ctx, it = make_context(gpu, s0, seed=3)
    g = Guarded(ctx, it)
    it.step(3)
    _capi.check(lib.constv_set_chains(it._kernel.handle, 3))
    it.step(3)
    g.check()
    ctx.close()
    # split path
    ctx, it = make_context(gpu, s0, seed=3, apply_constraints=lambda c, t: None)
    g = Guarded(ctx, it)
    it.step(3)
    g.check()
    ctx.close()
    # reordered slots, fused and split
    perm = np.random.default_rng(1).permutation(s0.num_atoms).astype(np.int32)
    sp, _ = H.permute_slab(s0, perm)
    for split in (False, True):
        ctx, it = make_context(gpu, sp, seed=3, apply_constraints=(lambda c, t: None) if split else None)
        g = Guarded(ctx, it)
        ctx.set_atom_order(perm)
        it.step(3)
        g.check()
        ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_real,num_cells", [(37, 2), (4099, 2), (3001, 4)])
def test_rigid_water_kernels_stay_inside_their_arrays(gpu, precision, num_real, num_cells):
    from openmm_constv_b200 import _capi
    lib = _capi.load()
    s0, atoms, params, cons = slabmod.make_water_slab(num_real, precision, num_cells=num_cells, seed=51)
    for variant in (0, 1):
        ctx, it = make_water_context(gpu, s0, cons, seed=4)
        g = Guarded(ctx, it)
        it._kernel.set_kernel_variant(variant)
        it.step(3)
        if variant == 0 and len(atoms) > 300:
            _capi.check(lib.constv_set_chains(it._kernel.handle, 2))
            it.step(2)
        g.check()
        ctx.close()
    ctx, it = make_water_context(gpu, s0, cons, seed=4, apply_constraints=lambda c, t: c.getIntegrator()._kernel.settle_positions())
    g = Guarded(ctx, it)
    it.step(3)
    g.check()
    ctx.close()
    perm = molecule_preserving_permutation(s0, np.random.default_rng(2))
    sp, _ = H.permute_slab(s0, perm)
    for variant in (0, 1):
        ctx, it = make_water_context(gpu, sp, cons, seed=4)
        g = Guarded(ctx, it)
        ctx.set_atom_order(perm)
        it._kernel.set_kernel_variant(variant)
        it.step(3)
        g.check()
        ctx.close()


@pytest.mark.parametrize("precision", slabmod.PRECISIONS)
@pytest.mark.parametrize("num_real,num_cells", [(41, 2), (4099, 2), (3001, 4)])
def test_drude_kernels_stay_inside_their_arrays(gpu, precision, num_real, num_cells):
    from openmm_constv_b200 import _capi
    from tests.test_gpu_drude import role_preserving_permutation
    lib = _capi.load()
    s0, normal, pairs = slabmod.make_drude_slab(num_real, precision, num_cells=num_cells, seed=52)
    far = np.stack([pairs[:, 0], np.roll(pairs[:, 1], -37)], axis=1).astype(np.int32) if len(pairs) > 40 else pairs
    for variant, pp, wall in ((0, pairs, 0.02), (1, pairs, 0.02), (0, far, 0.0)):
        ctx, it = make_drude_context(gpu, s0, pp, wall, seed=5)
        g = Guarded(ctx, it)
        it._kernel.set_kernel_variant(variant)
        it.step(3)
        if variant == 0 and num_real > 1000:
            _capi.check(lib.constv_set_chains(it._kernel.handle, 2))
            it.step(2)
        g.check()
        ctx.close()
    ctx, it = make_drude_context(gpu, s0, pairs, 0.02, seed=5, apply_constraints=lambda c, t: None)
    g = Guarded(ctx, it)
    it.step(3)
    g.check()
    ctx.close()
    perm = role_preserving_permutation(s0, np.random.default_rng(3))
    sp, _ = H.permute_slab(s0, perm)
    for variant in (0, 1):
        ctx, it = make_drude_context(gpu, sp, pairs, 0.02, seed=5)
        g = Guarded(ctx, it)
        ctx.set_atom_order(perm)
        it._kernel.set_kernel_variant(variant)
        it.step(3)
        g.check()
        ctx.close()
