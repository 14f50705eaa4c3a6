"""Measures the device Gaussian stream against the oracle (run on the GPU box)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from openmm_constv_b200 import _capi, integrator as H, slab as S
s = S.make_slab(1_000_000, "single", seed=1)
sysm = H.SlabSystem(); sysm.addParticles(s.masses)
sysm.setDefaultPeriodicBoxVectors((s.box[0],0,0),(0,s.box[1],0),(0,0,2*s.zmax))
it = H.ConstVLangevinIntegrator(300,1,0.001); it.setRandomNumberSeed(123456789)
ctx = H.SlabContext(sysm, it, padded=s.padded); ctx.load_slab(s)
out = torch.zeros((s.padded,4), dtype=torch.float32, device="cuda")
worst = 0.0
for step in range(4):
    _capi.check(_capi.load().constv_generate_noise(it._kernel.handle, step, out.data_ptr(), ctx.stream_ptr()))
    got = out.cpu().numpy()[:s.num_atoms]
    want = oracle.fill_noise(s.num_atoms, 123456789, step)
    err = np.abs(got - want)
    worst = max(worst, float(err.max()))
    i = np.unravel_index(err.argmax(), err.shape)
    print(f"step {step}: max abs err {err.max():.3e} at value {want[i]:.4f}; mean {got.mean():+.5f} std {got.std():.5f} "
          f"p99.9 err {np.quantile(err, 0.999):.2e}; small-|x| (<0.05) max err {err[np.abs(want)<0.05].max():.2e}")
print("WORST", worst)
