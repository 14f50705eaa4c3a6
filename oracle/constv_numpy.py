"""Second, independent CPU restatement of the ConstV Langevin step in numpy -- TEST INFRASTRUCTURE.

It exists to cross-check ``constv_oracle.c`` (the reference has no golden vectors for this path,
SURVEY.md 8c): two restatements written from the same reference lines must agree bit for bit.
Small cases only.

Reference lines restated (relative to /root/reference/):
  platforms/cuda/src/kernels/constVLangevin.cu:7-28, 34-75, 138-164, 170-177
  platforms/cuda/src/CudaConstVKernels.cpp:113-137
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble  # x87 80-bit: products of two floats and float fma are (all but) exact


def _split(a):
    c = 134217729.0 * a  # 2^27 + 1 (Veltkamp)
    hi = c - (c - a)
    return hi, a - hi


def _two_prod(a, b):
    p = a * b
    ah, al = _split(a)
    bh, bl = _split(b)
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def _two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def fma(a, b, c):
    """Correctly rounded a*b+c for float32 inputs (via 64-bit-mantissa long double) and, up to a
    ~2^-50 double-rounding hazard, for float64 inputs (error-free product + sum)."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        b = np.asarray(b, np.float32)
        c = np.asarray(c, np.float32)
        return (a.astype(LD) * b.astype(LD) + c.astype(LD)).astype(np.float32)
    a = a.astype(np.float64)
    b = np.asarray(b, np.float64)
    c = np.asarray(c, np.float64)
    p, e = _two_prod(a, b)
    s, t = _two_sum(p, c)
    return s + (t + e)


def params(temperature, friction, step_size, boltz):
    """CudaConstVKernels.cpp:113-137."""
    kT = boltz * temperature
    vscale = np.exp(-step_size * friction)
    fscale = step_size if friction == 0 else (1 - vscale) / friction
    noisescale = np.sqrt(kT * (1 - vscale * vscale))
    return np.array([vscale, fscale, noisescale], np.float64)


def part1(mixed, velm, force, pos_delta, prm, step_size, random4, random_index, num_atoms):
    """constVLangevin.cu:7-28.  velm/pos_delta are (P,4) `mixed`; force is (3,P) int64."""
    m = np.dtype(mixed).type
    vscale, fscale, noisescale = (m(x) for x in prm)
    fscale = m(fscale / m(4294967296.0))
    dt = m(step_size)
    idx = np.nonzero(velm[:num_atoms, 3] != 0)[0]
    w = velm[idx, 3]
    fw = (fscale * w).astype(m)
    ns = (noisescale * np.sqrt(w).astype(m)).astype(m)
    for k in range(3):
        f = force[k, idx].astype(m)
        xi = random4[random_index + idx, k].astype(m)
        t = (fw * f).astype(m)  # nvcc: FMUL t = (fscale*w)*F; FFMA(vscale, v, t); FFMA(ns, xi, .)
        velm[idx, k] = fma(ns, xi, fma(vscale, velm[idx, k], t))
        pos_delta[idx, k] = (dt * velm[idx, k]).astype(m)
    pos_delta[idx, 3] = 0


def part2(real, mixed, posq, posq_corr, pos_delta, velm, step_size, num_atoms):
    """constVLangevin.cu:34-75 (__CUDA_ARCH__ >= 130 branch)."""
    r, m = np.dtype(real).type, np.dtype(mixed).type
    inv_dt = np.float64(1.0) / np.float64(m(step_size))
    idx = np.nonzero(velm[:num_atoms, 3] != 0)[0]
    d = pos_delta[idx, :3]
    if r is np.float32 and m is np.float64:
        pos = posq[idx, :3].astype(m) + posq_corr[idx, :3].astype(m)
        pos = pos + d
        lo = pos.astype(r)
        posq[idx, :3] = lo
        posq_corr[idx, :3] = (pos - lo.astype(m)).astype(r)
        posq_corr[idx, 3] = 0
    else:
        posq[idx, :3] = (posq[idx, :3] + d).astype(r)
    velm[idx, :3] = (inv_dt * d.astype(np.float64)).astype(m)


def mirror(real, mixed, num_real, num_cells, zshift, posq, posq_corr, inv):
    """constVLangevin.cu:138-164."""
    r, m = np.dtype(real).type, np.dtype(mixed).type
    split = r is np.float32 and m is np.float64
    for a in range(num_real):
        s0 = inv[a]
        q = posq[s0, 3]
        if not (q != -q):
            continue
        if split:
            x, y, z = (np.float64(posq[s0, k]) + np.float64(posq_corr[s0, k]) for k in range(3))
        else:
            x, y, z = posq[s0, 0], posq[s0, 1], posq[s0, 2]
        for i in range(1, num_cells):
            si = inv[a + num_real * i]
            z64 = fma(np.float64(zshift), np.float64(2 * i), -np.float64(z))
            if split:
                z = z64
                p = np.array([x, y, z], np.float64)
                lo = p.astype(np.float32)
                posq[si, :3] = lo
                posq_corr[si, :3] = (p - lo.astype(np.float64)).astype(np.float32)
                posq_corr[si, 3] = 0
            else:
                z = r(z64)
                posq[si, 0], posq[si, 1], posq[si, 2] = x, y, z


def zero_images(num_atoms, num_real, velm, force, inv):
    slots = inv[num_real:num_atoms]
    velm[slots, :] = 0
    force[:, slots] = 0


def kinetic_energy(mixed, velm, force, time_shift, num_atoms):
    m = np.dtype(mixed).type
    scale = m(time_shift / 4294967296.0)
    idx = np.nonzero(velm[:num_atoms, 3] != 0)[0]
    w = velm[idx, 3]
    e = np.zeros(len(idx), np.float64)
    for k in range(3):
        sf = (scale * force[k, idx].astype(m)).astype(m)
        v = fma(sf, w, velm[idx, k]).astype(np.float64)
        e += v * v
    return 0.5 * float(np.sum(e / w.astype(np.float64)))


_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """Random123 Philox4x32 with 10 rounds, plain Python integers."""
    c = [int(x) for x in counter]
    k = [int(x) for x in key]
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + _W0) & 0xFFFFFFFF, (k[1] + _W1) & 0xFFFFFFFF]
    return np.array(c, np.uint32)
