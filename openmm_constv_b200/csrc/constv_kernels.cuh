// Device code of the B200-native ConstVLangevinIntegrator step path (sm_100a).
//
// What each kernel replaces in the reference (scychon/openmm_constV, platforms/cuda/src/kernels/
// constVLangevin.cu; host sequence platforms/cuda/src/CudaConstVKernels.cpp:139-172):
//   constv_fused_step_kernel          part 1 (:7-28) + part 2 (:34-75) + image rewrite (:138-164)
//                                     + explicit image zeroing + in-register Philox, ONE pass
//   constv_fused_step_indexed_kernel  same, for slots permuted by OpenMM's atom reordering
//   constv_part1_kernel               part 1 alone (constraint hand-off path)
//   constv_part2_mirror_kernel        part 2 + image rewrite + zeroing (after constraints)
//   constv_part2_indexed_kernel /     the same pair for reordered slots
//   constv_mirror_indexed_kernel
//   constv_inverse_order_kernel       reorderInverseAtomOrderIndexes (:170-177)
//   constv_kinetic_energy_kernel      OpenMM's computeKineticEnergy(0.5*dt) (Kernels.cpp:175-177)
//   constv_noise_kernel               the Gaussian stream that replaces OpenMM's random pool
//
// The path is HBM-bound integer/float streaming (about 20 flops against 148 bytes per real/image
// pair): no tensor cores.  What matters is that every record crosses HBM once, that all global
// accesses are coalesced 8/16-byte vectors, and that enough loads are in flight per SM.
//
// Arithmetic is written with explicit fma()/__f*_rn so that it is bit-identical to the CPU oracle
// (the file is also compiled with -fmad=false); the canonical operation order is stated in
// DESIGN.md "Canonical arithmetic".
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace constv {

enum { PREC_SINGLE = 0, PREC_MIXED = 1, PREC_DOUBLE = 2 };

template <int PREC> struct Prec;
template <> struct Prec<PREC_SINGLE> { using Real = float;  using Mixed = float;  static constexpr bool kSplitPos = false; };
template <> struct Prec<PREC_MIXED>  { using Real = float;  using Mixed = double; static constexpr bool kSplitPos = true;  };
template <> struct Prec<PREC_DOUBLE> { using Real = double; using Mixed = double; static constexpr bool kSplitPos = false; };

// Device-resident counters of one slab.
struct DevCounters {
    unsigned long long step;   // steps completed; the Philox counter of the next step
    unsigned int done;         // CTA ticket of the running step kernel (wraps to 0)
    unsigned int keTicket;     // CTA ticket of the kinetic-energy reduction
};

// Everything a step kernel needs, passed by value (__grid_constant__) or, for the batched
// ensemble launch, read from a table in global memory.
struct SlabDev {
    void* posq;                  // real4[P]
    void* corr;                  // real4[P]   (mixed precision) or nullptr
    void* velm;                  // mixed4[P]
    long long* force;            // int64[3*P]
    void* posDelta;              // mixed4[P]  (split path) or nullptr
    const int* atomIndex;        // slot -> original index, nullptr = identity
    const int* invAtomOrder;     // original index -> slot
    const void* imageCharge;     // real[(numCells-1)*R] snapshot, or nullptr = re-read posq.w
    DevCounters* counters;
    int numAtoms, numReal, padded, numCells;
    double zshift;               // zmax (reference mirror) or 0 (exact negation)
    double invDt;                // 1.0 / (double)(mixed)dt   (Langevin.cu:36)
    double vscale, fscaleFixed, noisescale, dt;  // already rounded to `mixed`; fscaleFixed = fscale/2^32
    unsigned long long seed, globalAtomOffset;
    unsigned int zeroVelm, zeroForce;
    int tileBase;                // batched launch: first tile of this slab in the global tile space
    int pad_;
};

// ---- small vector helpers ----------------------------------------------------------------------

template <typename T> struct Vec4 { T x, y, z, w; };

__device__ __forceinline__ Vec4<float> load4(const float* base, size_t i) {
    const float4 t = reinterpret_cast<const float4*>(base)[i];
    return {t.x, t.y, t.z, t.w};
}
__device__ __forceinline__ Vec4<double> load4(const double* base, size_t i) {
    const double2* p = reinterpret_cast<const double2*>(base) + 2 * i;
    const double2 a = p[0], b = p[1];
    return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void store4(float* base, size_t i, const Vec4<float>& v) {
    reinterpret_cast<float4*>(base)[i] = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void store4(double* base, size_t i, const Vec4<double>& v) {
    double2* p = reinterpret_cast<double2*>(base) + 2 * i;
    p[0] = make_double2(v.x, v.y);
    p[1] = make_double2(v.z, v.w);
}
// streaming (evict-first) 8-byte access for the fixed-point force planes: read once, dead after
__device__ __forceinline__ long long load_force(const long long* p) { return __ldcs(p); }

__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  fma_rn(float a, float b, float c)    { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float  sqrt_rn(float a)  { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float  from_fixed(long long f, float)  { return __ll2float_rn(f); }
__device__ __forceinline__ double from_fixed(long long f, double) { return __ll2double_rn(f); }

// ---- programmatic dependent launch -------------------------------------------------------------
// griddepcontrol.wait blocks until the preceding kernel in the stream has completed and its memory
// is visible; launch_dependents lets the next kernel's CTAs start occupying free SM slots.  Both
// are no-ops when the launch did not opt in.
__device__ __forceinline__ void pdl_wait()   { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- counter-based Gaussian stream -------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  One block per (atom, step):
//   counter = (atom_lo, atom_hi, step_lo, step_hi), key = (seed_lo, seed_hi)
// atom = GLOBAL ORIGINAL atom index, so the stream is independent of slot order and of the
// multi-GPU partition.  24-bit-exact uniforms u = ((w >> 9) + 0.5) * 2^-23 in (0,1), Box-Muller.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float uniform23(unsigned int w) {
    return ((float) (w >> 9) + 0.5f) * 1.1920928955078125e-07f;  // exact: 24 significant bits
}

template <bool WANT_W>
__device__ __forceinline__ float4 gaussian4(unsigned long long seed, unsigned long long atom,
                                            unsigned long long step) {
    const uint4 r = philox4x32_10(
        make_uint4((unsigned int) atom, (unsigned int) (atom >> 32), (unsigned int) step,
                   (unsigned int) (step >> 32)),
        make_uint2((unsigned int) seed, (unsigned int) (seed >> 32)));
    const float r1 = __fsqrt_rn(-2.0f * logf(uniform23(r.x)));
    const float r2 = __fsqrt_rn(-2.0f * logf(uniform23(r.z)));
    float s1, c1, s2, c2;
    sincospif(2.0f * uniform23(r.y), &s1, &c1);
    if (WANT_W) {
        sincospif(2.0f * uniform23(r.w), &s2, &c2);
    } else {
        c2 = cospif(2.0f * uniform23(r.w));
        s2 = 0.0f;
    }
    return make_float4(r1 * c1, r1 * s1, r2 * c2, r2 * s2);
}

// ---- per-atom arithmetic (canonical order) ------------------------------------------

template <typename M> struct StepScalars {
    M vscale, fs, noisescale, dt;
    double invDt, zshift;
};

template <typename M>
__device__ __forceinline__ StepScalars<M> scalars_of(const SlabDev& d) {
    return {(M) d.vscale, (M) d.fscaleFixed, (M) d.noisescale, (M) d.dt, d.invDt, d.zshift};
}

// Langevin.cu:17-23.  v <- fma(ns, xi, fma(fw, F, vscale*v)); delta <- dt*v.
template <typename M>
__device__ __forceinline__ void part1_update(Vec4<M>& v, Vec4<M>& delta, long long fx, long long fy,
                                             long long fz, float xix, float xiy, float xiz,
                                             const StepScalars<M>& s) {
    const M w = v.w;
    const M sqrtInvMass = sqrt_rn(w);
    const M fw = mul_rn(s.fs, w);
    const M ns = mul_rn(s.noisescale, sqrtInvMass);
    v.x = fma_rn(ns, (M) xix, fma_rn(fw, from_fixed(fx, M()), mul_rn(s.vscale, v.x)));
    v.y = fma_rn(ns, (M) xiy, fma_rn(fw, from_fixed(fy, M()), mul_rn(s.vscale, v.y)));
    v.z = fma_rn(ns, (M) xiz, fma_rn(fw, from_fixed(fz, M()), mul_rn(s.vscale, v.z)));
    delta.x = mul_rn(s.dt, v.x);
    delta.y = mul_rn(s.dt, v.y);
    delta.z = mul_rn(s.dt, v.z);
    delta.w = 0;
}

// Langevin.cu:43-71 (__CUDA_ARCH__ >= 130 branch).  p/c are posq and posqCorrection.
template <int PREC>
__device__ __forceinline__ void part2_update(Vec4<typename Prec<PREC>::Real>& p,
                                             Vec4<typename Prec<PREC>::Real>& c,
                                             Vec4<typename Prec<PREC>::Mixed>& v,
                                             const Vec4<typename Prec<PREC>::Mixed>& delta,
                                             double invDt) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    if (Prec<PREC>::kSplitPos) {
        const double px = __dadd_rn(__dadd_rn((double) p.x, (double) c.x), (double) delta.x);
        const double py = __dadd_rn(__dadd_rn((double) p.y, (double) c.y), (double) delta.y);
        const double pz = __dadd_rn(__dadd_rn((double) p.z, (double) c.z), (double) delta.z);
        p.x = (Real) px; p.y = (Real) py; p.z = (Real) pz;
        c.x = (Real) __dadd_rn(px, -(double) p.x);
        c.y = (Real) __dadd_rn(py, -(double) p.y);
        c.z = (Real) __dadd_rn(pz, -(double) p.z);
        c.w = 0;
    } else {
        p.x = (Real) add_rn((Mixed) p.x, delta.x);
        p.y = (Real) add_rn((Mixed) p.y, delta.y);
        p.z = (Real) add_rn((Mixed) p.z, delta.z);
    }
    v.x = (Mixed) __dmul_rn(invDt, (double) delta.x);
    v.y = (Mixed) __dmul_rn(invDt, (double) delta.y);
    v.z = (Mixed) __dmul_rn(invDt, (double) delta.z);
}

// Langevin.cu:144-160: running image position.  Holds x, y, z in the precision the reference
// iterates in (real for single/double, double for mixed) and emits posq/posqCorrection records.
template <int PREC> struct ImageCursor {
    using Real = typename Prec<PREC>::Real;
    using Iter = typename Prec<PREC>::Mixed;  // float | double | double == real for single/double
    Iter x, y, z;
    __device__ __forceinline__ ImageCursor(const Vec4<Real>& p, const Vec4<Real>& c) {
        if (Prec<PREC>::kSplitPos) {
            x = (Iter) __dadd_rn((double) p.x, (double) c.x);
            y = (Iter) __dadd_rn((double) p.y, (double) c.y);
            z = (Iter) __dadd_rn((double) p.z, (double) c.z);
        } else {
            x = (Iter) p.x; y = (Iter) p.y; z = (Iter) p.z;
        }
    }
    // z <- -z + zshift*(2i), evaluated in double with one rounding (Langevin.cu:152)
    __device__ __forceinline__ void advance(int i, double zshift) {
        z = (Iter) __fma_rn(zshift, (double) (2 * i), -(double) z);
    }
    __device__ __forceinline__ Vec4<Real> posq(Real q) const { return {(Real) x, (Real) y, (Real) z, q}; }
    __device__ __forceinline__ Vec4<Real> corr() const {
        return {(Real) __dadd_rn((double) x, -(double) (Real) x),
                (Real) __dadd_rn((double) y, -(double) (Real) y),
                (Real) __dadd_rn((double) z, -(double) (Real) z), (Real) 0};
    }
};

// Every step kernel takes a ticket when a CTA is done; the last CTA advances the device-resident
// step counter (so that CUDA-graph replays advance the Philox stream without host involvement).
// All threads of the CTA have read counters->step before the __syncthreads() that precedes their
// first tile, hence before this CTA's ticket.
__device__ __forceinline__ void finish_step(DevCounters* counters) {
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicInc(&counters->done, gridDim.x - 1);
        if (ticket == gridDim.x - 1) {
            counters->step += 1;
            __threadfence();
        }
    }
}

// ---- the fused step, identity slot order -------------------------------------------------------
// Thread <-> real/image pair.  For numCells == 2 a pair is slot i (real) and slot R + i (image):
// both streams are contiguous, so every access below is a fully coalesced 16-byte (posq, velm) or
// 8-byte (force plane) vector per lane.  All loads of a tile (ILP pairs per thread) are issued
// before the first use so that each thread keeps ILP*(60..100) bytes in flight.
template <int PREC, int BLOCK, int ILP>
__device__ __forceinline__ void fused_tile(const SlabDev& d, const int base, const unsigned long long step,
                                           const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);

    Vec4<Mixed> v[ILP];
    Vec4<Real> p[ILP], c[ILP];
    long long fx[ILP], fy[ILP], fz[ILP];
    Real q1[ILP];
    float4 xi[ILP];
    bool valid[ILP];

    // image charge of cell 1: side array (stride 1) or posq[R+i].w (stride 4)
    const Real* qbase = d.imageCharge ? static_cast<const Real*>(d.imageCharge) : posq + 4 * (size_t) R + 3;
    const size_t qstride = d.imageCharge ? 1 : 4;

#pragma unroll
    for (int j = 0; j < ILP; j++) {
        const int i = base + j * BLOCK + (int) threadIdx.x;
        valid[j] = i < R;
        if (valid[j]) {
            v[j] = load4(velm, (size_t) i);
            p[j] = load4(posq, (size_t) i);
            if (kSplit) c[j] = load4(corrA, (size_t) i);
            fx[j] = load_force(force + i);
            fy[j] = load_force(force + P + i);
            fz[j] = load_force(force + 2 * P + i);
            q1[j] = qbase[qstride * (size_t) i];
            if (random) xi[j] = __ldcs(random + (size_t) randomIndex + i);
        }
    }

#pragma unroll
    for (int j = 0; j < ILP; j++) {
        if (!valid[j]) continue;
        const int i = base + j * BLOCK + (int) threadIdx.x;
        if (v[j].w != 0) {
            if (!random) {
                xi[j] = (s.noisescale != 0)
                            ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) i, step)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            Vec4<Mixed> delta;
            part1_update(v[j], delta, fx[j], fy[j], fz[j], xi[j].x, xi[j].y, xi[j].z, s);
            part2_update<PREC>(p[j], c[j], v[j], delta, s.invDt);
            store4(velm, (size_t) i, v[j]);
            store4(posq, (size_t) i, p[j]);
            if (kSplit) store4(corrA, (size_t) i, c[j]);
        }
        // image rewrite (Langevin.cu:143: neutral atoms are skipped)
        if (p[j].w != -p[j].w) {
            ImageCursor<PREC> cur(p[j], c[j]);
            cur.advance(1, s.zshift);
            store4(posq, (size_t) i + R, cur.posq(q1[j]));
            if (kSplit) store4(corrA, (size_t) i + R, cur.corr());
            for (int cell = 2; cell < d.numCells; cell++) {
                const size_t slot = (size_t) i + (size_t) R * cell;
                const Real qc = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[(size_t) (cell - 1) * R + i]
                                              : posq[4 * slot + 3];
                cur.advance(cell, s.zshift);
                store4(posq, slot, cur.posq(qc));
                if (kSplit) store4(corrA, slot, cur.corr());
            }
        }
        // image invariants: velm = 0 (velocity and inverse mass), force = 0
        for (int cell = 1; cell < d.numCells; cell++) {
            const size_t slot = (size_t) i + (size_t) R * cell;
            if (d.zeroVelm) store4(velm, slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(force + slot, 0LL);
                __stcs(force + P + slot, 0LL);
                __stcs(force + 2 * P + slot, 0LL);
            }
        }
    }
}

template <int PREC, int BLOCK, int ILP>
__global__ void __launch_bounds__(BLOCK)
constv_fused_step_kernel(const __grid_constant__ SlabDev d, const float4* __restrict__ random,
                         const unsigned int randomIndex) {
    pdl_wait();
    const unsigned long long step = d.counters->step;
    __syncthreads();
    pdl_launch();
    constexpr int kTile = BLOCK * ILP;
    for (long long base = (long long) blockIdx.x * kTile; base < d.numReal; base += (long long) gridDim.x * kTile)
        fused_tile<PREC, BLOCK, ILP>(d, (int) base, step, random, randomIndex);
    finish_step(d.counters);
}

// Ensemble launch: `table` holds one SlabDev per slab, tileBase = exclusive prefix of the per-slab
// tile counts.  One grid covers every slab; blocks find their slab by binary search.
template <int PREC, int BLOCK, int ILP>
__global__ void __launch_bounds__(BLOCK)
constv_fused_step_batched_kernel(const SlabDev* __restrict__ table, const int numSlabs, const int totalTiles) {
    pdl_wait();
    __shared__ SlabDev d;
    __shared__ unsigned long long stepOfSlab;
    constexpr int kTile = BLOCK * ILP;
    int current = -1;
    for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
        int lo = 0, hi = numSlabs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid].tileBase <= tile) lo = mid; else hi = mid - 1;
        }
        if (lo != current) {
            __syncthreads();
            if (threadIdx.x == 0) {
                d = table[lo];
                stepOfSlab = table[lo].counters->step;
            }
            __syncthreads();
            current = lo;
        }
        fused_tile<PREC, BLOCK, ILP>(d, (tile - d.tileBase) * kTile, stepOfSlab, nullptr, 0u);
    }
    // every slab's counter advances once per launch: slab k is closed by the CTAs' shared ticket
    // on slab 0's counters; the last CTA bumps all step counters.
    __syncthreads();
    if (threadIdx.x == 0) {
        DevCounters* c0 = table[0].counters;
        const unsigned int ticket = atomicInc(&c0->done, gridDim.x - 1);
        if (ticket == gridDim.x - 1) {
            for (int k = 0; k < numSlabs; k++) table[k].counters->step += 1;
            __threadfence();
        }
    }
}

// ---- the fused step for reordered slots --------------------------------------------------------
// Behind OpenMM the slots are periodically re-sorted (cu.reorderAtoms(), Kernels.cpp:172); slot s
// holds original atom atomIndex[s].  Thread <-> slot: a real atom updates itself (coalesced) and
// scatters its images through invAtomOrder; an image slot zeroes itself (coalesced).
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_fused_step_indexed_kernel(const __grid_constant__ SlabDev d, const float4* __restrict__ random,
                                 const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    pdl_wait();
    const unsigned long long step = d.counters->step;
    __syncthreads();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    for (int slot = blockIdx.x * BLOCK + threadIdx.x; slot < d.numAtoms; slot += gridDim.x * BLOCK) {
        const int a = d.atomIndex[slot];
        if (a >= R) {
            if (d.zeroVelm) store4(velm, (size_t) slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(force + slot, 0LL);
                __stcs(force + P + slot, 0LL);
                __stcs(force + 2 * P + slot, 0LL);
            }
            continue;
        }
        Vec4<Mixed> v = load4(velm, (size_t) slot);
        Vec4<Real> p = load4(posq, (size_t) slot);
        Vec4<Real> c{};
        if (kSplit) c = load4(corrA, (size_t) slot);
        if (v.w != 0) {
            const long long fx = load_force(force + slot);
            const long long fy = load_force(force + P + slot);
            const long long fz = load_force(force + 2 * P + slot);
            float4 xi;
            if (random) xi = __ldcs(random + (size_t) randomIndex + slot);
            else xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            Vec4<Mixed> delta;
            part1_update(v, delta, fx, fy, fz, xi.x, xi.y, xi.z, s);
            part2_update<PREC>(p, c, v, delta, s.invDt);
            store4(velm, (size_t) slot, v);
            store4(posq, (size_t) slot, p);
            if (kSplit) store4(corrA, (size_t) slot, c);
        }
        if (p.w != -p.w) {
            ImageCursor<PREC> cur(p, c);
            for (int cell = 1; cell < d.numCells; cell++) {
                const size_t t = (size_t) d.invAtomOrder[a + R * cell];
                const Real qc = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[(size_t) (cell - 1) * R + a]
                                              : posq[4 * t + 3];
                cur.advance(cell, s.zshift);
                store4(posq, t, cur.posq(qc));
                if (kSplit) store4(corrA, t, cur.corr());
            }
        }
    }
    finish_step(d.counters);
}

// ---- split path: part 1 | constraints (caller) | part 2 + mirror -------------------------------

template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_part1_kernel(const __grid_constant__ SlabDev d, const float4* __restrict__ random,
                    const unsigned int randomIndex) {
    using Mixed = typename Prec<PREC>::Mixed;
    pdl_wait();
    const unsigned long long step = d.counters->step;
    pdl_launch();
    const size_t P = (size_t) d.padded;
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    Mixed* __restrict__ posDelta = static_cast<Mixed*>(d.posDelta);
    const long long* __restrict__ force = d.force;
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    for (int slot = blockIdx.x * BLOCK + threadIdx.x; slot < d.numAtoms; slot += gridDim.x * BLOCK) {
        Vec4<Mixed> v = load4(velm, (size_t) slot);
        if (v.w != 0) {
            const long long fx = load_force(force + slot);
            const long long fy = load_force(force + P + slot);
            const long long fz = load_force(force + 2 * P + slot);
            float4 xi;
            if (random) {
                xi = __ldcs(random + (size_t) randomIndex + slot);
            } else if (s.noisescale != 0) {
                const int a = d.atomIndex ? d.atomIndex[slot] : slot;
                xi = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step);
            } else {
                xi = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            Vec4<Mixed> delta;
            part1_update(v, delta, fx, fy, fz, xi.x, xi.y, xi.z, s);
            store4(velm, (size_t) slot, v);
            store4(posDelta, (size_t) slot, delta);
        }
    }
}

// identity order: thread <-> pair, same access pattern as the fused kernel
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_part2_mirror_kernel(const __grid_constant__ SlabDev d) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    pdl_wait();
    __syncthreads();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    const Mixed* __restrict__ posDelta = static_cast<const Mixed*>(d.posDelta);
    long long* __restrict__ force = d.force;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < R; i += gridDim.x * BLOCK) {
        Vec4<Mixed> v = load4(velm, (size_t) i);
        Vec4<Real> p = load4(posq, (size_t) i);
        Vec4<Real> c{};
        if (kSplit) c = load4(corrA, (size_t) i);
        if (v.w != 0) {
            const Vec4<Mixed> delta = load4(posDelta, (size_t) i);
            part2_update<PREC>(p, c, v, delta, d.invDt);
            store4(velm, (size_t) i, v);
            store4(posq, (size_t) i, p);
            if (kSplit) store4(corrA, (size_t) i, c);
        }
        if (p.w != -p.w) {
            ImageCursor<PREC> cur(p, c);
            for (int cell = 1; cell < d.numCells; cell++) {
                const size_t slot = (size_t) i + (size_t) R * cell;
                const Real qc = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[(size_t) (cell - 1) * R + i]
                                              : posq[4 * slot + 3];
                cur.advance(cell, d.zshift);
                store4(posq, slot, cur.posq(qc));
                if (kSplit) store4(corrA, slot, cur.corr());
            }
        }
        for (int cell = 1; cell < d.numCells; cell++) {
            const size_t slot = (size_t) i + (size_t) R * cell;
            if (d.zeroVelm) store4(velm, slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(force + slot, 0LL);
                __stcs(force + P + slot, 0LL);
                __stcs(force + 2 * P + slot, 0LL);
            }
        }
    }
    finish_step(d.counters);
}

// reordered slots: part 2 per slot (+ zeroing of image slots) ...
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_part2_indexed_kernel(const __grid_constant__ SlabDev d) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const size_t P = (size_t) d.padded;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    const Mixed* __restrict__ posDelta = static_cast<const Mixed*>(d.posDelta);
    long long* __restrict__ force = d.force;
    for (int slot = blockIdx.x * BLOCK + threadIdx.x; slot < d.numAtoms; slot += gridDim.x * BLOCK) {
        Vec4<Mixed> v = load4(velm, (size_t) slot);
        if (v.w != 0) {
            Vec4<Real> p = load4(posq, (size_t) slot);
            Vec4<Real> c{};
            if (kSplit) c = load4(corrA, (size_t) slot);
            const Vec4<Mixed> delta = load4(posDelta, (size_t) slot);
            part2_update<PREC>(p, c, v, delta, d.invDt);
            store4(velm, (size_t) slot, v);
            store4(posq, (size_t) slot, p);
            if (kSplit) store4(corrA, (size_t) slot, c);
        }
        if (d.atomIndex[slot] >= d.numReal) {
            if (d.zeroVelm) store4(velm, (size_t) slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(force + slot, 0LL);
                __stcs(force + P + slot, 0LL);
                __stcs(force + 2 * P + slot, 0LL);
            }
        }
    }
}

// ... then the image rewrite per original real index, as Langevin.cu:138-164 does
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_mirror_indexed_kernel(const __grid_constant__ SlabDev d) {
    using Real = typename Prec<PREC>::Real;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const int R = d.numReal;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    for (int a = blockIdx.x * BLOCK + threadIdx.x; a < R; a += gridDim.x * BLOCK) {
        const size_t s0 = (size_t) d.invAtomOrder[a];
        const Vec4<Real> p = load4(posq, s0);
        if (p.w != -p.w) {
            Vec4<Real> c{};
            if (kSplit) c = load4(corrA, s0);
            ImageCursor<PREC> cur(p, c);
            for (int cell = 1; cell < d.numCells; cell++) {
                const size_t t = (size_t) d.invAtomOrder[a + R * cell];
                const Real qc = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[(size_t) (cell - 1) * R + a]
                                              : posq[4 * t + 3];
                cur.advance(cell, d.zshift);
                store4(posq, t, cur.posq(qc));
                if (kSplit) store4(corrA, t, cur.corr());
            }
        }
    }
    finish_step(d.counters);
}

// ---- small kernels -----------------------------------------------------------------------------

// Langevin.cu:170-177
__global__ void constv_inverse_order_kernel(const int numAtoms, const int* __restrict__ atomIndex,
                                            int* __restrict__ invAtomOrder) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < numAtoms; slot += gridDim.x * blockDim.x)
        invAtomOrder[atomIndex[slot]] = slot;
}

template <typename Real>
__global__ void constv_image_charge_kernel(const int numReal, const int numCells, const Real* __restrict__ posq,
                                           const int* __restrict__ invAtomOrder, Real* __restrict__ imageCharge) {
    const int n = numReal * (numCells - 1);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        imageCharge[k] = posq[4 * (size_t) invAtomOrder[numReal + k] + 3];
}

__global__ void constv_noise_kernel(const int numAtoms, const int* __restrict__ atomIndex,
                                    const unsigned long long seed, const unsigned long long globalAtomOffset,
                                    const unsigned long long step, float4* __restrict__ out) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < numAtoms; slot += gridDim.x * blockDim.x) {
        const int a = atomIndex ? atomIndex[slot] : slot;
        out[slot] = gaussian4<true>(seed, globalAtomOffset + (unsigned long long) a, step);
    }
}

// KE = 0.5 * sum_{w != 0} |v + (timeShift/2^32) F w|^2 / w, accumulated in double.
// Warp shuffle -> shared -> one partial per CTA -> last CTA (ticket) adds the partials in index
// order, so the result is deterministic for a given launch shape.
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_kinetic_energy_kernel(const __grid_constant__ SlabDev d, const double scaleD, double* __restrict__ partials,
                             double* __restrict__ out) {
    using Mixed = typename Prec<PREC>::Mixed;
    const size_t P = (size_t) d.padded;
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const Mixed scale = (Mixed) scaleD;
    double e = 0.0;
    for (int slot = blockIdx.x * BLOCK + threadIdx.x; slot < d.numAtoms; slot += gridDim.x * BLOCK) {
        const Vec4<Mixed> v = load4(velm, (size_t) slot);
        if (v.w != 0) {
            const Mixed fx = from_fixed(force[slot], Mixed());
            const Mixed fy = from_fixed(force[P + slot], Mixed());
            const Mixed fz = from_fixed(force[2 * P + slot], Mixed());
            const double vx = (double) fma_rn(mul_rn(scale, fx), v.w, v.x);
            const double vy = (double) fma_rn(mul_rn(scale, fy), v.w, v.y);
            const double vz = (double) fma_rn(mul_rn(scale, fz), v.w, v.z);
            e += (vx * vx + vy * vy + vz * vz) / (double) v.w;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    __shared__ double warpSum[BLOCK / 32];
    __shared__ bool isLast;
    if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < BLOCK / 32; w++) b += warpSum[w];
        partials[blockIdx.x] = b;
        __threadfence();
        isLast = atomicInc(&d.counters->keTicket, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (isLast && threadIdx.x < 32) {
        __threadfence();
        double t = 0.0;
        for (int k = threadIdx.x; k < (int) gridDim.x; k += 32) t += __ldcg(partials + k);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (threadIdx.x == 0) *out = 0.5 * t;
    }
}

}  // namespace constv
