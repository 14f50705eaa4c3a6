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
    // reordered slots, staged kernels: the same two facts per real SLOT (constv_image_by_slot_kernel), or nullptr
    const int* imageSlotBySlot;      // int[(numCells-1)*R]: slot of the cell's image of the atom in real slot s (-1: s holds no real atom)
    const void* imageChargeBySlot;   // real[(numCells-1)*R]: that image's charge
    DevCounters* counters;
    int numAtoms, numReal, padded, numCells;
    double zshift;               // zmax (reference mirror) or 0 (exact negation)
    double invDt;                // 1.0 / (double)(mixed)dt   (Langevin.cu:36)
    double vscale, fscaleFixed, noisescale, dt;  // already rounded to `mixed`; fscaleFixed = fscale/2^32
    float vscaleF, fscaleFixedF, noisescaleF, dtF;  // the same four as float (single precision: no per-thread F2F)
    unsigned long long seed, globalAtomOffset;
    unsigned int zeroVelm, zeroForce;
    int tileBase;                // batched launch: first tile of this slab in the global tile space
    int rangeLo, rangeHi;        // real atoms [rangeLo, rangeHi) this launch advances (a chain, see
                                 // constv_set_chains; the whole slab = [0, numReal)); `counters` is the chain's
    int zeroWhen;                // fused step: when the image zeroing stores are issued (see fused_tile)
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
// Cache-operator variants.  STREAM = evict-first (ld.global.cs / st.global.cs): between two
// integrator steps of a real simulation the force kernels sweep far more than L2 through the cache,
// so nothing the step touches survives until the next step; marking it evict-first keeps it from
// displacing other data and measured +2-3 % on the step's own bandwidth.
template <bool STREAM> __device__ __forceinline__ Vec4<float> load4x(const float* base, size_t i) {
    const float4* p = reinterpret_cast<const float4*>(base) + i;
    const float4 t = STREAM ? __ldcs(p) : *p;
    return {t.x, t.y, t.z, t.w};
}
template <bool STREAM> __device__ __forceinline__ Vec4<double> load4x(const double* base, size_t i) {
    const double2* p = reinterpret_cast<const double2*>(base) + 2 * i;
    const double2 a = STREAM ? __ldcs(p) : p[0], b = STREAM ? __ldcs(p + 1) : p[1];
    return {a.x, a.y, b.x, b.y};
}
template <bool STREAM> __device__ __forceinline__ void store4x(float* base, size_t i, const Vec4<float>& v) {
    float4* p = reinterpret_cast<float4*>(base) + i;
    const float4 t = make_float4(v.x, v.y, v.z, v.w);
    if (STREAM) __stcs(p, t); else *p = t;
}
template <bool STREAM> __device__ __forceinline__ void store4x(double* base, size_t i, const Vec4<double>& v) {
    double2* p = reinterpret_cast<double2*>(base) + 2 * i;
    if (STREAM) { __stcs(p, make_double2(v.x, v.y)); __stcs(p + 1, make_double2(v.z, v.w)); }
    else { p[0] = make_double2(v.x, v.y); p[1] = make_double2(v.z, v.w); }
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

// ---- asynchronous copies global -> shared (cp.async, SASS LDGSTS) --------------------------------------
// No staging registers: a thread can have all the records of its slots in flight at once whatever its register budget.
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

// u = ((w >> 9) + 0.5) * 2^-23: 24 significant bits, exact in float, strictly inside (0,1)
__device__ __forceinline__ float uniform23(unsigned int w) {
    return __fmaf_rn((float) (w >> 9), 1.1920928955078125e-07f, 5.9604644775390625e-08f);
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// -2 ln(u) for u in (0,1).  MUFU lg2 has a 2^-22 ABSOLUTE error on (0.5, 2), which would be a large
// relative error where ln(u) -> 0, so the top 1/32 of the range uses the series of -2 ln(1 - t),
// t = 1 - u (exact), truncated after t^5 (relative error < 1e-8 for t < 2^-5).
__device__ __forceinline__ float neg2ln(float u) {
    const float t = __fadd_rn(1.0f, -u);
    const float fast = __fmul_rn(-1.38629436111989f, lg2_approx(u));
    float ser = __fmaf_rn(t, 0.4f, 0.5f);
    ser = __fmaf_rn(t, ser, 0.666666686534881591796875f);
    ser = __fmaf_rn(t, ser, 1.0f);
    ser = __fmaf_rn(t, ser, 2.0f);
    ser = __fmul_rn(t, ser);
    return t < 0.03125f ? ser : fast;
}

// SFU approximations, .ftz forms: every argument below is a normal number (u >= 2^-24,
// -2 ln u >= 1.19e-7, |angle| < pi), so flushing denormals only removes the guard code.
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Box-Muller on the four Philox words of (seed; atom, step).  The angle is
// a = u * fl(2 pi) - fl(pi) in (-pi, pi), where MUFU sin/cos are most accurate.  The device
// evaluates log/sqrt/sin/cos with the SFU (absolute error of a sample below 4e-6); the oracle
// evaluates the same definition in double.
template <bool WANT_W>
__device__ __forceinline__ float4 gaussian4(unsigned long long seed, unsigned long long atom,
                                            unsigned long long step) {
    const uint4 r = philox4x32_10(
        make_uint4((unsigned int) atom, (unsigned int) (atom >> 32), (unsigned int) step,
                   (unsigned int) (step >> 32)),
        make_uint2((unsigned int) seed, (unsigned int) (seed >> 32)));
    const float r1 = sqrt_approx(neg2ln(uniform23(r.x)));
    const float r2 = sqrt_approx(neg2ln(uniform23(r.z)));
    const float a1 = __fmaf_rn(uniform23(r.y), 6.2831854820251464844f, -3.1415927410125732422f);
    const float a2 = __fmaf_rn(uniform23(r.w), 6.2831854820251464844f, -3.1415927410125732422f);
    return make_float4(__fmul_rn(r1, cos_approx(a1)), __fmul_rn(r1, sin_approx(a1)), __fmul_rn(r2, cos_approx(a2)),
                       WANT_W ? __fmul_rn(r2, sin_approx(a2)) : 0.0f);
}

// ---- per-atom arithmetic (canonical order) ------------------------------------------

template <typename M> struct StepScalars {
    M vscale, fs, noisescale, dt;
    double invDt, zshift;
};

template <typename M> __device__ __forceinline__ StepScalars<M> scalars_of(const SlabDev& d);
template <> __device__ __forceinline__ StepScalars<float> scalars_of<float>(const SlabDev& d) {
    return {d.vscaleF, d.fscaleFixedF, d.noisescaleF, d.dtF, d.invDt, d.zshift};
}
template <> __device__ __forceinline__ StepScalars<double> scalars_of<double>(const SlabDev& d) {
    return {d.vscale, d.fscaleFixed, d.noisescale, d.dt, d.invDt, d.zshift};
}

// Langevin.cu:17-23.  v <- fma(ns, xi, fma(vscale, v, fw*F)); delta <- dt*v.  This is the contraction
// nvcc applies to the reference's expression vscale*v + fscale*w*F + noisescale*sqrtInvMass*xi
// (SASS of the reference kernel: FMUL t = fw*F; FFMA(vscale, v, t); FFMA(ns, xi, .)), so the
// result equals the reference kernel's bit for bit (tests/test_gpu_reference.py).
// fx, fy, fz: the fixed-point force components already converted to `mixed` (from_fixed)
template <typename M>
__device__ __forceinline__ void part1_update_converted(Vec4<M>& v, Vec4<M>& delta, const M fx, const M fy, const M fz,
                                                       float xix, float xiy, float xiz, const StepScalars<M>& s) {
    const M w = v.w;
    const M sqrtInvMass = sqrt_rn(w);
    const M fw = mul_rn(s.fs, w);
    const M ns = mul_rn(s.noisescale, sqrtInvMass);
    v.x = fma_rn(ns, (M) xix, fma_rn(s.vscale, v.x, mul_rn(fw, fx)));
    v.y = fma_rn(ns, (M) xiy, fma_rn(s.vscale, v.y, mul_rn(fw, fy)));
    v.z = fma_rn(ns, (M) xiz, fma_rn(s.vscale, v.z, mul_rn(fw, fz)));
    delta.x = mul_rn(s.dt, v.x);
    delta.y = mul_rn(s.dt, v.y);
    delta.z = mul_rn(s.dt, v.z);
    delta.w = 0;
}
template <typename M>
__device__ __forceinline__ void part1_update(Vec4<M>& v, Vec4<M>& delta, long long fx, long long fy,
                                             long long fz, float xix, float xiy, float xiz,
                                             const StepScalars<M>& s) {
    part1_update_converted(v, delta, from_fixed(fx, M()), from_fixed(fy, M()), from_fixed(fz, M()), xix, xiy, xiz, s);
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

// The step counter lives in device memory (so that CUDA-graph replays advance the Philox stream
// without host involvement) and is advanced by the step kernel itself.  Protocol, per unit of work
// ("ticket holder": a tile, or a CTA of a persistent kernel):
//   1. thread 0 reads the counter and publishes it to the CTA through shared memory (StepBox);
//   2. after the barrier that makes it visible -- i.e. once the value has arrived -- thread 0 takes
//      one ticket (atomicInc wrapping at holders-1, so the ticket counter is back at 0 for the next
//      launch); the atomic's round trip overlaps the unit's arithmetic and stores;
//   3. when its work is done, the holder of the LAST ticket stores step+1.  Every other holder took
//      its ticket after it had read the counter, so nobody can observe the new value early, and
//      no CTA ever waits for an atomic or re-reads the counter before it exits.
struct StepBox {
    unsigned long long step;
};

__device__ __forceinline__ void publish_step(StepBox& box, const DevCounters* counters) {
    if (threadIdx.x == 0) box.step = __ldcg(&counters->step);
}
__device__ __forceinline__ unsigned int take_ticket(DevCounters* counters, unsigned int holders) {
    return threadIdx.x == 0 ? atomicInc(&counters->done, holders - 1) : 0u;
}
__device__ __forceinline__ void close_step(DevCounters* counters, unsigned int ticket, unsigned int holders,
                                           unsigned long long step) {
    if (threadIdx.x == 0 && ticket == holders - 1) counters->step = step + 1;
}

// ---- the fused step, identity slot order -------------------------------------------------------
// Thread <-> real/image pair.  For numCells == 2 a pair is slot i (real) and slot R + i (image):
// both streams are contiguous, so every global access is a fully coalesced 16-byte (posq, velm) or
// 8-byte (force plane) vector per lane.

// Everything one pair reads, in registers.
template <int PREC> struct PairIn {
    Vec4<typename Prec<PREC>::Mixed> v;          // velm[i]
    Vec4<typename Prec<PREC>::Real> p, c;        // posq[i], posqCorrection[i] (mixed only)
    long long fx, fy, fz;                        // force planes
    typename Prec<PREC>::Real q1;                // charge of the cell-1 image
};

// What one real atom leaves in HBM at the end of a step (identity slot order): its own record (when
// it moved), the rewritten images (charged atoms only, Langevin.cu:143) and the zeroed image
// velocities / forces.  i = slot of the real atom, q1 = charge of its cell-1 image.
// image invariants of real atom i (identity slot order): velm = 0 (velocity and inverse mass), force = 0.
// Depends on nothing the step reads, so a kernel may issue it before its loads have returned.
template <int PREC, bool STREAM>
__device__ __forceinline__ void emit_zero(const SlabDev& d, const int i) {
    using Mixed = typename Prec<PREC>::Mixed;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;
#pragma unroll 1
    for (int cell = 1; cell < d.numCells; cell++) {
        const size_t slot = (size_t) i + (size_t) R * cell;
        if (d.zeroVelm) store4x<STREAM>(velm, slot, Vec4<Mixed>{0, 0, 0, 0});
        if (d.zeroForce) {
            __stcs(force + slot, 0LL);
            __stcs(force + P + slot, 0LL);
            __stcs(force + 2 * P + slot, 0LL);
        }
    }
}

template <int PREC, bool STREAM, bool ZERO = true>
__device__ __forceinline__ void emit_atom(const SlabDev& d, const double zshift, const int i, const bool moved,
                                          const Vec4<typename Prec<PREC>::Mixed>& v, const Vec4<typename Prec<PREC>::Real>& p,
                                          const Vec4<typename Prec<PREC>::Real>& c, const typename Prec<PREC>::Real q1) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;
    if (moved) {
        store4x<STREAM>(velm, (size_t) i, v);
        store4x<STREAM>(posq, (size_t) i, p);
        if (kSplit) store4x<STREAM>(corrA, (size_t) i, c);
    }
    // image rewrite (Langevin.cu:143: neutral atoms are skipped)
    if (p.w != -p.w) {
        ImageCursor<PREC> cur(p, c);
        cur.advance(1, zshift);
        store4x<STREAM>(posq, (size_t) i + R, cur.posq(q1));
        if (kSplit) store4x<STREAM>(corrA, (size_t) i + R, cur.corr());
        if (d.numCells > 2) {
#pragma unroll 1
            for (int cell = 2; cell < d.numCells; cell++) {
                const size_t slot = (size_t) i + (size_t) R * cell;
                const Real qc = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[(size_t) (cell - 1) * R + i]
                                              : posq[4 * slot + 3];
                cur.advance(cell, zshift);
                store4x<STREAM>(posq, slot, cur.posq(qc));
                if (kSplit) store4x<STREAM>(corrA, slot, cur.corr());
            }
        }
    }
    if (ZERO) emit_zero<PREC, STREAM>(d, i);
}

// Part 1 + part 2 + image rewrite + image zeroing for pair i; all stores go straight to global.
// ZERO = false: the caller has already issued emit_zero for this pair.
template <int PREC, bool STREAM, bool ZERO = true>
__device__ __forceinline__ void process_pair(const SlabDev& d, const StepScalars<typename Prec<PREC>::Mixed>& s,
                                             const int i, PairIn<PREC>& in, const bool havePool, float4 xi,
                                             const unsigned long long step) {
    using Mixed = typename Prec<PREC>::Mixed;
    const bool moved = in.v.w != 0;
    if (moved) {
        if (!havePool) {
            xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) i, step)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        Vec4<Mixed> delta;
        part1_update(in.v, delta, in.fx, in.fy, in.fz, xi.x, xi.y, xi.z, s);
        part2_update<PREC>(in.p, in.c, in.v, delta, s.invDt);
    }
    emit_atom<PREC, STREAM, ZERO>(d, s.zshift, i, moved, in.v, in.p, in.c, in.q1);
}

// LDG variant: all loads of a tile (ILP pairs per thread) are issued before the first use so that
// each thread keeps ILP*(60..100) bytes in flight.
// `stepOf()` is called once, after the tile's loads have been issued: it returns the step index (the
// Philox counter) and is where the caller synchronises the CTA on the published counter.
// EARLY: everything that does not depend on the tile's loads is done while they are in flight -- the image
// zeroing stores are issued right behind the loads, and the Philox block of every pair (a function of seed,
// atom index and step only; ~130 of the ~270 instructions of a pair) is drawn as soon as the step counter is
// known, for massless atoms too (their draw is discarded).  What remains once the records arrive is part 1,
// part 2 and the dependent stores.  Same arithmetic, same bits; it shortens the serial load -> compute ->
// store path of a CTA, which is what a launch of only 1-2 waves of CTAs (1M atoms) is made of.
template <int PREC, int BLOCK, int ILP, bool STREAM, bool EARLY, typename StepFn>
__device__ __forceinline__ void fused_tile(const SlabDev& d, const int base, const int end, StepFn stepOf,
                                           const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);

    PairIn<PREC> in[ILP];
    float4 xi[ILP];
    bool valid[ILP];

    // image charge of cell 1: side array (stride 1) or posq[R+i].w (stride 4)
    const Real* qbase = d.imageCharge ? static_cast<const Real*>(d.imageCharge) : posq + 4 * (size_t) R + 3;
    const size_t qstride = d.imageCharge ? 1 : 4;

#pragma unroll
    for (int j = 0; j < ILP; j++) {
        const int i = base + j * BLOCK + (int) threadIdx.x;
        valid[j] = i < end;
        xi[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid[j]) {
            in[j].v = load4x<STREAM>(velm, (size_t) i);
            in[j].p = load4x<STREAM>(posq, (size_t) i);
            if (kSplit) in[j].c = load4x<STREAM>(corrA, (size_t) i);
            in[j].fx = load_force(force + i);
            in[j].fy = load_force(force + P + i);
            in[j].fz = load_force(force + 2 * P + i);
            in[j].q1 = qbase[qstride * (size_t) i];
            if (random) xi[j] = __ldcs(random + (size_t) randomIndex + i);
        }
    }
    const int zeroWhen = d.zeroWhen & 3;  // 0 with the dependent stores, 1 right behind the loads (default), 2 after the barrier
    if (zeroWhen == 1) {
#pragma unroll
        for (int j = 0; j < ILP; j++)
            if (valid[j]) emit_zero<PREC, STREAM>(d, base + j * BLOCK + (int) threadIdx.x);
    }
    const unsigned long long step = stepOf();
    if (zeroWhen == 2) {
#pragma unroll
        for (int j = 0; j < ILP; j++)
            if (valid[j]) emit_zero<PREC, STREAM>(d, base + j * BLOCK + (int) threadIdx.x);
    }
    if (EARLY && random == nullptr && s.noisescale != 0) {
#pragma unroll
        for (int j = 0; j < ILP; j++)
            if (valid[j])
                xi[j] = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) (base + j * BLOCK + (int) threadIdx.x), step);
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        if (!valid[j]) continue;
        // with EARLY the noise is in xi[j] already (zero when the noise scale is zero: the reference's 0 * xi)
        process_pair<PREC, STREAM, false>(d, s, base + j * BLOCK + (int) threadIdx.x, in[j], EARLY || random != nullptr, xi[j], step);
        if (zeroWhen == 0) emit_zero<PREC, STREAM>(d, base + j * BLOCK + (int) threadIdx.x);
    }
}

template <int PREC, int BLOCK, int ILP, bool STREAM, bool EARLY = true>
__global__ void __launch_bounds__(BLOCK, (PREC == PREC_SINGLE && ILP == 1) ? 2048 / BLOCK : 1)  // single: 32 registers, 2048 threads per SM
constv_fused_step_kernel(const __grid_constant__ SlabDev d, const float4* __restrict__ random,
                         const unsigned int randomIndex) {
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    // One tile per CTA when the host launches a one-shot grid (the default: the hardware CTA
    // scheduler then keeps the memory access front sequential and balances the tail, which measured
    // 10-15 % faster than a persistent grid for this 11-stream pattern); grid-stride otherwise.
    constexpr int kTile = BLOCK * ILP;
    const int lo = d.rangeLo, hi = d.rangeHi;
    const unsigned int holders = (unsigned int) ((hi - lo + kTile - 1) / kTile);  // one ticket per tile
    for (long long base = lo + (long long) blockIdx.x * kTile; base < hi; base += (long long) gridDim.x * kTile) {
        unsigned int ticket = 0;
        unsigned long long step = 0;
        fused_tile<PREC, BLOCK, ILP, STREAM, EARLY>(d, (int) base, hi, [&]() {
            publish_step(box, d.counters);
            __syncthreads();
            step = box.step;
            ticket = take_ticket(d.counters, holders);
            return step;
        }, random, randomIndex);
        close_step(d.counters, ticket, holders, step);
        if (base + (long long) gridDim.x * kTile < hi) __syncthreads();  // box is rewritten by the next tile
    }
}

// ---- the fused step, TMA-pipelined (the B200 fast path) -----------------------------------------
// Persistent CTAs; each owns one contiguous range of pairs and walks it in chunks of BLOCK pairs.
// One elected thread streams every chunk's inputs (velm, posq, [posqCorrection], the three force
// planes, the image-charge side array) global -> shared with cp.async.bulk (TMA, 1-D) into a
// STAGES-deep ring guarded by mbarriers, so the bytes in flight per SM are set by the ring depth,
// not by how many warps are parked on a scoreboard.  All threads then read their pair from shared
// memory, do the arithmetic and store straight to global (stores are fire-and-forget).
// Requirements checked by the host: identity slot order, Philox noise, image-charge side array,
// padded_num_atoms even (16-byte aligned force planes).  Chunk starts are multiples of 4 pairs so
// every bulk copy is 16-byte aligned and sized.

__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
    unsigned int ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int PREC, int BLOCK> struct StageLayout {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    static constexpr int kVelm = 0;
    static constexpr int kPosq = kVelm + BLOCK * 4 * (int) sizeof(Mixed);
    static constexpr int kCorr = kPosq + BLOCK * 4 * (int) sizeof(Real);
    static constexpr int kFx = kCorr + (Prec<PREC>::kSplitPos ? BLOCK * 4 * (int) sizeof(Real) : 0);
    static constexpr int kFy = kFx + BLOCK * 8;
    static constexpr int kFz = kFy + BLOCK * 8;
    static constexpr int kQ = kFz + BLOCK * 8;
    static constexpr int kBytes = (kQ + BLOCK * (int) sizeof(Real) + 127) / 128 * 128;
};

template <int PREC, int BLOCK, int STAGES, bool STREAM>
__global__ void __launch_bounds__(BLOCK)
constv_fused_step_tma_kernel(const __grid_constant__ SlabDev d) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    using L = StageLayout<PREC, BLOCK>;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long fullBar[STAGES];

    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    // contiguous range of this CTA, boundaries on multiples of 4 pairs
    const int quads = (R + 3) >> 2;
    const int lo = min(R, (int) (((long long) quads * blockIdx.x / gridDim.x) << 2));
    const int hi = min(R, (int) (((long long) quads * (blockIdx.x + 1) / gridDim.x) << 2));
    const int numChunks = (hi - lo + BLOCK - 1) / BLOCK;

    __shared__ StepBox box;
    pdl_wait();
    publish_step(box, d.counters);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&fullBar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    pdl_launch();
    const unsigned long long step = box.step;
    const unsigned int ticket = take_ticket(d.counters, gridDim.x);  // one ticket per persistent CTA

    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const Real* __restrict__ charge = static_cast<const Real*>(d.imageCharge);

    auto issue = [&](int chunk) {
        const int stage = chunk % STAGES;
        const int first = lo + chunk * BLOCK;
        const int n = min(BLOCK, hi - first);
        const unsigned int n4 = (unsigned int) (n & ~3);  // whole quads go through TMA
        if (n4 == 0) return;
        const unsigned int bar = smem_u32(&fullBar[stage]);
        const unsigned int dst = smem_u32(ring + (size_t) stage * L::kBytes);
        const unsigned int bytes = n4 * (4 * sizeof(Mixed) + 4 * sizeof(Real) * (kSplit ? 2 : 1) + 24 + sizeof(Real));
        mbar_expect_tx(bar, bytes);
        bulk_g2s(dst + L::kVelm, velm + 4 * (size_t) first, n4 * 4 * sizeof(Mixed), bar);
        bulk_g2s(dst + L::kPosq, posq + 4 * (size_t) first, n4 * 4 * sizeof(Real), bar);
        if (kSplit) bulk_g2s(dst + L::kCorr, corrA + 4 * (size_t) first, n4 * 4 * sizeof(Real), bar);
        bulk_g2s(dst + L::kFx, force + first, n4 * 8, bar);
        bulk_g2s(dst + L::kFy, force + P + first, n4 * 8, bar);
        bulk_g2s(dst + L::kFz, force + 2 * P + first, n4 * 8, bar);
        bulk_g2s(dst + L::kQ, charge + first, n4 * sizeof(Real), bar);
    };

    if (threadIdx.x == 0) {
        for (int c = 0; c < STAGES && c < numChunks; c++) issue(c);
    }
    const StepScalars<Mixed> sc = scalars_of<Mixed>(d);
    for (int chunk = 0; chunk < numChunks; chunk++) {
        const int stage = chunk % STAGES;
        const int first = lo + chunk * BLOCK;
        const int n = min(BLOCK, hi - first);
        const int n4 = n & ~3;
        const int t = (int) threadIdx.x;
        if (n4 > 0) mbar_wait(smem_u32(&fullBar[stage]), (unsigned int) ((chunk / STAGES) & 1));
        if (t < n) {
            PairIn<PREC> in;
            const int i = first + t;
            if (t < n4) {
                const unsigned char* st = ring + (size_t) stage * L::kBytes;
                in.v = load4(reinterpret_cast<const Mixed*>(st + L::kVelm), (size_t) t);
                in.p = load4(reinterpret_cast<const Real*>(st + L::kPosq), (size_t) t);
                if (kSplit) in.c = load4(reinterpret_cast<const Real*>(st + L::kCorr), (size_t) t);
                in.fx = reinterpret_cast<const long long*>(st + L::kFx)[t];
                in.fy = reinterpret_cast<const long long*>(st + L::kFy)[t];
                in.fz = reinterpret_cast<const long long*>(st + L::kFz)[t];
                in.q1 = reinterpret_cast<const Real*>(st + L::kQ)[t];
            } else {  // ragged end of the slab (R % 4 pairs): plain loads
                in.v = load4(velm, (size_t) i);
                in.p = load4(posq, (size_t) i);
                if (kSplit) in.c = load4(corrA, (size_t) i);
                in.fx = force[i];
                in.fy = force[P + i];
                in.fz = force[2 * P + i];
                in.q1 = charge[i];
            }
            process_pair<PREC, STREAM>(d, sc, i, in, false, make_float4(0.f, 0.f, 0.f, 0.f), step);
        }
        __syncthreads();  // every thread has consumed its slice of this stage
        if (threadIdx.x == 0 && chunk + STAGES < numChunks) issue(chunk + STAGES);
    }
    close_step(d.counters, ticket, gridDim.x, step);
}

// Ensemble launch: one grid advances up to CAP independent slabs by one step each.  The per-slab
// descriptors travel in the kernel parameters (constant bank: no dependent global loads before a
// CTA can issue its data loads, nothing to upload per step, capturable in a CUDA graph);
// tileBase = exclusive prefix of the per-slab tile counts, blocks find their slab by binary search.
template <int CAP> struct BatchParams {
    int numSlabs, totalTiles;
    SlabDev slab[CAP];
};

template <int PREC, int BLOCK, int CAP>
__global__ void __launch_bounds__(BLOCK)
constv_fused_step_batched_kernel(const __grid_constant__ BatchParams<CAP> b) {
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    for (int tile = blockIdx.x; tile < b.totalTiles; tile += gridDim.x) {
        int lo = 0, hi = b.numSlabs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (b.slab[mid].tileBase <= tile) lo = mid; else hi = mid - 1;
        }
        const SlabDev& d = b.slab[lo];
        // one ticket per tile, on the tile's own slab: every slab's counter advances exactly once
        const unsigned int holders = (unsigned int) ((d.numReal + BLOCK - 1) / BLOCK);
        unsigned int ticket = 0;
        unsigned long long step = 0;
        fused_tile<PREC, BLOCK, 1, true, true>(d, (tile - d.tileBase) * BLOCK, d.numReal, [&]() {
            publish_step(box, d.counters);
            __syncthreads();
            step = box.step;
            ticket = take_ticket(d.counters, holders);
            return step;
        }, nullptr, 0u);
        close_step(d.counters, ticket, holders, step);
        if (tile + (int) gridDim.x < b.totalTiles) __syncthreads();
    }
}

// ---- the fused step for reordered slots --------------------------------------------------------
// Behind OpenMM the slots are periodically re-sorted (cu.reorderAtoms(), Kernels.cpp:172); slot s
// holds original atom atomIndex[s].  Thread i owns slot i and slots i + R*cell: a real atom is updated
// (coalesced) and scatters its images through invAtomOrder; an image slot is zeroed (coalesced).
template <int PREC, int BLOCK, int MINB = (PREC == PREC_SINGLE ? 0 : 4)>
__global__ void __launch_bounds__(BLOCK, MINB)  // mixed/double: 64 registers, 4 CTAs per SM; single: MINB = 8 caps it at 32 registers
constv_fused_step_indexed_kernel(const __grid_constant__ SlabDev d, const float4* __restrict__ random,
                                 const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);

    // Thread i < R owns slot i and slots i + R*cell (one-shot grid over R).  OpenMM's re-sorting (whole identical
    // molecules trade places) leaves the real atoms in the low slots and the images in the high ones, so a thread
    // normally updates one real atom and zeroes numCells-1 image slots -- no CTA is left with nothing but
    // zeroing to do -- but any slot may hold any atom and is treated by what atomIndex says it holds.
    // Everything the low slot needs is requested before the barrier that publishes the step counter.
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    const bool inRange = i < R;
    int a0 = R, a1 = R;
    Vec4<Mixed> v{0, 0, 0, 0};
    Vec4<Real> p{0, 0, 0, 0}, c{0, 0, 0, 0};
    long long fx = 0, fy = 0, fz = 0;
    // per-slot image tables (SlabDev::imageSlotBySlot), when the host has them: the low slot's cell-1 image slot and charge
    // are requested with everything else instead of after atomIndex has arrived
    const bool tables = d.imageSlotBySlot != nullptr;
    int t1Low = -1;
    Real q1Low = 0;
    if (inRange) {
        a0 = d.atomIndex[i];
        a1 = d.atomIndex[i + R];
        if (tables) {
            t1Low = d.imageSlotBySlot[i];
            q1Low = static_cast<const Real*>(d.imageChargeBySlot)[i];
        }
        v = load4x<true>(velm, (size_t) i);
        p = load4x<true>(posq, (size_t) i);
        if (kSplit) c = load4x<true>(corrA, (size_t) i);
        fx = load_force(force + i);
        fy = load_force(force + P + i);
        fz = load_force(force + 2 * P + i);
    }
    publish_step(box, d.counters);
    __syncthreads();
    const unsigned long long step = box.step;
    const unsigned int ticket = take_ticket(d.counters, gridDim.x);  // one ticket per CTA

    auto zeroSlot = [&](int slot) {
        if (d.zeroVelm) store4x<true>(velm, (size_t) slot, Vec4<Mixed>{0, 0, 0, 0});
        if (d.zeroForce) {
            __stcs(force + slot, 0LL);
            __stcs(force + P + slot, 0LL);
            __stcs(force + 2 * P + slot, 0LL);
        }
    };
    // haveXi: the caller drew this atom's Philox block already (while the record was in flight)
    auto updateReal = [&](int slot, int a, Vec4<Mixed> v, Vec4<Real> p, Vec4<Real> c, long long fx, long long fy, long long fz,
                          bool haveXi, float4 xi, bool low) {
        const bool charged = p.w != -p.w;
        size_t t1 = 0;
        Real q1 = 0;
        if (charged) {  // the cell-1 image's slot and charge: requested before the arithmetic
            if (low && tables) {
                t1 = (size_t) t1Low;
                q1 = q1Low;
            } else {
                t1 = (size_t) d.invAtomOrder[a + R];
                q1 = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[a] : posq[4 * t1 + 3];
            }
        }
        if (v.w != 0) {
            if (random) xi = __ldcs(random + (size_t) randomIndex + slot);
            else if (!haveXi) xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            Vec4<Mixed> delta;
            part1_update(v, delta, fx, fy, fz, xi.x, xi.y, xi.z, s);
            part2_update<PREC>(p, c, v, delta, s.invDt);
            store4x<true>(velm, (size_t) slot, v);
            store4x<true>(posq, (size_t) slot, p);
            if (kSplit) store4x<true>(corrA, (size_t) slot, c);
        }
        if (charged) {
            ImageCursor<PREC> cur(p, c);
            cur.advance(1, s.zshift);
            store4(posq, t1, cur.posq(q1));
            if (kSplit) store4(corrA, t1, cur.corr());
            for (int cell = 2; cell < d.numCells; cell++) {
                const size_t t = (size_t) d.invAtomOrder[a + R * cell];
                const Real qc = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[(size_t) (cell - 1) * R + a]
                                              : posq[4 * t + 3];
                cur.advance(cell, s.zshift);
                store4(posq, t, cur.posq(qc));
                if (kSplit) store4(corrA, t, cur.corr());
            }
        }
    };

    if (inRange) {
        // What needs the atom indices only is done first, while the low slot's record is still in flight (the
        // index loads were issued ahead of it): the zeroing of the image slots this thread owns and the Philox
        // block of the low slot's atom (drawn for a massless atom too, and discarded).
        if (a1 >= R) zeroSlot(i + R);
        float4 xi0 = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool drawn = a0 < R && random == nullptr;
        if (drawn && s.noisescale != 0) xi0 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a0, step);
        if (a0 >= R) zeroSlot(i);
        else updateReal(i, a0, v, p, c, fx, fy, fz, drawn, xi0, true);
#pragma unroll 1
        for (int cell = 1; cell < d.numCells; cell++) {
            const int slot = i + R * cell;
            const int a = cell == 1 ? a1 : d.atomIndex[slot];
            if (a >= R) {
                if (cell > 1) zeroSlot(slot);
            } else {  // a real atom in a high slot
                Vec4<Real> c2{0, 0, 0, 0};
                if (kSplit) c2 = load4x<true>(corrA, (size_t) slot);
                updateReal(slot, a, load4x<true>(velm, (size_t) slot), load4x<true>(posq, (size_t) slot), c2,
                           load_force(force + slot), load_force(force + P + slot), load_force(force + 2 * P + slot),
                           false, make_float4(0.f, 0.f, 0.f, 0.f), false);
            }
        }
    }
    close_step(d.counters, ticket, gridDim.x, step);
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
    // no thread of this launch reads the counter (part 1 did): one thread advances it
    if (blockIdx.x == 0 && threadIdx.x == 0) d.counters->step = __ldcg(&d.counters->step) + 1;
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
        if ((d.atomIndex ? d.atomIndex[slot] : slot) >= d.numReal) {
            if (d.zeroVelm) store4(velm, (size_t) slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(force + slot, 0LL);
                __stcs(force + P + slot, 0LL);
                __stcs(force + 2 * P + slot, 0LL);
            }
        }
    }
}

// ... then the image rewrite per original real index, as Langevin.cu:138-164 does.  ADVANCE = this launch
// closes a split step (advances the step counter); false = the image rewrite alone (constv_mirror: the pass
// the reference runs after integration.computeVirtualSites(), CudaConstVKernels.cpp:161-166).  With
// atomIndex == nullptr invAtomOrder is the identity the library initialised it to: same kernel, same bits.
template <int PREC, int BLOCK, bool ADVANCE>
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
    // no thread of this launch reads the counter (part 1 did): one thread advances it
    if (ADVANCE && blockIdx.x == 0 && threadIdx.x == 0) d.counters->step = __ldcg(&d.counters->step) + 1;
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

// Reordered slots: per real slot and image cell, where the image of the atom in that slot sits and what it is charged --
// what the staged kernels would otherwise gather through atomIndex -> invAtomOrder / imageCharge at the end of every tile.
// Rebuilt whenever the slot order or the charge snapshot changes.
template <typename Real>
__global__ void constv_image_by_slot_kernel(const int numReal, const int numCells, const int* __restrict__ atomIndex,
                                            const int* __restrict__ invAtomOrder, const Real* __restrict__ imageCharge,
                                            int* __restrict__ imageSlot, Real* __restrict__ imageChargeBySlot) {
    const int n = numReal * (numCells - 1);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int cell1 = k / numReal, slot = k - cell1 * numReal, a = atomIndex[slot];
        const bool real = a < numReal;
        imageSlot[k] = real ? invAtomOrder[a + numReal * (cell1 + 1)] : -1;
        imageChargeBySlot[k] = real ? imageCharge[(size_t) cell1 * numReal + a] : (Real) 0;
    }
}

// Host-force staging -> OpenMM's force planes (constv_step_host_submit).  FROM_FLOAT converts as
// OpenMM's force kernels do when they accumulate: (long long)(f * 0x100000000), truncating.
template <bool FROM_FLOAT>
__global__ void constv_stage_forces_kernel(const int numReal, const int padded, const void* __restrict__ staged,
                                           long long* __restrict__ force) {
    const int total = 3 * numReal;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int plane = k / numReal, i = k - plane * numReal;
        long long v;
        if (FROM_FLOAT) v = __float2ll_rz(__fmul_rn(__ldcs(static_cast<const float*>(staged) + k), 4294967296.0f));
        else v = __ldcs(static_cast<const long long*>(staged) + k);
        force[(size_t) plane * padded + i] = v;
    }
}

__global__ void constv_noise_kernel(const int numAtoms, const int* __restrict__ atomIndex,
                                    const unsigned long long seed, const unsigned long long globalAtomOffset,
                                    const unsigned long long step, float4* __restrict__ out) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < numAtoms; slot += gridDim.x * blockDim.x) {
        const int a = atomIndex ? atomIndex[slot] : slot;
        out[slot] = gaussian4<true>(seed, globalAtomOffset + (unsigned long long) a, step);
    }
}

// the raw Philox block of (seed; atom, step): what gaussian4 turns into four Gaussians
__global__ void constv_philox_words_kernel(const int numAtoms, const int* __restrict__ atomIndex,
                                           const unsigned long long seed, const unsigned long long globalAtomOffset,
                                           const unsigned long long step, uint4* __restrict__ out) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < numAtoms; slot += gridDim.x * blockDim.x) {
        const unsigned long long atom = globalAtomOffset + (unsigned long long) (atomIndex ? atomIndex[slot] : slot);
        out[slot] = philox4x32_10(make_uint4((unsigned int) atom, (unsigned int) (atom >> 32), (unsigned int) step,
                                             (unsigned int) (step >> 32)),
                                  make_uint2((unsigned int) seed, (unsigned int) (seed >> 32)));
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
