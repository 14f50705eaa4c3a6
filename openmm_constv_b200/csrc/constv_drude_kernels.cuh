// Device code of the B200-native ConstVDrudeLangevinIntegrator step (sm_100a).
//
// What it replaces in the reference (scychon/openmm_constV), "Drude.cu" =
// platforms/cuda/src/kernels/constVDrudeLangevin.cu, host sequence
// platforms/cuda/src/CudaConstVKernels.cpp:252-349:
//   integrateConstVDrudeLangevinPart1 (Drude.cu:5-66)    ordinary particles: the Langevin update;
//                                                        (Drude particle, parent) pairs: centre of mass
//                                                        thermalised at (T, friction), relative motion
//                                                        at (T_drude, friction_drude)
//   integrateConstVDrudeLangevinPart2 (Drude.cu:72-103)  x += delta, v = delta/dt over the real atoms
//   applyHardWallConstraints (Drude.cu:108-211)          the Drude displacement bounces off a hard wall
//   updateImageParticlePositions (Langevin.cu:138-164)   image rewrite (+ the explicit image zeroing)
// The reference launches four kernels that each stream the state through HBM; here ONE kernel does the
// whole constraint-free step (MODE_FUSED) and the constraint hand-off path is two (MODE_PART1, then
// MODE_PART2 = part 2 + hard wall + image rewrite + zeroing).
//
// Work decomposition: thread <-> slot of a real atom, so every thread's own record is a coalesced
// 8/16-byte vector access.  A per-slot role table (built on the host from the pair list) says whether the
// slot is an ordinary particle, the Drude particle of pair k (it then OWNS the pair: gathers the parent's
// record -- in practice the neighbouring 16 bytes -- computes both atoms and stores both) or a parent
// (nothing to do: its owner handles it).  One thread per pair, not two, because both atoms' new state
// depends on both atoms' old state: two threads would race on each other's records.
//
// The reference's Drude kernels use the particle indices of the DrudeForce directly as slots
// (normalParticles / pairParticles are never translated through the atom order; OpenMM only swaps whole
// identical molecules, so a slot keeps its role) -- only the image rewrite goes through invAtomOrder.
// Same here: roles are per slot; INDEXED adds the atomIndex / invAtomOrder gathers of the image rewrite.
//
// Canonical arithmetic: the contraction nvcc applies to the reference's expressions, written out with
// explicit fma / __f*_rn (file compiled with -fmad=false); restated by the CPU oracle (test infrastructure)
// and checked bit for bit against the reference's kernels on the GPU (tests/test_gpu_drude.py).
#pragma once

#include "constv_kernels.cuh"

namespace constv {

// role of a real slot: >= 0 owner of that pair (the Drude particle); -1 ordinary particle;
// <= -2 owned (the parent): role = DRUDE_ROLE_PASSIVE - (slot of its owner)
enum { DRUDE_ROLE_NORMAL = -1, DRUDE_ROLE_PASSIVE = -2 };
enum { DRUDE_FUSED = 0, DRUDE_PART1 = 1, DRUDE_PART2 = 2 };

struct DrudeDev {
    const int* role;     // int[numReal]: DRUDE_ROLE_NORMAL | pair index (slot is particles.x) | <= DRUDE_ROLE_PASSIVE (slot is particles.y)
    const int2* pairs;   // (Drude particle, parent) slots, CudaConstVKernels.cpp:222-228
    int numNormal, numPairs;
    double vscale, fscaleFixed, noisescale;        // the Drude (relative-motion) thermostat, rounded to `mixed`
    double maxDistance, hardwallScale;             // maxDrudeDistance, sqrt(kB T_drude)
    float vscaleF, fscaleFixedF, noisescaleF, maxDistanceF, hardwallScaleF;
    int pad_;
};

template <typename M> struct DrudeScalars { M vscale, fs, noisescale, maxDistance, hardwallScale; };
template <typename M> __device__ __forceinline__ DrudeScalars<M> drude_scalars_of(const DrudeDev& d);
template <> __device__ __forceinline__ DrudeScalars<float> drude_scalars_of<float>(const DrudeDev& d) {
    return {d.vscaleF, d.fscaleFixedF, d.noisescaleF, d.maxDistanceF, d.hardwallScaleF};
}
template <> __device__ __forceinline__ DrudeScalars<double> drude_scalars_of<double>(const DrudeDev& d) {
    return {d.vscale, d.fscaleFixed, d.noisescale, d.maxDistance, d.hardwallScale};
}

__device__ __forceinline__ float  rcp_rn(float a)  { return __frcp_rn(a); }
__device__ __forceinline__ double rcp_rn(double a) { return __drcp_rn(a); }
__device__ __forceinline__ float  div_rn(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float  abs_of(float a)  { return fabsf(a); }
__device__ __forceinline__ double abs_of(double a) { return fabs(a); }

template <typename T> __device__ __forceinline__ T& comp(Vec4<T>& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
template <typename T> __device__ __forceinline__ const T& comp(const Vec4<T>& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// Drude.cu:31-63.  f1/f2 = fixed-point forces of the Drude particle / the parent, r1/r2 = the pair's
// two Gaussian float4s.  Writes the new velocities (w untouched) and posDelta = dt * v (w = 0).
// f1, f2: the fixed-point force components already converted to `mixed` (from_fixed)
template <typename M>
__device__ __forceinline__ void drude_pair_part1_converted(Vec4<M>& v1, Vec4<M>& v2, Vec4<M>& d1, Vec4<M>& d2,
                                                           const M (&f1)[3], const M (&f2)[3],
                                                           const float4 r1, const float4 r2,
                                                           const StepScalars<M>& s, const DrudeScalars<M>& ds) {
    const M mass1 = rcp_rn(v1.w);
    const M mass2 = rcp_rn(v2.w);
    const M total = add_rn(mass1, mass2);
    const M invTotalMass = rcp_rn(total);
    const M invReducedMass = mul_rn(mul_rn(total, v1.w), v2.w);
    const M mass1fract = mul_rn(invTotalMass, mass1);
    const M mass2fract = mul_rn(invTotalMass, mass2);
    const M fcm = mul_rn(s.fs, invTotalMass);
    const M frel = mul_rn(ds.fs, invReducedMass);
    const M ncm = mul_rn(s.noisescale, sqrt_rn(invTotalMass));
    const M nrel = mul_rn(ds.noisescale, sqrt_rn(invReducedMass));
    const float xi1[3] = {r1.x, r1.y, r1.z}, xi2[3] = {r2.x, r2.y, r2.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const M a = f1[k], b = f2[k];
        M cm = fma_rn(comp(v1, k), mass1fract, mul_rn(comp(v2, k), mass2fract));
        M rel = add_rn(comp(v2, k), -comp(v1, k));
        const M cmForce = add_rn(a, b);
        const M relForce = fma_rn(mass1fract, b, -mul_rn(mass2fract, a));
        cm = fma_rn((M) xi1[k], ncm, fma_rn(cm, s.vscale, mul_rn(cmForce, fcm)));
        rel = fma_rn((M) xi2[k], nrel, fma_rn(rel, ds.vscale, mul_rn(relForce, frel)));
        comp(v1, k) = fma_rn(-mass2fract, rel, cm);
        comp(v2, k) = fma_rn(mass1fract, rel, cm);
        comp(d1, k) = mul_rn(s.dt, comp(v1, k));
        comp(d2, k) = mul_rn(s.dt, comp(v2, k));
    }
    d1.w = 0;
    d2.w = 0;
}
template <typename M>
__device__ __forceinline__ void drude_pair_part1(Vec4<M>& v1, Vec4<M>& v2, Vec4<M>& d1, Vec4<M>& d2,
                                                 const long long (&f1)[3], const long long (&f2)[3],
                                                 const float4 r1, const float4 r2,
                                                 const StepScalars<M>& s, const DrudeScalars<M>& ds) {
    const M a[3] = {from_fixed(f1[0], M()), from_fixed(f1[1], M()), from_fixed(f1[2], M())};
    const M b[3] = {from_fixed(f2[0], M()), from_fixed(f2[1], M()), from_fixed(f2[2], M())};
    drude_pair_part1_converted(v1, v2, d1, d2, a, b, r1, r2, s, ds);
}

// position of one particle as the `mixed` value the hard wall works on (Drude.cu:113-123) and back
// (:150-155): posq + posqCorrection in mixed precision, posq alone otherwise
template <int PREC>
__device__ __forceinline__ void drude_pos_load(const Vec4<typename Prec<PREC>::Real>& p, const Vec4<typename Prec<PREC>::Real>& c,
                                               typename Prec<PREC>::Mixed (&out)[3]) {
    using M = typename Prec<PREC>::Mixed;
#pragma unroll
    for (int k = 0; k < 3; k++)
        out[k] = Prec<PREC>::kSplitPos ? add_rn((M) comp(p, k), (M) comp(c, k)) : (M) comp(p, k);
}
template <int PREC>
__device__ __forceinline__ void drude_pos_store(Vec4<typename Prec<PREC>::Real>& p, Vec4<typename Prec<PREC>::Real>& c,
                                                const typename Prec<PREC>::Mixed (&in)[3]) {
    using M = typename Prec<PREC>::Mixed;
    using R = typename Prec<PREC>::Real;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        comp(p, k) = (R) in[k];
        if (Prec<PREC>::kSplitPos) comp(c, k) = (R) add_rn(in[k], -(M) comp(p, k));
    }
    if (Prec<PREC>::kSplitPos) c.w = 0;
}

// Drude.cu:108-211 for one pair (1 = Drude particle, 2 = parent).  Returns true when the pair bounced.
template <int PREC>
__device__ __forceinline__ bool drude_hardwall(Vec4<typename Prec<PREC>::Mixed>& v1, Vec4<typename Prec<PREC>::Mixed>& v2,
                                               Vec4<typename Prec<PREC>::Real>& p1, Vec4<typename Prec<PREC>::Real>& c1,
                                               Vec4<typename Prec<PREC>::Real>& p2, Vec4<typename Prec<PREC>::Real>& c2,
                                               const typename Prec<PREC>::Mixed stepSize,
                                               const typename Prec<PREC>::Mixed maxDistance,
                                               const typename Prec<PREC>::Mixed hardwallScale) {
    using M = typename Prec<PREC>::Mixed;
    M pos1[3], pos2[3], delta[3], b[3];
    drude_pos_load<PREC>(p1, c1, pos1);
    drude_pos_load<PREC>(p2, c2, pos2);
#pragma unroll
    for (int k = 0; k < 3; k++) delta[k] = add_rn(pos1[k], -pos2[k]);
    const M r = sqrt_rn(fma_rn(delta[2], delta[2], fma_rn(delta[0], delta[0], mul_rn(delta[1], delta[1]))));
    const M rInv = rcp_rn(r);
    if (!(mul_rn(rInv, maxDistance) < 1)) return false;
#pragma unroll
    for (int k = 0; k < 3; k++) b[k] = mul_rn(delta[k], rInv);
    const M mass1 = rcp_rn(v1.w);
    const M mass2 = rcp_rn(v2.w);
    const M deltaR = add_rn(r, -maxDistance);
    M deltaT = stepSize;
    M dotvr1 = fma_rn(b[2], v1.z, fma_rn(b[0], v1.x, mul_rn(b[1], v1.y)));
    M vp1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) vp1[k] = fma_rn(-b[k], dotvr1, comp(v1, k));
    if (v2.w == 0) {
        // the parent is massless: only the Drude particle moves (:141-165)
        if (dotvr1 != 0) deltaT = div_rn(deltaR, abs_of(dotvr1));
        if (deltaT > stepSize) deltaT = stepSize;
        dotvr1 = div_rn(mul_rn(-dotvr1, hardwallScale), mul_rn(abs_of(dotvr1), sqrt_rn(mass1)));
        const M dr = fma_rn(deltaT, dotvr1, -deltaR);
#pragma unroll
        for (int k = 0; k < 3; k++) pos1[k] = fma_rn(b[k], dr, pos1[k]);
        drude_pos_store<PREC>(p1, c1, pos1);
#pragma unroll
        for (int k = 0; k < 3; k++) comp(v1, k) = fma_rn(b[k], dotvr1, vp1[k]);
    } else {
        // both particles move (:166-208)
        const M invTotalMass = rcp_rn(add_rn(mass1, mass2));
        M dotvr2 = fma_rn(b[2], v2.z, fma_rn(b[1], v2.y, mul_rn(b[0], v2.x)));
        M vp2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) vp2[k] = fma_rn(-b[k], dotvr2, comp(v2, k));
        const M vbCMass = mul_rn(fma_rn(dotvr1, mass1, mul_rn(dotvr2, mass2)), invTotalMass);
        dotvr1 = add_rn(dotvr1, -vbCMass);
        dotvr2 = add_rn(dotvr2, -vbCMass);
        if (dotvr1 != dotvr2) deltaT = div_rn(deltaR, abs_of(add_rn(dotvr1, -dotvr2)));
        if (deltaT > stepSize) deltaT = stepSize;
        const M vBond = div_rn(hardwallScale, sqrt_rn(mass1));
        dotvr1 = div_rn(mul_rn(mul_rn(mul_rn(-dotvr1, vBond), mass2), invTotalMass), abs_of(dotvr1));
        dotvr2 = div_rn(mul_rn(mul_rn(mul_rn(-dotvr2, vBond), mass1), invTotalMass), abs_of(dotvr2));
        const M dr1 = fma_rn(deltaT, dotvr1, -mul_rn(mul_rn(deltaR, mass2), invTotalMass));
        const M dr2 = fma_rn(mul_rn(deltaR, mass1), invTotalMass, mul_rn(deltaT, dotvr2));
        dotvr1 = add_rn(dotvr1, vbCMass);
        dotvr2 = add_rn(dotvr2, vbCMass);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            pos1[k] = fma_rn(b[k], dr1, pos1[k]);
            pos2[k] = fma_rn(b[k], dr2, pos2[k]);
        }
        drude_pos_store<PREC>(p1, c1, pos1);
        drude_pos_store<PREC>(p2, c2, pos2);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            comp(v1, k) = fma_rn(b[k], dotvr1, vp1[k]);
            comp(v2, k) = fma_rn(b[k], dotvr2, vp2[k]);
        }
    }
    return true;
}

// One real atom's registers.
template <int PREC> struct DrudeAtom {
    Vec4<typename Prec<PREC>::Mixed> v, delta;
    Vec4<typename Prec<PREC>::Real> p, c;
    long long f[3];
    typename Prec<PREC>::Real q1;  // charge of the cell-1 image (identity order: fetched with the record)
};

template <int PREC, int MODE, bool INDEXED, bool STREAM>
__device__ __forceinline__ void drude_load(const SlabDev& d, const int slot, DrudeAtom<PREC>& a) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    const size_t P = (size_t) d.padded;
    a.v = load4x<STREAM>(static_cast<const Mixed*>(d.velm), (size_t) slot);
    a.c = Vec4<Real>{0, 0, 0, 0};
    a.delta = Vec4<Mixed>{0, 0, 0, 0};
    a.q1 = 0;
    if (MODE != DRUDE_PART1) {
        a.p = load4x<STREAM>(static_cast<const Real*>(d.posq), (size_t) slot);
        if (Prec<PREC>::kSplitPos) a.c = load4x<STREAM>(static_cast<const Real*>(d.corr), (size_t) slot);
        if (!INDEXED)
            a.q1 = d.imageCharge ? static_cast<const Real*>(d.imageCharge)[slot]
                                 : static_cast<const Real*>(d.posq)[4 * ((size_t) d.numReal + slot) + 3];
    }
    if (MODE != DRUDE_PART2) {
        a.f[0] = load_force(d.force + slot);
        a.f[1] = load_force(d.force + P + slot);
        a.f[2] = load_force(d.force + 2 * P + slot);
    } else {
        a.delta = load4x<STREAM>(static_cast<const Mixed*>(d.posDelta), (size_t) slot);
    }
}

// What one real atom leaves in HBM after part 2 (+ hard wall): its record, its rewritten images
// (charged atoms only, Langevin.cu:143) and, identity order only, its zeroed image slots (reordered
// slots zero themselves).  `a` = original index of the atom held by `slot`.
template <int PREC, bool INDEXED, bool STREAM>
__device__ __forceinline__ void drude_emit(const SlabDev& d, const int slot, const int a, const bool moved,
                                           const DrudeAtom<PREC>& r) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    if (!INDEXED) {
        emit_atom<PREC, STREAM>(d, d.zshift, slot, moved, r.v, r.p, r.c, r.q1);
        return;
    }
    const int R = d.numReal;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    if (moved) {
        store4(static_cast<Mixed*>(d.velm), (size_t) slot, r.v);
        store4(posq, (size_t) slot, r.p);
        if (kSplit) store4(corrA, (size_t) slot, r.c);
    }
    if (r.p.w != -r.p.w && a < R) {  // a < R: a neutral wall atom and its image may have traded slots; nothing else can
        ImageCursor<PREC> cur(r.p, r.c);
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

// The same for the staged kernels when the per-slot image tables exist (SlabDev::imageSlotBySlot): the cell-1 image's slot
// and charge came in with the tile (img1 < 0: the slot holds no real atom), further cells are read by slot -- no gather through
// atomIndex -> invAtomOrder / imageCharge at the end of the tile.
template <int PREC>
__device__ __forceinline__ void drude_emit_tables(const SlabDev& d, const int slot, const bool moved,
                                                  const Vec4<typename Prec<PREC>::Mixed>& v, const Vec4<typename Prec<PREC>::Real>& p,
                                                  const Vec4<typename Prec<PREC>::Real>& c, const int img1, const typename Prec<PREC>::Real q1) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const int R = d.numReal;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    if (moved) {
        store4(static_cast<Mixed*>(d.velm), (size_t) slot, v);
        store4(posq, (size_t) slot, p);
        if (kSplit) store4(corrA, (size_t) slot, c);
    }
    if (p.w != -p.w && img1 >= 0) {
        ImageCursor<PREC> cur(p, c);
        cur.advance(1, d.zshift);
        store4(posq, (size_t) img1, cur.posq(q1));
        if (kSplit) store4(corrA, (size_t) img1, cur.corr());
        for (int cell = 2; cell < d.numCells; cell++) {
            const size_t k = (size_t) (cell - 1) * R + slot;
            const size_t t = (size_t) d.imageSlotBySlot[k];
            const Real qc = static_cast<const Real*>(d.imageChargeBySlot)[k];
            cur.advance(cell, d.zshift);
            store4(posq, t, cur.posq(qc));
            if (kSplit) store4(corrA, t, cur.corr());
        }
    }
}

// The Drude step.  MODE_FUSED: the whole constraint-free step; MODE_PART1: velocities + posDelta;
// MODE_PART2: part 2 + hard wall + image rewrite + zeroing (after the caller's constraints).
// Gaussians: random != nullptr reads the caller's pool with the reference's addressing (ordinary
// particle `index`: random[randomIndex + index], Drude.cu:18; pair i: random[randomIndex +
// NUM_NORMAL_PARTICLES + 2i] and the next one, :30,46-47); otherwise the Philox stream, keyed by the
// original atom index: an ordinary particle uses its own block, a pair the Drude particle's block for
// the centre of mass and the parent's block for the relative motion.
template <int PREC, int BLOCK, int MODE, bool INDEXED, bool STREAM>
__global__ void __launch_bounds__(BLOCK)
constv_drude_step_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ DrudeDev dd,
                         const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    // reordered slots: image slots take part (they zero themselves), so the slot space is all atoms
    const int numSlots = (INDEXED && MODE != DRUDE_PART1) ? d.numAtoms : R;
    const unsigned int holders = (unsigned int) ((numSlots + BLOCK - 1) / BLOCK);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    const DrudeScalars<Mixed> ds = drude_scalars_of<Mixed>(dd);

    for (long long base = (long long) blockIdx.x * BLOCK; base < numSlots; base += (long long) gridDim.x * BLOCK) {
        const int slot = (int) base + (int) threadIdx.x;
        const bool inRange = slot < numSlots;
        // original index of the atom in this slot; real atoms keep slots [0, R) (roles are per slot)
        const int a = (INDEXED && inRange && d.atomIndex) ? d.atomIndex[slot] : slot;
        const bool isReal = inRange && slot < R;
        const int role = isReal ? dd.role[slot] : DRUDE_ROLE_PASSIVE;
        DrudeAtom<PREC> me, other;
        int2 pair = make_int2(slot, slot);
        if (role > DRUDE_ROLE_PASSIVE) drude_load<PREC, MODE, INDEXED, STREAM>(d, slot, me);
        if (role >= 0) {
            pair = dd.pairs[role];
            drude_load<PREC, MODE, INDEXED, STREAM>(d, pair.y, other);
        }
        float4 xi1 = make_float4(0.f, 0.f, 0.f, 0.f), xi2 = xi1;
        if (MODE != DRUDE_PART2 && random != nullptr) {
            if (role == DRUDE_ROLE_NORMAL) xi1 = __ldcs(random + (size_t) randomIndex + slot);
            if (role >= 0) {
                const size_t at = (size_t) randomIndex + (size_t) dd.numNormal + 2 * (size_t) role;
                xi1 = __ldcs(random + at);
                xi2 = __ldcs(random + at + 1);
            }
        }

        // step index (the Philox counter) and this tile's ticket
        unsigned long long step = 0;
        unsigned int ticket = 0;
        if (MODE == DRUDE_FUSED) {
            publish_step(box, d.counters);
            __syncthreads();
            step = box.step;
            ticket = take_ticket(d.counters, holders);
        } else if (MODE == DRUDE_PART1) {
            step = __ldcg(&d.counters->step);
        }

        if (INDEXED && MODE != DRUDE_PART1 && inRange && a >= R) {  // an image slot zeroes itself
            if (d.zeroVelm) store4(static_cast<Mixed*>(d.velm), (size_t) slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(d.force + slot, 0LL);
                __stcs(d.force + P + slot, 0LL);
                __stcs(d.force + 2 * P + slot, 0LL);
            }
        }

        if (role == DRUDE_ROLE_NORMAL) {
            const bool moved = me.v.w != 0;
            if (moved) {
                if (MODE != DRUDE_PART2) {
                    if (random == nullptr && s.noisescale != 0)
                        xi1 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step);
                    part1_update(me.v, me.delta, me.f[0], me.f[1], me.f[2], xi1.x, xi1.y, xi1.z, s);
                }
                if (MODE != DRUDE_PART1) part2_update<PREC>(me.p, me.c, me.v, me.delta, s.invDt);
            }
            if (MODE == DRUDE_PART1) {
                if (moved) {
                    store4(static_cast<Mixed*>(d.velm), (size_t) slot, me.v);
                    store4(static_cast<Mixed*>(d.posDelta), (size_t) slot, me.delta);
                }
            } else {
                drude_emit<PREC, INDEXED, STREAM>(d, slot, a, moved, me);
            }
        } else if (role >= 0) {
            const int a2 = (INDEXED && d.atomIndex) ? d.atomIndex[pair.y] : pair.y;
            if (MODE != DRUDE_PART2) {
                if (random == nullptr) {
                    xi1 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step);
                    xi2 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a2, step);
                }
                drude_pair_part1(me.v, other.v, me.delta, other.delta, me.f, other.f, xi1, xi2, s, ds);
            }
            if (MODE == DRUDE_PART1) {  // Drude.cu:60-63: unconditional
                store4(static_cast<Mixed*>(d.velm), (size_t) slot, me.v);
                store4(static_cast<Mixed*>(d.velm), (size_t) pair.y, other.v);
                store4(static_cast<Mixed*>(d.posDelta), (size_t) slot, me.delta);
                store4(static_cast<Mixed*>(d.posDelta), (size_t) pair.y, other.delta);
            } else {
                if (me.v.w != 0) part2_update<PREC>(me.p, me.c, me.v, me.delta, s.invDt);
                if (other.v.w != 0) part2_update<PREC>(other.p, other.c, other.v, other.delta, s.invDt);
                if (ds.maxDistance > 0)  // CudaConstVKernels.cpp:336
                    drude_hardwall<PREC>(me.v, other.v, me.p, me.c, other.p, other.c, s.dt, ds.maxDistance, ds.hardwallScale);
                // a pair's records are always written back: part 1 stores the velocities unconditionally,
                // and records that neither part 2 nor the wall touched are stored with the bits they had
                drude_emit<PREC, INDEXED, STREAM>(d, slot, a, true, me);
                drude_emit<PREC, INDEXED, STREAM>(d, pair.y, a2, true, other);
            }
        }

        if (MODE == DRUDE_FUSED) {
            close_step(d.counters, ticket, holders, step);
            if (base + (long long) gridDim.x * BLOCK < numSlots) __syncthreads();  // box is rewritten by the next tile
        }
    }
    // part 2 closes a split step: nobody in this launch reads the counter (part 1 did)
    if (MODE == DRUDE_PART2 && blockIdx.x == 0 && threadIdx.x == 0) d.counters->step = __ldcg(&d.counters->step) + 1;
}


// ---- the fused Drude step staged through shared memory (identity slot order) -----------------------------
// The kernel above runs thread <-> slot: in a warp the pair owners (~500 instructions: two Philox blocks, the
// centre-of-mass / relative-motion transform with its IEEE reciprocals and square roots, the hard wall), the
// ordinary particles (~200) and the owned parents (nothing) sit side by side, so every warp pays for both
// branches -- 12.5 M warp instructions per step at 1 M atoms, issue-bound at 22 us -- and an owner's gather of
// its parent fills a quarter of each sector.  Here a CTA
//   1. stages its 256-slot tile in shared memory, every thread loading ITS OWN slot (all roles: full
//      coalesced vectors);
//   2. compacts the work: a block-wide ballot scan lists the tile's pair owners first, then its ordinary
//      particles, so that thread t takes work item t and whole warps run one branch;
//   3. computes item t from the stage (a parent inside the tile is read from and written back to the stage;
//      a parent outside it -- pairs straddling a tile boundary, or far apart -- is handled by the owner in
//      global memory, and that parent's own thread, which finds its owner's slot in its role, stays silent);
//   4. after a barrier every thread emits its own slot (record, rewritten images, zeroed images), coalesced.
template <int PREC, int BLOCK> struct DrudeStage {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    Vec4<Mixed> v[BLOCK];
    Vec4<Real> p[BLOCK];
    Vec4<Real> c[Prec<PREC>::kSplitPos ? BLOCK : 1];
    long long f[3][BLOCK];
    int role[BLOCK];
    int orig[BLOCK];             // original index of the atom in the slot (= the slot in identity order)
    int list[BLOCK];             // work items: tile-local slot of each owner, then of each ordinary particle
    int warpOwners[BLOCK / 32], warpNormals[BLOCK / 32];
};

// INDEXED = reordered slots: same tiles and arithmetic (the real atoms keep slots [0, R) and a slot keeps its
// role); the Philox key is atomIndex[slot], images are written wherever invAtomOrder says, and the image slots
// are zeroed in slot order (every thread zeroes the slots R * cell above its own: coalesced, whatever atoms they hold).
template <int PREC, int BLOCK, bool INDEXED, bool STREAM>
__global__ void __launch_bounds__(BLOCK, PREC == PREC_SINGLE ? 0 : 5)
constv_drude_step_tiled_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ DrudeDev dd,
                               const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    constexpr int kWarps = BLOCK / 32;
    __shared__ __align__(16) DrudeStage<PREC, BLOCK> st;  // 16-byte vector accesses into its records
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    // a chain (constv_set_chains) advances tiles [rangeLo, rangeHi) of the tile space with its own counters;
    // the whole step = [0, realTiles)
    const int tileLo = d.rangeLo, tileHi = d.rangeHi;
    const unsigned int holders = (unsigned int) (tileHi - tileLo);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    const DrudeScalars<Mixed> ds = drude_scalars_of<Mixed>(dd);
    const int tid = (int) threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int tile = tileLo + (int) blockIdx.x; tile < tileHi; tile += (int) gridDim.x) {
        const long long base = (long long) tile * BLOCK;
        const int slot = (int) base + tid;
        const bool inRange = slot < R;
        const int role = inRange ? dd.role[slot] : DRUDE_ROLE_PASSIVE;
        typename Prec<PREC>::Real q1 = 0;
        // 1. stage in
        if (inRange) {
            DrudeAtom<PREC> me;
            drude_load<PREC, DRUDE_FUSED, INDEXED, STREAM>(d, slot, me);
            st.orig[tid] = (INDEXED && d.atomIndex) ? d.atomIndex[slot] : slot;
            st.v[tid] = me.v;
            st.p[tid] = me.p;
            if (kSplit) st.c[tid] = me.c;
            st.f[0][tid] = me.f[0];
            st.f[1][tid] = me.f[1];
            st.f[2][tid] = me.f[2];
            q1 = me.q1;
        }
        st.role[tid] = role;
        // 2. compaction: owners first, then ordinary particles
        const bool isOwner = role >= 0, isNormal = inRange && role == DRUDE_ROLE_NORMAL;
        const unsigned int ownerMask = __ballot_sync(0xffffffffu, isOwner), normalMask = __ballot_sync(0xffffffffu, isNormal);
        if (lane == 0) {
            st.warpOwners[warp] = __popc(ownerMask);
            st.warpNormals[warp] = __popc(normalMask);
        }
        publish_step(box, d.counters);
        __syncthreads();
        const unsigned long long step = box.step;
        const unsigned int ticket = take_ticket(d.counters, holders);
        int ownersBefore = 0, normalsBefore = 0, numOwners = 0, numNormals = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const int o = st.warpOwners[w], n = st.warpNormals[w];
            if (w < warp) { ownersBefore += o; normalsBefore += n; }
            numOwners += o;
            numNormals += n;
        }
        const unsigned int below = (1u << lane) - 1u;
        if (isOwner) st.list[ownersBefore + __popc(ownerMask & below)] = tid;
        if (isNormal) st.list[numOwners + normalsBefore + __popc(normalMask & below)] = tid;
        __syncthreads();

        // 3. compute work item `tid`
        if (tid < numOwners) {
            const int a = st.list[tid];                       // tile-local slot of the Drude particle
            const int pairIndex = st.role[a];
            const int2 pair = dd.pairs[pairIndex];
            const int b = pair.y - (int) base;                // tile-local slot of the parent, if inside
            const bool partnerInTile = b >= 0 && b < BLOCK;
            Vec4<Mixed> v1 = st.v[a], v2, delta1{0, 0, 0, 0}, delta2{0, 0, 0, 0};
            Vec4<Real> p1 = st.p[a], p2, c1 = kSplit ? st.c[a] : Vec4<Real>{0, 0, 0, 0}, c2{0, 0, 0, 0};
            long long f1[3] = {st.f[0][a], st.f[1][a], st.f[2][a]}, f2[3];
            Real q2 = 0;
            int orig2;
            if (partnerInTile) {
                v2 = st.v[b];
                p2 = st.p[b];
                if (kSplit) c2 = st.c[b];
                f2[0] = st.f[0][b]; f2[1] = st.f[1][b]; f2[2] = st.f[2][b];
                orig2 = st.orig[b];
            } else {
                DrudeAtom<PREC> far;
                drude_load<PREC, DRUDE_FUSED, INDEXED, STREAM>(d, pair.y, far);
                v2 = far.v; p2 = far.p; c2 = far.c; q2 = far.q1;
                f2[0] = far.f[0]; f2[1] = far.f[1]; f2[2] = far.f[2];
                orig2 = (INDEXED && d.atomIndex) ? d.atomIndex[pair.y] : pair.y;
            }
            float4 xi1, xi2;
            if (random != nullptr) {
                const size_t at = (size_t) randomIndex + (size_t) dd.numNormal + 2 * (size_t) pairIndex;
                xi1 = __ldcs(random + at);
                xi2 = __ldcs(random + at + 1);
            } else {
                xi1 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) st.orig[a], step);
                xi2 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) orig2, step);
            }
            drude_pair_part1(v1, v2, delta1, delta2, f1, f2, xi1, xi2, s, ds);
            if (v1.w != 0) part2_update<PREC>(p1, c1, v1, delta1, s.invDt);
            if (v2.w != 0) part2_update<PREC>(p2, c2, v2, delta2, s.invDt);
            if (ds.maxDistance > 0)
                drude_hardwall<PREC>(v1, v2, p1, c1, p2, c2, s.dt, ds.maxDistance, ds.hardwallScale);
            st.v[a] = v1;
            st.p[a] = p1;
            if (kSplit) st.c[a] = c1;
            if (partnerInTile) {
                st.v[b] = v2;
                st.p[b] = p2;
                if (kSplit) st.c[b] = c2;
            } else if (INDEXED) {
                DrudeAtom<PREC> rec;
                rec.v = v2; rec.p = p2; rec.c = c2;
                drude_emit<PREC, true, STREAM>(d, pair.y, orig2, true, rec);
            } else {
                emit_atom<PREC, STREAM>(d, d.zshift, pair.y, true, v2, p2, c2, q2);
            }
        } else if (tid < numOwners + numNormals) {
            const int a = st.list[tid];
            Vec4<Mixed> v = st.v[a];
            if (v.w != 0) {
                Vec4<Real> p = st.p[a], c = kSplit ? st.c[a] : Vec4<Real>{0, 0, 0, 0};
                Vec4<Mixed> delta;
                float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
                if (random != nullptr) xi = __ldcs(random + (size_t) randomIndex + (size_t) base + a);
                else if (s.noisescale != 0) xi = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) st.orig[a], step);
                part1_update(v, delta, st.f[0][a], st.f[1][a], st.f[2][a], xi.x, xi.y, xi.z, s);
                part2_update<PREC>(p, c, v, delta, s.invDt);
                st.v[a] = v;
                st.p[a] = p;
                if (kSplit) st.c[a] = c;
            }
        }
        __syncthreads();

        // 4. every thread emits its own slot; a parent whose owner sits in another tile was emitted by that owner
        bool mine = inRange;
        if (role <= DRUDE_ROLE_PASSIVE && inRange) {
            const int owner = DRUDE_ROLE_PASSIVE - role;
            mine = owner >= (int) base && owner < (int) base + BLOCK;
        }
        if (mine) {
            const Vec4<Mixed> v = st.v[tid];
            const Vec4<Real> p = st.p[tid];
            const Vec4<Real> c = kSplit ? st.c[tid] : Vec4<Real>{0, 0, 0, 0};
            // a pair's records are always written back (constVDrudeLangevin.cu:60-63 stores unconditionally)
            const bool moved = role != DRUDE_ROLE_NORMAL || v.w != 0;
            if (INDEXED) {
                DrudeAtom<PREC> rec;
                rec.v = v; rec.p = p; rec.c = c;
                drude_emit<PREC, true, STREAM>(d, slot, st.orig[tid], moved, rec);
                emit_zero<PREC, STREAM>(d, slot);  // the image slots, in slot order: coalesced whatever atoms they hold
            } else {
                emit_atom<PREC, STREAM>(d, d.zshift, slot, moved, v, p, c, q1);
            }
        }
        close_step(d.counters, ticket, holders, step);
        if (tile + (int) gridDim.x < tileHi) __syncthreads();  // stage and box are rewritten by the next tile
    }
}

// ---- the staged fused step, balanced ---------------------------------------------------------------------
// In constv_drude_step_tiled_kernel thread t takes work item t, so a 128-slot tile of 4-site molecules
// (32 pairs + 64 ordinary particles) is one warp of owners (~500 instructions), two warps of ordinary particles
// (~200) and one idle warp: the three barriers of the tile make everybody wait for the owners, and fewer than half of
// the resident warps compute at any time.  Here a CTA of BLOCK threads stages SPT * BLOCK slots (every thread loads
// and later emits SPT slots, coalesced as before) and its warps draw CHUNKS of 32 work items from a counter in
// shared memory -- owner chunks first, then ordinary ones -- until the tile is done: with SPT = 2 the same tile shape
// gives two owner chunks and four ordinary chunks over four warps, 500 | 500 | 400 | 400 instructions.  Same
// arithmetic per work item, same bits.  The stage differs too:
//   * the three fixed-point force components are converted to `mixed` once, when staged (from_fixed is the first
//     thing part 1 does with them), and travel with the image charge as ONE 4-vector: 16 bytes instead of 28 in
//     single precision, one shared-memory load per work item instead of four;
//   * records sit at position slot ^ ((slot >> 3) & 7): the work items of one kind are every 4th slot (O D H H), a
//     stride that put 4 lanes of every quarter-warp on the same banks; the XOR keeps the coalesced phases (8
//     consecutive slots stay in one aligned 128-byte window) conflict-free and spreads the strided ones;
//   * a parent staged in the same tile leaves its position at its owner's entry of `partner` (checked against the
//     parent's role, so a stale entry cannot pass): no dependent load of the pair list before the arithmetic.
template <int PREC, int BLOCK, int SPT, bool INDEXED> struct DrudeStageB {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    static constexpr int kTile = BLOCK * SPT;
    Vec4<Mixed> v[kTile];
    Vec4<Real> p[kTile];
    Vec4<Real> c[Prec<PREC>::kSplitPos ? kTile : 1];
    Vec4<Mixed> fq[kTile];       // force components as `mixed`, w = image charge of cell 1 (identity order)
    int role[kTile];
    int orig[INDEXED ? kTile : 1];
    int img[INDEXED ? kTile : 1];   // reordered slots: slot of the cell-1 image (per-slot tables), -1 = the slot holds no real atom
    unsigned short list[kTile];     // work items: tile-local slot of each owner, then of each ordinary particle
    unsigned short partner[kTile];  // [owner] = tile-local slot of its parent, when that parent is staged in this tile
    int warpOwners[SPT * (BLOCK / 32)], warpNormals[SPT * (BLOCK / 32)];
    int next;                    // next chunk of 32 work items
};
#ifdef CONSTV_DRUDE_NO_SWIZZLE  // experiment (session 68): the stage laid out linearly, as bulk copies would leave it
__device__ __forceinline__ int drude_swizzle(const int i) { return i; }
#else
__device__ __forceinline__ int drude_swizzle(const int i) { return i ^ ((i >> 3) & 7); }
#endif

// The zeroing stores of the image slots are issued in the stage-in phase, behind the loads (they depend on nothing):
// 13.4 -> 13.0 us per step at 1M atoms with six chains, 19.4 -> 18.7 us with one launch per step
// (profiles/r02_ab_staged_zero_timing.jsonl).
// Mixed precision (what the reference's example runs): 5 CTAs per SM at 96 registers measured 27.8 -> 24.5 us per step at 1M
// atoms with six chains against 4 CTAs at 120 registers (31.6 -> 32.9 us with one launch per step;
// profiles/r02_ab_staged_mixed_ctas_per_sm.jsonl).
// TABLES (reordered slots): the cell-1 image's slot and charge come in with the tile from the per-slot tables
// (SlabDev::imageSlotBySlot) instead of being gathered through atomIndex -> invAtomOrder / imageCharge at the end of the tile.
template <int PREC, int BLOCK, int SPT, bool INDEXED, bool STREAM, bool TABLES = false>
__global__ void __launch_bounds__(BLOCK, PREC == PREC_SINGLE ? 1024 / BLOCK : (PREC == PREC_MIXED ? 5 : 4))  // single: 64 registers, 32 warps per SM
constv_drude_step_balanced_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ DrudeDev dd,
                                  const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    constexpr int kWarps = BLOCK / 32;
    constexpr int kTile = BLOCK * SPT;
    static_assert(kTile <= 65536 && (kTile & 63) == 0, "tile-local slots are 16-bit; the swizzle permutes aligned groups of 64");
    __shared__ __align__(16) DrudeStageB<PREC, BLOCK, SPT, INDEXED> st;  // 16-byte vector accesses and cp.async into its records
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    const int tileLo = d.rangeLo, tileHi = d.rangeHi;
    const unsigned int holders = (unsigned int) (tileHi - tileLo);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    const DrudeScalars<Mixed> ds = drude_scalars_of<Mixed>(dd);
    const int tid = (int) threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int tile = tileLo + (int) blockIdx.x; tile < tileHi; tile += (int) gridDim.x) {
        const int base = tile * kTile;
        // 1. stage in: thread t takes slots base + t + i * BLOCK.  Every load of the thread is issued before the first
        // use: the velocity and position records as asynchronous copies straight into the stage (cp.async, no registers),
        // roles, forces and image charges through registers (the forces are converted on the way).
        unsigned int ownerMask[SPT], normalMask[SPT];
        {
            int role[SPT];
            long long fr[SPT][3];
            Real q1[SPT];
            const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
            const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
            const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
            // register loads first, unconditional (past the end of the real range they read the last real slot: discarded
            // below), so that they form one block the compiler has no reason to split with the conversions
#pragma unroll
            for (int i = 0; i < SPT; i++) {
                const int ls = min(base + tid + i * BLOCK, R - 1);
                role[i] = dd.role[ls];
                fr[i][0] = load_force(d.force + ls);
                fr[i][1] = load_force(d.force + P + ls);
                fr[i][2] = load_force(d.force + 2 * P + ls);
                q1[i] = INDEXED ? (TABLES ? static_cast<const Real*>(d.imageChargeBySlot)[ls] : (Real) 0)
                                : (d.imageCharge ? static_cast<const Real*>(d.imageCharge)[ls] : posq[4 * ((size_t) R + ls) + 3]);
            }
#pragma unroll
            for (int i = 0; i < SPT; i++) {
                const int idx = tid + i * BLOCK, slot = base + idx, at = drude_swizzle(idx);
                if (slot < R) {
                    cp_async16(&st.v[at], velm + 4 * (size_t) slot);
                    if (sizeof(Vec4<Mixed>) == 32) cp_async16(reinterpret_cast<char*>(&st.v[at]) + 16, velm + 4 * (size_t) slot + 2);
                    cp_async16(&st.p[at], posq + 4 * (size_t) slot);
                    if (sizeof(Vec4<Real>) == 32) cp_async16(reinterpret_cast<char*>(&st.p[at]) + 16, posq + 4 * (size_t) slot + 2);
                    if (kSplit) cp_async16(&st.c[at], corrA + 4 * (size_t) slot);
                    if (INDEXED) {
                        if (d.atomIndex) cp_async4(&st.orig[idx], d.atomIndex + slot);
                        else st.orig[idx] = slot;
                        if (TABLES) cp_async4(&st.img[idx], d.imageSlotBySlot + slot);
                    }
                    emit_zero<PREC, STREAM>(d, slot);
                } else {
                    role[i] = DRUDE_ROLE_NORMAL;  // which no parent test below can mistake for "owned by slot x"
                }
            }
#pragma unroll
            for (int i = 0; i < SPT; i++) {
                const int idx = tid + i * BLOCK, slot = base + idx, at = drude_swizzle(idx);
                const bool inRange = slot < R;
                if (inRange) {
                    st.fq[at] = Vec4<Mixed>{from_fixed(fr[i][0], Mixed()), from_fixed(fr[i][1], Mixed()), from_fixed(fr[i][2], Mixed()), (Mixed) q1[i]};
                    if (role[i] <= DRUDE_ROLE_PASSIVE) {  // a parent: tell the owner, if it is staged here
                        const int owner = DRUDE_ROLE_PASSIVE - role[i] - base;
                        if (owner >= 0 && owner < kTile) st.partner[owner] = (unsigned short) idx;
                    }
                }
                st.role[idx] = role[i];
                ownerMask[i] = __ballot_sync(0xffffffffu, role[i] >= 0);
                normalMask[i] = __ballot_sync(0xffffffffu, inRange && role[i] == DRUDE_ROLE_NORMAL);
                if (lane == 0) {
                    st.warpOwners[i * kWarps + warp] = __popc(ownerMask[i]);
                    st.warpNormals[i * kWarps + warp] = __popc(normalMask[i]);
                }
            }
            cp_async_wait_all();
        }
        if (tid == 0) st.next = 0;
        publish_step(box, d.counters);
        __syncthreads();
        const unsigned long long step = box.step;
        const unsigned int ticket = take_ticket(d.counters, holders);
        // 2. compaction: owners first, then ordinary particles
        int numOwners = 0, numNormals = 0;
        {
            int ownersBefore[SPT], normalsBefore[SPT];
#pragma unroll
            for (int i = 0; i < SPT; i++)
#pragma unroll
                for (int w = 0; w < kWarps; w++) {
                    if (w == warp) { ownersBefore[i] = numOwners; normalsBefore[i] = numNormals; }
                    numOwners += st.warpOwners[i * kWarps + w];
                    numNormals += st.warpNormals[i * kWarps + w];
                }
            const unsigned int below = (1u << lane) - 1u;
#pragma unroll
            for (int i = 0; i < SPT; i++) {
                const int idx = tid + i * BLOCK;
                if ((ownerMask[i] >> lane) & 1u) st.list[ownersBefore[i] + __popc(ownerMask[i] & below)] = (unsigned short) idx;
                if ((normalMask[i] >> lane) & 1u) st.list[numOwners + normalsBefore[i] + __popc(normalMask[i] & below)] = (unsigned short) idx;
            }
        }
        __syncthreads();

        // 3. the warps draw chunks of 32 work items
        const int ownerChunks = (numOwners + 31) >> 5, chunks = ownerChunks + ((numNormals + 31) >> 5);
        for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&st.next, 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= chunks) break;
            if (chunk < ownerChunks) {
                const int item = chunk * 32 + lane;
                if (item < numOwners) {
                    const int a = (int) st.list[item];                // tile-local slot of the Drude particle
                    const int aAt = drude_swizzle(a);
                    const int pairIndex = st.role[a];
                    int b = (int) st.partner[a];                      // tile-local slot of the parent, if staged here
                    const bool partnerInTile = b < kTile && st.role[b] == DRUDE_ROLE_PASSIVE - (base + a);
                    Vec4<Mixed> v1 = st.v[aAt], v2, delta1{0, 0, 0, 0}, delta2{0, 0, 0, 0};
                    Vec4<Real> p1 = st.p[aAt], p2, c1 = kSplit ? st.c[aAt] : Vec4<Real>{0, 0, 0, 0}, c2{0, 0, 0, 0};
                    const Vec4<Mixed> fq1 = st.fq[aAt];
                    const Mixed f1[3] = {fq1.x, fq1.y, fq1.z};
                    Mixed f2[3];
                    Real q2 = 0;
                    int orig2, far = 0, bAt = 0;
                    if (partnerInTile) {
                        bAt = drude_swizzle(b);
                        v2 = st.v[bAt];
                        p2 = st.p[bAt];
                        if (kSplit) c2 = st.c[bAt];
                        const Vec4<Mixed> fq2 = st.fq[bAt];
                        f2[0] = fq2.x; f2[1] = fq2.y; f2[2] = fq2.z;
                        orig2 = INDEXED ? st.orig[b] : base + b;
                    } else {  // a pair that straddles two tiles: the owner handles its parent in global memory
                        far = dd.pairs[pairIndex].y;
                        DrudeAtom<PREC> rec;
                        drude_load<PREC, DRUDE_FUSED, INDEXED, STREAM>(d, far, rec);
                        v2 = rec.v; p2 = rec.p; c2 = rec.c; q2 = rec.q1;
                        f2[0] = from_fixed(rec.f[0], Mixed()); f2[1] = from_fixed(rec.f[1], Mixed()); f2[2] = from_fixed(rec.f[2], Mixed());
                        orig2 = (INDEXED && d.atomIndex) ? d.atomIndex[far] : far;
                    }
                    float4 xi1, xi2;
                    if (random != nullptr) {
                        const size_t at = (size_t) randomIndex + (size_t) dd.numNormal + 2 * (size_t) pairIndex;
                        xi1 = __ldcs(random + at);
                        xi2 = __ldcs(random + at + 1);
                    } else {
                        const int orig1 = INDEXED ? st.orig[a] : base + a;
                        xi1 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) orig1, step);
                        xi2 = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) orig2, step);
                    }
                    drude_pair_part1_converted(v1, v2, delta1, delta2, f1, f2, xi1, xi2, s, ds);
                    if (v1.w != 0) part2_update<PREC>(p1, c1, v1, delta1, s.invDt);
                    if (v2.w != 0) part2_update<PREC>(p2, c2, v2, delta2, s.invDt);
                    if (ds.maxDistance > 0)
                        drude_hardwall<PREC>(v1, v2, p1, c1, p2, c2, s.dt, ds.maxDistance, ds.hardwallScale);
                    st.v[aAt] = v1;
                    st.p[aAt] = p1;
                    if (kSplit) st.c[aAt] = c1;
                    if (partnerInTile) {
                        st.v[bAt] = v2;
                        st.p[bAt] = p2;
                        if (kSplit) st.c[bAt] = c2;
                    } else if (INDEXED) {
                        DrudeAtom<PREC> rec;
                        rec.v = v2; rec.p = p2; rec.c = c2;
                        drude_emit<PREC, true, STREAM>(d, far, orig2, true, rec);
                    } else {
                        emit_atom<PREC, STREAM, false>(d, d.zshift, far, true, v2, p2, c2, q2);  // its image slots are zeroed by the thread that staged it
                    }
                }
            } else {
                const int item = (chunk - ownerChunks) * 32 + lane;
                if (item < numNormals) {
                    const int a = (int) st.list[numOwners + item];
                    const int aAt = drude_swizzle(a);
                    Vec4<Mixed> v = st.v[aAt];
                    if (v.w != 0) {
                        Vec4<Real> p = st.p[aAt], c = kSplit ? st.c[aAt] : Vec4<Real>{0, 0, 0, 0};
                        const Vec4<Mixed> fq = st.fq[aAt];
                        Vec4<Mixed> delta;
                        float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (random != nullptr) xi = __ldcs(random + (size_t) randomIndex + (size_t) base + a);
                        else if (s.noisescale != 0)
                            xi = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) (INDEXED ? st.orig[a] : base + a), step);
                        part1_update_converted(v, delta, fq.x, fq.y, fq.z, xi.x, xi.y, xi.z, s);
                        part2_update<PREC>(p, c, v, delta, s.invDt);
                        st.v[aAt] = v;
                        st.p[aAt] = p;
                        if (kSplit) st.c[aAt] = c;
                    }
                }
            }
        }
        __syncthreads();

        // 4. every thread emits its own slots; a parent whose owner sits in another tile was emitted by that owner
#pragma unroll
        for (int i = 0; i < SPT; i++) {
            const int idx = tid + i * BLOCK, slot = base + idx, at = drude_swizzle(idx);
            const int role = st.role[idx];
            bool mine = slot < R;
            if (role <= DRUDE_ROLE_PASSIVE && mine) {
                const int owner = DRUDE_ROLE_PASSIVE - role;
                mine = owner >= base && owner < base + kTile;
            }
            if (mine) {
                const Vec4<Mixed> v = st.v[at];
                const Vec4<Real> p = st.p[at];
                const Vec4<Real> c = kSplit ? st.c[at] : Vec4<Real>{0, 0, 0, 0};
                // a pair's records are always written back (constVDrudeLangevin.cu:60-63 stores unconditionally)
                const bool moved = role != DRUDE_ROLE_NORMAL || v.w != 0;
                if (INDEXED && TABLES) {
                    drude_emit_tables<PREC>(d, slot, moved, v, p, c, st.img[idx], (Real) st.fq[at].w);
                } else if (INDEXED) {
                    DrudeAtom<PREC> rec;
                    rec.v = v; rec.p = p; rec.c = c;
                    drude_emit<PREC, true, STREAM>(d, slot, st.orig[idx], moved, rec);
                } else {
                    emit_atom<PREC, STREAM, false>(d, d.zshift, slot, moved, v, p, c, (Real) st.fq[at].w);
                }
            }
        }
        close_step(d.counters, ticket, holders, step);
        if (tile + (int) gridDim.x < tileHi) __syncthreads();  // stage, partner entries and box are rewritten by the next tile
    }
}

// applyHardWallConstraints alone (Drude.cu:108-211, launch CudaConstVKernels.cpp:336-340): thread <-> pair.
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_drude_hardwall_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ DrudeDev dd) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    const DrudeScalars<Mixed> ds = drude_scalars_of<Mixed>(dd);
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < dd.numPairs; i += gridDim.x * BLOCK) {
        const int2 pair = dd.pairs[i];
        Vec4<Mixed> v1 = load4(velm, (size_t) pair.x), v2 = load4(velm, (size_t) pair.y);
        Vec4<Real> p1 = load4(posq, (size_t) pair.x), p2 = load4(posq, (size_t) pair.y);
        Vec4<Real> c1{0, 0, 0, 0}, c2{0, 0, 0, 0};
        if (kSplit) { c1 = load4(corrA, (size_t) pair.x); c2 = load4(corrA, (size_t) pair.y); }
        const bool parentMoves = v2.w != 0;
        if (drude_hardwall<PREC>(v1, v2, p1, c1, p2, c2, s.dt, ds.maxDistance, ds.hardwallScale)) {
            store4(posq, (size_t) pair.x, p1);
            if (kSplit) store4(corrA, (size_t) pair.x, c1);
            store4(velm, (size_t) pair.x, v1);
            if (parentMoves) {
                store4(posq, (size_t) pair.y, p2);
                if (kSplit) store4(corrA, (size_t) pair.y, c2);
                store4(velm, (size_t) pair.y, v2);
            }
        }
    }
}

}  // namespace constv
