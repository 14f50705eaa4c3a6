// Device code for rigid three-site molecules inside the ConstV Langevin step (sm_100a): SURVEY.md 8f.3.
//
// What it replaces in the reference (scychon/openmm_constV): the constraint hand-off of
// platforms/cuda/src/CudaConstVKernels.cpp:146-166 -- part 1 | integration.applyConstraints(tol) | part 2 |
// image rewrite -- for the systems the reference's example runs, SPC/E water with constraints=AllBonds
// (example/nacl_1m_spce/nacl_constv.py:56).  Three distance constraints on a rigid 3-site molecule are what
// OpenMM's CUDA platform solves analytically with SETTLE (Miyamoto & Kollman, J. Comput. Chem. 13, 952
// (1992)); the solver lives in OpenMM, not in the reference tree, so it is restated here from the published
// algorithm in the formulation OpenMM's kernel uses [recalled]: displacements (posDelta) relative to the
// apex atom's old position, float parameters (d(apex-b) = d(apex-c), d(b-c)).
//
// The reference streams posDelta and velm through HBM three times around the solver (part 1 writes them,
// the solver reads and rewrites posDelta, part 2 reads them back).  Here the thread that owns a molecule's
// apex does part 1 for its three atoms, solves the constraints in registers, does part 2, rewrites the three
// images and zeroes them: ONE kernel, each record crossing HBM once, posDelta never leaving registers.
// constv_settle_positions_kernel is the solver alone (thread <-> molecule), for the split path.
//
// Arithmetic: a canonical sequence of individually rounded operations and explicit fused multiply-adds
// (struct RN: operators on __f*_rn / __d*_rn intrinsics, which the compiler never contracts or reorders),
// divisions, reciprocals and square roots IEEE-correct; the CPU oracle spells the same sequence, so results
// agree bit for bit.
#pragma once

#include "constv_drude_kernels.cuh"

namespace constv {

enum { SETTLE_ROLE_NORMAL = -1, SETTLE_ROLE_PASSIVE = -2 };  // item kinds; >= 0: that molecule

// When the work items form a few runs -- ordinary atoms in consecutive slots, or molecules laid out
// (apex, apex+1, apex+2) one after the other with the same parameters: ions, then water, then wall atoms --
// a thread finds its slots arithmetically instead of through two dependent loads (item list, molecule list).
constexpr int kSettleMaxSegments = 8;
struct SettleSegments {
    int count;                              // 0 = not segmented: use the item list
    int itemStart[kSettleMaxSegments + 1];  // first work item of each run; [count] = numItems
    int slotStart[kSettleMaxSegments];      // slot of that item
    int atomsPerItem[kSettleMaxSegments];   // 1 (ordinary atoms) or 3 (molecules)
    float2 params[kSettleMaxSegments];      // molecule runs: (d(apex-b), d(b-c))
    int tileStart[kSettleMaxSegments + 1];  // tiled kernel: first CTA of each run (tiles of kSettleTile items)
};
constexpr int kSettleTile = 128;            // work items per CTA of the tiled kernel

struct SettleDev {
    const int2* items;    // work items of the fused step in slot order: (slot, SETTLE_ROLE_NORMAL) = an ordinary
                          // real atom, (apex slot, molecule index) = a molecule; b and c atoms are not listed
    const int4* atoms;    // (apex, b, c, 0) slots
    const float2* params; // (d(apex-b) = d(apex-c), d(b-c))
    int numItems, numClusters;
};

// A value whose +, -, *, / are individually rounded IEEE operations, whatever the compiler flags.
template <typename M> struct RN {
    M v;
    __device__ __forceinline__ RN() : v(0) {}
    __device__ __forceinline__ RN(M x) : v(x) {}
};
template <typename M> __device__ __forceinline__ RN<M> operator+(RN<M> a, RN<M> b) { return RN<M>(add_rn(a.v, b.v)); }
template <typename M> __device__ __forceinline__ RN<M> operator-(RN<M> a, RN<M> b) { return RN<M>(add_rn(a.v, -b.v)); }
template <typename M> __device__ __forceinline__ RN<M> operator*(RN<M> a, RN<M> b) { return RN<M>(mul_rn(a.v, b.v)); }
template <typename M> __device__ __forceinline__ RN<M> operator/(RN<M> a, RN<M> b) { return RN<M>(div_rn(a.v, b.v)); }
template <typename M> __device__ __forceinline__ RN<M> operator-(RN<M> a) { return RN<M>(-a.v); }
template <typename M> __device__ __forceinline__ RN<M> sqrt_of(RN<M> a) { return RN<M>(sqrt_rn(a.v)); }

template <typename M> __device__ __forceinline__ RN<M> fma_of(RN<M> a, RN<M> b, RN<M> c) { return RN<M>(fma_rn(a.v, b.v, c.v)); }
template <typename M> __device__ __forceinline__ RN<M> rcp_of(RN<M> a) { return RN<M>(rcp_rn(a.v)); }
// a1*b1 + a2*b2 + a3*b3 as fma(a3, b3, fma(a2, b2, a1*b1))
template <typename M> __device__ __forceinline__ RN<M> dot3(RN<M> a1, RN<M> b1, RN<M> a2, RN<M> b2, RN<M> a3, RN<M> b3) {
    return fma_of(a3, b3, fma_of(a2, b2, a1 * b1));
}

// SETTLE for one molecule.  pos* = old positions as `mixed`; d* = displacements to the unconstrained new
// positions, overwritten with the constrained ones (w untouched); m* = masses.  The canonical sequence:
// fused multiply-adds exactly where written, everything else rounded separately, one correctly rounded
// reciprocal per frame axis.
template <typename M>
__device__ __forceinline__ void settle_molecule(const M (&pos0)[3], const M (&pos1)[3], const M (&pos2)[3],
                                                Vec4<M>& d0, Vec4<M>& d1, Vec4<M>& d2,
                                                const M mass0, const M mass1, const M mass2,
                                                const float distAB, const float distBC) {
    using R = RN<M>;
    const R m0(mass0), m1(mass1), m2(mass2), one((M) 1), two((M) 2), four((M) 4), half((M) 0.5);
    // old bond vectors, from the apex
    const R xb0 = R(pos1[0]) - R(pos0[0]), yb0 = R(pos1[1]) - R(pos0[1]), zb0 = R(pos1[2]) - R(pos0[2]);
    const R xc0 = R(pos2[0]) - R(pos0[0]), yc0 = R(pos2[1]) - R(pos0[1]), zc0 = R(pos2[2]) - R(pos0[2]);

    const R invTotalMass = rcp_of(m0 + m1 + m2);
    const R xcom = fma_of(xc0 + R(d2.x), m2, fma_of(xb0 + R(d1.x), m1, R(d0.x) * m0)) * invTotalMass;
    const R ycom = fma_of(yc0 + R(d2.y), m2, fma_of(yb0 + R(d1.y), m1, R(d0.y) * m0)) * invTotalMass;
    const R zcom = fma_of(zc0 + R(d2.z), m2, fma_of(zb0 + R(d1.z), m1, R(d0.z) * m0)) * invTotalMass;

    // unconstrained new positions relative to the new centre of mass
    const R xa1 = R(d0.x) - xcom, ya1 = R(d0.y) - ycom, za1 = R(d0.z) - zcom;
    const R xb1 = xb0 + R(d1.x) - xcom, yb1 = yb0 + R(d1.y) - ycom, zb1 = zb0 + R(d1.z) - zcom;
    const R xc1 = xc0 + R(d2.x) - xcom, yc1 = yc0 + R(d2.y) - ycom, zc1 = zc0 + R(d2.z) - zcom;

    // orthonormal frame: Z = normal of the old molecular plane, X = a1 x Z, Y = Z x X
    const R xaksZd = fma_of(yb0, zc0, -(zb0 * yc0));
    const R yaksZd = fma_of(zb0, xc0, -(xb0 * zc0));
    const R zaksZd = fma_of(xb0, yc0, -(yb0 * xc0));
    const R xaksXd = fma_of(ya1, zaksZd, -(za1 * yaksZd));
    const R yaksXd = fma_of(za1, xaksZd, -(xa1 * zaksZd));
    const R zaksXd = fma_of(xa1, yaksZd, -(ya1 * xaksZd));
    const R xaksYd = fma_of(yaksZd, zaksXd, -(zaksZd * yaksXd));
    const R yaksYd = fma_of(zaksZd, xaksXd, -(xaksZd * zaksXd));
    const R zaksYd = fma_of(xaksZd, yaksXd, -(yaksZd * xaksXd));

    const R invX = rcp_of(sqrt_of(fma_of(zaksXd, zaksXd, fma_of(yaksXd, yaksXd, xaksXd * xaksXd))));
    const R invY = rcp_of(sqrt_of(fma_of(zaksYd, zaksYd, fma_of(yaksYd, yaksYd, xaksYd * xaksYd))));
    const R invZ = rcp_of(sqrt_of(fma_of(zaksZd, zaksZd, fma_of(yaksZd, yaksZd, xaksZd * xaksZd))));
    const R trns11 = xaksXd * invX, trns21 = yaksXd * invX, trns31 = zaksXd * invX;
    const R trns12 = xaksYd * invY, trns22 = yaksYd * invY, trns32 = zaksYd * invY;
    const R trns13 = xaksZd * invZ, trns23 = yaksZd * invZ, trns33 = zaksZd * invZ;

    const R xb0d = dot3(trns11, xb0, trns21, yb0, trns31, zb0);
    const R yb0d = dot3(trns12, xb0, trns22, yb0, trns32, zb0);
    const R xc0d = dot3(trns11, xc0, trns21, yc0, trns31, zc0);
    const R yc0d = dot3(trns12, xc0, trns22, yc0, trns32, zc0);
    const R za1d = dot3(trns13, xa1, trns23, ya1, trns33, za1);
    const R xb1d = dot3(trns11, xb1, trns21, yb1, trns31, zb1);
    const R yb1d = dot3(trns12, xb1, trns22, yb1, trns32, zb1);
    const R zb1d = dot3(trns13, xb1, trns23, yb1, trns33, zb1);
    const R xc1d = dot3(trns11, xc1, trns21, yc1, trns31, zc1);
    const R yc1d = dot3(trns12, xc1, trns22, yc1, trns32, zc1);
    const R zc1d = dot3(trns13, xc1, trns23, yc1, trns33, zc1);

    // step 2: the canonical triangle (apex at +ra on Y, b and c at -rb, x = -+rc), tilted by phi and psi.
    // The squared parameters are float arithmetic, as in the formulation restated (float2 parameters).
    const float rcf = __fmul_rn(0.5f, distBC);
    const R rc((M) rcf);
    R rb = sqrt_of(R((M) __fadd_rn(__fmul_rn(distAB, distAB), -__fmul_rn(rcf, rcf))));
    const R ra = rb * (m1 + m2) * invTotalMass;
    rb = rb - ra;
    const R sinphi = za1d / ra;
    const R cosphi = sqrt_of(fma_of(-sinphi, sinphi, one));
    const R sinpsi = (zb1d - zc1d) / (two * rc * cosphi);
    const R cospsi = sqrt_of(fma_of(-sinpsi, sinpsi, one));

    const R ya2d = ra * cosphi;
    R xb2d = -(rc * cospsi);
    const R tilt = rc * sinpsi * sinphi;
    const R yb2d = -(rb * cosphi) - tilt;
    const R yc2d = -(rb * cosphi) + tilt;
    const R xb2d2 = xb2d * xb2d;
    const R hh2 = fma_of(zb1d - zc1d, zb1d - zc1d, fma_of(yb2d - yc2d, yb2d - yc2d, four * xb2d2));
    const R deltx = fma_of(two, xb2d, sqrt_of(fma_of(four, xb2d2, -hh2) + R((M) __fmul_rn(distBC, distBC))));
    xb2d = fma_of(-deltx, half, xb2d);

    // step 3: the in-plane rotation theta that conserves the angular momentum about Z
    const R alpha = fma_of(yc0d, yc2d, fma_of(yb0d, yb2d, xb2d * (xb0d - xc0d)));
    const R beta = fma_of(xc0d, yc2d, fma_of(xb0d, yb2d, xb2d * (yc0d - yb0d)));
    const R gamma = fma_of(xc0d, yc1d, -(xc1d * yc0d)) + fma_of(xb0d, yb1d, -(xb1d * yb0d));
    const R al2be2 = fma_of(beta, beta, alpha * alpha);
    const R sintheta = fma_of(alpha, gamma, -(beta * sqrt_of(fma_of(-gamma, gamma, al2be2)))) / al2be2;

    // step 4
    const R costheta = sqrt_of(fma_of(-sintheta, sintheta, one));
    const R xa3d = -(ya2d * sintheta);
    const R ya3d = ya2d * costheta;
    const R za3d = za1d;
    const R xb3d = fma_of(xb2d, costheta, -(yb2d * sintheta));
    const R yb3d = fma_of(xb2d, sintheta, yb2d * costheta);
    const R zb3d = zb1d;
    const R xc3d = -fma_of(xb2d, costheta, yc2d * sintheta);
    const R yc3d = fma_of(-xb2d, sintheta, yc2d * costheta);
    const R zc3d = zc1d;

    // step 5: back to the laboratory frame, and to displacements
    d0.x = (xcom + dot3(trns11, xa3d, trns12, ya3d, trns13, za3d)).v;
    d0.y = (ycom + dot3(trns21, xa3d, trns22, ya3d, trns23, za3d)).v;
    d0.z = (zcom + dot3(trns31, xa3d, trns32, ya3d, trns33, za3d)).v;
    d1.x = (xcom + dot3(trns11, xb3d, trns12, yb3d, trns13, zb3d) - xb0).v;
    d1.y = (ycom + dot3(trns21, xb3d, trns22, yb3d, trns23, zb3d) - yb0).v;
    d1.z = (zcom + dot3(trns31, xb3d, trns32, yb3d, trns33, zb3d) - zb0).v;
    d2.x = (xcom + dot3(trns11, xc3d, trns12, yc3d, trns13, zc3d) - xc0).v;
    d2.y = (ycom + dot3(trns21, xc3d, trns22, yc3d, trns23, zc3d) - yc0).v;
    d2.z = (zcom + dot3(trns31, xc3d, trns32, yc3d, trns33, zc3d) - zc0).v;
}

// The solver alone on posDelta (the constraint step of the split path): thread <-> molecule.
template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_settle_positions_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleDev sd) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    Mixed* __restrict__ posDelta = static_cast<Mixed*>(d.posDelta);
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < sd.numClusters; i += gridDim.x * BLOCK) {
        const int4 at = sd.atoms[i];
        const float2 prm = sd.params[i];
        const int slot[3] = {at.x, at.y, at.z};
        Mixed pos[3][3], mass[3];
        Vec4<Mixed> delta[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const Vec4<Real> p = load4(posq, (size_t) slot[k]);
            Vec4<Real> c{0, 0, 0, 0};
            if (kSplit) c = load4(corrA, (size_t) slot[k]);
            drude_pos_load<PREC>(p, c, pos[k]);
            delta[k] = load4(posDelta, (size_t) slot[k]);
            mass[k] = rcp_rn(velm[4 * (size_t) slot[k] + 3]);
        }
        settle_molecule<Mixed>(pos[0], pos[1], pos[2], delta[0], delta[1], delta[2], mass[0], mass[1], mass[2], prm.x, prm.y);
#pragma unroll
        for (int k = 0; k < 3; k++) store4(posDelta, (size_t) slot[k], delta[k]);
    }
}

// The whole step of a system whose constraints are all rigid 3-site molecules, in one kernel.
// Thread <-> work item, in slot order: an ordinary real atom (the plain Langevin step) or a molecule (part 1
// for its three atoms, SETTLE, part 2, the image rewrite and the zeroing).  Items, not slots, so that the
// lanes of a warp inside the water region ALL solve a molecule (a thread-per-slot version with two idle
// lanes in three was issue-bound at 3x the unconstrained step's time).  A lane's three records are
// neighbours in memory (stride 48 B across the warp): the three gathers of a warp cover the same lines.
// Reordered slots (INDEXED): threads past the item list are the image slots, which zero themselves.
// Gaussians: the caller's pool with the reference's addressing (random[randomIndex + slot],
// constVLangevin.cu:14) or the Philox stream keyed by the original atom index.
template <int PREC, int BLOCK, bool INDEXED, bool STREAM, bool SEGMENTED>
__global__ void __launch_bounds__(BLOCK)
constv_settle_step_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleDev sd,
                          const __grid_constant__ SettleSegments seg,
                          const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    const int numWork = sd.numItems + (INDEXED ? d.numAtoms - R : 0);
    const unsigned int holders = (unsigned int) ((numWork + BLOCK - 1) / BLOCK);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);

    for (long long base = (long long) blockIdx.x * BLOCK; base < numWork; base += (long long) gridDim.x * BLOCK) {
        const int t = (int) base + (int) threadIdx.x;
        int2 item = make_int2(0, SETTLE_ROLE_PASSIVE);
        float2 prm = make_float2(0.f, 0.f);
        int slots[3] = {0, 0, 0};
        if (SEGMENTED) {
            if (t < sd.numItems) {
                int k = 0;
#pragma unroll
                for (int q = 1; q < kSettleMaxSegments; q++)
                    if (q < seg.count && t >= seg.itemStart[q]) k = q;
                const int per = seg.atomsPerItem[k];
                item.x = seg.slotStart[k] + (t - seg.itemStart[k]) * per;
                item.y = per == 3 ? 0 : SETTLE_ROLE_NORMAL;
                prm = seg.params[k];
            }
            slots[0] = item.x;
            slots[1] = item.x + 1;
            slots[2] = item.x + 2;
        } else {
            if (t < sd.numItems) item = sd.items[t];
            slots[0] = slots[1] = slots[2] = item.x;
            if (item.y >= 0) {
                const int4 at = sd.atoms[item.y];
                slots[1] = at.y;
                slots[2] = at.z;
                prm = sd.params[item.y];
            }
        }
        const int role = item.y;
        DrudeAtom<PREC> atom[3];
        float4 xi[3];
        xi[0] = xi[1] = xi[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int count = role >= 0 ? 3 : (role == SETTLE_ROLE_NORMAL ? 1 : 0);
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < count) {
                drude_load<PREC, DRUDE_FUSED, INDEXED, STREAM>(d, slots[k], atom[k]);
                if (random != nullptr) xi[k] = __ldcs(random + (size_t) randomIndex + slots[k]);
            }

        publish_step(box, d.counters);
        __syncthreads();
        const unsigned long long step = box.step;
        const unsigned int ticket = take_ticket(d.counters, holders);

        if (INDEXED && t >= sd.numItems && t < numWork) {  // an image slot zeroes itself
            const int slot = R + (t - sd.numItems);
            if (d.zeroVelm) store4(static_cast<Mixed*>(d.velm), (size_t) slot, Vec4<Mixed>{0, 0, 0, 0});
            if (d.zeroForce) {
                __stcs(d.force + slot, 0LL);
                __stcs(d.force + P + slot, 0LL);
                __stcs(d.force + 2 * P + slot, 0LL);
            }
        }

        int orig[3] = {slots[0], slots[1], slots[2]};
        bool moved[3] = {false, false, false};
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < count) {
                if (INDEXED && d.atomIndex) orig[k] = d.atomIndex[slots[k]];
                moved[k] = atom[k].v.w != 0;
                if (moved[k]) {
                    if (random == nullptr)
                        xi[k] = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) orig[k], step)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                    part1_update(atom[k].v, atom[k].delta, atom[k].f[0], atom[k].f[1], atom[k].f[2], xi[k].x, xi[k].y, xi[k].z, s);
                }
            }
        if (role >= 0) {
            Mixed pos[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++) drude_pos_load<PREC>(atom[k].p, atom[k].c, pos[k]);
            settle_molecule<Mixed>(pos[0], pos[1], pos[2], atom[0].delta, atom[1].delta, atom[2].delta,
                                   rcp_rn(atom[0].v.w), rcp_rn(atom[1].v.w), rcp_rn(atom[2].v.w), prm.x, prm.y);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < count) {
                if (moved[k]) part2_update<PREC>(atom[k].p, atom[k].c, atom[k].v, atom[k].delta, s.invDt);
                drude_emit<PREC, INDEXED, STREAM>(d, slots[k], orig[k], moved[k], atom[k]);
            }

        close_step(d.counters, ticket, holders, step);
        if (base + (long long) gridDim.x * BLOCK < numWork) __syncthreads();  // box is rewritten by the next tile
    }
}



// ---- the same step, staged through shared memory (identity slot order, segmented item list) -----------
// A molecule's three records are neighbours in memory, so a thread that gathers "its" three records makes
// every warp-wide access a stride-48-byte pattern: half-filled sectors on the way in and, worse, on the
// way out (measured: 20 us per step at 1 M atoms with the arithmetic REMOVED, against 12.5 us for the
// unconstrained step).  Here a CTA owns a tile of BLOCK work items of one run = BLOCK*per consecutive
// slots and moves them like the unconstrained kernel does -- thread j loads slots base + j + i*BLOCK, every
// access a full 16-byte (8-byte for force planes) coalesced vector -- into a structure-of-arrays stage in
// shared memory; after a barrier thread j computes item j from stage entries per*j + k (stride 3: free of
// bank conflicts), writes the new velocity and position back, and after a second barrier every thread
// emits "its" slots again: own record, rewritten images, zeroed images, all coalesced.
template <int PREC, int BLOCK, bool INDEXED> struct SettleStage {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    static constexpr int kSlots = 3 * BLOCK;
    Vec4<Mixed> v[kSlots];
    Vec4<Real> p[kSlots];
    Vec4<Real> c[Prec<PREC>::kSplitPos ? kSlots : 1];
    long long f[3][kSlots];
    Real q1[kSlots];                 // charge of the cell-1 image, fetched with the record (reordered slots: when the per-slot tables exist)
    int orig[INDEXED ? kSlots : 1];  // reordered slots: original index of the atom in the slot
    int img[INDEXED ? kSlots : 1];   // reordered slots: slot of the cell-1 image (per-slot tables), -1 = the slot holds no real atom
};

// Stage in slots [base, base + slotsInTile): thread j takes slots base + j + i * BLOCK (coalesced).
// ASYNC: asynchronous copies global -> shared (cp.async, SASS LDGSTS) that never pass through registers.  Under the
// 64-register cap the plain loads of the three slots of a thread were serialised by the compiler (load, wait, store
// to shared memory, three times: three DRAM round trips per tile, 44 % of the stall samples of the kernel); the copies
// are all in flight at once and are waited for once.  Measured (in-process A/B, profiles/r02_ab_settle_cp_async.jsonl):
// 15.5 -> 14.7 us per step at 1M atoms with three chains, 52.7 -> 46.9 us at 4M with two (0.88 of the copy peak).
template <int PREC, int BLOCK, bool INDEXED, bool STREAM, bool ASYNC>
__device__ __forceinline__ void settle_stage_in(const SlabDev& d, SettleStage<PREC, BLOCK, INDEXED>& st, const int base,
                                                const int slotsInTile, const int j) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const size_t P = (size_t) d.padded;
    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const Real* qbase = d.imageCharge ? static_cast<const Real*>(d.imageCharge) : posq + 4 * (size_t) d.numReal + 3;
    const size_t qstride = d.imageCharge ? 1 : 4;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int idx = j + i * BLOCK;
        if (idx < slotsInTile) {
            const size_t slot = (size_t) (base + idx);
            if (ASYNC) {
                cp_async16(&st.v[idx], velm + 4 * slot);
                if (sizeof(Vec4<Mixed>) == 32) cp_async16(reinterpret_cast<char*>(&st.v[idx]) + 16, velm + 4 * slot + 2);
                cp_async16(&st.p[idx], posq + 4 * slot);
                if (sizeof(Vec4<Real>) == 32) cp_async16(reinterpret_cast<char*>(&st.p[idx]) + 16, posq + 4 * slot + 2);
                if (kSplit) cp_async16(&st.c[idx], corrA + 4 * slot);
                cp_async8(&st.f[0][idx], force + slot);
                cp_async8(&st.f[1][idx], force + P + slot);
                cp_async8(&st.f[2][idx], force + 2 * P + slot);
                if (INDEXED) cp_async4(&st.orig[idx], d.atomIndex + slot);
                else if (sizeof(Real) == 4) cp_async4(&st.q1[idx], qbase + qstride * slot);
                else cp_async8(&st.q1[idx], qbase + qstride * slot);
            } else {
                st.v[idx] = load4x<STREAM>(velm, slot);
                st.p[idx] = load4x<STREAM>(posq, slot);
                if (kSplit) st.c[idx] = load4x<STREAM>(corrA, slot);
                st.f[0][idx] = load_force(force + slot);
                st.f[1][idx] = load_force(force + P + slot);
                st.f[2][idx] = load_force(force + 2 * P + slot);
                if (INDEXED) st.orig[idx] = d.atomIndex[slot];
                else st.q1[idx] = qbase[qstride * slot];
            }
        }
    }
    if (ASYNC) cp_async_wait_all();
}

// INDEXED = reordered slots (behind OpenMM after cu.reorderAtoms()).  OpenMM swaps whole identical molecules,
// so the real atoms keep slots [0, R) and a slot keeps its role: the tiles, the stage and the arithmetic are the
// same; what changes is that the Philox key is the ORIGINAL index atomIndex[slot], that an atom's images sit
// wherever invAtomOrder says (a scattered 16-byte store each, as in the reference, constVLangevin.cu:155-158),
// and that the image slots are zeroed in slot order (every thread zeroes the slots R * cell above its own: coalesced,
// whatever atoms they hold) instead of by their partners.
// One tile of the staged step for one slab (`tile` in the slab's own tile space), for the batched ensemble
// kernel.  The caller owns the loop over tiles and the barrier between two tiles (the stage and the step box are
// rewritten by the next tile).  Same text as the body of constv_settle_step_tiled_kernel.
template <int PREC, int BLOCK, bool INDEXED, bool STREAM>
__device__ __forceinline__ void settle_tiled_tile(const SlabDev& d, const SettleSegments& seg, SettleStage<PREC, BLOCK, INDEXED>& st,
                                                  StepBox& box, const float4* __restrict__ random, const unsigned int randomIndex,
                                                  const int tile) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    // a chain (constv_set_chains) advances tiles [rangeLo, rangeHi) of the tile space with its own counters;
    // the whole step = [0, realTiles)
    const unsigned int holders = (unsigned int) (d.rangeHi - d.rangeLo);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const int j = (int) threadIdx.x;

    int q = 0;
#pragma unroll
    for (int k = 1; k < kSettleMaxSegments; k++)
        if (k < seg.count && tile >= seg.tileStart[k]) q = k;
    const int per = seg.atomsPerItem[q];
    const int firstItem = (tile - seg.tileStart[q]) * BLOCK;
    const int items = min(BLOCK, seg.itemStart[q + 1] - seg.itemStart[q] - firstItem);
    const int base = seg.slotStart[q] + firstItem * per;
    const int slotsInTile = items * per;
    const float2 prm = seg.params[q];

    settle_stage_in<PREC, BLOCK, INDEXED, STREAM, true>(d, st, base, slotsInTile, j);
    publish_step(box, d.counters);
    __syncthreads();
    const unsigned long long step = box.step;
    const unsigned int ticket = take_ticket(d.counters, holders);

    // compute item j from the stage
    if (j < items) {
        Vec4<Mixed> v[3], delta[3];
        Vec4<Real> p[3], c[3];
        bool moved[3] = {false, false, false};
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < per) {
                const int idx = per * j + k;
                v[k] = st.v[idx];
                p[k] = st.p[idx];
                c[k] = kSplit ? st.c[idx] : Vec4<Real>{0, 0, 0, 0};
                delta[k] = Vec4<Mixed>{0, 0, 0, 0};
                moved[k] = v[k].w != 0;
                if (moved[k]) {
                    float4 xi;
                    const int a = INDEXED ? st.orig[idx] : base + idx;
                    if (random != nullptr) xi = __ldcs(random + (size_t) randomIndex + base + idx);
                    else xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                    part1_update(v[k], delta[k], st.f[0][idx], st.f[1][idx], st.f[2][idx], xi.x, xi.y, xi.z, s);
                }
            }
        if (per == 3) {
            Mixed pos[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++) drude_pos_load<PREC>(p[k], c[k], pos[k]);
            settle_molecule<Mixed>(pos[0], pos[1], pos[2], delta[0], delta[1], delta[2], rcp_rn(v[0].w), rcp_rn(v[1].w),
                                   rcp_rn(v[2].w), prm.x, prm.y);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < per && moved[k]) {
                const int idx = per * j + k;
                part2_update<PREC>(p[k], c[k], v[k], delta[k], s.invDt);
                st.v[idx] = v[k];
                st.p[idx] = p[k];
                if (kSplit) st.c[idx] = c[k];
            }
    }
    __syncthreads();

    // stage out: own record (coalesced), rewritten images, and (identity order) zeroed images
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int idx = j + i * BLOCK;
        if (idx < slotsInTile) {
            const Vec4<Mixed> v = st.v[idx];
            const Vec4<Real> p = st.p[idx];
            const Vec4<Real> c = kSplit ? st.c[idx] : Vec4<Real>{0, 0, 0, 0};
            if (INDEXED) {
                DrudeAtom<PREC> rec;
                rec.v = v; rec.p = p; rec.c = c;
                drude_emit<PREC, true, STREAM>(d, base + idx, st.orig[idx], v.w != 0, rec);
                emit_zero<PREC, STREAM>(d, base + idx);  // the image slots, in slot order: coalesced whatever atoms they hold
            } else {
                emit_atom<PREC, STREAM>(d, s.zshift, base + idx, v.w != 0, v, p, c, st.q1[idx]);
            }
        }
    }
    close_step(d.counters, ticket, holders, step);
}

// The single-slab kernel keeps its own copy of the tile body (below) rather than calling settle_tiled_tile: the
// call, though inlined, cost a spill under the 64-register cap and 8 % of the step.
// ASYNC: see settle_stage_in (false = the round-1 stage-in through registers, kernel variant 3).
// Reordered slots: the zeroing stores of the image slots are issued in the stage-in phase, behind the tile's loads (they
// depend on nothing): 18.7 -> 18.3 us per step at 1M atoms; in identity order they stay with the dependent stores at the
// end of the tile (issued early: 13.7 -> 14.0 us with six chains; profiles/r02_ab_staged_zero_timing.jsonl).
// TABLES (reordered slots): the cell-1 image's slot and charge come in with the tile from the per-slot tables
// (SlabDev::imageSlotBySlot) instead of being gathered through atomIndex -> invAtomOrder / imageCharge at the end of the tile.
template <int PREC, int BLOCK, bool INDEXED, bool STREAM, bool ASYNC = true, bool TABLES = false>
__global__ void __launch_bounds__(BLOCK, PREC == PREC_SINGLE ? 8 : (INDEXED ? 0 : 4))  // single precision: 64 registers, 8 CTAs per SM
constv_settle_step_tiled_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleSegments seg,
                                const float4* __restrict__ random, const unsigned int randomIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    constexpr bool ZEARLY = INDEXED;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SettleStage<PREC, BLOCK, INDEXED>& st = *reinterpret_cast<SettleStage<PREC, BLOCK, INDEXED>*>(smemRaw);
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
    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const Real* qbase = d.imageCharge ? static_cast<const Real*>(d.imageCharge) : posq + 4 * (size_t) R + 3;
    const size_t qstride = d.imageCharge ? 1 : 4;
    const int j = (int) threadIdx.x;

    for (int tile = tileLo + (int) blockIdx.x; tile < tileHi; tile += (int) gridDim.x) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < kSettleMaxSegments; k++)
            if (k < seg.count && tile >= seg.tileStart[k]) q = k;
        const int per = seg.atomsPerItem[q];
        const int firstItem = (tile - seg.tileStart[q]) * BLOCK;
        const int items = min(BLOCK, seg.itemStart[q + 1] - seg.itemStart[q] - firstItem);
        const int base = seg.slotStart[q] + firstItem * per;
        const int slotsInTile = items * per;
        const float2 prm = seg.params[q];

        // stage in (the text of settle_stage_in: calling it, though inlined, costs a spill under the 64-register cap)
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int idx = j + i * BLOCK;
            if (idx < slotsInTile) {
                const size_t slot = (size_t) (base + idx);
                if (ASYNC) {
                    cp_async16(&st.v[idx], velm + 4 * slot);
                    if (sizeof(Vec4<Mixed>) == 32) cp_async16(reinterpret_cast<char*>(&st.v[idx]) + 16, velm + 4 * slot + 2);
                    cp_async16(&st.p[idx], posq + 4 * slot);
                    if (sizeof(Vec4<Real>) == 32) cp_async16(reinterpret_cast<char*>(&st.p[idx]) + 16, posq + 4 * slot + 2);
                    if (kSplit) cp_async16(&st.c[idx], corrA + 4 * slot);
                    cp_async8(&st.f[0][idx], force + slot);
                    cp_async8(&st.f[1][idx], force + P + slot);
                    cp_async8(&st.f[2][idx], force + 2 * P + slot);
                    if (INDEXED) {
                        cp_async4(&st.orig[idx], d.atomIndex + slot);
                        if (TABLES) {
                            cp_async4(&st.img[idx], d.imageSlotBySlot + slot);
                            if (sizeof(Real) == 4) cp_async4(&st.q1[idx], static_cast<const Real*>(d.imageChargeBySlot) + slot);
                            else cp_async8(&st.q1[idx], static_cast<const Real*>(d.imageChargeBySlot) + slot);
                        }
                    } else if (sizeof(Real) == 4) cp_async4(&st.q1[idx], qbase + qstride * slot);
                    else cp_async8(&st.q1[idx], qbase + qstride * slot);
                } else {
                    st.v[idx] = load4x<STREAM>(velm, slot);
                    st.p[idx] = load4x<STREAM>(posq, slot);
                    if (kSplit) st.c[idx] = load4x<STREAM>(corrA, slot);
                    st.f[0][idx] = load_force(force + slot);
                    st.f[1][idx] = load_force(force + P + slot);
                    st.f[2][idx] = load_force(force + 2 * P + slot);
                    if (INDEXED) {
                        st.orig[idx] = d.atomIndex[slot];
                        if (TABLES) {
                            st.img[idx] = d.imageSlotBySlot[slot];
                            st.q1[idx] = static_cast<const Real*>(d.imageChargeBySlot)[slot];
                        }
                    } else st.q1[idx] = qbase[qstride * slot];
                }
                if (ZEARLY) emit_zero<PREC, STREAM>(d, base + idx);  // the image slots, in slot order: coalesced whatever atoms they hold
            }
        }
        if (ASYNC) cp_async_wait_all();
        publish_step(box, d.counters);
        __syncthreads();
        const unsigned long long step = box.step;
        const unsigned int ticket = take_ticket(d.counters, holders);

        // compute item j from the stage
        if (j < items) {
            Vec4<Mixed> v[3], delta[3];
            Vec4<Real> p[3], c[3];
            bool moved[3] = {false, false, false};
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (k < per) {
                    const int idx = per * j + k;
                    v[k] = st.v[idx];
                    p[k] = st.p[idx];
                    c[k] = kSplit ? st.c[idx] : Vec4<Real>{0, 0, 0, 0};
                    delta[k] = Vec4<Mixed>{0, 0, 0, 0};
                    moved[k] = v[k].w != 0;
                    if (moved[k]) {
                        float4 xi;
                        const int a = INDEXED ? st.orig[idx] : base + idx;
                        if (random != nullptr) xi = __ldcs(random + (size_t) randomIndex + base + idx);
                        else xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                        part1_update(v[k], delta[k], st.f[0][idx], st.f[1][idx], st.f[2][idx], xi.x, xi.y, xi.z, s);
                    }
                }
            if (per == 3) {
                Mixed pos[3][3];
#pragma unroll
                for (int k = 0; k < 3; k++) drude_pos_load<PREC>(p[k], c[k], pos[k]);
                settle_molecule<Mixed>(pos[0], pos[1], pos[2], delta[0], delta[1], delta[2], rcp_rn(v[0].w), rcp_rn(v[1].w),
                                       rcp_rn(v[2].w), prm.x, prm.y);
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (k < per && moved[k]) {
                    const int idx = per * j + k;
                    part2_update<PREC>(p[k], c[k], v[k], delta[k], s.invDt);
                    st.v[idx] = v[k];
                    st.p[idx] = p[k];
                    if (kSplit) st.c[idx] = c[k];
                }
        }
        __syncthreads();

        // stage out: own record (coalesced), rewritten images, and (identity order) zeroed images
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int idx = j + i * BLOCK;
            if (idx < slotsInTile) {
                const Vec4<Mixed> v = st.v[idx];
                const Vec4<Real> p = st.p[idx];
                const Vec4<Real> c = kSplit ? st.c[idx] : Vec4<Real>{0, 0, 0, 0};
                if (INDEXED && TABLES) {
                    drude_emit_tables<PREC>(d, base + idx, v.w != 0, v, p, c, st.img[idx], st.q1[idx]);
                } else if (INDEXED) {
                    DrudeAtom<PREC> rec;
                    rec.v = v; rec.p = p; rec.c = c;
                    drude_emit<PREC, true, STREAM>(d, base + idx, st.orig[idx], v.w != 0, rec);
                } else {
                    emit_atom<PREC, STREAM, !ZEARLY>(d, s.zshift, base + idx, v.w != 0, v, p, c, st.q1[idx]);
                }
            }
        }
        close_step(d.counters, ticket, holders, step);
        if (tile + (int) gridDim.x < tileHi) __syncthreads();  // stage and box are rewritten by the next tile
    }
}


// An ensemble of slabs (identity slot order, Philox noise) in ONE launch: the CTA looks its tile up in the table
// of slabs and runs it with that slab's descriptor, runs and step counter (one ticket per tile, on the tile's own
// slab: every slab's counter advances exactly once).
template <int CAP> struct SettleBatchParams {
    int numSlabs, totalTiles;
    int tileBase[CAP];
    SlabDev slab[CAP];
    SettleSegments seg[CAP];
};

template <int PREC, int BLOCK, int CAP>
__global__ void __launch_bounds__(BLOCK, PREC == PREC_SINGLE ? 8 : 4)
constv_settle_step_tiled_batched_kernel(const __grid_constant__ SettleBatchParams<CAP> b) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SettleStage<PREC, BLOCK, false>& st = *reinterpret_cast<SettleStage<PREC, BLOCK, false>*>(smemRaw);
    __shared__ StepBox box;
    pdl_wait();
    pdl_launch();
    for (int tile = (int) blockIdx.x; tile < b.totalTiles; tile += (int) gridDim.x) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < CAP; q++)
            if (q < b.numSlabs && tile >= b.tileBase[q]) k = q;
        settle_tiled_tile<PREC, BLOCK, false, true>(b.slab[k], b.seg[k], st, box, nullptr, 0u, tile - b.tileBase[k]);
        if (tile + (int) gridDim.x < b.totalTiles) __syncthreads();
    }
}

// ---- kinetic energy of a system with rigid molecules ------------------------------------------------------
// OpenMM's integration.computeKineticEnergy (CudaConstVKernels.cpp:175-177) projects the half-step-shifted
// velocities onto the constraints before summing; for rigid 3-site molecules that is a 3x3 linear system in the
// three Lagrange multipliers, solved here by Cramer's rule in double (an on-demand report).  Thread <-> work
// item; reduction as constv_kinetic_energy_kernel (warp shuffle -> shared -> one partial per CTA -> the last
// CTA adds the partials in index order).
__device__ __forceinline__ void settle_velocity_molecule(const double (&x)[3][3], double (&v)[3][3], const double (&w)[3]) {
    using D = RN<double>;
    D r[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r[0][k] = D(x[0][k]) - D(x[1][k]);
        r[1][k] = D(x[0][k]) - D(x[2][k]);
        r[2][k] = D(x[1][k]) - D(x[2][k]);
    }
    D g[3][3];
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
        for (int q = 0; q < 3; q++) g[p][q] = r[p][0] * r[q][0] + r[p][1] * r[q][1] + r[p][2] * r[q][2];
    const D w0(w[0]), w1(w[1]), w2(w[2]);
    const D a00 = (w0 + w1) * g[0][0], a01 = w0 * g[1][0], a02 = (-w1) * g[2][0];
    const D a10 = w0 * g[0][1], a11 = (w0 + w2) * g[1][1], a12 = w2 * g[2][1];
    const D a20 = (-w1) * g[0][2], a21 = w2 * g[1][2], a22 = (w1 + w2) * g[2][2];
    const D b0 = -((D(v[0][0]) - D(v[1][0])) * r[0][0] + (D(v[0][1]) - D(v[1][1])) * r[0][1] + (D(v[0][2]) - D(v[1][2])) * r[0][2]);
    const D b1 = -((D(v[0][0]) - D(v[2][0])) * r[1][0] + (D(v[0][1]) - D(v[2][1])) * r[1][1] + (D(v[0][2]) - D(v[2][2])) * r[1][2]);
    const D b2 = -((D(v[1][0]) - D(v[2][0])) * r[2][0] + (D(v[1][1]) - D(v[2][1])) * r[2][1] + (D(v[1][2]) - D(v[2][2])) * r[2][2]);
    const D c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const D det = a00 * c00 + a01 * c01 + a02 * c02;
    const D l0 = (b0 * c00 + a01 * (a12 * b2 - b1 * a22) + a02 * (b1 * a21 - a11 * b2)) / det;
    const D l1 = (a00 * (b1 * a22 - a12 * b2) + b0 * c01 + a02 * (a10 * b2 - b1 * a20)) / det;
    const D l2 = (a00 * (a11 * b2 - b1 * a21) + a01 * (b1 * a20 - a10 * b2) + b0 * c02) / det;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        v[0][k] = (D(v[0][k]) + w0 * (l0 * r[0][k] + l1 * r[1][k])).v;
        v[1][k] = (D(v[1][k]) + w1 * (l2 * r[2][k] - l0 * r[0][k])).v;
        v[2][k] = (D(v[2][k]) - w2 * (l1 * r[1][k] + l2 * r[2][k])).v;
    }
}

template <int PREC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
constv_kinetic_energy_rigid_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleDev sd, const double scaleD,
                                   double* __restrict__ partials, double* __restrict__ out) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    const size_t P = (size_t) d.padded;
    const Real* __restrict__ posq = static_cast<const Real*>(d.posq);
    const Real* __restrict__ corrA = static_cast<const Real*>(d.corr);
    const Mixed* __restrict__ velm = static_cast<const Mixed*>(d.velm);
    const long long* __restrict__ force = d.force;
    const Mixed scale = (Mixed) scaleD;
    double e = 0.0;
    for (int t = blockIdx.x * BLOCK + threadIdx.x; t < sd.numItems; t += gridDim.x * BLOCK) {
        const int2 item = sd.items[t];
        if (item.y < 0) {
            const Vec4<Mixed> v = load4(velm, (size_t) item.x);
            if (v.w != 0) {
                const double vx = (double) fma_rn(mul_rn(scale, from_fixed(force[item.x], Mixed())), v.w, v.x);
                const double vy = (double) fma_rn(mul_rn(scale, from_fixed(force[P + item.x], Mixed())), v.w, v.y);
                const double vz = (double) fma_rn(mul_rn(scale, from_fixed(force[2 * P + item.x], Mixed())), v.w, v.z);
                e += (vx * vx + vy * vy + vz * vz) / (double) v.w;
            }
        } else {
            const int4 at = sd.atoms[item.y];
            const int slots[3] = {at.x, at.y, at.z};
            double x[3][3], v[3][3], w[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const size_t s = (size_t) slots[a];
                const Vec4<Mixed> vm = load4(velm, s);
                const Vec4<Real> p = load4(posq, s);
                Vec4<Real> c{0, 0, 0, 0};
                if (kSplit) c = load4(corrA, s);
                Mixed pos[3];
                drude_pos_load<PREC>(p, c, pos);
                w[a] = (double) vm.w;
                x[a][0] = (double) pos[0]; x[a][1] = (double) pos[1]; x[a][2] = (double) pos[2];
                v[a][0] = (double) fma_rn(mul_rn(scale, from_fixed(force[s], Mixed())), vm.w, vm.x);
                v[a][1] = (double) fma_rn(mul_rn(scale, from_fixed(force[P + s], Mixed())), vm.w, vm.y);
                v[a][2] = (double) fma_rn(mul_rn(scale, from_fixed(force[2 * P + s], Mixed())), vm.w, vm.z);
            }
            settle_velocity_molecule(x, v, w);
#pragma unroll
            for (int a = 0; a < 3; a++) e += (v[a][0] * v[a][0] + v[a][1] * v[a][1] + v[a][2] * v[a][2]) / w[a];
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
