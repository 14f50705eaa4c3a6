// The constrained step (part 1 + SETTLE + part 2 + image rewrite + zeroing) as a
// warp-specialised TMA pipeline for sm_100a: NO thread of the kernel issues a global load or store in the
// common case -- every byte enters and leaves the SM as a bulk asynchronous copy (cp.async.bulk, SASS UBLKCP)
// between global memory and a ring of stages in shared memory, tracked by mbarriers.
//
// What it replaces: constv_settle_step_tiled_kernel (same tiles, same arithmetic, same bits), whose CTAs did
// load -> barrier -> compute -> barrier -> store one after the other with 8 CTAs per SM to hide each phase
// behind other CTAs' (profile r01: issue slots 46 % busy, 0.64 of the copy peak at 1M atoms, the barrier the
// first stall reason).  Here ONE persistent CTA per SM runs three roles concurrently:
//
//   producer (1 thread)    walks this CTA's tiles (round robin over the grid: the access front stays
//                          sequential across the slab, walls -- the cheap tiles -- last), waits for a free
//                          stage and issues the tile's 6-7 bulk loads: velm, posq, [posqCorrection], the three
//                          force planes, the image charges (side array, or the image posq records themselves);
//   consumers (NG groups   wait for the stage to fill, run part 1, the constraints and part 2 for "their" work
//   of 128 threads)        item out of shared memory, write the new records back IN PLACE and the rewritten
//                          image records next to them -- registers are never a staging area, so the solver
//                          keeps its ~110 registers without costing occupancy;
//   storer (1 warp)        waits until a tile has been computed and issues its 6-10 bulk stores: own records,
//                          image records, and the zeroing of the image velocities / forces straight out of a
//                          block of zeros in shared memory; releases the stage once the copies have read it.
//
// A tile is up to 128 work items of one run of the slab = up to 384 consecutive slots.  Bulk copies need
// 16-byte aligned addresses and sizes: the load window is widened to multiples of 4 slots (the few extra slots
// are read and ignored); stores cover exactly the tile (16-byte records), except the 8-byte force planes, whose
// odd edge elements the storer writes with plain stores.  Tiles whose atoms are not uniformly massive /
// massless or charged / neutral (never the case for runs of ions, water and wall atoms, but legal) take a
// per-slot store path in the storer warp.
//
// Requirements checked by the host (constv_settle_capi.inc: settleWsApplies): identity slot order, work items
// in arithmetic runs, numCells == 2, in-kernel Philox noise.
#pragma once

#include "constv_settle_kernels.cuh"

namespace constv {

constexpr int kWsItems = 128;                 // work items per tile = threads per consumer group
constexpr int kWsSlots = 3 * kWsItems + 4;    // slots of a stage: a tile plus the alignment margin of the load window
constexpr int kWsProducerWarps = 1, kWsStorerWarps = 1;

template <int PREC, bool STAGE_IMAGE_POSQ> struct WsStage {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    static constexpr bool kSplit = Prec<PREC>::kSplitPos;
    Vec4<Mixed> v[kWsSlots];                  // velm, rewritten in place
    Vec4<Real> p[kWsSlots];                   // posq, rewritten in place
    Vec4<Real> c[kSplit ? kWsSlots : 1];      // posqCorrection
    long long f[3][kWsSlots];                 // force planes
    Real q1[kWsSlots];                        // image charges (side array)
    Vec4<Real> ip[kWsSlots];                  // image posq records: loaded when there is no side array (w is kept), stored
    Vec4<Real> ic[kSplit ? kWsSlots : 1];     // image posqCorrection records, stored
    unsigned int counts[4];                   // [0] atoms that moved, [1] atoms that are charged (consumers -> storer)
};

template <int PREC> struct WsShape {
    static constexpr int kGroups = PREC == PREC_SINGLE ? 4 : 2;
    static constexpr int kThreads = 32 * (kWsProducerWarps + kWsStorerWarps) + kWsItems * kGroups;
    static constexpr size_t kStageBytes = (sizeof(WsStage<PREC, true>) + 127) / 128 * 128;
    static constexpr size_t kZeroBytes = (kWsSlots * sizeof(Vec4<typename Prec<PREC>::Mixed>) + 127) / 128 * 128;
    static constexpr int kStages = (int) ((227 * 1024 - kZeroBytes - 1024) / kStageBytes);
    static constexpr size_t kSmemBytes = kZeroBytes + kStages * kStageBytes + 128;  // + alignment slack
};

__device__ __forceinline__ void mbar_arrive(unsigned int bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global bulk copy of the calling thread's current bulk group
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned int src, unsigned int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Where tile `tile` of the slab lies: run q, slots [base, base + n), `per` atoms per item.
struct WsTile {
    int q, per, items, base, n;
    float2 prm;
};
__device__ __forceinline__ WsTile ws_tile_of(const SettleSegments& seg, const int tile) {
    WsTile t;
    t.q = 0;
#pragma unroll
    for (int k = 1; k < kSettleMaxSegments; k++)
        if (k < seg.count && tile >= seg.tileStart[k]) t.q = k;
    t.per = seg.atomsPerItem[t.q];
    const int firstItem = (tile - seg.tileStart[t.q]) * kWsItems;
    t.items = min(kWsItems, seg.itemStart[t.q + 1] - seg.itemStart[t.q] - firstItem);
    t.base = seg.slotStart[t.q] + firstItem * t.per;
    t.n = t.items * t.per;
    t.prm = seg.params[t.q];
    return t;
}

template <int PREC>
__global__ void __launch_bounds__(WsShape<PREC>::kThreads, 1)
constv_settle_step_ws_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleSegments seg) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    using Shape = WsShape<PREC>;
    using Stage = WsStage<PREC, true>;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    constexpr int kStages = Shape::kStages, kGroups = Shape::kGroups;
    static_assert(kStages >= kGroups + 1, "the ring must hold one tile per consumer group and one in flight");

    extern __shared__ unsigned char smemRawWs[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((size_t) smemRawWs + 127) & ~(size_t) 127);
    unsigned char* zeros = smem;
    auto stageAt = [&](int s) -> Stage& { return *reinterpret_cast<Stage*>(smem + Shape::kZeroBytes + (size_t) s * Shape::kStageBytes); };
    __shared__ __align__(8) unsigned long long fullBar[kStages], computedBar[kStages], emptyBar[kStages];
    __shared__ StepBox box;

    const int tid = (int) threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    const int tileLo = d.rangeLo, tileHi = d.rangeHi;

    // ---- set-up: zeros, barriers, the step counter ---------------------------------------------------------------
    for (int i = tid; i < (int) (Shape::kZeroBytes / 16); i += Shape::kThreads) reinterpret_cast<uint4*>(zeros)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(smem_u32(&fullBar[s]), 1);
            mbar_init(smem_u32(&computedBar[s]), kWsItems);
            mbar_init(smem_u32(&emptyBar[s]), 1);
            stageAt(s).counts[0] = 0;
            stageAt(s).counts[1] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    publish_step(box, d.counters);
    fence_proxy_async_smem();  // the zeros (generic proxy) are read by bulk stores (async proxy)
    __syncthreads();
    pdl_launch();
    const unsigned long long step = box.step;

    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;
    const Real* __restrict__ charge = static_cast<const Real*>(d.imageCharge);  // may be null: image posq records are staged instead
    const int firstTile = tileLo + (int) blockIdx.x, stride = (int) gridDim.x;

    if (warp == 0) {
        // ================================= producer =================================
        if (lane == 0) {
            int n = 0;
            for (int tile = firstTile; tile < tileHi; tile += stride, n++) {
                const int s = n % kStages;
                if (n >= kStages) mbar_wait(smem_u32(&emptyBar[s]), (unsigned int) ((n / kStages - 1) & 1));
                const WsTile t = ws_tile_of(seg, tile);
                const int lo = t.base & ~3;                                  // load window: multiples of 4 slots
                const int hi = min((t.base + t.n + 3) & ~3, (int) P);
                const unsigned int w = (unsigned int) (hi - lo);
                Stage& st = stageAt(s);
                const unsigned int bar = smem_u32(&fullBar[s]);
                unsigned int bytes = w * (unsigned int) (sizeof(Vec4<Mixed>) + sizeof(Vec4<Real>) * (kSplit ? 2 : 1) + 24);
                bytes += charge ? w * (unsigned int) sizeof(Real) : (unsigned int) t.n * (unsigned int) sizeof(Vec4<Real>);
                mbar_expect_tx(bar, bytes);
                bulk_g2s(smem_u32(st.v), velm + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Mixed>), bar);
                bulk_g2s(smem_u32(st.p), posq + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
                if (kSplit) bulk_g2s(smem_u32(st.c), corrA + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
                bulk_g2s(smem_u32(st.f[0]), force + lo, w * 8u, bar);
                bulk_g2s(smem_u32(st.f[1]), force + P + lo, w * 8u, bar);
                bulk_g2s(smem_u32(st.f[2]), force + 2 * P + lo, w * 8u, bar);
                if (charge) bulk_g2s(smem_u32(st.q1), charge + lo, w * (unsigned int) sizeof(Real), bar);
                else bulk_g2s(smem_u32(st.ip + (t.base - lo)), posq + 4 * ((size_t) R + t.base), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Real>), bar);
            }
        }
    } else if (warp == 1) {
        // ================================= storer =================================
        int n = 0;
        for (int tile = firstTile; tile < tileHi; tile += stride, n++) {
            const int s = n % kStages;
            mbar_wait(smem_u32(&computedBar[s]), (unsigned int) ((n / kStages) & 1));
            const WsTile t = ws_tile_of(seg, tile);
            const int off = t.base - (t.base & ~3);
            Stage& st = stageAt(s);
            const unsigned int moved = st.counts[0], charged = st.counts[1];
            __syncwarp();
            if (lane == 0) {
                st.counts[0] = 0;
                st.counts[1] = 0;
                const unsigned int nb = (unsigned int) t.n;
                if (moved == nb) {  // every atom of the tile moved: its records go out as they lie
                    bulk_s2g(velm + 4 * (size_t) t.base, smem_u32(st.v + off), nb * (unsigned int) sizeof(Vec4<Mixed>));
                    bulk_s2g(posq + 4 * (size_t) t.base, smem_u32(st.p + off), nb * (unsigned int) sizeof(Vec4<Real>));
                    if (kSplit) bulk_s2g(corrA + 4 * (size_t) t.base, smem_u32(st.c + off), nb * (unsigned int) sizeof(Vec4<Real>));
                }
                if (charged == nb) {  // every atom is charged: the whole range of image records is rewritten
                    bulk_s2g(posq + 4 * ((size_t) R + t.base), smem_u32(st.ip + off), nb * (unsigned int) sizeof(Vec4<Real>));
                    if (kSplit) bulk_s2g(corrA + 4 * ((size_t) R + t.base), smem_u32(st.ic + off), nb * (unsigned int) sizeof(Vec4<Real>));
                }
                if (d.zeroVelm) bulk_s2g(velm + 4 * ((size_t) R + t.base), smem_u32(zeros), nb * (unsigned int) sizeof(Vec4<Mixed>));
                if (d.zeroForce) {
                    // 8-byte elements: the 16-byte aligned interior as bulk copies, the odd edges as plain stores
                    const int a = R + t.base, b = a + t.n;
                    const int a2 = (a + 1) & ~1, b2 = b & ~1;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        long long* plane = force + (size_t) k * P;
                        if (b2 > a2) bulk_s2g(plane + a2, smem_u32(zeros), (unsigned int) (b2 - a2) * 8u);
                        if (a2 != a) __stcs(plane + a, 0LL);
                        if (b2 != b && b2 >= a2) __stcs(plane + b2, 0LL);
                    }
                }
                bulk_commit();
            }
            // tiles that are not uniform: per-slot stores from the stage (generic proxy), one warp
            if ((moved != 0 && moved != (unsigned int) t.n) || (charged != 0 && charged != (unsigned int) t.n)) {
                for (int i = lane; i < t.n; i += 32) {
                    const int idx = off + i;
                    const Vec4<Mixed> v = st.v[idx];
                    const Vec4<Real> p = st.p[idx];
                    if (moved != (unsigned int) t.n && v.w != 0) {
                        store4(velm, (size_t) (t.base + i), v);
                        store4(posq, (size_t) (t.base + i), p);
                        if (kSplit) store4(corrA, (size_t) (t.base + i), st.c[idx]);
                    }
                    if (charged != (unsigned int) t.n && p.w != -p.w) {
                        store4(posq, (size_t) R + t.base + i, st.ip[idx]);
                        if (kSplit) store4(corrA, (size_t) R + t.base + i, st.ic[idx]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && n >= 1) {
                // a stage is free once the copies that read it have done so: release the PREVIOUS tile's stage and
                // leave this tile's copies in flight while the next tile is being waited for
                bulk_wait_read<1>();
                mbar_arrive(smem_u32(&emptyBar[(n - 1) % kStages]));
            }
        }
        if (lane == 0) {
            bulk_wait_all<0>();
            // one ticket per CTA; the last CTA to finish advances the step counter
            const unsigned int ticket = atomicInc(&d.counters->done, gridDim.x - 1);
            if (ticket == gridDim.x - 1) d.counters->step = step + 1;
        }
    } else {
        // ================================= consumers =================================
        const StepScalars<Mixed> sc = scalars_of<Mixed>(d);
        const int group = (warp - 2) / (kWsItems / 32);
        const int j = tid - 64 - group * kWsItems;  // work item of the tile
        int n = 0;
        for (int tile = firstTile; tile < tileHi; tile += stride, n++) {
            if (n % kGroups != group) continue;
            const int s = n % kStages;
            const WsTile t = ws_tile_of(seg, tile);
            const int off = t.base - (t.base & ~3);
            Stage& st = stageAt(s);
            mbar_wait(smem_u32(&fullBar[s]), (unsigned int) ((n / kStages) & 1));
            unsigned int nMoved = 0, nCharged = 0;
            if (j < t.items) {
                Vec4<Mixed> v[3], delta[3];
                Vec4<Real> p[3], c[3];
                bool moved[3] = {false, false, false};
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if (k < t.per) {
                        const int idx = off + t.per * j + k;
                        v[k] = st.v[idx];
                        p[k] = st.p[idx];
                        c[k] = kSplit ? st.c[idx] : Vec4<Real>{0, 0, 0, 0};
                        delta[k] = Vec4<Mixed>{0, 0, 0, 0};
                        moved[k] = v[k].w != 0;
                        if (moved[k]) {
                            const int a = t.base + t.per * j + k;
                            const float4 xi = (sc.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                            part1_update(v[k], delta[k], st.f[0][idx], st.f[1][idx], st.f[2][idx], xi.x, xi.y, xi.z, sc);
                        }
                    }
                if (t.per == 3) {
                    Mixed pos[3][3];
#pragma unroll
                    for (int k = 0; k < 3; k++) drude_pos_load<PREC>(p[k], c[k], pos[k]);
                    settle_molecule<Mixed>(pos[0], pos[1], pos[2], delta[0], delta[1], delta[2], rcp_rn(v[0].w), rcp_rn(v[1].w),
                                           rcp_rn(v[2].w), t.prm.x, t.prm.y);
                }
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if (k < t.per) {
                        const int idx = off + t.per * j + k;
                        if (moved[k]) {
                            part2_update<PREC>(p[k], c[k], v[k], delta[k], sc.invDt);
                            st.v[idx] = v[k];
                            st.p[idx] = p[k];
                            if (kSplit) st.c[idx] = c[k];
                            nMoved++;
                        }
                        if (p[k].w != -p[k].w) {  // the image rewrite (Langevin.cu:143: neutral atoms are skipped)
                            const Real q1 = charge ? st.q1[idx] : st.ip[idx].w;
                            ImageCursor<PREC> cur(p[k], c[k]);
                            cur.advance(1, sc.zshift);
                            st.ip[idx] = cur.posq(q1);
                            if (kSplit) st.ic[idx] = cur.corr();
                            nCharged++;
                        }
                    }
            }
            nMoved = __reduce_add_sync(0xffffffffu, nMoved);
            nCharged = __reduce_add_sync(0xffffffffu, nCharged);
            if (lane == 0) {
                if (nMoved) atomicAdd(&st.counts[0], nMoved);
                if (nCharged) atomicAdd(&st.counts[1], nCharged);
            }
            fence_proxy_async_smem();  // this thread's writes to the stage, before the storer's bulk copies read them
            mbar_arrive(smem_u32(&computedBar[s]));
        }
    }
}

}  // namespace constv
