// The constrained step (part 1 + SETTLE + part 2 + image rewrite + zeroing) as a
// warp-specialised TMA pipeline for sm_100a: NO thread of the kernel issues a global load or store in the
// common case -- every byte enters and leaves the SM as a bulk asynchronous copy (cp.async.bulk, SASS UBLKCP)
// between global memory and two rings in shared memory, tracked by mbarriers.
//
// What it replaces: constv_settle_step_tiled_kernel (same tiles, same arithmetic, same bits), whose CTAs did
// load -> barrier -> compute -> barrier -> store one after the other with 8 CTAs per SM to hide each phase
// behind other CTAs'.  Here ONE persistent CTA per SM runs three roles concurrently:
//
//   producer (1 thread)    walks this CTA's tiles (round robin over the grid: the access front stays
//                          sequential across the slab, walls -- the cheap tiles -- last), waits for a free
//                          INPUT stage and issues the tile's 6-7 bulk loads: velm, posq, [posqCorrection], the
//                          three force planes, the image charges (side array);
//   consumers (NG groups   wait for the stage to fill, pull "their" work item's records into registers and hand the
//   of 128 threads)        stage straight back to the producer; run part 1, the constraints and part 2; then take
//                          an OUTPUT buffer and leave the new records and the rewritten image records in it;
//   storer (1 warp)        waits until a tile has been computed and issues its 6-10 bulk stores: own records,
//                          image records, and the zeroing of the image velocities / forces straight out of a
//                          block of zeros in shared memory; releases the buffer once the copies have read it.
//
// Every group owns ONE input stage and ONE output buffer.  The input stage is busy for a load latency plus a few
// hundred cycles and the output buffer for the time its copies take to leave -- neither for the ~9000 cycles the
// group spends on the arithmetic of a tile, during which the next tile's records arrive and the last tile's copies
// drain (first cut, profiles/r02_settle_ws_first_cut_ncu.txt: one ring of stages held from load to store starved
// the consumers).  Private stages also make every mbarrier a strict two-party alternation, which the parity
// protocol requires.
//
// A tile is up to 128 work items of one run of the slab = up to 384 consecutive slots.  Bulk copies need
// 16-byte aligned addresses and sizes: the load window is widened to multiples of 4 slots (the few extra slots
// are read and ignored); stores cover exactly the tile (16-byte records), except the 8-byte force planes, whose
// odd edge elements the storer writes with plain stores.  Tiles whose atoms are not uniformly massive /
// massless or charged / neutral (never the case for runs of ions, water and wall atoms, but legal) take a
// per-slot store path in the storer warp.
//
// Requirements checked by the host (constv_settle_capi.inc: settleWsApplies): identity slot order, work items
// in arithmetic runs, numCells == 2, in-kernel Philox noise, the image-charge side array.
#pragma once

#include "constv_settle_kernels.cuh"

namespace constv {

constexpr int kWsItems = 128;                 // work items per tile = threads per consumer group
constexpr int kWsSlots = 3 * kWsItems + 4;    // slots of a stage: a tile plus the alignment margin of the load window
constexpr int kWsProducerWarps = 1, kWsStorerWarps = 1;

template <int PREC> struct WsIn {             // what a tile reads
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    static constexpr bool kSplit = Prec<PREC>::kSplitPos;
    Vec4<Mixed> v[kWsSlots];
    Vec4<Real> p[kWsSlots];
    Vec4<Real> c[kSplit ? kWsSlots : 1];
    long long f[3][kWsSlots];
    Real q1[kWsSlots];                        // image charges (side array)
};

template <int PREC> struct WsOut {            // what a tile writes (indexed by slot - base)
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    static constexpr bool kSplit = Prec<PREC>::kSplitPos;
    Vec4<Mixed> v[3 * kWsItems];
    Vec4<Real> p[3 * kWsItems];
    Vec4<Real> c[kSplit ? 3 * kWsItems : 1];
    Vec4<Real> ip[3 * kWsItems];              // image posq records
    Vec4<Real> ic[kSplit ? 3 * kWsItems : 1]; // image posqCorrection records
    unsigned int counts[4];                   // [0] atoms that moved, [1] atoms that are charged (consumers -> storer)
};

// Consumer groups x ring depths.  The arithmetic of an item is one long dependent chain (~1400 instructions, 25
// IEEE divisions / square roots) that wants ~100 registers: 4 groups (16 warps) in single precision, 2 otherwise.
template <int PREC, int GROUPS> struct WsShape {
    static constexpr int kGroups = GROUPS;
    static constexpr int kThreads = 32 * (kWsProducerWarps + kWsStorerWarps) + kWsItems * kGroups;
    static constexpr size_t kInBytes = (sizeof(WsIn<PREC>) + 127) / 128 * 128;
    static constexpr size_t kOutBytes = (sizeof(WsOut<PREC>) + 127) / 128 * 128;
    static constexpr size_t kZeroBytes = (3 * kWsItems * sizeof(Vec4<typename Prec<PREC>::Mixed>) + 127) / 128 * 128;
    // one input stage and one output buffer PER GROUP: every mbarrier then has one party on each side in strict
    // alternation, which is what the parity protocol needs (a waiter may be at most one phase ahead; on a ring shared
    // by several groups nothing guarantees that).  One stage each way is also all a group needs: the next tile's
    // records arrive and the last tile's copies leave while it spends ~9000 cycles on the current one.
    static constexpr int kIn = GROUPS, kOut = GROUPS;
    static constexpr size_t kSmemBytes = kZeroBytes + kIn * kInBytes + kOut * kOutBytes + 128;  // + alignment slack
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

// Part 1, the constraints and part 2 for one work item whose records are in registers (v, p, c, f); leaves the new
// records in v / p / c and returns in `moved` which atoms moved.  Same operations as the LDG-staged kernel: same bits.
template <int PREC>
__device__ __forceinline__ void ws_compute_item(const SlabDev& d, const StepScalars<typename Prec<PREC>::Mixed>& sc, const WsTile& t,
                                                const int j, const unsigned long long step,
                                                Vec4<typename Prec<PREC>::Mixed> (&v)[3], Vec4<typename Prec<PREC>::Real> (&p)[3],
                                                Vec4<typename Prec<PREC>::Real> (&c)[3], const long long (&f)[3][3], bool (&moved)[3]) {
    using Mixed = typename Prec<PREC>::Mixed;
    const bool three = t.per == 3;  // uniform over the tile
    const unsigned long long a0 = d.globalAtomOffset + (unsigned long long) (t.base + t.per * j);
    Vec4<Mixed> delta[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (k == 0 || three) {
            delta[k] = Vec4<Mixed>{0, 0, 0, 0};
            moved[k] = v[k].w != 0;
            if (moved[k]) {
                const float4 xi = (sc.noisescale != 0) ? gaussian4<false>(d.seed, a0 + k, step) : make_float4(0.f, 0.f, 0.f, 0.f);
                part1_update(v[k], delta[k], f[k][0], f[k][1], f[k][2], xi.x, xi.y, xi.z, sc);
            }
        }
    if (three) {
        Mixed pos[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++) drude_pos_load<PREC>(p[k], c[k], pos[k]);
        settle_molecule<Mixed>(pos[0], pos[1], pos[2], delta[0], delta[1], delta[2], rcp_rn(v[0].w), rcp_rn(v[1].w),
                               rcp_rn(v[2].w), t.prm.x, t.prm.y);
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
        if ((k == 0 || three) && moved[k]) part2_update<PREC>(p[k], c[k], v[k], delta[k], sc.invDt);
}

template <int PREC, int GROUPS>
__global__ void __launch_bounds__(WsShape<PREC, GROUPS>::kThreads, 1)
constv_settle_step_ws_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleSegments seg) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    using Shape = WsShape<PREC, GROUPS>;
    using In = WsIn<PREC>;
    using Out = WsOut<PREC>;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    constexpr int kIn = Shape::kIn, kOut = Shape::kOut, kGroups = Shape::kGroups;
    static_assert(Shape::kSmemBytes <= 227 * 1024 - 1024, "stages of all groups must fit in shared memory");

    extern __shared__ unsigned char smemRawWs[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((size_t) smemRawWs + 127) & ~(size_t) 127);
    unsigned char* zeros = smem;
    auto inAt = [&](int s) -> In& { return *reinterpret_cast<In*>(smem + Shape::kZeroBytes + (size_t) s * Shape::kInBytes); };
    auto outAt = [&](int s) -> Out& {
        return *reinterpret_cast<Out*>(smem + Shape::kZeroBytes + (size_t) kIn * Shape::kInBytes + (size_t) s * Shape::kOutBytes);
    };
    __shared__ __align__(8) unsigned long long fullIn[kIn], emptyIn[kIn], computedOut[kOut], emptyOut[kOut];
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
        for (int s = 0; s < kIn; s++) {
            mbar_init(smem_u32(&fullIn[s]), 1);
            mbar_init(smem_u32(&emptyIn[s]), kWsItems);
        }
#pragma unroll
        for (int s = 0; s < kOut; s++) {
            mbar_init(smem_u32(&computedOut[s]), kWsItems);
            mbar_init(smem_u32(&emptyOut[s]), 1);
            outAt(s).counts[0] = 0;
            outAt(s).counts[1] = 0;
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
    const Real* __restrict__ charge = static_cast<const Real*>(d.imageCharge);  // the side array (the host checks it is there)
    const int firstTile = tileLo + (int) blockIdx.x, stride = (int) gridDim.x;

    if (warp == 0) {
        // ================================= producer =================================
        if (lane == 0) {
            int n = 0;
            for (int tile = firstTile; tile < tileHi; tile += stride, n++) {
                const int s = n % kGroups, use = n / kGroups;  // the group's own stage, its use-th tile
                if (use >= 1) mbar_wait(smem_u32(&emptyIn[s]), (unsigned int) ((use - 1) & 1));
                const WsTile t = ws_tile_of(seg, tile);
                const int lo = t.base & ~3;                                  // load window: multiples of 4 slots
                const int hi = min((t.base + t.n + 3) & ~3, (int) P);
                const unsigned int w = (unsigned int) (hi - lo);
                In& st = inAt(s);
                const unsigned int bar = smem_u32(&fullIn[s]);
                const unsigned int bytes = w * (unsigned int) (sizeof(Vec4<Mixed>) + sizeof(Vec4<Real>) * (kSplit ? 2 : 1) + 24 + sizeof(Real));
                mbar_expect_tx(bar, bytes);
                bulk_g2s(smem_u32(st.v), velm + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Mixed>), bar);
                bulk_g2s(smem_u32(st.p), posq + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
                if (kSplit) bulk_g2s(smem_u32(st.c), corrA + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
                bulk_g2s(smem_u32(st.f[0]), force + lo, w * 8u, bar);
                bulk_g2s(smem_u32(st.f[1]), force + P + lo, w * 8u, bar);
                bulk_g2s(smem_u32(st.f[2]), force + 2 * P + lo, w * 8u, bar);
                bulk_g2s(smem_u32(st.q1), charge + lo, w * (unsigned int) sizeof(Real), bar);
            }
        }
    } else if (warp == 1) {
        // ================================= storer =================================
        int n = 0;
        for (int tile = firstTile; tile < tileHi; tile += stride, n++) {
            const int s = n % kGroups, use = n / kGroups;
            mbar_wait(smem_u32(&computedOut[s]), (unsigned int) (use & 1));
            const WsTile t = ws_tile_of(seg, tile);
            Out& st = outAt(s);
            const unsigned int moved = st.counts[0], charged = st.counts[1];
            __syncwarp();
            if (lane == 0) {
                st.counts[0] = 0;
                st.counts[1] = 0;
                const unsigned int nb = (unsigned int) t.n;
                if (moved == nb) {  // every atom of the tile moved: its records go out as they lie
                    bulk_s2g(velm + 4 * (size_t) t.base, smem_u32(st.v), nb * (unsigned int) sizeof(Vec4<Mixed>));
                    bulk_s2g(posq + 4 * (size_t) t.base, smem_u32(st.p), nb * (unsigned int) sizeof(Vec4<Real>));
                    if (kSplit) bulk_s2g(corrA + 4 * (size_t) t.base, smem_u32(st.c), nb * (unsigned int) sizeof(Vec4<Real>));
                }
                if (charged == nb) {  // every atom is charged: the whole range of image records is rewritten
                    bulk_s2g(posq + 4 * ((size_t) R + t.base), smem_u32(st.ip), nb * (unsigned int) sizeof(Vec4<Real>));
                    if (kSplit) bulk_s2g(corrA + 4 * ((size_t) R + t.base), smem_u32(st.ic), nb * (unsigned int) sizeof(Vec4<Real>));
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
            // tiles that are not uniform: per-slot stores from the buffer (generic proxy), one warp
            if ((moved != 0 && moved != (unsigned int) t.n) || (charged != 0 && charged != (unsigned int) t.n)) {
                for (int i = lane; i < t.n; i += 32) {
                    const Vec4<Mixed> v = st.v[i];
                    const Vec4<Real> p = st.p[i];
                    if (moved != (unsigned int) t.n && v.w != 0) {
                        store4(velm, (size_t) (t.base + i), v);
                        store4(posq, (size_t) (t.base + i), p);
                        if (kSplit) store4(corrA, (size_t) (t.base + i), st.c[i]);
                    }
                    if (charged != (unsigned int) t.n && p.w != -p.w) {
                        store4(posq, (size_t) R + t.base + i, st.ip[i]);
                        if (kSplit) store4(corrA, (size_t) R + t.base + i, st.ic[i]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && n >= 1) {
                // a buffer is free once the copies that read it have done so: release the PREVIOUS tile's buffer and
                // leave this tile's copies in flight while the next tile is being waited for
                bulk_wait_read<1>();
                mbar_arrive(smem_u32(&emptyOut[(n - 1) % kGroups]));
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
            const int si = group, so = group, use = n / kGroups;
            const WsTile t = ws_tile_of(seg, tile);
            const int off = t.base - (t.base & ~3);
            const bool active = j < t.items;
            const bool three = t.per == 3;  // uniform over the tile
            Vec4<Mixed> v[3];
            Vec4<Real> p[3], c[3];
            long long f[3][3];
            Real q1[3];
            {   // records -> registers, and the input stage goes back to the producer
                In& in = inAt(si);
                mbar_wait(smem_u32(&fullIn[si]), (unsigned int) (use & 1));
                if (active) {
                    const int idx0 = off + t.per * j;
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        if (k == 0 || three) {
                            v[k] = in.v[idx0 + k];
                            p[k] = in.p[idx0 + k];
                            c[k] = kSplit ? in.c[idx0 + k] : Vec4<Real>{0, 0, 0, 0};
                            f[k][0] = in.f[0][idx0 + k];
                            f[k][1] = in.f[1][idx0 + k];
                            f[k][2] = in.f[2][idx0 + k];
                            q1[k] = in.q1[idx0 + k];
                        }
                }
                mbar_arrive(smem_u32(&emptyIn[si]));
            }
            bool moved[3] = {false, false, false};
            if (active) ws_compute_item<PREC>(d, sc, t, j, step, v, p, c, f, moved);
            // results -> an output buffer (free once the copies of the tile that used it before have read it)
            Out& out = outAt(so);
            if (use >= 1) mbar_wait(smem_u32(&emptyOut[so]), (unsigned int) ((use - 1) & 1));
            unsigned int nMoved = 0, nCharged = 0;
            if (active) {
                const int o0 = t.per * j;
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if (k == 0 || three) {
                        out.v[o0 + k] = v[k];
                        out.p[o0 + k] = p[k];
                        if (kSplit) out.c[o0 + k] = c[k];
                        nMoved += moved[k] ? 1u : 0u;
                        if (p[k].w != -p[k].w) {  // the image rewrite (Langevin.cu:143: neutral atoms are skipped)
                            ImageCursor<PREC> cur(p[k], c[k]);
                            cur.advance(1, sc.zshift);
                            out.ip[o0 + k] = cur.posq(q1[k]);
                            if (kSplit) out.ic[o0 + k] = cur.corr();
                            nCharged++;
                        }
                    }
            }
            nMoved = __reduce_add_sync(0xffffffffu, nMoved);
            nCharged = __reduce_add_sync(0xffffffffu, nCharged);
            if (lane == 0) {
                if (nMoved) atomicAdd(&out.counts[0], nMoved);
                if (nCharged) atomicAdd(&out.counts[1], nCharged);
            }
            fence_proxy_async_smem();  // this thread's writes to the buffer, before the storer's bulk copies read them
            mbar_arrive(smem_u32(&computedOut[so]));
        }
    }
}

// ---- the staged step with a bulk stage-in (kernel variant 6) ------------------------------------------------------------
// The shape of constv_settle_step_tiled_kernel -- one tile of 128 work items per CTA, 128 threads, eight CTAs per SM in
// single precision, every thread computes one work item -- with the staging of the pipeline above: the tile enters shared
// memory as 6-7 bulk copies issued by ONE thread (same 4-slot-aligned window as the producer's) and, when every atom
// of the tile moved, the new velm / posq records leave as bulk stores.  What it removes is the per-thread staging code:
// 18 asynchronous copies with their address arithmetic in front of every tile (~220 of the ~1050 instructions a warp
// issues per tile) and two of the seven stores per slot behind it; the arithmetic and the image / zeroing stores are
// the staged kernel's.  Same requirements as the pipeline (settleWsApplies).  One tile per CTA, so the mbarrier is used
// once (phase 0) and needs no ring.
// INDEXED = reordered slots with the per-slot image tables (SlabDev::imageSlotBySlot): the tile's atomIndex, image-slot and
// image-charge entries come in as three more bulk copies (everything a tile reads is in slot order), the Philox key
// is atomIndex[slot], the image records go to the tables' slots and the image slots are zeroed in slot order behind
// the loads, as in the staged kernel.
template <int PREC, bool INDEXED> struct BulkIn : WsIn<PREC> {
    int orig[INDEXED ? kWsSlots : 1];  // original index of the atom in the slot
    int img[INDEXED ? kWsSlots : 1];   // slot of its cell-1 image, -1 = the slot holds no real atom
};

// EARLY: a molecule's three Philox blocks (functions of seed, atom index and step only) are drawn while the tile's bulk
// copies are in flight -- nothing else is live in the thread at that point, and no CTA barrier lies between the draws
// and their use (the cp.async kernel's early draws had to cross one: DESIGN.md 3c).  Runs of single atoms draw late.
template <int PREC, bool INDEXED, bool EARLY = false>
__global__ void __launch_bounds__(kWsItems, PREC == PREC_SINGLE ? 8 : 4)
constv_settle_step_bulk_kernel(const __grid_constant__ SlabDev d, const __grid_constant__ SettleSegments seg) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    using In = BulkIn<PREC, INDEXED>;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;
    extern __shared__ unsigned char smemRawBulk[];
    In& st = *reinterpret_cast<In*>(((size_t) smemRawBulk + 127) & ~(size_t) 127);
    __shared__ __align__(8) unsigned long long full;
    __shared__ StepBox box;
    __shared__ unsigned int counts[1];  // atoms of the tile that moved

    const int j = (int) threadIdx.x;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    const unsigned int holders = (unsigned int) (d.rangeHi - d.rangeLo);
    const WsTile t = ws_tile_of(seg, d.rangeLo + (int) blockIdx.x);
    const int lo = t.base & ~3;  // load window: multiples of 4 slots
    const int off = t.base - lo;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;

    if (j == 0) {
        mbar_init(smem_u32(&full), 1);
        counts[0] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    pdl_launch();
    if (j == 0) {
        // image charges: the side array, by atom (identity order) or by slot (the host checks they are there)
        const Real* __restrict__ charge = static_cast<const Real*>(INDEXED ? d.imageChargeBySlot : d.imageCharge);
        const int hi = min((t.base + t.n + 3) & ~3, (int) P);
        const unsigned int w = (unsigned int) (hi - lo);
        const unsigned int bar = smem_u32(&full);
        mbar_expect_tx(bar, w * (unsigned int) (sizeof(Vec4<Mixed>) + sizeof(Vec4<Real>) * (kSplit ? 2 : 1) + 24 + sizeof(Real) + (INDEXED ? 8 : 0)));
        bulk_g2s(smem_u32(st.v), velm + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Mixed>), bar);
        bulk_g2s(smem_u32(st.p), posq + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
        if (kSplit) bulk_g2s(smem_u32(st.c), corrA + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
        bulk_g2s(smem_u32(st.f[0]), force + lo, w * 8u, bar);
        bulk_g2s(smem_u32(st.f[1]), force + P + lo, w * 8u, bar);
        bulk_g2s(smem_u32(st.f[2]), force + 2 * P + lo, w * 8u, bar);
        bulk_g2s(smem_u32(st.q1), charge + lo, w * (unsigned int) sizeof(Real), bar);
        if (INDEXED) {
            bulk_g2s(smem_u32(st.orig), d.atomIndex + lo, w * 4u, bar);
            bulk_g2s(smem_u32(st.img), d.imageSlotBySlot + lo, w * 4u, bar);
        }
    }
    if (INDEXED) {  // the image slots, in slot order: coalesced whatever atoms they hold, and dependent on nothing
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (j + i * kWsItems < t.n) emit_zero<PREC, true>(d, t.base + j + i * kWsItems);
    }
    publish_step(box, d.counters);
    __syncthreads();  // the barrier's initialisation, the counts and the step index are visible to every thread
    const unsigned long long step = box.step;
    const unsigned int ticket = take_ticket(d.counters, holders);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    float4 early[3];
    const bool drawn = EARLY && !INDEXED && t.per == 3 && s.noisescale != 0 && j < t.items;
    if (EARLY && drawn) {
#pragma unroll
        for (int k = 0; k < 3; k++)
            early[k] = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) (t.base + 3 * j + k), step);
    }
    mbar_wait(smem_u32(&full), 0u);

    // compute item j in place (the text of the staged kernel's compute phase)
    unsigned int nMoved = 0;
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
                    nMoved++;
                    const int a = INDEXED ? st.orig[idx] : t.base + t.per * j + k;
                    float4 xi;
                    if (EARLY && drawn) xi = early[k];
                    else xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                    part1_update(v[k], delta[k], st.f[0][idx], st.f[1][idx], st.f[2][idx], xi.x, xi.y, xi.z, s);
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
            if (k < t.per && moved[k]) {
                const int idx = off + t.per * j + k;
                part2_update<PREC>(p[k], c[k], v[k], delta[k], s.invDt);
                st.v[idx] = v[k];
                st.p[idx] = p[k];
                if (kSplit) st.c[idx] = c[k];
            }
    }
    nMoved = __reduce_add_sync(0xffffffffu, nMoved);
    if ((j & 31) == 0 && nMoved) atomicAdd(&counts[0], nMoved);
    fence_proxy_async_smem();  // this thread's writes to the stage, before the bulk stores read them
    __syncthreads();

    // stage out.  Every atom moved (runs of ions and of water): the own records go out as they lie, two or three bulk
    // stores by one thread; otherwise (wall atoms do not move: nothing to store; mixed tiles: legal, rare) per slot.
    const bool allMoved = counts[0] == (unsigned int) t.n;
    if (j == 0 && allMoved) {
        bulk_s2g(velm + 4 * (size_t) t.base, smem_u32(&st.v[off]), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Mixed>));
        bulk_s2g(posq + 4 * (size_t) t.base, smem_u32(&st.p[off]), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Real>));
        if (kSplit) bulk_s2g(corrA + 4 * (size_t) t.base, smem_u32(&st.c[off]), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Real>));
        bulk_commit();
    }
    const bool someMoved = counts[0] != 0 && !allMoved;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int idx = j + i * kWsItems;
        if (idx < t.n) {
            const int slot = t.base + idx;
            if (someMoved) {
                const Vec4<Mixed> v = st.v[off + idx];
                if (v.w != 0) {
                    store4x<true>(velm, (size_t) slot, v);
                    store4x<true>(posq, (size_t) slot, st.p[off + idx]);
                    if (kSplit) store4x<true>(corrA, (size_t) slot, st.c[off + idx]);
                }
            }
            const Vec4<Real> p = st.p[off + idx];
            if (INDEXED) {
                // the image records go where the tables say (every cell); the own record went out above
                drude_emit_tables<PREC>(d, slot, false, Vec4<Mixed>{0, 0, 0, 0}, p, kSplit ? st.c[off + idx] : Vec4<Real>{0, 0, 0, 0},
                                        st.img[off + idx], st.q1[off + idx]);
            } else {
                if (p.w != -p.w) {  // the image rewrite (constVLangevin.cu:143: neutral atoms are skipped)
                    ImageCursor<PREC> cur(p, kSplit ? st.c[off + idx] : Vec4<Real>{0, 0, 0, 0});
                    cur.advance(1, s.zshift);
                    store4x<true>(posq, (size_t) slot + R, cur.posq(st.q1[off + idx]));
                    if (kSplit) store4x<true>(corrA, (size_t) slot + R, cur.corr());
                }
                emit_zero<PREC, true>(d, slot);
            }
        }
    }
    close_step(d.counters, ticket, holders, step);
    if (j == 0 && allMoved) bulk_wait_read<0>();  // the stage must outlive the bulk stores' reads of it
}

// The same tile as a device function, for the ensemble kernel below (`full`, `box`, `counts` and the stage are the CTA's
// own).  The single-slab kernel above keeps its own copy of this text: calling the function from it, though inlined,
// costs 16 B of stack under the 64-register cap (as with settle_tiled_tile and the staged kernel).
template <int PREC, bool INDEXED, bool EARLY>
__device__ __forceinline__ void settle_bulk_tile(const SlabDev& d, const SettleSegments& seg, BulkIn<PREC, INDEXED>& st,
                                                 unsigned long long& full, StepBox& box, unsigned int (&counts)[1], const int tileIndex) {
    using Real = typename Prec<PREC>::Real;
    using Mixed = typename Prec<PREC>::Mixed;
    constexpr bool kSplit = Prec<PREC>::kSplitPos;

    const int j = (int) threadIdx.x;
    const int R = d.numReal;
    const size_t P = (size_t) d.padded;
    const unsigned int holders = (unsigned int) (d.rangeHi - d.rangeLo);
    const WsTile t = ws_tile_of(seg, tileIndex);
    const int lo = t.base & ~3;  // load window: multiples of 4 slots
    const int off = t.base - lo;
    Real* __restrict__ posq = static_cast<Real*>(d.posq);
    Real* __restrict__ corrA = static_cast<Real*>(d.corr);
    Mixed* __restrict__ velm = static_cast<Mixed*>(d.velm);
    long long* __restrict__ force = d.force;

    if (j == 0) {
        mbar_init(smem_u32(&full), 1);
        counts[0] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    pdl_launch();
    if (j == 0) {
        // image charges: the side array, by atom (identity order) or by slot (the host checks they are there)
        const Real* __restrict__ charge = static_cast<const Real*>(INDEXED ? d.imageChargeBySlot : d.imageCharge);
        const int hi = min((t.base + t.n + 3) & ~3, (int) P);
        const unsigned int w = (unsigned int) (hi - lo);
        const unsigned int bar = smem_u32(&full);
        mbar_expect_tx(bar, w * (unsigned int) (sizeof(Vec4<Mixed>) + sizeof(Vec4<Real>) * (kSplit ? 2 : 1) + 24 + sizeof(Real) + (INDEXED ? 8 : 0)));
        bulk_g2s(smem_u32(st.v), velm + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Mixed>), bar);
        bulk_g2s(smem_u32(st.p), posq + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
        if (kSplit) bulk_g2s(smem_u32(st.c), corrA + 4 * (size_t) lo, w * (unsigned int) sizeof(Vec4<Real>), bar);
        bulk_g2s(smem_u32(st.f[0]), force + lo, w * 8u, bar);
        bulk_g2s(smem_u32(st.f[1]), force + P + lo, w * 8u, bar);
        bulk_g2s(smem_u32(st.f[2]), force + 2 * P + lo, w * 8u, bar);
        bulk_g2s(smem_u32(st.q1), charge + lo, w * (unsigned int) sizeof(Real), bar);
        if (INDEXED) {
            bulk_g2s(smem_u32(st.orig), d.atomIndex + lo, w * 4u, bar);
            bulk_g2s(smem_u32(st.img), d.imageSlotBySlot + lo, w * 4u, bar);
        }
    }
    if (INDEXED) {  // the image slots, in slot order: coalesced whatever atoms they hold, and dependent on nothing
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (j + i * kWsItems < t.n) emit_zero<PREC, true>(d, t.base + j + i * kWsItems);
    }
    publish_step(box, d.counters);
    __syncthreads();  // the barrier's initialisation, the counts and the step index are visible to every thread
    const unsigned long long step = box.step;
    const unsigned int ticket = take_ticket(d.counters, holders);
    const StepScalars<Mixed> s = scalars_of<Mixed>(d);
    float4 early[3];
    const bool drawn = EARLY && !INDEXED && t.per == 3 && s.noisescale != 0 && j < t.items;
    if (EARLY && drawn) {
#pragma unroll
        for (int k = 0; k < 3; k++)
            early[k] = gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) (t.base + 3 * j + k), step);
    }
    mbar_wait(smem_u32(&full), 0u);

    // compute item j in place (the text of the staged kernel's compute phase)
    unsigned int nMoved = 0;
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
                    nMoved++;
                    const int a = INDEXED ? st.orig[idx] : t.base + t.per * j + k;
                    float4 xi;
                    if (EARLY && drawn) xi = early[k];
                    else xi = (s.noisescale != 0) ? gaussian4<false>(d.seed, d.globalAtomOffset + (unsigned long long) a, step)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                    part1_update(v[k], delta[k], st.f[0][idx], st.f[1][idx], st.f[2][idx], xi.x, xi.y, xi.z, s);
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
            if (k < t.per && moved[k]) {
                const int idx = off + t.per * j + k;
                part2_update<PREC>(p[k], c[k], v[k], delta[k], s.invDt);
                st.v[idx] = v[k];
                st.p[idx] = p[k];
                if (kSplit) st.c[idx] = c[k];
            }
    }
    nMoved = __reduce_add_sync(0xffffffffu, nMoved);
    if ((j & 31) == 0 && nMoved) atomicAdd(&counts[0], nMoved);
    fence_proxy_async_smem();  // this thread's writes to the stage, before the bulk stores read them
    __syncthreads();

    // stage out.  Every atom moved (runs of ions and of water): the own records go out as they lie, two or three bulk
    // stores by one thread; otherwise (wall atoms do not move: nothing to store; mixed tiles: legal, rare) per slot.
    const bool allMoved = counts[0] == (unsigned int) t.n;
    if (j == 0 && allMoved) {
        bulk_s2g(velm + 4 * (size_t) t.base, smem_u32(&st.v[off]), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Mixed>));
        bulk_s2g(posq + 4 * (size_t) t.base, smem_u32(&st.p[off]), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Real>));
        if (kSplit) bulk_s2g(corrA + 4 * (size_t) t.base, smem_u32(&st.c[off]), (unsigned int) t.n * (unsigned int) sizeof(Vec4<Real>));
        bulk_commit();
    }
    const bool someMoved = counts[0] != 0 && !allMoved;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int idx = j + i * kWsItems;
        if (idx < t.n) {
            const int slot = t.base + idx;
            if (someMoved) {
                const Vec4<Mixed> v = st.v[off + idx];
                if (v.w != 0) {
                    store4x<true>(velm, (size_t) slot, v);
                    store4x<true>(posq, (size_t) slot, st.p[off + idx]);
                    if (kSplit) store4x<true>(corrA, (size_t) slot, st.c[off + idx]);
                }
            }
            const Vec4<Real> p = st.p[off + idx];
            if (INDEXED) {
                // the image records go where the tables say (every cell); the own record went out above
                drude_emit_tables<PREC>(d, slot, false, Vec4<Mixed>{0, 0, 0, 0}, p, kSplit ? st.c[off + idx] : Vec4<Real>{0, 0, 0, 0},
                                        st.img[off + idx], st.q1[off + idx]);
            } else {
                if (p.w != -p.w) {  // the image rewrite (constVLangevin.cu:143: neutral atoms are skipped)
                    ImageCursor<PREC> cur(p, kSplit ? st.c[off + idx] : Vec4<Real>{0, 0, 0, 0});
                    cur.advance(1, s.zshift);
                    store4x<true>(posq, (size_t) slot + R, cur.posq(st.q1[off + idx]));
                    if (kSplit) store4x<true>(corrA, (size_t) slot + R, cur.corr());
                }
                emit_zero<PREC, true>(d, slot);
            }
        }
    }
    close_step(d.counters, ticket, holders, step);
    if (j == 0 && allMoved) bulk_wait_read<0>();  // the stage must outlive the bulk stores' reads of it
}

// An ensemble of slabs (identity slot order, Philox noise, image-charge side arrays) in ONE launch, one tile per CTA: the
// CTA looks its tile up in the table of slabs (constv_settle_step_tiled_batched_kernel's table) and runs it with that
// slab's descriptor, runs and step counter.
template <int PREC, int CAP>
__global__ void __launch_bounds__(kWsItems, PREC == PREC_SINGLE ? 8 : 4)
constv_settle_step_bulk_batched_kernel(const __grid_constant__ SettleBatchParams<CAP> b) {
    extern __shared__ unsigned char smemRawBulkB[];
    BulkIn<PREC, false>& st = *reinterpret_cast<BulkIn<PREC, false>*>(((size_t) smemRawBulkB + 127) & ~(size_t) 127);
    __shared__ __align__(8) unsigned long long full;
    __shared__ StepBox box;
    __shared__ unsigned int counts[1];
    const int tile = (int) blockIdx.x;
    int k = 0;
#pragma unroll
    for (int q = 1; q < CAP; q++)
        if (q < b.numSlabs && tile >= b.tileBase[q]) k = q;
    settle_bulk_tile<PREC, false, PREC != PREC_SINGLE>(b.slab[k], b.seg[k], st, full, box, counts, tile - b.tileBase[k]);
}

}  // namespace constv
