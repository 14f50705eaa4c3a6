// Host side of the C ABI declared in include/constv_b200.h: argument checking, parameter
// computation, launch configuration and the CUDA runtime calls.  The reference code this replaces
// is platforms/cuda/src/CudaConstVKernels.cpp:53-177 (scychon/openmm_constV); every entry point
// cites its lines in the header.  No CPU fallback exists: without a device the calls fail.
#include "constv_b200.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "constv_kernels.cuh"
#include "constv_drude_kernels.cuh"
#include "constv_settle_kernels.cuh"
#include "constv_settle_ws_kernels.cuh"

using namespace constv;

namespace {

thread_local std::string g_lastError;
thread_local const char* g_lastKernel = "";  // constv_last_kernel: which kernel the calling thread's last launch was
std::atomic<uint64_t> g_launchCount{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return fail(CONSTV_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(err__));    \
    } while (0)

constexpr double kBoltzDefault = 8.31446261815324e-3;  // OpenMM SimTKOpenMMRealType.h BOLTZ (>= 7.4)
constexpr int kBlock = 256;
constexpr int kBatchMax = 128;  // slabs per batched launch (descriptors travel as kernel parameters)
constexpr int kMaxChains = 16;  // independent real-atom ranges of one slab (constv_set_chains)

size_t realSize(int prec) { return prec == CONSTV_PRECISION_DOUBLE ? 8 : 4; }
size_t mixedSize(int prec) { return prec == CONSTV_PRECISION_SINGLE ? 4 : 8; }

}  // namespace

struct constv_slab {
    constv_config cfg{};
    int numReal = 0;
    int numSMs = 0;
    // bound arrays
    void* posq = nullptr;
    void* corr = nullptr;
    void* velm = nullptr;
    void* force = nullptr;
    void* posDelta = nullptr;
    const int32_t* atomIndex = nullptr;
    bool owned = false;
    // library-owned device state
    int* invAtomOrder = nullptr;
    void* imageCharge = nullptr;
    bool imageChargeValid = false;
    // reordered slots: image slot and image charge per real SLOT, for the staged kernels (constv_image_by_slot_kernel)
    int* imageSlotBySlot = nullptr;
    void* imageChargeBySlot = nullptr;
    bool imageBySlotValid = false;
    bool imageTables = true;  // CONSTV_IMAGE_TABLES=0: never use them (comparison)
    DevCounters* counters = nullptr;
    double* kePartials = nullptr;
    double* keOut = nullptr;
    double* keHost = nullptr;  // pinned
    std::vector<SlabDev> batchScratch;  // descriptor table of the ensemble launch led by this slab
    // parameters
    bool paramsSet = false;
    double temperature = 0, friction = 0, stepSize = -1;
    double vscale = 0, fscaleFixed = 0, noisescale = 0, dt = 0, invDt = 0;
    double time = 0;
    // launch shape
    int ilp = 1;
    int fusedBlock = 256;    // threads per CTA of the fused step (128 | 256); experiments: CONSTV_FUSED_BLOCK
    // fused step: the Philox blocks are drawn and the image zeroing stores issued while the tile's loads are in flight
    // (in-process A/B on a B200, profiles/r02_ab_fused.jsonl: 13.76 -> 13.46 us with one launch per step, 12.27 ->
    // 11.64 us with two chains).  CONSTV_FUSED_EARLY=0 / CONSTV_FUSED_ZERO=0|2 restore the round-1 order.
    bool fusedEarly = true;
    int zeroWhen = 1;         // 0 with the dependent stores, 1 right behind the loads, 2 after the step-counter barrier
    bool settleWs = false;    // rigid-water step: the warp-specialised TMA pipeline instead of the LDG-staged kernel (CONSTV_SETTLE_WS=1, or kernel variant 4)
    // staged Drude step: slots per thread.  2 | 3 = the balanced kernel (warps draw chunks of work items: 17.4 -> 15.2 us per step at 1M atoms
    // with two chains, profiles/r02_ab_drude_balanced.jsonl); 1 = thread t takes work item t (CONSTV_DRUDE_SPT)
    int drudeSlotsPerThread = 2;
    int indexedMinBlocks = 0; // reordered-slot fused kernel, single precision: 8 = capped at 32 registers (experiment: CONSTV_INDEXED_MINB=8)
    // bulk-staged rigid-water step: a molecule's Philox draws made while the bulk copies fly.  -1 = automatic: on in mixed / double
    // precision (128 registers: 26.8 -> 25.8 us per step at 1M atoms with one launch, 20.8 -> 20.2 with three chains), off in single
    // (64-register cap: 24 B of stack, 12.7 -> 13.4 us with six chains; profiles/r02_ab_settle_bulk_early.jsonl).  CONSTV_SETTLE_EARLY=0|1 forces.
    int settleEarly = -1;
    int wsGroups = 4;         // its consumer groups in single precision (CONSTV_WS_GROUPS=3..5: experiments)
    int ctasPerSM = 0;   // 0 = one-shot grid (one tile per CTA); > 0 = persistent grid of that many CTAs per SM
    int variant = 0;     // 0 auto (= the one-shot LDG kernel, the faster one as measured), 1 LDG, 2 TMA ring
    int tmaBlock = 256;  // pairs per chunk = threads per CTA of the TMA kernel
    int tmaStages = 4;
    int tmaCtasPerSM = 0;  // 0 = as many as shared memory allows
    // ConstVDrudeLangevinIntegrator (constv_drude_*): per-slot roles, the pair list, the second thermostat
    struct Drude {
        int* role = nullptr;     // device int[numReal]
        int2* pairs = nullptr;   // device int2[numPairs]
        int numNormal = 0, numPairs = 0;
        bool configured = false;
        double temperature = 0, friction = 0;
        double vscale = 0, fscaleFixed = 0, noisescale = 0, maxDistance = 0, hardwallScale = 0;
    } drude;
    // rigid three-site molecules solved inside the step (constv_settle_*)
    struct Settle {
        int2* items = nullptr;     // device int2[numItems]: (slot, -1 | molecule index), slot order
        int4* atoms = nullptr;     // device int4[numClusters]: (apex, b, c, 0)
        float2* params = nullptr;  // device float2[numClusters]
        int numItems = 0, numClusters = 0;
        SettleSegments segments{};  // count > 0: the work items form a few arithmetic runs
        bool configured = false;
    } settle;
    int paramsKind = 0;  // which integrator's coefficients vscale/fscaleFixed/noisescale hold: 1 Langevin, 2 Drude
    // chains (constv_set_chains): chain q advances real atoms [chainLo[q], chainLo[q+1]) with counters[q]
    int chains = 1;
    int chainLo[kMaxChains + 1] = {0};
    // constv_step_fused_multi: one side stream per chain beyond the first, fork/join events
    cudaStream_t chainStream[kMaxChains] = {nullptr};
    cudaEvent_t chainFork = nullptr, chainJoin[kMaxChains] = {nullptr};
    // pipelined host stepping (constv_step_host_submit / _collect): two slots
    struct HostPipe {
        bool ready = false;
        cudaStream_t copyStream = nullptr;
        void* stage[2] = {nullptr, nullptr};
        cudaEvent_t h2dDone[2] = {nullptr, nullptr}, stageFree[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
        double* keDev = nullptr;   // 2 doubles
        double* keHost = nullptr;  // 2 doubles, pinned
        uint64_t submitted = 0, collected = 0;
        // constv_step_host_submit_state: read-back of posq / velm on its own stream (the other PCIe direction)
        cudaStream_t readStream = nullptr;
        cudaEvent_t stepDone[2] = {nullptr, nullptr}, readDone[2] = {nullptr, nullptr};
        bool hasRead[2] = {false, false};
        bool readOutstanding = false;  // a read-back of the bound arrays may still be running: the next step waits for it
        cudaEvent_t lastRead = nullptr;
    } pipe;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    int rc = CONSTV_OK;
    explicit DeviceGuard(int device) {
        if (device < 0) return;  // caller's current context (OpenMM interop): never switch
        if (cudaGetDevice(&prev) != cudaSuccess) { rc = fail(CONSTV_ERR_CUDA, "cudaGetDevice failed: no CUDA device"); return; }
        if (prev != device) {
            cudaError_t e = cudaSetDevice(device);
            if (e != cudaSuccess) { rc = fail(CONSTV_ERR_CUDA, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e)); return; }
            switched = true;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

#define GUARD(slab)                              \
    if (!(slab)) return fail(CONSTV_ERR_INVALID, "null slab"); \
    DeviceGuard guard__((slab)->cfg.device);     \
    if (guard__.rc) return guard__.rc

SlabDev makeDev(const constv_slab* s) {
    SlabDev d{};
    d.posq = s->posq;
    d.corr = s->corr;
    d.velm = s->velm;
    d.force = static_cast<long long*>(s->force);
    d.posDelta = s->posDelta;
    d.atomIndex = s->atomIndex;
    d.invAtomOrder = s->invAtomOrder;
    d.imageCharge = ((s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && s->imageChargeValid) ? s->imageCharge : nullptr;
    const bool bySlot = d.imageCharge && s->atomIndex && s->imageBySlotValid && s->imageTables;
    d.imageSlotBySlot = bySlot ? s->imageSlotBySlot : nullptr;
    d.imageChargeBySlot = bySlot ? s->imageChargeBySlot : nullptr;
    d.counters = s->counters;
    d.numAtoms = s->cfg.num_atoms;
    d.numReal = s->numReal;
    d.padded = s->cfg.padded_num_atoms;
    d.numCells = s->cfg.num_cells;
    d.zshift = s->cfg.mirror_mode == CONSTV_MIRROR_NEGATE ? 0.0 : s->cfg.zmax;
    d.invDt = s->invDt;
    d.vscale = s->vscale;
    d.fscaleFixed = s->fscaleFixed;
    d.noisescale = s->noisescale;
    d.dt = s->dt;
    d.vscaleF = (float) s->vscale;
    d.fscaleFixedF = (float) s->fscaleFixed;
    d.noisescaleF = (float) s->noisescale;
    d.dtF = (float) s->dt;
    d.seed = s->cfg.seed;
    d.globalAtomOffset = s->cfg.global_atom_offset;
    d.zeroVelm = (s->cfg.flags & CONSTV_FLAG_ZERO_IMAGE_VELM) ? 1u : 0u;
    d.zeroForce = (s->cfg.flags & CONSTV_FLAG_ZERO_IMAGE_FORCE) ? 1u : 0u;
    d.tileBase = 0;
    d.rangeLo = 0;
    d.rangeHi = s->numReal;
    d.zeroWhen = s->zeroWhen;
    return d;
}

// the descriptor of chain q: its range of real atoms and its own step counter / ticket
SlabDev makeChainDev(const constv_slab* s, int q) {
    SlabDev d = makeDev(s);
    d.counters = s->counters + q;
    d.rangeLo = s->chainLo[q];
    d.rangeHi = s->chainLo[q + 1];
    return d;
}

int checkReady(const constv_slab* s, bool needDelta) {
    if (!s->posq || !s->velm || !s->force) return fail(CONSTV_ERR_STATE, "device arrays are not bound (constv_bind_arrays / constv_alloc_arrays)");
    if (s->cfg.precision == CONSTV_PRECISION_MIXED && !s->corr) return fail(CONSTV_ERR_STATE, "mixed precision needs posq_correction");
    if (needDelta && !s->posDelta) return fail(CONSTV_ERR_STATE, "the split path needs pos_delta");
    if (!s->paramsSet) return fail(CONSTV_ERR_STATE, "constv_set_parameters has not been called");
    return CONSTV_OK;
}

// the plain Langevin entry points refuse a slab whose coefficients were set by constv_drude_set_parameters
int checkLangevin(const constv_slab* s) {
    if (s->paramsKind != 1) return fail(CONSTV_ERR_STATE, "the slab holds ConstVDrudeLangevinIntegrator coefficients: call constv_set_parameters");
    return CONSTV_OK;
}

template <typename K, typename... Args>
int launch(K kernel, int grid, int block, bool pdl, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned) grid);
    cfg.blockDim = dim3((unsigned) block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, args...));
    g_lastKernel = "";  // launch sites that have a choice of kernels name theirs afterwards
    g_launchCount.fetch_add(1, std::memory_order_relaxed);
    return CONSTV_OK;
}

int gridFor(const constv_slab* s, long long work, int perBlock, int ctasPerSM) {
    long long blocks = (work + perBlock - 1) / perBlock;
    long long cap = ctasPerSM > 0 ? (long long) s->numSMs * ctasPerSM : blocks;  // 0 = one-shot grid
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int) blocks;
}

// ---- dispatch over (precision, ILP) --------------------------------------------------------------

template <int PREC, bool STREAM, bool EARLY>
int launchFusedS(constv_slab* s, const SlabDev& d, const float4* random, unsigned randomIndex, cudaStream_t st) {
    const bool pdl = s->cfg.flags & CONSTV_FLAG_PDL;
    if (s->fusedBlock == 128)
        return launch(constv_fused_step_kernel<PREC, 128, 1, STREAM, EARLY>, gridFor(s, d.rangeHi - d.rangeLo, 128, s->ctasPerSM), 128, pdl, st, d, random, randomIndex);
    switch (s->ilp) {
        case 1: return launch(constv_fused_step_kernel<PREC, kBlock, 1, STREAM, EARLY>, gridFor(s, d.rangeHi - d.rangeLo, kBlock * 1, s->ctasPerSM), kBlock, pdl, st, d, random, randomIndex);
        case 4: return launch(constv_fused_step_kernel<PREC, kBlock, 4, STREAM, EARLY>, gridFor(s, d.rangeHi - d.rangeLo, kBlock * 4, s->ctasPerSM), kBlock, pdl, st, d, random, randomIndex);
        default: return launch(constv_fused_step_kernel<PREC, kBlock, 2, STREAM, EARLY>, gridFor(s, d.rangeHi - d.rangeLo, kBlock * 2, s->ctasPerSM), kBlock, pdl, st, d, random, randomIndex);
    }
}

template <int PREC>
int launchFused(constv_slab* s, const SlabDev& d, const float4* random, unsigned randomIndex, cudaStream_t st) {
    if (s->fusedEarly)
        return (s->cfg.flags & CONSTV_FLAG_KEEP_IN_L2) ? launchFusedS<PREC, false, true>(s, d, random, randomIndex, st)
                                                       : launchFusedS<PREC, true, true>(s, d, random, randomIndex, st);
    return (s->cfg.flags & CONSTV_FLAG_KEEP_IN_L2) ? launchFusedS<PREC, false, false>(s, d, random, randomIndex, st)
                                                   : launchFusedS<PREC, true, false>(s, d, random, randomIndex, st);
}

template <int PREC, int BLOCK, int STAGES>
int launchFusedTmaShape(constv_slab* s, const SlabDev& d, cudaStream_t st) {
    auto kernel = constv_fused_step_tma_kernel<PREC, BLOCK, STAGES, true>;
    const int smem = STAGES * StageLayout<PREC, BLOCK>::kBytes;
    static thread_local int configuredDevice = -1;
    static thread_local int maxCtas = 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (configuredDevice != dev) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxCtas, kernel, BLOCK, smem));
        configuredDevice = dev;
    }
    if (maxCtas < 1) return fail(CONSTV_ERR_CUDA, "TMA kernel does not fit on this device (%d bytes of shared memory)", smem);
    int per = s->tmaCtasPerSM > 0 && s->tmaCtasPerSM < maxCtas ? s->tmaCtasPerSM : maxCtas;
    long long grid = (long long) s->numSMs * per;
    const long long chunks = ((long long) s->numReal + BLOCK - 1) / BLOCK;
    if (grid > chunks) grid = chunks;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned) grid);
    cfg.blockDim = dim3(BLOCK);
    cfg.dynamicSmemBytes = (size_t) smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (s->cfg.flags & CONSTV_FLAG_PDL) ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, d));
    g_lastKernel = "constv_fused_step_tma_kernel";
    g_launchCount.fetch_add(1, std::memory_order_relaxed);
    return CONSTV_OK;
}

template <int PREC>
int launchFusedTma(constv_slab* s, const SlabDev& d, cudaStream_t st) {
    const int key = s->tmaBlock * 10 + s->tmaStages;
    switch (key) {
        case 2562: return launchFusedTmaShape<PREC, 256, 2>(s, d, st);
        case 2563: return launchFusedTmaShape<PREC, 256, 3>(s, d, st);
        case 2564: return launchFusedTmaShape<PREC, 256, 4>(s, d, st);
        case 5122: return launchFusedTmaShape<PREC, 512, 2>(s, d, st);
        case 5123: return launchFusedTmaShape<PREC, 512, 3>(s, d, st);
        case 1286: return launchFusedTmaShape<PREC, 128, 6>(s, d, st);
        default: return fail(CONSTV_ERR_INVALID, "unsupported TMA shape block=%d stages=%d", s->tmaBlock, s->tmaStages);
    }
}

bool tmaEligible(const constv_slab* s, const void* random4) {
    return !s->atomIndex && !random4 && (s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && s->imageChargeValid &&
           (s->cfg.padded_num_atoms % 2 == 0);
}

template <int PREC>
int launchFusedIndexed(constv_slab* s, const SlabDev& d, const float4* random, unsigned randomIndex, cudaStream_t st) {
    // thread i owns slot i and its numCells-1 image-cell slots: a one-shot grid over the real-atom count
    if (PREC == PREC_SINGLE && s->indexedMinBlocks == 8)
        return launch(constv_fused_step_indexed_kernel<PREC_SINGLE, kBlock, 8>, gridFor(s, s->numReal, kBlock, 0), kBlock,
                      (bool) (s->cfg.flags & CONSTV_FLAG_PDL), st, d, random, randomIndex);
    return launch(constv_fused_step_indexed_kernel<PREC, kBlock>, gridFor(s, s->numReal, kBlock, 0), kBlock,
                  (bool) (s->cfg.flags & CONSTV_FLAG_PDL), st, d, random, randomIndex);
}

template <int PREC>
int launchPart1(constv_slab* s, const SlabDev& d, const float4* random, unsigned randomIndex, cudaStream_t st) {
    return launch(constv_part1_kernel<PREC, kBlock>, gridFor(s, s->cfg.num_atoms, kBlock, s->ctasPerSM), kBlock,
                  (bool) (s->cfg.flags & CONSTV_FLAG_PDL), st, d, random, randomIndex);
}

template <int PREC>
int launchPart2(constv_slab* s, const SlabDev& d, cudaStream_t st) {
    if (!s->atomIndex)
        return launch(constv_part2_mirror_kernel<PREC, kBlock>, gridFor(s, s->numReal, kBlock, s->ctasPerSM), kBlock,
                      (bool) (s->cfg.flags & CONSTV_FLAG_PDL), st, d);
    int rc = launch(constv_part2_indexed_kernel<PREC, kBlock>, gridFor(s, s->cfg.num_atoms, kBlock, s->ctasPerSM), kBlock, false, st, d);
    if (rc) return rc;
    return launch(constv_mirror_indexed_kernel<PREC, kBlock, true>, gridFor(s, s->numReal, kBlock, s->ctasPerSM), kBlock, false, st, d);
}

template <int PREC>
int launchKE(constv_slab* s, const SlabDev& d, double scale, double* out, cudaStream_t st) {
    const int grid = gridFor(s, s->cfg.num_atoms, kBlock, 4);
    return launch(constv_kinetic_energy_kernel<PREC, kBlock>, grid, kBlock, false, st, d, scale, s->kePartials, out);
}

template <int PREC, int CAP>
int launchBatchedCap(constv_slab* lead, const SlabDev* table, int n, int totalTiles, cudaStream_t st) {
    static thread_local BatchParams<CAP> b;  // up to 24 KB: kept off the stack
    b.numSlabs = n;
    b.totalTiles = totalTiles;
    for (int k = 0; k < n; k++) b.slab[k] = table[k];
    const bool pdl = lead->cfg.flags & CONSTV_FLAG_PDL;
    const long long cap = lead->ctasPerSM > 0 ? (long long) lead->numSMs * lead->ctasPerSM : totalTiles;
    const int grid = (int) (totalTiles < cap ? totalTiles : cap);
    return launch(constv_fused_step_batched_kernel<PREC, kBlock, CAP>, grid, kBlock, pdl, st, b);
}

template <int PREC>
int launchBatched(constv_slab* lead, const SlabDev* table, int n, int totalTiles, cudaStream_t st) {
    if (n <= 8) return launchBatchedCap<PREC, 8>(lead, table, n, totalTiles, st);
    if (n <= 32) return launchBatchedCap<PREC, 32>(lead, table, n, totalTiles, st);
    return launchBatchedCap<PREC, kBatchMax>(lead, table, n, totalTiles, st);
}

#define DISPATCH_PREC(prec, call)                                       \
    switch (prec) {                                                     \
        case CONSTV_PRECISION_SINGLE: { constexpr int PR = PREC_SINGLE; rc = call; break; } \
        case CONSTV_PRECISION_MIXED:  { constexpr int PR = PREC_MIXED;  rc = call; break; } \
        default:                      { constexpr int PR = PREC_DOUBLE; rc = call; break; } \
    }

void advanceTime(constv_slab* s) { s->time += s->stepSize; }

// A chain entry point that finds the image-charge snapshot stale takes it for the whole slab, and must not let
// any chain -- the others run on other streams and rewrite posq[image] -- start before the snapshot is complete:
// the device is drained before and after (a rare path: first step after binding or uploading posq; it cannot
// happen inside a stream capture, where the error says to step once first).
int refreshImageChargesForChains(constv_slab* s, void* stream);

}  // namespace

// =================================================================================================

extern "C" {

const char* constv_last_error(void) { return g_lastError.c_str(); }
int constv_version(void) { return CONSTV_VERSION; }
uint64_t constv_launch_count(void) { return g_launchCount.load(std::memory_order_relaxed); }
const char* constv_last_kernel(void) { return g_lastKernel; }

int constv_validate_system(int32_t num_atoms, int32_t num_cells, const double* masses, double box_z,
                           double cell_size, double* zmax_out) {
    if (num_atoms < 0 || num_cells <= 0 || (!masses && num_atoms > 0))
        return fail(CONSTV_ERR_INVALID, "constv_validate_system: bad arguments");
    if (num_cells % 2 != 0) return fail(CONSTV_ERR_SYSTEM, "Not even number of cells");
    if (num_atoms % num_cells != 0) return fail(CONSTV_ERR_SYSTEM, "Number of particles is not multiple of number of cells");
    for (int i = num_atoms / num_cells; i < num_atoms; i++)
        if (masses[i] != 0.0) return fail(CONSTV_ERR_SYSTEM, "Image Particle has nonzero mass");
    double zmax = cell_size;
    if (zmax < 0)
        zmax = box_z / num_cells;
    else if (zmax * num_cells != box_z)
        return fail(CONSTV_ERR_SYSTEM, "Unit cell dimension does not match with the provided zmax value");
    if (zmax_out) *zmax_out = zmax;
    return CONSTV_OK;
}

int constv_create(const constv_config* config, constv_slab** out) {
    if (!config || !out) return fail(CONSTV_ERR_INVALID, "constv_create: null argument");
    if (config->struct_size != sizeof(constv_config)) return fail(CONSTV_ERR_INVALID, "constv_create: struct_size %u != %zu", config->struct_size, sizeof(constv_config));
    if (config->num_cells < 2 || config->num_cells % 2 != 0) return fail(CONSTV_ERR_SYSTEM, "Not even number of cells");
    if (config->num_atoms <= 0 || config->num_atoms % config->num_cells != 0) return fail(CONSTV_ERR_SYSTEM, "Number of particles is not multiple of number of cells");
    if (config->padded_num_atoms < config->num_atoms) return fail(CONSTV_ERR_INVALID, "padded_num_atoms < num_atoms");
    if (config->precision < 0 || config->precision > 2) return fail(CONSTV_ERR_INVALID, "unknown precision %d", config->precision);
    if (config->mirror_mode != CONSTV_MIRROR_REFERENCE && config->mirror_mode != CONSTV_MIRROR_NEGATE) return fail(CONSTV_ERR_INVALID, "unknown mirror mode %d", config->mirror_mode);
    if (config->mirror_mode == CONSTV_MIRROR_NEGATE && config->num_cells != 2) return fail(CONSTV_ERR_INVALID, "CONSTV_MIRROR_NEGATE requires num_cells == 2");
    if (!(config->zmax > 0)) return fail(CONSTV_ERR_INVALID, "zmax must be resolved (> 0); see constv_validate_system");

    constv_slab* s = new (std::nothrow) constv_slab();
    if (!s) return fail(CONSTV_ERR_INVALID, "out of host memory");
    s->cfg = *config;
    if (s->cfg.boltz == 0) s->cfg.boltz = kBoltzDefault;
    s->numReal = config->num_atoms / config->num_cells;
    if (const char* e = std::getenv("CONSTV_FUSED_BLOCK")) s->fusedBlock = std::atoi(e) == 128 ? 128 : 256;
    if (const char* e = std::getenv("CONSTV_FUSED_EARLY")) s->fusedEarly = std::atoi(e) != 0;
    if (const char* e = std::getenv("CONSTV_FUSED_ZERO")) s->zeroWhen = std::atoi(e);
    if (const char* e = std::getenv("CONSTV_SETTLE_WS")) s->settleWs = std::atoi(e) != 0;
    if (const char* e = std::getenv("CONSTV_SETTLE_EARLY")) s->settleEarly = std::atoi(e) != 0 ? 1 : 0;
    if (const char* e = std::getenv("CONSTV_INDEXED_MINB")) s->indexedMinBlocks = std::atoi(e);
    if (const char* e = std::getenv("CONSTV_IMAGE_TABLES")) s->imageTables = std::atoi(e) != 0;
    if (const char* e = std::getenv("CONSTV_DRUDE_SPT")) s->drudeSlotsPerThread = std::min(3, std::max(1, std::atoi(e)));
    if (const char* e = std::getenv("CONSTV_WS_GROUPS")) { const int g = std::atoi(e); if (g >= 3 && g <= 5) s->wsGroups = g; }
    DeviceGuard guard(config->device);
    if (guard.rc) { delete s; return guard.rc; }
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->numSMs, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { delete s; return fail(CONSTV_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e)); }

    const int P = config->padded_num_atoms;
    auto cleanup = [&](cudaError_t err, const char* what) {
        constv_destroy(s);
        return fail(CONSTV_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
    };
    if ((e = cudaMalloc(&s->invAtomOrder, sizeof(int) * (size_t) P)) != cudaSuccess) return cleanup(e, "cudaMalloc(invAtomOrder)");
    if ((e = cudaMalloc(&s->imageCharge, realSize(config->precision) * ((size_t) (config->num_atoms - s->numReal) + 4))) != cudaSuccess) return cleanup(e, "cudaMalloc(imageCharge)");  // + 4: the TMA kernels read whole 4-slot windows
    if ((e = cudaMalloc(&s->imageSlotBySlot, sizeof(int) * ((size_t) (config->num_atoms - s->numReal) + 4))) != cudaSuccess) return cleanup(e, "cudaMalloc(imageSlotBySlot)");
    if ((e = cudaMalloc(&s->imageChargeBySlot, realSize(config->precision) * ((size_t) (config->num_atoms - s->numReal) + 4))) != cudaSuccess) return cleanup(e, "cudaMalloc(imageChargeBySlot)");
    if ((e = cudaMalloc(&s->counters, sizeof(DevCounters) * kMaxChains)) != cudaSuccess) return cleanup(e, "cudaMalloc(counters)");
    if ((e = cudaMemset(s->counters, 0, sizeof(DevCounters) * kMaxChains)) != cudaSuccess) return cleanup(e, "cudaMemset(counters)");
    s->chainLo[1] = s->numReal;
    const int maxBlocks = s->numSMs * 16;
    if ((e = cudaMalloc(&s->kePartials, sizeof(double) * (size_t) maxBlocks)) != cudaSuccess) return cleanup(e, "cudaMalloc(kePartials)");
    if ((e = cudaMalloc(&s->keOut, sizeof(double))) != cudaSuccess) return cleanup(e, "cudaMalloc(keOut)");
    if ((e = cudaMallocHost(&s->keHost, sizeof(double))) != cudaSuccess) return cleanup(e, "cudaMallocHost(keHost)");
    // identity inverse order (the reference leaves this array uninitialised until the first reorder)
    {
        std::vector<int> ident((size_t) P);
        for (int i = 0; i < P; i++) ident[(size_t) i] = i;
        if ((e = cudaMemcpy(s->invAtomOrder, ident.data(), sizeof(int) * (size_t) P, cudaMemcpyHostToDevice)) != cudaSuccess) return cleanup(e, "cudaMemcpy(invAtomOrder)");
    }
    *out = s;
    return CONSTV_OK;
}

int constv_destroy(constv_slab* s) {
    if (!s) return CONSTV_OK;
    DeviceGuard guard(s->cfg.device);
    if (s->owned) {
        cudaFree(s->posq); cudaFree(s->corr); cudaFree(s->velm); cudaFree(s->force); cudaFree(s->posDelta);
    }
    for (int q = 0; q < kMaxChains; q++) {
        if (s->chainStream[q]) cudaStreamDestroy(s->chainStream[q]);
        if (s->chainJoin[q]) cudaEventDestroy(s->chainJoin[q]);
    }
    if (s->chainFork) cudaEventDestroy(s->chainFork);
    cudaFree(s->invAtomOrder);
    cudaFree(s->drude.role);
    cudaFree(s->drude.pairs);
    cudaFree(s->settle.items);
    cudaFree(s->settle.atoms);
    cudaFree(s->settle.params);
    cudaFree(s->imageCharge);
    cudaFree(s->imageSlotBySlot);
    cudaFree(s->imageChargeBySlot);
    cudaFree(s->counters);
    cudaFree(s->kePartials);
    cudaFree(s->keOut);
    if (s->pipe.ready) {
        cudaStreamSynchronize(s->pipe.copyStream);
        for (int k = 0; k < 2; k++) {
            cudaFree(s->pipe.stage[k]);
            cudaEventDestroy(s->pipe.h2dDone[k]);
            cudaEventDestroy(s->pipe.stageFree[k]);
            cudaEventDestroy(s->pipe.done[k]);
        }
        cudaFree(s->pipe.keDev);
        cudaFreeHost(s->pipe.keHost);
        cudaStreamDestroy(s->pipe.copyStream);
        if (s->pipe.readStream) {
            cudaStreamSynchronize(s->pipe.readStream);
            for (int k = 0; k < 2; k++) {
                cudaEventDestroy(s->pipe.stepDone[k]);
                cudaEventDestroy(s->pipe.readDone[k]);
            }
            cudaStreamDestroy(s->pipe.readStream);
        }
    }
    if (s->keHost) cudaFreeHost(s->keHost);
    delete s;
    return CONSTV_OK;
}

int constv_bind_arrays(constv_slab* s, void* posq, void* posq_correction, void* velm, void* force,
                       void* pos_delta, const int32_t* atom_index) {
    if (!s) return fail(CONSTV_ERR_INVALID, "null slab");
    if (!posq || !velm || !force) return fail(CONSTV_ERR_INVALID, "posq, velm and force are required");
    if (s->cfg.precision == CONSTV_PRECISION_MIXED && !posq_correction) return fail(CONSTV_ERR_INVALID, "mixed precision needs posq_correction");
    if (((uintptr_t) posq | (uintptr_t) posq_correction | (uintptr_t) velm | (uintptr_t) force | (uintptr_t) pos_delta) & 15u)
        return fail(CONSTV_ERR_INVALID, "device arrays must be 16-byte aligned");
    if (s->owned) {
        DeviceGuard guard(s->cfg.device);
        cudaFree(s->posq); cudaFree(s->corr); cudaFree(s->velm); cudaFree(s->force); cudaFree(s->posDelta);
        s->owned = false;
    }
    s->posq = posq;
    s->corr = s->cfg.precision == CONSTV_PRECISION_MIXED ? posq_correction : nullptr;
    s->velm = velm;
    s->force = force;
    s->posDelta = pos_delta;
    s->atomIndex = atom_index;
    s->imageChargeValid = false;
    s->imageBySlotValid = false;
    return CONSTV_OK;
}

int constv_alloc_arrays(constv_slab* s) {
    GUARD(s);
    if (s->owned) return CONSTV_OK;
    const size_t P = (size_t) s->cfg.padded_num_atoms;
    const size_t rb = 4 * realSize(s->cfg.precision) * P, mb = 4 * mixedSize(s->cfg.precision) * P;
    void *posq = nullptr, *corr = nullptr, *velm = nullptr, *force = nullptr, *delta = nullptr;
    CUDA_TRY(cudaMalloc(&posq, rb));
    CUDA_TRY(cudaMemset(posq, 0, rb));
    if (s->cfg.precision == CONSTV_PRECISION_MIXED) {
        CUDA_TRY(cudaMalloc(&corr, rb));
        CUDA_TRY(cudaMemset(corr, 0, rb));
    }
    CUDA_TRY(cudaMalloc(&velm, mb));
    CUDA_TRY(cudaMemset(velm, 0, mb));
    CUDA_TRY(cudaMalloc(&force, 24 * P));
    CUDA_TRY(cudaMemset(force, 0, 24 * P));
    CUDA_TRY(cudaMalloc(&delta, mb));
    CUDA_TRY(cudaMemset(delta, 0, mb));
    s->posq = posq; s->corr = corr; s->velm = velm; s->force = force; s->posDelta = delta;
    s->atomIndex = nullptr;
    s->owned = true;
    s->imageChargeValid = false;
    return CONSTV_OK;
}

int constv_get_arrays(constv_slab* s, void** posq, void** posq_correction, void** velm, void** force, void** pos_delta) {
    if (!s) return fail(CONSTV_ERR_INVALID, "null slab");
    if (posq) *posq = s->posq;
    if (posq_correction) *posq_correction = s->corr;
    if (velm) *velm = s->velm;
    if (force) *force = s->force;
    if (pos_delta) *pos_delta = s->posDelta;
    return CONSTV_OK;
}

int constv_set_parameters(constv_slab* s, double temperature, double friction, double step_size) {
    if (!s) return fail(CONSTV_ERR_INVALID, "null slab");
    if (!(step_size > 0)) return fail(CONSTV_ERR_INVALID, "step size must be positive");
    if (s->paramsSet && s->paramsKind == 1 && temperature == s->temperature && friction == s->friction && step_size == s->stepSize)
        return CONSTV_OK;  // Kernels.cpp:113: recompute only on change
    const double kT = s->cfg.boltz * temperature;
    const double vscale = std::exp(-step_size * friction);
    const double fscale = (friction == 0 ? step_size : (1 - vscale) / friction);
    const double noisescale = std::sqrt(kT * (1 - vscale * vscale));
    if (s->cfg.precision == CONSTV_PRECISION_SINGLE) {
        const float dtf = (float) step_size;
        s->vscale = (double) (float) vscale;
        s->fscaleFixed = (double) ((float) fscale / (float) 4294967296.0);  // Langevin.cu:10
        s->noisescale = (double) (float) noisescale;
        s->dt = (double) dtf;
        s->invDt = 1.0 / (double) dtf;  // Langevin.cu:36
    } else {
        s->vscale = vscale;
        s->fscaleFixed = fscale / 4294967296.0;
        s->noisescale = noisescale;
        s->dt = step_size;
        s->invDt = 1.0 / step_size;
    }
    s->temperature = temperature;
    s->friction = friction;
    s->stepSize = step_size;
    s->paramsSet = true;
    s->paramsKind = 1;
    return CONSTV_OK;
}

int constv_get_parameters(const constv_slab* s, double out[4]) {
    if (!s || !out) return fail(CONSTV_ERR_INVALID, "null argument");
    if (!s->paramsSet) return fail(CONSTV_ERR_STATE, "constv_set_parameters has not been called");
    out[0] = s->vscale;
    out[1] = s->fscaleFixed * 4294967296.0;
    out[2] = s->noisescale;
    out[3] = s->dt;
    return CONSTV_OK;
}

int constv_update_inverse_order(constv_slab* s, void* stream) {
    GUARD(s);
    if (!s->atomIndex) return fail(CONSTV_ERR_STATE, "no atom_index array is bound (identity order)");
    const int n = s->cfg.num_atoms;
    int rc = launch(constv_inverse_order_kernel, gridFor(s, n, 256, 8), 256, false, (cudaStream_t) stream, n,
                    (const int*) s->atomIndex, s->invAtomOrder);
    s->imageBySlotValid = false;  // the per-slot image tables follow the order: rebuilt before the next staged step
    return rc;
}

int constv_refresh_image_charges(constv_slab* s, void* stream) {
    GUARD(s);
    if (!s->posq) return fail(CONSTV_ERR_STATE, "device arrays are not bound");
    const int n = s->cfg.num_atoms - s->numReal;
    int rc;
    if (s->cfg.precision == CONSTV_PRECISION_DOUBLE)
        rc = launch(constv_image_charge_kernel<double>, gridFor(s, n, 256, 8), 256, false, (cudaStream_t) stream,
                    s->numReal, s->cfg.num_cells, (const double*) s->posq, (const int*) s->invAtomOrder, (double*) s->imageCharge);
    else
        rc = launch(constv_image_charge_kernel<float>, gridFor(s, n, 256, 8), 256, false, (cudaStream_t) stream,
                    s->numReal, s->cfg.num_cells, (const float*) s->posq, (const int*) s->invAtomOrder, (float*) s->imageCharge);
    if (rc == CONSTV_OK) s->imageChargeValid = true;
    s->imageBySlotValid = false;
    return rc;
}

// Reordered slots: the per-slot image tables of the staged kernels, built when they are stale (after a re-sort or a new
// charge snapshot) and the snapshot exists.
static int ensureImageBySlot(constv_slab* s, cudaStream_t stream) {
    if (!s->atomIndex || s->imageBySlotValid || !(s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) || !s->imageChargeValid) return CONSTV_OK;
    const int n = s->cfg.num_atoms - s->numReal;
    int rc;
    if (s->cfg.precision == CONSTV_PRECISION_DOUBLE)
        rc = launch(constv_image_by_slot_kernel<double>, gridFor(s, n, 256, 8), 256, false, stream, s->numReal, s->cfg.num_cells,
                    (const int*) s->atomIndex, (const int*) s->invAtomOrder, (const double*) s->imageCharge, s->imageSlotBySlot, (double*) s->imageChargeBySlot);
    else
        rc = launch(constv_image_by_slot_kernel<float>, gridFor(s, n, 256, 8), 256, false, stream, s->numReal, s->cfg.num_cells,
                    (const int*) s->atomIndex, (const int*) s->invAtomOrder, (const float*) s->imageCharge, s->imageSlotBySlot, (float*) s->imageChargeBySlot);
    if (rc == CONSTV_OK) s->imageBySlotValid = true;
    return rc;
}

int constv_step_fused(constv_slab* s, const void* random4, uint32_t random_index, void* stream) {
    GUARD(s);
    int rc = checkReady(s, false);
    if (rc) return rc;
    if ((rc = checkLangevin(s))) return rc;
    if ((s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && !s->imageChargeValid) {
        rc = constv_refresh_image_charges(s, stream);
        if (rc) return rc;
    }
    if ((rc = ensureImageBySlot(s, (cudaStream_t) stream))) return rc;
    const SlabDev d = makeDev(s);
    const float4* rnd = static_cast<const float4*>(random4);
    cudaStream_t st = (cudaStream_t) stream;
    const bool tma = s->variant == 2 && tmaEligible(s, random4);
    if (s->variant == 2 && !tma)
        return fail(CONSTV_ERR_STATE, "the TMA kernel needs identity order, Philox noise, CONSTV_FLAG_CACHE_IMAGE_CHARGE and even padded_num_atoms");
    if (s->chains > 1) {  // every chain in turn on this stream: same result, one step counter per chain
        if (s->atomIndex || tma) return fail(CONSTV_ERR_STATE, "chains need identity slot order and the plain-load kernel");
        for (int q = 0; q < s->chains; q++) {
            const SlabDev dq = makeChainDev(s, q);
            DISPATCH_PREC(s->cfg.precision, launchFused<PR>(s, dq, rnd, random_index, st));
            if (rc) return rc;
        }
    } else if (s->atomIndex) {
        DISPATCH_PREC(s->cfg.precision, launchFusedIndexed<PR>(s, d, rnd, random_index, st));
    } else if (tma) {
        DISPATCH_PREC(s->cfg.precision, launchFusedTma<PR>(s, d, st));
    } else {
        DISPATCH_PREC(s->cfg.precision, launchFused<PR>(s, d, rnd, random_index, st));
    }
    if (rc == CONSTV_OK) advanceTime(s);
    return rc;
}

int constv_set_chains(constv_slab* s, int32_t num_chains) {
    GUARD(s);
    if (num_chains < 1 || num_chains > kMaxChains) return fail(CONSTV_ERR_INVALID, "num_chains must be 1..%d", kMaxChains);
    uint64_t step = 0;
    int rc = constv_get_step_count(s, &step);  // synchronises; refuses if the current chains are out of step
    if (rc) return rc;
    const long long tiles = ((long long) s->numReal + kBlock - 1) / kBlock;
    if (num_chains > tiles) num_chains = (int) (tiles < 1 ? 1 : tiles);
    s->chains = num_chains;
    for (int q = 0; q <= num_chains; q++) {
        const long long lo = tiles * q / num_chains * kBlock;  // chain boundaries on tile boundaries
        s->chainLo[q] = (int) (lo < s->numReal ? lo : s->numReal);
    }
    return constv_set_step_count(s, step);
}

int constv_get_chains(const constv_slab* s, int32_t* num_chains) {
    if (!s || !num_chains) return fail(CONSTV_ERR_INVALID, "null argument");
    *num_chains = s->chains;
    return CONSTV_OK;
}

int constv_step_fused_chain(constv_slab* s, int32_t chain, const void* random4, uint32_t random_index, void* stream) {
    GUARD(s);
    int rc = checkReady(s, false);
    if (rc) return rc;
    if ((rc = checkLangevin(s))) return rc;
    if (chain < 0 || chain >= s->chains) return fail(CONSTV_ERR_INVALID, "chain %d out of range (the slab has %d)", chain, s->chains);
    if (s->atomIndex || s->variant == 2) return fail(CONSTV_ERR_STATE, "chains need identity slot order and the plain-load kernel");
    if ((rc = refreshImageChargesForChains(s, stream))) return rc;
    const SlabDev d = makeChainDev(s, chain);
    DISPATCH_PREC(s->cfg.precision, launchFused<PR>(s, d, static_cast<const float4*>(random4), random_index, (cudaStream_t) stream));
    if (rc == CONSTV_OK && chain == 0) advanceTime(s);
    return rc;
}

int constv_step_fused_multi(constv_slab* s, int32_t num_steps, void* stream) {
    GUARD(s);
    if (num_steps < 0) return fail(CONSTV_ERR_INVALID, "num_steps must be >= 0");
    int rc = checkReady(s, false);
    if (rc) return rc;
    if ((rc = checkLangevin(s))) return rc;
    if (s->chains == 1 || num_steps < 2) {
        for (int k = 0; k < num_steps; k++)
            if ((rc = constv_step_fused(s, nullptr, 0, stream))) return rc;
        return CONSTV_OK;
    }
    if (s->atomIndex || s->variant == 2) return fail(CONSTV_ERR_STATE, "chains need identity slot order and the plain-load kernel");
    if ((s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && !s->imageChargeValid)
        if ((rc = constv_refresh_image_charges(s, stream))) return rc;
    cudaStream_t main = (cudaStream_t) stream;
    if (!s->chainFork) CUDA_TRY(cudaEventCreateWithFlags(&s->chainFork, cudaEventDisableTiming));
    for (int q = 1; q < s->chains; q++) {
        if (!s->chainStream[q]) CUDA_TRY(cudaStreamCreateWithFlags(&s->chainStream[q], cudaStreamNonBlocking));
        if (!s->chainJoin[q]) CUDA_TRY(cudaEventCreateWithFlags(&s->chainJoin[q], cudaEventDisableTiming));
    }
    // fork: chain q's steps, in order, on its own stream (chain 0 stays on the caller's); join back
    CUDA_TRY(cudaEventRecord(s->chainFork, main));
    for (int q = 0; q < s->chains; q++) {
        cudaStream_t st = q == 0 ? main : s->chainStream[q];
        if (q > 0) CUDA_TRY(cudaStreamWaitEvent(st, s->chainFork, 0));
        const SlabDev dq = makeChainDev(s, q);
        for (int k = 0; k < num_steps; k++) {
            DISPATCH_PREC(s->cfg.precision, launchFused<PR>(s, dq, nullptr, 0u, st));
            if (rc) return rc;
        }
        if (q > 0) CUDA_TRY(cudaEventRecord(s->chainJoin[q], st));
    }
    for (int q = 1; q < s->chains; q++) CUDA_TRY(cudaStreamWaitEvent(main, s->chainJoin[q], 0));
    for (int k = 0; k < num_steps; k++) advanceTime(s);
    return CONSTV_OK;
}

}  // extern "C"

namespace {
int refreshImageChargesForChains(constv_slab* s, void* stream) {
    if (!(s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) || s->imageChargeValid) return CONSTV_OK;
    if (s->chains > 1) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess)
            return fail(CONSTV_ERR_CUDA, "the image-charge snapshot is stale and cannot be taken here (%s): step the slab once "
                                         "outside stream capture first", cudaGetErrorString(e));
    }
    int rc = constv_refresh_image_charges(s, stream);
    if (rc) return rc;
    if (s->chains > 1) CUDA_TRY(cudaDeviceSynchronize());
    return CONSTV_OK;
}
}  // namespace

extern "C" {

int constv_step_part1(constv_slab* s, const void* random4, uint32_t random_index, void* stream) {
    GUARD(s);
    int rc = checkReady(s, true);
    if (rc) return rc;
    if (s->chains > 1) return fail(CONSTV_ERR_STATE, "chains are for the fused step only (constv_set_chains(slab, 1) first)");
    if ((rc = checkLangevin(s))) return rc;
    const SlabDev d = makeDev(s);
    DISPATCH_PREC(s->cfg.precision, launchPart1<PR>(s, d, static_cast<const float4*>(random4), random_index, (cudaStream_t) stream));
    return rc;
}

int constv_step_part2_mirror(constv_slab* s, void* stream) {
    GUARD(s);
    int rc = checkReady(s, true);
    if (rc) return rc;
    if ((rc = checkLangevin(s))) return rc;
    if ((s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && !s->imageChargeValid) {
        rc = constv_refresh_image_charges(s, stream);
        if (rc) return rc;
    }
    const SlabDev d = makeDev(s);
    DISPATCH_PREC(s->cfg.precision, launchPart2<PR>(s, d, (cudaStream_t) stream));
    if (rc == CONSTV_OK) advanceTime(s);
    return rc;
}

int constv_mirror(constv_slab* s, void* stream) {
    GUARD(s);
    if (!s->posq) return fail(CONSTV_ERR_STATE, "device arrays are not bound (constv_bind_arrays / constv_alloc_arrays)");
    if (s->cfg.precision == CONSTV_PRECISION_MIXED && !s->corr) return fail(CONSTV_ERR_STATE, "mixed precision needs posq_correction");
    int rc;
    if ((s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && !s->imageChargeValid)
        if ((rc = constv_refresh_image_charges(s, stream))) return rc;
    const SlabDev d = makeDev(s);
    DISPATCH_PREC(s->cfg.precision, launch(constv_mirror_indexed_kernel<PR, kBlock, false>, gridFor(s, s->numReal, kBlock, 0), kBlock, false,
                                           (cudaStream_t) stream, d));
    return rc;
}

int constv_step_fused_batched(constv_slab* const* slabs, int32_t n, void* stream) {
    if (!slabs || n <= 0) return fail(CONSTV_ERR_INVALID, "constv_step_fused_batched: empty batch");
    constv_slab* lead = slabs[0];
    GUARD(lead);
    int rc;
    for (int k = 0; k < n; k++) {
        constv_slab* s = slabs[k];
        if (!s) return fail(CONSTV_ERR_INVALID, "null slab in batch");
        if (s->cfg.precision != lead->cfg.precision || s->cfg.device != lead->cfg.device)
            return fail(CONSTV_ERR_INVALID, "batched slabs must share device and precision");
        if (s->atomIndex) return fail(CONSTV_ERR_INVALID, "batched launch needs identity slot order");
        if (s->chains > 1) return fail(CONSTV_ERR_STATE, "chains are for the fused step only (constv_set_chains(slab, 1) first)");
        if ((rc = checkReady(s, false))) return rc;
        if ((rc = checkLangevin(s))) return rc;
        if ((s->cfg.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE) && !s->imageChargeValid)
            if ((rc = constv_refresh_image_charges(s, stream))) return rc;
    }
    // one launch per group of up to kBatchMax slabs; each group's ticket lives on its first slab
    std::vector<SlabDev>& table = lead->batchScratch;
    for (int first = 0; first < n; first += kBatchMax) {
        const int m = n - first < kBatchMax ? n - first : kBatchMax;
        table.resize((size_t) m);
        int totalTiles = 0;
        for (int k = 0; k < m; k++) {
            SlabDev d = makeDev(slabs[first + k]);
            d.tileBase = totalTiles;
            totalTiles += (slabs[first + k]->numReal + kBlock - 1) / kBlock;
            table[(size_t) k] = d;
        }
        DISPATCH_PREC(lead->cfg.precision, launchBatched<PR>(slabs[first], table.data(), m, totalTiles, (cudaStream_t) stream));
        if (rc) return rc;
    }
    for (int k = 0; k < n; k++) advanceTime(slabs[k]);
    return CONSTV_OK;
}

int constv_kinetic_energy_device(constv_slab* s, double time_shift, double* ke_device, void* stream) {
    GUARD(s);
    if (!s->velm || !s->force) return fail(CONSTV_ERR_STATE, "device arrays are not bound");
    if (!ke_device) return fail(CONSTV_ERR_INVALID, "null output");
    const SlabDev d = makeDev(s);
    const double scale = time_shift / 4294967296.0;
    int rc;
    if (s->settle.configured && s->settle.numClusters > 0 && s->paramsKind != 2) {
        // rigid molecules: the shifted velocities are projected onto the constraints first, as OpenMM's
        // computeKineticEnergy does for a constrained System
        SettleDev sd{};
        sd.items = s->settle.items;
        sd.atoms = s->settle.atoms;
        sd.params = s->settle.params;
        sd.numItems = s->settle.numItems;
        sd.numClusters = s->settle.numClusters;
        const int grid = gridFor(s, sd.numItems, kBlock, 4);
        DISPATCH_PREC(s->cfg.precision, launch(constv_kinetic_energy_rigid_kernel<PR, kBlock>, grid, kBlock, false, (cudaStream_t) stream,
                                               d, sd, scale, s->kePartials, ke_device));
        return rc;
    }
    DISPATCH_PREC(s->cfg.precision, launchKE<PR>(s, d, scale, ke_device, (cudaStream_t) stream));
    return rc;
}

int constv_kinetic_energy(constv_slab* s, double time_shift, double* ke_out, void* stream) {
    if (!ke_out) return fail(CONSTV_ERR_INVALID, "null output");
    int rc = constv_kinetic_energy_device(s, time_shift, s ? s->keOut : nullptr, stream);
    if (rc) return rc;
    GUARD(s);
    CUDA_TRY(cudaMemcpyAsync(s->keHost, s->keOut, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t) stream));
    *ke_out = *s->keHost;
    return CONSTV_OK;
}

int constv_generate_noise(constv_slab* s, uint64_t step, void* out4, void* stream) {
    GUARD(s);
    if (!out4) return fail(CONSTV_ERR_INVALID, "null output");
    const int n = s->cfg.num_atoms;
    return launch(constv_noise_kernel, gridFor(s, n, 256, 8), 256, false, (cudaStream_t) stream, n,
                  (const int*) s->atomIndex, (unsigned long long) s->cfg.seed,
                  (unsigned long long) s->cfg.global_atom_offset, (unsigned long long) step, static_cast<float4*>(out4));
}

int constv_generate_philox_words(constv_slab* s, uint64_t step, void* words4, void* stream) {
    GUARD(s);
    if (!words4) return fail(CONSTV_ERR_INVALID, "null output");
    const int n = s->cfg.num_atoms;
    return launch(constv_philox_words_kernel, gridFor(s, n, 256, 8), 256, false, (cudaStream_t) stream, n,
                  (const int*) s->atomIndex, (unsigned long long) s->cfg.seed,
                  (unsigned long long) s->cfg.global_atom_offset, (unsigned long long) step, static_cast<uint4*>(words4));
}

int constv_get_step_count(constv_slab* s, uint64_t* count) {
    GUARD(s);
    if (!count) return fail(CONSTV_ERR_INVALID, "null output");
    DevCounters host[kMaxChains];
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(host, s->counters, sizeof(DevCounters) * (size_t) s->chains, cudaMemcpyDeviceToHost));
    for (int q = 1; q < s->chains; q++)
        if (host[q].step != host[0].step)
            return fail(CONSTV_ERR_STATE, "chains are out of step (chain 0 at %llu, chain %d at %llu): advance every chain equally",
                        host[0].step, q, host[q].step);
    *count = host[0].step;
    return CONSTV_OK;
}

int constv_set_step_count(constv_slab* s, uint64_t count) {
    GUARD(s);
    DevCounters host[kMaxChains];
    CUDA_TRY(cudaDeviceSynchronize());
    for (int q = 0; q < kMaxChains; q++) host[q] = DevCounters{(unsigned long long) count, 0u, 0u};
    CUDA_TRY(cudaMemcpy(s->counters, host, sizeof host, cudaMemcpyHostToDevice));
    return CONSTV_OK;
}

int constv_get_time(const constv_slab* s, double* time) {
    if (!s || !time) return fail(CONSTV_ERR_INVALID, "null argument");
    *time = s->time;
    return CONSTV_OK;
}

int constv_set_time(constv_slab* s, double time) {
    if (!s) return fail(CONSTV_ERR_INVALID, "null slab");
    s->time = time;
    return CONSTV_OK;
}

int constv_upload_state(constv_slab* s, const void* posq, const void* posq_correction, const void* velm,
                        const int64_t* force, void* stream) {
    GUARD(s);
    if (!s->posq) return fail(CONSTV_ERR_STATE, "device arrays are not bound");
    const size_t P = (size_t) s->cfg.padded_num_atoms;
    const size_t rb = 4 * realSize(s->cfg.precision) * P, mb = 4 * mixedSize(s->cfg.precision) * P;
    cudaStream_t st = (cudaStream_t) stream;
    if (posq) { CUDA_TRY(cudaMemcpyAsync(s->posq, posq, rb, cudaMemcpyHostToDevice, st)); s->imageChargeValid = false; }
    if (posq_correction && s->corr) CUDA_TRY(cudaMemcpyAsync(s->corr, posq_correction, rb, cudaMemcpyHostToDevice, st));
    if (velm) CUDA_TRY(cudaMemcpyAsync(s->velm, velm, mb, cudaMemcpyHostToDevice, st));
    if (force) CUDA_TRY(cudaMemcpyAsync(s->force, force, 24 * P, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CONSTV_OK;
}

int constv_download_state(constv_slab* s, void* posq, void* posq_correction, void* velm, int64_t* force, void* stream) {
    GUARD(s);
    if (!s->posq) return fail(CONSTV_ERR_STATE, "device arrays are not bound");
    const size_t P = (size_t) s->cfg.padded_num_atoms;
    const size_t rb = 4 * realSize(s->cfg.precision) * P, mb = 4 * mixedSize(s->cfg.precision) * P;
    cudaStream_t st = (cudaStream_t) stream;
    if (posq) CUDA_TRY(cudaMemcpyAsync(posq, s->posq, rb, cudaMemcpyDeviceToHost, st));
    if (posq_correction && s->corr) CUDA_TRY(cudaMemcpyAsync(posq_correction, s->corr, rb, cudaMemcpyDeviceToHost, st));
    if (velm) CUDA_TRY(cudaMemcpyAsync(velm, s->velm, mb, cudaMemcpyDeviceToHost, st));
    if (force) CUDA_TRY(cudaMemcpyAsync(force, s->force, 24 * P, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CONSTV_OK;
}

int constv_step_host(constv_slab* s, const int64_t* host_force_real, void* host_posq_out, void* host_velm_out,
                     double* host_ke_out, void* stream) {
    GUARD(s);
    int rc = checkReady(s, false);
    if (rc) return rc;
    if (s->atomIndex) return fail(CONSTV_ERR_STATE, "constv_step_host needs identity slot order");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t P = (size_t) s->cfg.padded_num_atoms, R = (size_t) s->numReal, N = (size_t) s->cfg.num_atoms;
    if (host_force_real)
        CUDA_TRY(cudaMemcpy2DAsync(s->force, 8 * P, host_force_real, 8 * R, 8 * R, 3, cudaMemcpyHostToDevice, st));
    rc = constv_step_fused(s, nullptr, 0, stream);
    if (rc) return rc;
    if (host_ke_out) {
        rc = constv_kinetic_energy_device(s, 0.5 * s->stepSize, s->keOut, stream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(s->keHost, s->keOut, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (host_posq_out)
        CUDA_TRY(cudaMemcpyAsync(host_posq_out, s->posq, 4 * realSize(s->cfg.precision) * N, cudaMemcpyDeviceToHost, st));
    if (host_velm_out)
        CUDA_TRY(cudaMemcpyAsync(host_velm_out, s->velm, 4 * mixedSize(s->cfg.precision) * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (host_ke_out) *host_ke_out = *s->keHost;
    return CONSTV_OK;
}

int constv_step_host_submit(constv_slab* s, const void* host_force_real, int32_t force_format, uint64_t* ticket_out,
                            void* stream) {
    return constv_step_host_submit_state(s, host_force_real, force_format, nullptr, nullptr, ticket_out, stream);
}

int constv_step_host_submit_state(constv_slab* s, const void* host_force_real, int32_t force_format, void* host_posq_out,
                                  void* host_velm_out, uint64_t* ticket_out, void* stream) {
    GUARD(s);
    int rc = checkReady(s, false);
    if (rc) return rc;
    if (s->atomIndex) return fail(CONSTV_ERR_STATE, "constv_step_host_submit needs identity slot order");
    if (!host_force_real || !ticket_out) return fail(CONSTV_ERR_INVALID, "null argument");
    if (force_format != CONSTV_FORCE_FIXED64 && force_format != CONSTV_FORCE_FLOAT32) return fail(CONSTV_ERR_INVALID, "unknown force format %d", force_format);
    constv_slab::HostPipe& p = s->pipe;
    if (p.submitted - p.collected >= 2) return fail(CONSTV_ERR_STATE, "two tickets are outstanding: collect one first");
    const size_t R = (size_t) s->numReal;
    if (!p.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p.copyStream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CUDA_TRY(cudaMalloc(&p.stage[k], 24 * R));
            CUDA_TRY(cudaEventCreateWithFlags(&p.h2dDone[k], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.stageFree[k], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.done[k], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaMalloc(&p.keDev, 2 * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&p.keHost, 2 * sizeof(double)));
        p.ready = true;
    }
    const bool wantRead = host_posq_out != nullptr || host_velm_out != nullptr;
    if (wantRead && !p.readStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p.readStream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CUDA_TRY(cudaEventCreateWithFlags(&p.stepDone[k], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.readDone[k], cudaEventDisableTiming));
        }
    }
    const int slot = (int) (p.submitted & 1);
    cudaStream_t st = (cudaStream_t) stream;
    const size_t bytes = (force_format == CONSTV_FORCE_FLOAT32 ? 12 : 24) * R;
    // upload on the copy stream as soon as the staging slot's previous contents have been consumed
    if (p.submitted >= 2) CUDA_TRY(cudaStreamWaitEvent(p.copyStream, p.stageFree[slot], 0));
    CUDA_TRY(cudaMemcpyAsync(p.stage[slot], host_force_real, bytes, cudaMemcpyHostToDevice, p.copyStream));
    CUDA_TRY(cudaEventRecord(p.h2dDone[slot], p.copyStream));
    // compute stream: staging -> force planes, fused step, kinetic energy, read-back
    CUDA_TRY(cudaStreamWaitEvent(st, p.h2dDone[slot], 0));
    const int grid = gridFor(s, (long long) (3 * R), 256, 16);
    if (force_format == CONSTV_FORCE_FLOAT32)
        rc = launch(constv_stage_forces_kernel<true>, grid, 256, false, st, s->numReal, s->cfg.padded_num_atoms, (const void*) p.stage[slot], (long long*) s->force);
    else
        rc = launch(constv_stage_forces_kernel<false>, grid, 256, false, st, s->numReal, s->cfg.padded_num_atoms, (const void*) p.stage[slot], (long long*) s->force);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(p.stageFree[slot], st));
    // the previous ticket's read-back copies posq / velm straight out of the bound arrays: this step must not
    // rewrite them before it has finished (the upload and the force staging above do not wait for it)
    if (p.readOutstanding) {
        CUDA_TRY(cudaStreamWaitEvent(st, p.lastRead, 0));
        p.readOutstanding = false;
    }
    rc = constv_step_fused(s, nullptr, 0, stream);
    if (rc) return rc;
    rc = constv_kinetic_energy_device(s, 0.5 * s->stepSize, p.keDev + slot, stream);
    if (rc) return rc;
    p.hasRead[slot] = wantRead;
    if (wantRead) {
        const size_t N = (size_t) s->cfg.num_atoms;
        CUDA_TRY(cudaEventRecord(p.stepDone[slot], st));
        CUDA_TRY(cudaStreamWaitEvent(p.readStream, p.stepDone[slot], 0));
        if (host_posq_out)
            CUDA_TRY(cudaMemcpyAsync(host_posq_out, s->posq, 4 * realSize(s->cfg.precision) * N, cudaMemcpyDeviceToHost, p.readStream));
        if (host_velm_out)
            CUDA_TRY(cudaMemcpyAsync(host_velm_out, s->velm, 4 * mixedSize(s->cfg.precision) * N, cudaMemcpyDeviceToHost, p.readStream));
        CUDA_TRY(cudaEventRecord(p.readDone[slot], p.readStream));
        p.lastRead = p.readDone[slot];
        p.readOutstanding = true;
    }
    CUDA_TRY(cudaMemcpyAsync(p.keHost + slot, p.keDev + slot, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(p.done[slot], st));
    *ticket_out = p.submitted++;
    return CONSTV_OK;
}

int constv_step_host_collect(constv_slab* s, uint64_t ticket, double* host_ke_out) {
    GUARD(s);
    constv_slab::HostPipe& p = s->pipe;
    if (!p.ready || ticket != p.collected || ticket >= p.submitted)
        return fail(CONSTV_ERR_STATE, "tickets are collected in submission order (next: %llu)", (unsigned long long) p.collected);
    const int slot = (int) (ticket & 1);
    CUDA_TRY(cudaEventSynchronize(p.done[slot]));
    if (p.hasRead[slot]) CUDA_TRY(cudaEventSynchronize(p.readDone[slot]));
    if (host_ke_out) *host_ke_out = p.keHost[slot];
    p.collected++;
    return CONSTV_OK;
}

int constv_set_launch_shape(constv_slab* s, int32_t pairs_per_thread, int32_t ctas_per_sm) {
    if (!s) return fail(CONSTV_ERR_INVALID, "null slab");
    if (pairs_per_thread != 0) {
        if (pairs_per_thread != 1 && pairs_per_thread != 2 && pairs_per_thread != 4)
            return fail(CONSTV_ERR_INVALID, "pairs_per_thread must be 1, 2 or 4");
        s->ilp = pairs_per_thread;
    }
    if (ctas_per_sm != 0) {
        if (ctas_per_sm < -1 || ctas_per_sm > 32) return fail(CONSTV_ERR_INVALID, "ctas_per_sm must be -1 (one-shot grid) or 1..32");
        s->ctasPerSM = ctas_per_sm < 0 ? 0 : ctas_per_sm;
    }
    return CONSTV_OK;
}

int constv_set_kernel_variant(constv_slab* s, int32_t variant, int32_t block_size, int32_t stages, int32_t ctas_per_sm) {
    if (!s) return fail(CONSTV_ERR_INVALID, "null slab");
    if (variant < 0 || variant > 7)
        return fail(CONSTV_ERR_INVALID, "variant must be 0 (auto), 1 (LDG / gather), 2 (TMA ring), 3 (LDG-staged rigid-water kernel), 4 (TMA-pipelined "
                                        "rigid-water kernel), 5 (Drude step staged with one work item per thread), 6 (rigid-water kernel with a bulk "
                                        "stage-in) or 7 (rigid-water kernel staged by cp.async)");
    s->variant = variant;
    if (block_size) s->tmaBlock = block_size;
    if (stages) s->tmaStages = stages;
    if (ctas_per_sm >= 0) s->tmaCtasPerSM = ctas_per_sm;
    return CONSTV_OK;
}

}  // extern "C"

#include "constv_drude_capi.inc"
#include "constv_settle_capi.inc"
