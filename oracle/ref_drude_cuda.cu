/*
 * ref_drude_cuda.cu -- TEST INFRASTRUCTURE.  The reference's Drude kernels
 * (/root/reference/platforms/cuda/src/kernels/vectorOps.cu + constVDrudeLangevin.cu, included from
 * where they lie) compiled UNMODIFIED by nvcc for sm_100a behind plain C launchers, one translation
 * unit per precision model -- the Drude counterpart of ref_cuda.cu.  OpenMM injects NUM_ATOMS,
 * PADDED_NUM_ATOMS, NUM_NORMAL_PARTICLES and NUM_PAIRS as literals when it compiles the module at
 * run time (CudaConstVKernels.cpp:236-240); here they are __constant__ ints set by
 * constv_ref_drude_cuda_configure_<prec>, which is arithmetically the same.
 * Launch shape: cu.executeKernel(kernel, args, workUnits) with OpenMM's default 64-thread blocks
 * [recalled] for part 1 / part 2 / hard wall (CudaConstVKernels.cpp:321,332,339).
 */
#include <cuda_runtime.h>

#include "ref_preamble.h"

__constant__ int CONSTV_REF_NAME(constv_ref_num_atoms);
__constant__ int CONSTV_REF_NAME(constv_ref_padded_num_atoms);
__constant__ int CONSTV_REF_NAME(constv_ref_num_normal);
__constant__ int CONSTV_REF_NAME(constv_ref_num_pairs);
#define NUM_ATOMS CONSTV_REF_NAME(constv_ref_num_atoms)
#define PADDED_NUM_ATOMS CONSTV_REF_NAME(constv_ref_padded_num_atoms)
#define NUM_NORMAL_PARTICLES CONSTV_REF_NAME(constv_ref_num_normal)
#define NUM_PAIRS CONSTV_REF_NAME(constv_ref_num_pairs)

#include CONSTV_REF_VECTOROPS_SOURCE
#include CONSTV_REF_DRUDE_SOURCE

namespace {
int gridFor(int work, int block, int maxBlocks) {
    int blocks = (work + block - 1) / block;
    if (maxBlocks > 0 && blocks > maxBlocks) blocks = maxBlocks;
    return blocks < 1 ? 1 : blocks;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_drude_cuda_configure)(int numRealAtoms, int paddedNumAtoms,
                                                                                        int numNormal, int numPairs) {
    cudaError_t e = cudaMemcpyToSymbol(NUM_ATOMS, &numRealAtoms, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(PADDED_NUM_ATOMS, &paddedNumAtoms, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(NUM_NORMAL_PARTICLES, &numNormal, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(NUM_PAIRS, &numPairs, sizeof(int));
    return (int) e;
}

/* scalars arrive as doubles and are narrowed to `mixed` exactly as CudaConstVKernels.cpp:284-312 does */
__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_drude_cuda_part1)(
        int numRealAtoms, void* velm, const long long* force, void* posDelta, const int* normalParticles, const void* pairParticles,
        const void* dt, double vscale, double fscale, double noisescale, double vscaleDrude, double fscaleDrude,
        double noisescaleDrude, const void* random, unsigned int randomIndex, int maxBlocks, void* stream) {
    integrateConstVDrudeLangevinPart1<<<gridFor(numRealAtoms, 64, maxBlocks), 64, 0, (cudaStream_t) stream>>>(
        (mixed4*) velm, force, (mixed4*) posDelta, normalParticles, (const int2*) pairParticles, (const mixed2*) dt, (mixed) vscale,
        (mixed) fscale, (mixed) noisescale, (mixed) vscaleDrude, (mixed) fscaleDrude, (mixed) noisescaleDrude,
        (const float4*) random, randomIndex);
    return (int) cudaGetLastError();
}

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_drude_cuda_part2)(
        int numRealAtoms, void* posq, void* posqCorrection, const void* posDelta, void* velm, const void* dt, int maxBlocks,
        void* stream) {
    integrateConstVDrudeLangevinPart2<<<gridFor(numRealAtoms, 64, maxBlocks), 64, 0, (cudaStream_t) stream>>>(
        (real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta, (mixed4*) velm, (const mixed2*) dt);
    return (int) cudaGetLastError();
}

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_drude_cuda_hardwall)(
        int numPairs, void* posq, void* posqCorrection, void* velm, const void* pairParticles, const void* dt,
        double maxDrudeDistance, double hardwallscaleDrude, int maxBlocks, void* stream) {
    applyHardWallConstraints<<<gridFor(numPairs, 64, maxBlocks), 64, 0, (cudaStream_t) stream>>>(
        (real4*) posq, (real4*) posqCorrection, (mixed4*) velm, (const int2*) pairParticles, (const mixed2*) dt,
        (mixed) maxDrudeDistance, (mixed) hardwallscaleDrude);
    return (int) cudaGetLastError();
}

}  // extern "C"
