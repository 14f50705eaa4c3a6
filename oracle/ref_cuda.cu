/*
 * ref_cuda.cu -- TEST INFRASTRUCTURE.  The reference's own CUDA kernels
 * (/root/reference/platforms/cuda/src/kernels/constVLangevin.cu, included from where it lies via
 * -DCONSTV_REF_KERNEL_SOURCE=...) compiled by nvcc for sm_100a behind plain C launchers, one
 * translation unit per precision model.  This is what OpenMM would build at run time on a B200
 * ("compiled at run time for whatever device OpenMM finds"); it is used
 *   - by tests/ to pin the product kernels and the oracle against the reference's arithmetic as
 *     nvcc contracts it, on the GPU;
 *   - by bench.py to time the reference's three launches per step beside the fused kernel.
 * Launch shape follows cu.executeKernel(kernel, args, workUnits, 128)
 * (CudaConstVKernels.cpp:149,160,166): 128-thread blocks, grid = min(ceil(work/128), maxBlocks),
 * maxBlocks <= 0 meaning uncapped.  Nothing under openmm_constv_b200/ links this file.
 */
#include <cuda_runtime.h>

#include "ref_preamble.h"

#include CONSTV_REF_KERNEL_SOURCE

namespace {
int gridFor(int work, int maxBlocks) {
    int blocks = (work + 127) / 128;
    if (maxBlocks > 0 && blocks > maxBlocks) blocks = maxBlocks;
    return blocks < 1 ? 1 : blocks;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_cuda_part1)(
        int numAtoms, int paddedNumAtoms, void* velm, const long long* force, void* posDelta, const void* paramBuffer,
        const void* dt, const void* random, unsigned int randomIndex, int maxBlocks, void* stream) {
    integrateConstVLangevinPart1<<<gridFor(numAtoms, maxBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, paddedNumAtoms, (mixed4*) velm, force, (mixed4*) posDelta, (const mixed*) paramBuffer, (const mixed2*) dt,
        (const float4*) random, randomIndex);
    return (int) cudaGetLastError();
}

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_cuda_part2)(
        int numAtoms, void* posq, void* posqCorrection, const void* posDelta, void* velm, const void* dt, int maxBlocks,
        void* stream) {
    integrateConstVLangevinPart2<<<gridFor(numAtoms, maxBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, (real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta, (mixed4*) velm, (const mixed2*) dt);
    return (int) cudaGetLastError();
}

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_cuda_mirror)(
        int numRealAtoms, int numCells, double zmax, void* posq, void* posqCorrection, int* invAtomOrder, int maxBlocks,
        void* stream) {
    updateImageParticlePositions<<<gridFor(numRealAtoms, maxBlocks), 128, 0, (cudaStream_t) stream>>>(
        numRealAtoms, numCells, zmax, (real4*) posq, (real4*) posqCorrection, invAtomOrder);
    return (int) cudaGetLastError();
}

__attribute__((visibility("default"))) int CONSTV_REF_NAME(constv_ref_cuda_inverse_order)(
        int numAtoms, int* atomIndex, int* invAtomOrder, int maxBlocks, void* stream) {
    reorderInverseAtomOrderIndexes<<<gridFor(numAtoms, maxBlocks), 128, 0, (cudaStream_t) stream>>>(numAtoms, atomIndex,
                                                                                                 invAtomOrder);
    return (int) cudaGetLastError();
}

}  // extern "C"
