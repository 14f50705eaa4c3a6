/*
 * ref_host.cpp -- TEST INFRASTRUCTURE.  The reference's kernel source
 * (/root/reference/platforms/cuda/src/kernels/constVLangevin.cu, included from where it lies via
 * -DCONSTV_REF_KERNEL_SOURCE=...) compiled UNMODIFIED by g++ for the host, with ref_preamble.h
 * supplying the CUDA dialect and OpenMM's typedef preamble.  The reference ships no CPU
 * implementation of this path; this is its only implementation run on host cores.
 *
 * Execution model: a launch is emulated as a grid of `work` one-thread blocks (blockDim.x = 1,
 * gridDim.x = work), so every emulated thread executes exactly one iteration of the kernel's
 * grid-stride loop; the blocks are spread over OpenMP threads in contiguous chunks, so every host
 * thread walks a contiguous range of atoms.  A sub-range [lo, hi) of the blocks can be run alone
 * (bench.py times bounded windows of a large slab that way).  The kernels have no inter-thread
 * communication (selectConstVLangevinStepSize, which has, is never launched by the reference:
 * CudaConstVKernels.cpp:67-70), so this is exact.
 *
 * Compiled with -ffp-contract=off: every operation is rounded separately.  nvcc fuses some of them
 * (see ref_cuda.cu and DESIGN.md "Canonical arithmetic"), so this build pins structure, operand
 * order, types and indexing (against the oracle built with -DORACLE_NO_CONTRACT), while the GPU
 * build pins the contraction.
 */
#include <omp.h>

#include "ref_preamble.h"

extern "C" {
mixed CONSTV_REF_NAME(constv_ref_dynamic_shared)[4096];
}

#include CONSTV_REF_KERNEL_SOURCE

namespace {

template <typename Call>
void emulateLaunch(int work, int lo, int hi, Call call) {
    blockDim.x = 1;
    gridDim.x = (unsigned int) (work < 1 ? 1 : work);
#pragma omp parallel for schedule(static)
    for (int block = lo; block < hi; block++) {
        blockIdx.x = (unsigned int) block;
        threadIdx.x = 0;
        call();
    }
}

template <typename Call>
void emulateLaunch(int work, Call call) {
    emulateLaunch(work, 0, work, call);
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_host_part1)(
        int numAtoms, int paddedNumAtoms, void* velm, const long long* force, void* posDelta, const void* paramBuffer,
        const void* dt, const void* random, unsigned int randomIndex) {
    emulateLaunch(numAtoms, [=]() {
        integrateConstVLangevinPart1(numAtoms, paddedNumAtoms, (mixed4*) velm, force, (mixed4*) posDelta,
                                     (const mixed*) paramBuffer, (const mixed2*) dt, (const float4*) random, randomIndex);
    });
}

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_host_part2)(
        int numAtoms, void* posq, void* posqCorrection, const void* posDelta, void* velm, const void* dt) {
    emulateLaunch(numAtoms, [=]() {
        integrateConstVLangevinPart2(numAtoms, (real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta, (mixed4*) velm,
                                     (const mixed2*) dt);
    });
}

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_host_mirror)(
        int numRealAtoms, int numCells, double zmax, void* posq, void* posqCorrection, int* invAtomOrder) {
    emulateLaunch(numRealAtoms, [=]() {
        updateImageParticlePositions(numRealAtoms, numCells, zmax, (real4*) posq, (real4*) posqCorrection, invAtomOrder);
    });
}

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_host_inverse_order)(
        int numAtoms, int* atomIndex, int* invAtomOrder) {
    emulateLaunch(numAtoms, [=]() { reorderInverseAtomOrderIndexes(numAtoms, atomIndex, invAtomOrder); });
}

/* One constraint-free step in the reference's launch order (CudaConstVKernels.cpp:146-166). */
__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_host_step)(
        int numAtoms, int paddedNumAtoms, int numCells, double zmax, void* posq, void* posqCorrection, void* velm,
        const long long* force, void* posDelta, const void* paramBuffer, const void* dt, const void* random,
        unsigned int randomIndex, int* invAtomOrder) {
    CONSTV_REF_NAME(constv_ref_host_part1)(numAtoms, paddedNumAtoms, velm, force, posDelta, paramBuffer, dt, random, randomIndex);
    CONSTV_REF_NAME(constv_ref_host_part2)(numAtoms, posq, posqCorrection, posDelta, velm, dt);
    CONSTV_REF_NAME(constv_ref_host_mirror)(numAtoms / numCells, numCells, zmax, posq, posqCorrection, invAtomOrder);
}

/* The same three launches restricted to the real atoms [lo, hi) and their images (identity slot
 * order): part 1 and part 2 visit the real slots and the image slots of the window, the image
 * rewrite the window's real indices.  lo = 0, hi = numAtoms/numCells is a whole step. */
__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_host_step_range)(
        int numAtoms, int paddedNumAtoms, int numCells, double zmax, void* posq, void* posqCorrection, void* velm,
        const long long* force, void* posDelta, const void* paramBuffer, const void* dt, const void* random,
        unsigned int randomIndex, int* invAtomOrder, int lo, int hi) {
    const int numReal = numAtoms / numCells;
    for (int cell = 0; cell < numCells; cell++)
        emulateLaunch(numAtoms, cell * numReal + lo, cell * numReal + hi, [=]() {
            integrateConstVLangevinPart1(numAtoms, paddedNumAtoms, (mixed4*) velm, force, (mixed4*) posDelta,
                                         (const mixed*) paramBuffer, (const mixed2*) dt, (const float4*) random, randomIndex);
        });
    for (int cell = 0; cell < numCells; cell++)
        emulateLaunch(numAtoms, cell * numReal + lo, cell * numReal + hi, [=]() {
            integrateConstVLangevinPart2(numAtoms, (real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta,
                                         (mixed4*) velm, (const mixed2*) dt);
        });
    emulateLaunch(numReal, lo, hi, [=]() {
        updateImageParticlePositions(numReal, numCells, zmax, (real4*) posq, (real4*) posqCorrection, invAtomOrder);
    });
}

}  // extern "C"

#if CONSTV_REF_PREC == 0
extern "C" __attribute__((visibility("default"))) int constv_ref_host_num_threads(void) { return omp_get_max_threads(); }
extern "C" __attribute__((visibility("default"))) void constv_ref_host_set_num_threads(int n) { omp_set_num_threads(n); }
#endif
