/*
 * ref_drude_host.cpp -- TEST INFRASTRUCTURE.  The reference's Drude kernels (vectorOps.cu +
 * constVDrudeLangevin.cu, included from where they lie under /root/reference) compiled UNMODIFIED by
 * g++ for the host; CUDA execution model emulated as in ref_host.cpp (one emulated thread per work
 * item), -ffp-contract=off.  NUM_ATOMS, PADDED_NUM_ATOMS, NUM_NORMAL_PARTICLES and NUM_PAIRS, which
 * OpenMM injects as literals (CudaConstVKernels.cpp:236-240), are plain ints set per call.
 */
#include <omp.h>

#include "ref_preamble.h"

static int CONSTV_REF_NAME(constv_ref_num_atoms), CONSTV_REF_NAME(constv_ref_padded_num_atoms),
    CONSTV_REF_NAME(constv_ref_num_normal), CONSTV_REF_NAME(constv_ref_num_pairs);
#define NUM_ATOMS CONSTV_REF_NAME(constv_ref_num_atoms)
#define PADDED_NUM_ATOMS CONSTV_REF_NAME(constv_ref_padded_num_atoms)
#define NUM_NORMAL_PARTICLES CONSTV_REF_NAME(constv_ref_num_normal)
#define NUM_PAIRS CONSTV_REF_NAME(constv_ref_num_pairs)

#include CONSTV_REF_VECTOROPS_SOURCE
#include CONSTV_REF_DRUDE_SOURCE

namespace {
template <typename Call>
void emulateLaunch(int work, Call call) {
    blockDim.x = 1;
    gridDim.x = (unsigned int) (work < 1 ? 1 : work);
#pragma omp parallel for schedule(static)
    for (int block = 0; block < (int) gridDim.x; block++) {
        blockIdx.x = (unsigned int) block;
        threadIdx.x = 0;
        call();
    }
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_drude_host_configure)(int numRealAtoms, int paddedNumAtoms,
                                                                                         int numNormal, int numPairs) {
    NUM_ATOMS = numRealAtoms;
    PADDED_NUM_ATOMS = paddedNumAtoms;
    NUM_NORMAL_PARTICLES = numNormal;
    NUM_PAIRS = numPairs;
}

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_drude_host_part1)(
        void* velm, const long long* force, void* posDelta, const int* normalParticles, const void* pairParticles, const void* dt,
        double vscale, double fscale, double noisescale, double vscaleDrude, double fscaleDrude, double noisescaleDrude,
        const void* random, unsigned int randomIndex) {
    const int work = NUM_NORMAL_PARTICLES > NUM_PAIRS ? NUM_NORMAL_PARTICLES : NUM_PAIRS;
    emulateLaunch(work, [=]() {
        integrateConstVDrudeLangevinPart1((mixed4*) velm, force, (mixed4*) posDelta, normalParticles, (const int2*) pairParticles,
                                          (const mixed2*) dt, (mixed) vscale, (mixed) fscale, (mixed) noisescale, (mixed) vscaleDrude,
                                          (mixed) fscaleDrude, (mixed) noisescaleDrude, (const float4*) random, randomIndex);
    });
}

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_drude_host_part2)(void* posq, void* posqCorrection,
                                                                                     const void* posDelta, void* velm, const void* dt) {
    emulateLaunch(NUM_ATOMS, [=]() {
        integrateConstVDrudeLangevinPart2((real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta, (mixed4*) velm,
                                          (const mixed2*) dt);
    });
}

__attribute__((visibility("default"))) void CONSTV_REF_NAME(constv_ref_drude_host_hardwall)(
        void* posq, void* posqCorrection, void* velm, const void* pairParticles, const void* dt, double maxDrudeDistance,
        double hardwallscaleDrude) {
    emulateLaunch(NUM_PAIRS, [=]() {
        applyHardWallConstraints((real4*) posq, (real4*) posqCorrection, (mixed4*) velm, (const int2*) pairParticles,
                                 (const mixed2*) dt, (mixed) maxDrudeDistance, (mixed) hardwallscaleDrude);
    });
}

}  // extern "C"
