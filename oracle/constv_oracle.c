/*
 * constv_oracle.c -- CPU ORACLE for the ConstVLangevinIntegrator step path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Nothing under openmm_constv_b200/ links,
 * imports or executes it; the product path has no CPU fallback.
 *
 * What it restates (all paths relative to /root/reference/):
 *   platforms/cuda/src/kernels/constVLangevin.cu:7-28     integrateConstVLangevinPart1
 *   platforms/cuda/src/kernels/constVLangevin.cu:34-75    integrateConstVLangevinPart2
 *   platforms/cuda/src/kernels/constVLangevin.cu:138-164  updateImageParticlePositions
 *   platforms/cuda/src/kernels/constVLangevin.cu:170-177  reorderInverseAtomOrderIndexes
 *   platforms/cuda/src/CudaConstVKernels.cpp:113-137      integration parameters
 *   platforms/cuda/src/CudaConstVKernels.cpp:76-96        system validation, zmax derivation
 *   platforms/cuda/src/CudaConstVKernels.cpp:139-172      per-step launch sequence
 *   platforms/reference/src/ReferenceConstVKernels.cpp:70-98  shifted kinetic energy formula
 *   platforms/cuda/src/kernels/constVDrudeLangevin.cu         the Drude integrator (constv_oracle_drude_impl.h)
 *   platforms/cuda/src/CudaConstVKernels.cpp:153              the constraint hand-off for rigid 3-site molecules
 *                                                             (SETTLE, constv_oracle_settle_impl.h; lives in OpenMM)
 *
 * PARITY STATUS.  The reference ships no CPU implementation of this path, no golden vectors and
 * no test that touches an image particle, and neither it nor OpenMM can
 * be built in this container.  The integrator arithmetic is pinned by the analytic properties the
 * reference's own CUDA test asserts (damped oscillator, friction=0 energy conservation,
 * equipartition, massless v == 0, seed determinism:
 * platforms/cuda/tests/TestCudaConstVLangevinIntegrator.cpp:52-268), re-expressed in
 * tests/test_oracle_properties.py.  The mirror / image handling is "PARITY UNPINNED": this
 * restatement, reviewed line by line against the .cu, is the de-facto specification.
 *
 * The Philox4x32-10 generator is not in the reference (it consumes OpenMM's random pool); it is
 * the published Random123 algorithm (Salmon et al., SC'11) and is pinned by the three Random123
 * known-answer vectors in tests/test_oracle_philox.py.
 */

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- three precision models --------------------------------------------------------------- */

#define REAL float
#define MIXED float
#define SUFFIX f32
#define MIXED_POS 0
#define MIXED_IS_DOUBLE 0
#include "constv_oracle_impl.h"
#include "constv_oracle_drude_impl.h"
#include "constv_oracle_settle_impl.h"
#undef REAL
#undef MIXED
#undef SUFFIX
#undef MIXED_POS
#undef MIXED_IS_DOUBLE

#define REAL float
#define MIXED double
#define SUFFIX mixed
#define MIXED_POS 1
#define MIXED_IS_DOUBLE 1
#include "constv_oracle_impl.h"
#include "constv_oracle_drude_impl.h"
#include "constv_oracle_settle_impl.h"
#undef REAL
#undef MIXED
#undef SUFFIX
#undef MIXED_POS
#undef MIXED_IS_DOUBLE

#define REAL double
#define MIXED double
#define SUFFIX f64
#define MIXED_POS 0
#define MIXED_IS_DOUBLE 1
#include "constv_oracle_impl.h"
#include "constv_oracle_drude_impl.h"
#include "constv_oracle_settle_impl.h"
#undef REAL
#undef MIXED
#undef SUFFIX
#undef MIXED_POS
#undef MIXED_IS_DOUBLE

/* ---- host-side pieces --------------------------------------------------------------------- */

/*
 * Integration parameters, CudaConstVKernels.cpp:113-137 (computed in double; the caller casts to
 * float for single precision exactly as :98-103 does).  boltz is OpenMM's BOLTZ (kJ/mol/K) from
 * SimTKOpenMMRealType.h, which is not in the reference tree; 8.31446261815324e-3 in OpenMM >= 7.4.
 */
void oracle_params(double temperature, double friction, double stepSize, double boltz,
                   double out[3]) {
    const double kT = boltz * temperature;
    const double vscale = exp(-stepSize * friction);
    const double fscale = (friction == 0 ? stepSize : (1 - vscale) / friction);
    const double noisescale = sqrt(kT * (1 - vscale * vscale));
    out[0] = vscale;
    out[1] = fscale;
    out[2] = noisescale;
}

/*
 * Parameters of the Drude integrator, CudaConstVKernels.cpp:262-270, in double (the caller narrows
 * them to float for single precision as :284-312 does).  NOTE: unlike the plain integrator there
 * is no friction == 0 special case in the reference: fscale = (1-vscale)/friction/2^32.
 * out[8] = vscale, fscale/2^32, noisescale, vscaleDrude, fscaleDrude/2^32, noisescaleDrude,
 *          maxDrudeDistance, hardwallscaleDrude = sqrt(kB*T_drude).
 */
void oracle_drude_params(double temperature, double friction, double drudeTemperature, double drudeFriction,
                         double maxDrudeDistance, double stepSize, double boltz, double out[8]) {
    const double vscale = exp(-stepSize * friction);
    const double fscale = (1 - vscale) / friction / (double) 0x100000000;
    const double noisescale = sqrt(2 * boltz * temperature * friction) * sqrt(0.5 * (1 - vscale * vscale) / friction);
    const double vscaleDrude = exp(-stepSize * drudeFriction);
    const double fscaleDrude = (1 - vscaleDrude) / drudeFriction / (double) 0x100000000;
    const double noisescaleDrude = sqrt(2 * boltz * drudeTemperature * drudeFriction) *
                                   sqrt(0.5 * (1 - vscaleDrude * vscaleDrude) / drudeFriction);
    out[0] = vscale; out[1] = fscale; out[2] = noisescale;
    out[3] = vscaleDrude; out[4] = fscaleDrude; out[5] = noisescaleDrude;
    out[6] = maxDrudeDistance;
    out[7] = sqrt(boltz * drudeTemperature);
}

/*
 * System validation + zmax derivation, CudaConstVKernels.cpp:76-96.
 * Returns 0 on success, else the 1-based index of the reference's error message:
 *   1 "Not even number of cells"                                        (:78)
 *   2 "Number of particles is not multiple of number of cells"          (:81)
 *   3 "Image Particle has nonzero mass"                                 (:84)
 *   4 "Unit cell dimension does not match with the provided zmax value" (:95)
 */
int oracle_validate(int numAtoms, int numCells, const double* masses, double boxZ, double cellSize,
                    double* zmaxOut) {
    if (numCells % 2 != 0)
        return 1;
    if (numAtoms % numCells != 0)
        return 2;
    for (int i = numAtoms / numCells; i < numAtoms; i++)
        if (masses[i] != 0.0)
            return 3;
    double zmax = cellSize;
    if (zmax < 0)
        zmax = boxZ / numCells;
    else if (zmax * numCells != boxZ)
        return 4;
    *zmaxOut = zmax;
    return 0;
}

/* constVLangevin.cu:170-177: invAtomOrder[atomIndex[slot]] = slot. */
void oracle_inverse_order(int numAtoms, const int* atomIndex, int* invAtomOrder) {
    for (int index = 0; index < numAtoms; index++)
        invAtomOrder[atomIndex[index]] = index;
}

/* ---- Philox4x32-10 (Random123) ------------------------------------------------------------ */

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t) 0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t) 0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t) p1;
    const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t) p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void oracle_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {counter[0], counter[1], counter[2], counter[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/*
 * The Gaussian stream of the B200 path (DESIGN.md "RNG"): one Philox block per (atom, step),
 * counter = (atom_lo, atom_hi, step_lo, step_hi), key = (seed_lo, seed_hi), atom = GLOBAL ORIGINAL
 * atom index (independent of slot order and of the multi-GPU partition).  Each 32-bit word gives a
 * uniform u = ((w >> 9) + 0.5) * 2^-23 in (0,1), exact in float; two Box-Muller pairs:
 *   a1 = u2*fl(2 pi) - fl(pi), a2 = u4*fl(2 pi) - fl(pi)   (fl = the float constants 0x40C90FDB /
 *        0x40490FDB, so the angle lies in (-pi, pi) and is defined identically on both sides)
 *   xi.x = r1*cos(a1), xi.y = r1*sin(a1), r1 = sqrt(-2 ln u1)
 *   xi.z = r2*cos(a2), xi.w = r2*sin(a2), r2 = sqrt(-2 ln u3)
 * Evaluated here in double and rounded to float; the device evaluates log/sqrt/sin/cos on its
 * special-function unit in float, so the comparison in tests carries a stated absolute tolerance.
 * The 32-bit Philox words themselves are bit-exact.
 */
void oracle_gaussian4(uint64_t seed, uint64_t atom, uint64_t step, float out[4]) {
    const uint32_t ctr[4] = {(uint32_t) atom, (uint32_t) (atom >> 32), (uint32_t) step,
                             (uint32_t) (step >> 32)};
    const uint32_t key[2] = {(uint32_t) seed, (uint32_t) (seed >> 32)};
    uint32_t w[4];
    oracle_philox4x32_10(ctr, key, w);
    double u[4];
    for (int i = 0; i < 4; i++)
        u[i] = ((double) (w[i] >> 9) + 0.5) * (1.0 / 8388608.0);
    const double twoPiF = (double) 6.2831854820251464844f;
    const double piF = (double) 3.1415927410125732422f;
    const double r1 = sqrt(-2.0 * log(u[0]));
    const double r2 = sqrt(-2.0 * log(u[2]));
    const double a1 = u[1] * twoPiF - piF;
    const double a2 = u[3] * twoPiF - piF;
    out[0] = (float) (r1 * cos(a1));
    out[1] = (float) (r1 * sin(a1));
    out[2] = (float) (r2 * cos(a2));
    out[3] = (float) (r2 * sin(a2));
}

/* Fill random4[slot] for slots [0, numAtoms) with the stream above; atomIndex may be NULL. */
void oracle_fill_noise(int numAtoms, const int* atomIndex, uint64_t seed, uint64_t globalAtomOffset,
                       uint64_t step, float* random4) {
#pragma omp parallel for schedule(static)
    for (int slot = 0; slot < numAtoms; slot++) {
        const uint64_t a = (uint64_t) (atomIndex ? atomIndex[slot] : slot) + globalAtomOffset;
        oracle_gaussian4(seed, a, step, random4 + 4 * (size_t) slot);
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}
