/*
 * TEST INFRASTRUCTURE ONLY -- included by constv_oracle.c once per precision mode (REAL, MIXED,
 * SUFFIX, MIXED_POS, MIXED_IS_DOUBLE as for constv_oracle_impl.h).
 *
 * Restates the ConstVDrudeLangevinIntegrator kernels of the reference:
 *   /root/reference/platforms/cuda/src/kernels/constVDrudeLangevin.cu:5-66    part 1 (normal particles + Drude pairs)
 *   /root/reference/platforms/cuda/src/kernels/constVDrudeLangevin.cu:72-103  part 2 (= the Langevin part 2 over the real atoms)
 *   /root/reference/platforms/cuda/src/kernels/constVDrudeLangevin.cu:108-211 hard-wall constraint on the Drude displacement
 * and their host sequence, /root/reference/platforms/cuda/src/CudaConstVKernels.cpp:252-349.
 *
 * The kernels address velm/posq/force by ORIGINAL particle index (normalParticles, pairParticles):
 * the reference does not translate them through atom reordering; only the image rewrite does.
 *
 * Canonical operation order: read off the SASS of the reference kernels compiled by nvcc for
 * sm_100a in single precision (oracle/_ref/libconstv_ref_cuda.so); M_FMA marks the fused
 * operations, every other operation is rounded separately; 1/x, a/b and sqrt are IEEE-correct
 * (rcp.rn / div.rn / sqrt.rn).  The ORACLE_NO_CONTRACT build un-fuses them (= the host build of
 * the reference, oracle/ref_drude_host.cpp).
 */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#if ORACLE_NO_CONTRACT
#define M_FMA(a, b, c) ((a) * (b) + (c))
#elif MIXED_IS_DOUBLE
#define M_FMA(a, b, c) fma((a), (b), (c))
#else
#define M_FMA(a, b, c) fmaf((a), (b), (c))
#endif
#if MIXED_IS_DOUBLE
#define M_SQRT(a) sqrt(a)
#define M_FABS(a) fabs(a)
#else
#define M_SQRT(a) sqrtf(a)
#define M_FABS(a) fabsf(a)
#endif

/*
 * Part 1, constVDrudeLangevin.cu:5-66.  s[6] = vscale, fscale (already divided by 2^32 on the host,
 * CudaConstVKernels.cpp:263,266), noisescale, vscaleDrude, fscaleDrude, noisescaleDrude, as `mixed`.
 *   :15-26  ordinary particles: the Langevin update; the Gaussian of particle `index` is
 *           random[randomIndex + index] (indexed by the particle, not by its rank in the list)
 *   :30     pairs read random[randomIndex + NUM_NORMAL_PARTICLES + 2i] and the next one
 *   :31-63  pair (particles.x = Drude particle, particles.y = parent): centre-of-mass motion
 *           thermalised at (T, friction), relative motion at (T_drude, friction_drude)
 */
void FN(oracle_drude_part1)(int numNormal, int numPairs, int paddedNumAtoms, MIXED* velm, const long long* force,
                            MIXED* posDelta, const int* normalParticles, const int* pairParticles, MIXED stepSize,
                            const MIXED* s, const float* random4, unsigned int randomIndex) {
    const MIXED vscale = s[0], fscale = s[1], noisescale = s[2], vscaleDrude = s[3], fscaleDrude = s[4],
                noisescaleDrude = s[5];
    const size_t P = (size_t) paddedNumAtoms;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < numNormal; i++) {
        const int index = normalParticles[i];
        MIXED* vel = velm + 4 * (size_t) index;
        const MIXED w = vel[3];
        if (w != 0) {
            const float* xi = random4 + 4 * ((size_t) randomIndex + (size_t) index);
            const MIXED fw = fscale * w;
            const MIXED ns = noisescale * M_SQRT(w);
            MIXED* d = posDelta + 4 * (size_t) index;
            for (int k = 0; k < 3; k++) {
                const MIXED t = fw * (MIXED) force[(size_t) index + P * k];
                vel[k] = M_FMA(ns, (MIXED) xi[k], M_FMA(vscale, vel[k], t));
                d[k] = stepSize * vel[k];
            }
            d[3] = 0;
        }
    }
    const unsigned int pairBase = randomIndex + (unsigned int) numNormal;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < numPairs; i++) {
        const int p1 = pairParticles[2 * i], p2 = pairParticles[2 * i + 1];
        MIXED* v1 = velm + 4 * (size_t) p1;
        MIXED* v2 = velm + 4 * (size_t) p2;
        const MIXED mass1 = (MIXED) 1 / v1[3];
        const MIXED mass2 = (MIXED) 1 / v2[3];
        const MIXED total = mass1 + mass2;
        const MIXED invTotalMass = (MIXED) 1 / total;
        const MIXED invReducedMass = (total * v1[3]) * v2[3];
        const MIXED mass1fract = invTotalMass * mass1;
        const MIXED mass2fract = invTotalMass * mass2;
        const MIXED sqrtInvTotalMass = M_SQRT(invTotalMass);
        const MIXED sqrtInvReducedMass = M_SQRT(invReducedMass);
        const float* rand1 = random4 + 4 * ((size_t) pairBase + 2 * (size_t) i);
        const float* rand2 = rand1 + 4;
        const MIXED fcm = fscale * invTotalMass;
        const MIXED frel = fscaleDrude * invReducedMass;
        const MIXED ncm = noisescale * sqrtInvTotalMass;
        const MIXED nrel = noisescaleDrude * sqrtInvReducedMass;
        MIXED* d1 = posDelta + 4 * (size_t) p1;
        MIXED* d2 = posDelta + 4 * (size_t) p2;
        for (int k = 0; k < 3; k++) {
            const MIXED f1 = (MIXED) force[(size_t) p1 + P * k];
            const MIXED f2 = (MIXED) force[(size_t) p2 + P * k];
            MIXED cm = M_FMA(v1[k], mass1fract, v2[k] * mass2fract);
            MIXED rel = v2[k] - v1[k];
            const MIXED cmForce = f1 + f2;
            const MIXED relForce = M_FMA(mass1fract, f2, -(mass2fract * f1));
            cm = M_FMA((MIXED) rand1[k], ncm, M_FMA(cm, vscale, cmForce * fcm));
            rel = M_FMA((MIXED) rand2[k], nrel, M_FMA(rel, vscaleDrude, relForce * frel));
            v1[k] = M_FMA(-mass2fract, rel, cm);
            v2[k] = M_FMA(mass1fract, rel, cm);
            d1[k] = stepSize * v1[k];
            d2[k] = stepSize * v2[k];
        }
        d1[3] = 0;
        d2[3] = 0;
    }
}

/* position (+ correction) of one particle as the `mixed` value the kernels work on */
static inline void FN(drude_load_pos)(const REAL* posq, const REAL* corr, size_t a, MIXED out[3]) {
#if MIXED_POS
    for (int k = 0; k < 3; k++) out[k] = (MIXED) posq[4 * a + k] + (MIXED) corr[4 * a + k];
#else
    (void) corr;
    for (int k = 0; k < 3; k++) out[k] = (MIXED) posq[4 * a + k];
#endif
}
static inline void FN(drude_store_pos)(REAL* posq, REAL* corr, size_t a, const MIXED in[3]) {
#if MIXED_POS
    for (int k = 0; k < 3; k++) {
        posq[4 * a + k] = (REAL) in[k];
        corr[4 * a + k] = (REAL) (in[k] - (MIXED) posq[4 * a + k]);
    }
    corr[4 * a + 3] = 0;
#else
    (void) corr;
    for (int k = 0; k < 3; k++) posq[4 * a + k] = (REAL) in[k];
#endif
}

/*
 * Hard wall, constVDrudeLangevin.cu:108-211: when a Drude particle is further than maxDrudeDistance
 * from its parent, the pair "bounces": positions are put back on the wall along the bond and the
 * bond-parallel velocities are re-drawn from the Drude temperature scale (hardwallscaleDrude =
 * sqrt(kB T_drude), CudaConstVKernels.cpp:270).  :141-165 massless parent, :166-208 both move.
 * In single and double precision the positions are read straight from posq (`mixed4 pos1 =
 * posq[...]`, :121-122); in mixed precision from posq + posqCorrection (:114-119).
 */
void FN(oracle_drude_hardwall)(int numPairs, REAL* posq, REAL* posqCorrection, MIXED* velm, const int* pairParticles,
                               MIXED stepSize, MIXED maxDrudeDistance, MIXED hardwallscaleDrude) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < numPairs; i++) {
        const size_t a1 = (size_t) pairParticles[2 * i], a2 = (size_t) pairParticles[2 * i + 1];
        MIXED pos1[3], pos2[3], delta[3], b[3];
        FN(drude_load_pos)(posq, posqCorrection, a1, pos1);
        FN(drude_load_pos)(posq, posqCorrection, a2, pos2);
        for (int k = 0; k < 3; k++) delta[k] = pos1[k] - pos2[k];
        const MIXED r = M_SQRT(M_FMA(delta[2], delta[2], M_FMA(delta[0], delta[0], delta[1] * delta[1])));
        const MIXED rInv = (MIXED) 1 / r;
        if (!(rInv * maxDrudeDistance < 1)) continue;
        for (int k = 0; k < 3; k++) b[k] = delta[k] * rInv;
        MIXED* vel1 = velm + 4 * a1;
        MIXED* vel2 = velm + 4 * a2;
        const MIXED mass1 = (MIXED) 1 / vel1[3];
        const MIXED mass2 = (MIXED) 1 / vel2[3];
        const MIXED deltaR = r - maxDrudeDistance;
        MIXED deltaT = stepSize;
        MIXED dotvr1 = M_FMA(b[2], vel1[2], M_FMA(b[0], vel1[0], b[1] * vel1[1]));
        MIXED vp1[3];
        for (int k = 0; k < 3; k++) vp1[k] = M_FMA(-b[k], dotvr1, vel1[k]);
        if (vel2[3] == 0) {
            /* the parent is massless: only the Drude particle moves */
            if (dotvr1 != 0) deltaT = deltaR / M_FABS(dotvr1);
            if (deltaT > stepSize) deltaT = stepSize;
            dotvr1 = (-dotvr1 * hardwallscaleDrude) / (M_FABS(dotvr1) * M_SQRT(mass1));
            const MIXED dr = M_FMA(deltaT, dotvr1, -deltaR);
            for (int k = 0; k < 3; k++) pos1[k] = M_FMA(b[k], dr, pos1[k]);
            FN(drude_store_pos)(posq, posqCorrection, a1, pos1);
            for (int k = 0; k < 3; k++) vel1[k] = M_FMA(b[k], dotvr1, vp1[k]);
        } else {
            const MIXED invTotalMass = (MIXED) 1 / (mass1 + mass2);
            MIXED dotvr2 = M_FMA(b[2], vel2[2], M_FMA(b[1], vel2[1], b[0] * vel2[0]));
            MIXED vp2[3];
            for (int k = 0; k < 3; k++) vp2[k] = M_FMA(-b[k], dotvr2, vel2[k]);
            const MIXED vbCMass = M_FMA(dotvr1, mass1, dotvr2 * mass2) * invTotalMass;
            dotvr1 = dotvr1 - vbCMass;
            dotvr2 = dotvr2 - vbCMass;
            if (dotvr1 != dotvr2) deltaT = deltaR / M_FABS(dotvr1 - dotvr2);
            if (deltaT > stepSize) deltaT = stepSize;
            const MIXED vBond = hardwallscaleDrude / M_SQRT(mass1);
            dotvr1 = (((-dotvr1 * vBond) * mass2) * invTotalMass) / M_FABS(dotvr1);
            dotvr2 = (((-dotvr2 * vBond) * mass1) * invTotalMass) / M_FABS(dotvr2);
            const MIXED dr1 = M_FMA(deltaT, dotvr1, -((deltaR * mass2) * invTotalMass));
            const MIXED dr2 = M_FMA(deltaR * mass1, invTotalMass, deltaT * dotvr2);
            dotvr1 = dotvr1 + vbCMass;
            dotvr2 = dotvr2 + vbCMass;
            for (int k = 0; k < 3; k++) {
                pos1[k] = M_FMA(b[k], dr1, pos1[k]);
                pos2[k] = M_FMA(b[k], dr2, pos2[k]);
            }
            FN(drude_store_pos)(posq, posqCorrection, a1, pos1);
            FN(drude_store_pos)(posq, posqCorrection, a2, pos2);
            for (int k = 0; k < 3; k++) {
                vel1[k] = M_FMA(b[k], dotvr1, vp1[k]);
                vel2[k] = M_FMA(b[k], dotvr2, vp2[k]);
            }
        }
    }
}

/*
 * One whole step without constraints in the reference's launch order, CudaConstVKernels.cpp:317-349:
 * part 1 -> [constraints: none] -> part 2 over the real atoms -> hard wall (only if
 * maxDrudeDistance > 0, :336) -> image rewrite; then the explicit image zeroing of the B200 path.
 * s[8] = the six part-1 scalars, maxDrudeDistance, hardwallscaleDrude.
 */
void FN(oracle_drude_step)(int numAtoms, int paddedNumAtoms, int numCells, double zshiftPerCell, REAL* posq,
                           REAL* posqCorrection, MIXED* velm, long long* force, MIXED* posDelta, int numNormal,
                           const int* normalParticles, int numPairs, const int* pairParticles, const MIXED* s,
                           MIXED stepSize, const float* random4, unsigned int randomIndex, const int* invAtomOrder,
                           int zeroVelm, int zeroForce) {
    const int numRealAtoms = numAtoms / numCells;
    FN(oracle_drude_part1)(numNormal, numPairs, paddedNumAtoms, velm, force, posDelta, normalParticles, pairParticles,
                           stepSize, s, random4, randomIndex);
    FN(oracle_part2)(numRealAtoms, posq, posqCorrection, posDelta, velm, stepSize);
    if (s[6] > 0)
        FN(oracle_drude_hardwall)(numPairs, posq, posqCorrection, velm, pairParticles, stepSize, s[6], s[7]);
    FN(oracle_mirror)(numRealAtoms, numCells, zshiftPerCell, posq, posqCorrection, invAtomOrder);
    FN(oracle_zero_images)(numAtoms, numRealAtoms, paddedNumAtoms, velm, force, invAtomOrder, zeroVelm, zeroForce);
}

#undef M_FMA
#undef M_SQRT
#undef M_FABS
#undef FN
#undef CAT
#undef CAT_
