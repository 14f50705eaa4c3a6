/*
 * TEST INFRASTRUCTURE ONLY -- included by constv_oracle.c once per precision mode.
 *
 * Before inclusion define:
 *   REAL   : float | double          (type of posq / posqCorrection)
 *   MIXED  : float | double          (type of velm / posDelta / params / dt)
 *   SUFFIX : f32 | mixed | f64
 *   MIXED_POS : 1 when REAL=float and MIXED=double (posqCorrection carries the low bits)
 *
 * The three modes are OpenMM's "single", "mixed" and "double" precision models, i.e. the
 * `real` / `mixed` typedefs that OpenMM's createModule injects in front of
 * /root/reference/platforms/cuda/src/kernels/constVLangevin.cu.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* ORACLE_NO_CONTRACT builds the variant in which every operation is rounded separately: what an
 * uncontracted host compilation of the reference's kernel source computes (oracle/ref_host.cpp).
 * The default variant fuses exactly the operations nvcc fuses in that source (checked in the SASS
 * of oracle/_ref/libconstv_ref_cuda.so and, bit for bit, on the GPU by tests/test_gpu_reference.py). */
#if ORACLE_NO_CONTRACT
#define M_FMA(a, b, c) ((a) * (b) + (c))
#define D_FMA(a, b, c) ((a) * (b) + (c))
#elif MIXED_IS_DOUBLE
#define M_FMA(a, b, c) fma((a), (b), (c))
#define D_FMA(a, b, c) fma((a), (b), (c))
#else
#define M_FMA(a, b, c) fmaf((a), (b), (c))
#define D_FMA(a, b, c) fma((a), (b), (c))
#endif
#if MIXED_IS_DOUBLE
#define M_SQRT(a) sqrt(a)
#else
#define M_SQRT(a) sqrtf(a)
#endif

/*
 * Part 1: velocity update and posDelta.
 * Follows constVLangevin.cu:7-28.
 *   :9-12  parameter fetch, fscale divided by 2^32 once
 *   :14,25 the random index advances with the slot index (one float4 per slot, massless included)
 *   :17    slots whose inverse mass is zero are skipped entirely
 *   :18-21 v <- vscale*v + fscale*invMass*F + noisescale*sqrt(invMass)*xi
 *   :23    posDelta <- dt*v, w = 0
 * Canonical operation order = what nvcc's default fmad contraction makes of the left-associated
 * source expression vscale*v + fscale*w*F + noisescale*sqrtInvMass*xi (SASS of the reference
 * kernel: FMUL t = (fscale*w)*F; FFMA(vscale, v, t); FFMA(ns, xi, .)):
 *   v = fma(ns, xi, fma(vscale, v, fw*F)), fw = fscale*w, ns = noisescale*sqrt(w).
 */
void FN(oracle_part1)(int numAtoms, int paddedNumAtoms, MIXED* velm, const long long* force,
                      MIXED* posDelta, const MIXED* params, MIXED stepSize, const float* random4,
                      unsigned int randomIndex) {
    const MIXED vscale = params[0];
    const MIXED fscale = params[1] / (MIXED) 4294967296.0;
    const MIXED noisescale = params[2];
#pragma omp parallel for schedule(static)
    for (int index = 0; index < numAtoms; index++) {
        MIXED* vel = velm + 4 * (size_t) index;
        const MIXED w = vel[3];
        if (w != 0) {
            const float* xi = random4 + 4 * ((size_t) randomIndex + (size_t) index);
            const MIXED sqrtInvMass = M_SQRT(w);
            const MIXED fw = fscale * w;
            const MIXED ns = noisescale * sqrtInvMass;
            for (int k = 0; k < 3; k++) {
                const MIXED f = (MIXED) force[(size_t) index + (size_t) paddedNumAtoms * k];
                const MIXED t = fw * f;
                vel[k] = M_FMA(ns, (MIXED) xi[k], M_FMA(vscale, vel[k], t));
            }
            MIXED* d = posDelta + 4 * (size_t) index;
            d[0] = stepSize * vel[0];
            d[1] = stepSize * vel[1];
            d[2] = stepSize * vel[2];
            d[3] = 0;
        }
    }
}

/*
 * Part 2: position update; velocity recomputed from the (possibly constrained) displacement.
 * Follows constVLangevin.cu:34-75, taking the __CUDA_ARCH__ >= 130 branch (:36, :57-59):
 * 1/dt is formed in double and the product is rounded once to `mixed`.
 * Mixed mode carries the position low bits in posqCorrection (:45-48, :65-67).
 */
void FN(oracle_part2)(int numAtoms, REAL* posq, REAL* posqCorrection, const MIXED* posDelta,
                      MIXED* velm, MIXED stepSize) {
    const double invStepSize = 1.0 / (double) stepSize;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < numAtoms; index++) {
        MIXED* vel = velm + 4 * (size_t) index;
        if (vel[3] != 0) {
            REAL* p = posq + 4 * (size_t) index;
            const MIXED* d = posDelta + 4 * (size_t) index;
#if MIXED_POS
            REAL* c = posqCorrection + 4 * (size_t) index;
            for (int k = 0; k < 3; k++) {
                MIXED pos = (MIXED) p[k] + (MIXED) c[k];
                pos += d[k];
                p[k] = (REAL) pos;
                c[k] = (REAL) (pos - (MIXED) p[k]);
            }
            c[3] = 0;
#else
            (void) posqCorrection;
            for (int k = 0; k < 3; k++)
                p[k] = (REAL) (p[k] + d[k]);
#endif
            for (int k = 0; k < 3; k++)
                vel[k] = (MIXED) (invStepSize * (double) d[k]);
        }
    }
}

/*
 * Image rewrite ("mirror").  Follows constVLangevin.cu:138-164.
 *   :142      the real atom is addressed through invAtomOrder (original index -> slot)
 *   :143      neutral atoms (q == -q) are skipped: their images keep whatever position they had
 *   :151-152  iterative z <- -z + zmax*(2i); zmax is a double kernel argument, so in single mode
 *             the expression is evaluated in double and rounded to float at the assignment,
 *             once per cell; in mixed mode `pos` is double and is only rounded at the store
 *   :153      the image keeps its own charge
 *   :155-158  store through invAtomOrder
 * zshiftPerCell is zmax for the reference behaviour; 0 gives the exact negation (x, y, -z)
 * named by BASELINE.json north_star (only meaningful for numCells == 2).
 * Canonical order: z = fma(zshiftPerCell, (double)(2i), -(double)z)  (one rounding in double).
 */
void FN(oracle_mirror)(int numRealAtoms, int numCells, double zshiftPerCell, REAL* posq,
                       REAL* posqCorrection, const int* invAtomOrder) {
#pragma omp parallel for schedule(static)
    for (int index = 0; index < numRealAtoms; index++) {
        const size_t slot0 = (size_t) invAtomOrder[index];
        const REAL* p0 = posq + 4 * slot0;
        if (p0[3] != -p0[3]) {
#if MIXED_POS
            const REAL* c0 = posqCorrection + 4 * slot0;
            MIXED x = (MIXED) p0[0] + (MIXED) c0[0];
            MIXED y = (MIXED) p0[1] + (MIXED) c0[1];
            MIXED z = (MIXED) p0[2] + (MIXED) c0[2];
#else
            (void) posqCorrection;
            REAL x = p0[0], y = p0[1], z = p0[2];
#endif
            for (int i = 1; i < numCells; i++) {
                const size_t slot = (size_t) invAtomOrder[(size_t) index + (size_t) numRealAtoms * i];
                REAL* pi = posq + 4 * slot;
#if MIXED_POS
                z = D_FMA(zshiftPerCell, (double) (2 * i), -z);
                REAL* ci = posqCorrection + 4 * slot;
                const REAL qi = pi[3];
                pi[0] = (REAL) x; pi[1] = (REAL) y; pi[2] = (REAL) z; pi[3] = qi;
                ci[0] = (REAL) (x - (MIXED) pi[0]);
                ci[1] = (REAL) (y - (MIXED) pi[1]);
                ci[2] = (REAL) (z - (MIXED) pi[2]);
                ci[3] = 0;
#else
                z = (REAL) D_FMA(zshiftPerCell, (double) (2 * i), -(double) z);
                pi[0] = x; pi[1] = y; pi[2] = z;
#endif
            }
        }
    }
}

/*
 * Image invariants made explicit (BASELINE.json north_star: "zeroing of image-particle velocities,
 * masses and forces").  The reference only *checks* mass == 0 at initialize
 * (CudaConstVKernels.cpp:76-85) and relies on parts 1/2 skipping w == 0 slots; the B200 path
 * rewrites velm of every image to (0,0,0,0) and its three fixed-point force entries to 0.
 */
void FN(oracle_zero_images)(int numAtoms, int numRealAtoms, int paddedNumAtoms, MIXED* velm,
                            long long* force, const int* invAtomOrder, int zeroVelm, int zeroForce) {
#pragma omp parallel for schedule(static)
    for (int a = numRealAtoms; a < numAtoms; a++) {
        const size_t slot = (size_t) invAtomOrder[a];
        if (zeroVelm) {
            MIXED* v = velm + 4 * slot;
            v[0] = 0; v[1] = 0; v[2] = 0; v[3] = 0;
        }
        if (zeroForce) {
            force[slot] = 0;
            force[slot + (size_t) paddedNumAtoms] = 0;
            force[slot + 2 * (size_t) paddedNumAtoms] = 0;
        }
    }
}

/*
 * Half-step shifted kinetic energy.  The reference delegates to OpenMM
 * (CudaConstVKernels.cpp:175-177, timeShift = 0.5*dt); the formula as written in-tree is
 * platforms/reference/src/ReferenceConstVKernels.cpp:70-98:
 *   v' = v + F*(timeShift*invMass) for massive atoms;  KE = 0.5 * sum (v'.v')/invMass.
 * (Velocity constraints are outside this path: no constraints in the standalone slabs.)
 * Canonical: scale = (mixed)(timeShift/2^32); v'_k = fma(scale*(mixed)F_k, w, v_k); sum in double.
 */
double FN(oracle_kinetic_energy)(int numAtoms, int paddedNumAtoms, const MIXED* velm,
                                 const long long* force, double timeShift) {
    const MIXED scale = (MIXED) (timeShift / 4294967296.0);
    double energy = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : energy)
    for (int index = 0; index < numAtoms; index++) {
        const MIXED* vel = velm + 4 * (size_t) index;
        const MIXED w = vel[3];
        if (w != 0) {
            double e = 0.0;
            for (int k = 0; k < 3; k++) {
                const MIXED f = (MIXED) force[(size_t) index + (size_t) paddedNumAtoms * k];
                const MIXED v = M_FMA(scale * f, w, vel[k]);
                e += (double) v * (double) v;
            }
            energy += e / (double) w;
        }
    }
    return 0.5 * energy;
}

/*
 * One whole step without constraints, in the reference's launch order
 * (CudaConstVKernels.cpp:139-172): part 1 -> [constraints: none] -> part 2 -> mirror, followed by
 * the explicit image zeroing.  posDelta is scratch of 4*paddedNumAtoms MIXED.
 */
void FN(oracle_step)(int numAtoms, int paddedNumAtoms, int numCells, double zshiftPerCell,
                     REAL* posq, REAL* posqCorrection, MIXED* velm, long long* force,
                     MIXED* posDelta, const MIXED* params, MIXED stepSize, const float* random4,
                     unsigned int randomIndex, const int* invAtomOrder, int zeroVelm,
                     int zeroForce) {
    const int numRealAtoms = numAtoms / numCells;
    FN(oracle_part1)(numAtoms, paddedNumAtoms, velm, force, posDelta, params, stepSize, random4,
                     randomIndex);
    FN(oracle_part2)(numAtoms, posq, posqCorrection, posDelta, velm, stepSize);
    FN(oracle_mirror)(numRealAtoms, numCells, zshiftPerCell, posq, posqCorrection, invAtomOrder);
    FN(oracle_zero_images)(numAtoms, numRealAtoms, paddedNumAtoms, velm, force, invAtomOrder,
                           zeroVelm, zeroForce);
}


/*
 * The same step restricted to the real atoms [lo, hi) and their images (identity slot order only).
 * Used by bench.py's CPU legs to time a bounded window of a large slab; a full step is
 * lo = 0, hi = numAtoms/numCells and gives exactly what oracle_step gives.
 */
void FN(oracle_step_range)(int numAtoms, int paddedNumAtoms, int numCells, double zshiftPerCell,
                           REAL* posq, REAL* posqCorrection, MIXED* velm, long long* force,
                           const MIXED* params, MIXED stepSize, const float* random4,
                           unsigned int randomIndex, int zeroVelm, int zeroForce, int lo, int hi) {
    const int R = numAtoms / numCells;
    const MIXED vscale = params[0];
    const MIXED fscale = params[1] / (MIXED) 4294967296.0;
    const MIXED noisescale = params[2];
    const double invStepSize = 1.0 / (double) stepSize;
    const size_t P = (size_t) paddedNumAtoms;
#pragma omp parallel for schedule(static)
    for (int index = lo; index < hi; index++) {
        MIXED* vel = velm + 4 * (size_t) index;
        REAL* p = posq + 4 * (size_t) index;
        const MIXED w = vel[3];
        if (w != 0) {
            const float* xi = random4 + 4 * ((size_t) randomIndex + (size_t) index);
            const MIXED fw = fscale * w;
            const MIXED ns = noisescale * M_SQRT(w);
            MIXED d[3];
            for (int k = 0; k < 3; k++) {
                const MIXED f = (MIXED) force[(size_t) index + P * k];
                vel[k] = M_FMA(ns, (MIXED) xi[k], M_FMA(vscale, vel[k], fw * f));
                d[k] = stepSize * vel[k];
            }
#if MIXED_POS
            REAL* c = posqCorrection + 4 * (size_t) index;
            for (int k = 0; k < 3; k++) {
                MIXED pos = (MIXED) p[k] + (MIXED) c[k];
                pos += d[k];
                p[k] = (REAL) pos;
                c[k] = (REAL) (pos - (MIXED) p[k]);
            }
            c[3] = 0;
#else
            for (int k = 0; k < 3; k++)
                p[k] = (REAL) (p[k] + d[k]);
#endif
            for (int k = 0; k < 3; k++)
                vel[k] = (MIXED) (invStepSize * (double) d[k]);
        }
        if (p[3] != -p[3]) {
#if MIXED_POS
            const REAL* c0 = posqCorrection + 4 * (size_t) index;
            MIXED x = (MIXED) p[0] + (MIXED) c0[0], y = (MIXED) p[1] + (MIXED) c0[1],
                  z = (MIXED) p[2] + (MIXED) c0[2];
#else
            REAL x = p[0], y = p[1], z = p[2];
#endif
            for (int i = 1; i < numCells; i++) {
                REAL* pi = posq + 4 * ((size_t) index + (size_t) R * i);
#if MIXED_POS
                z = D_FMA(zshiftPerCell, (double) (2 * i), -z);
                REAL* ci = posqCorrection + 4 * ((size_t) index + (size_t) R * i);
                pi[0] = (REAL) x; pi[1] = (REAL) y; pi[2] = (REAL) z;
                ci[0] = (REAL) (x - (MIXED) pi[0]);
                ci[1] = (REAL) (y - (MIXED) pi[1]);
                ci[2] = (REAL) (z - (MIXED) pi[2]);
                ci[3] = 0;
#else
                z = (REAL) D_FMA(zshiftPerCell, (double) (2 * i), -(double) z);
                pi[0] = x; pi[1] = y; pi[2] = z;
#endif
            }
        }
        for (int i = 1; i < numCells; i++) {
            const size_t slot = (size_t) index + (size_t) R * i;
            if (zeroVelm) {
                MIXED* v = velm + 4 * slot;
                v[0] = 0; v[1] = 0; v[2] = 0; v[3] = 0;
            }
            if (zeroForce) {
                force[slot] = 0; force[slot + P] = 0; force[slot + 2 * P] = 0;
            }
        }
    }
}

#undef M_FMA
#undef D_FMA
#undef M_SQRT
#undef FN
#undef CAT
#undef CAT_
