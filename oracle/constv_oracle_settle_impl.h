/*
 * TEST INFRASTRUCTURE ONLY -- included by constv_oracle.c once per precision mode (REAL, MIXED,
 * SUFFIX, MIXED_POS, MIXED_IS_DOUBLE as for constv_oracle_impl.h).
 *
 * The constraint hand-off of the reference, /root/reference/platforms/cuda/src/CudaConstVKernels.cpp:153
 * (`integration.applyConstraints(tol)` between part 1 and part 2), for the systems the reference's
 * example runs: SPC/E water with constraints=AllBonds (example/nacl_1m_spce/nacl_constv.py:56) is
 * three distance constraints per rigid 3-site molecule, which OpenMM's CUDA platform solves
 * analytically with SETTLE (Miyamoto & Kollman, J. Comput. Chem. 13, 952 (1992)).
 *
 * PARITY UNPINNED.  The solver lives in OpenMM (un-vendored, unpinned, absent here: SURVEY.md 8c), not
 * in the reference tree.  What is restated below is the published algorithm in the formulation OpenMM's
 * CUDA kernel `applySettleToPositions` uses [recalled]: positions relative to the apex atom's old
 * position, the new positions given as displacements (posDelta), masses from velm.w, parameters
 * (d(apex-b) = d(apex-c), d(b-c)) as floats.  It is validated against an INDEPENDENT solution of the same
 * constraint problem (iterative SHAKE in float64, oracle/settle_numpy.py) and by its invariants
 * (distances restored, centre of mass and the molecule's angular momentum direction unchanged) in
 * tests/test_oracle_settle.py -- not against OpenMM.
 *
 * Arithmetic: the canonical sequence below -- fused multiply-adds exactly where S_FMA says, everything else
 * rounded separately, square roots / divisions / reciprocals IEEE-correct, one reciprocal per frame axis; the
 * CUDA kernel spells the same sequence with explicit intrinsics.
 */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#if MIXED_IS_DOUBLE
#define M_SQRT(a) sqrt(a)
#define S_FMA(a, b, c) fma((a), (b), (c))
#else
#define M_SQRT(a) sqrtf(a)
#define S_FMA(a, b, c) fmaf((a), (b), (c))
#endif

/*
 * One molecule.  pos0/pos1/pos2 = old (constrained) positions as `mixed`; d0/d1/d2 = displacements to the
 * unconstrained new positions, overwritten with the constrained displacements; m = masses; distAB =
 * apex-b = apex-c distance, distBC = b-c distance.
 */
static inline void FN(settle_molecule)(const MIXED pos0[3], const MIXED pos1[3], const MIXED pos2[3], MIXED d0[3],
                                       MIXED d1[3], MIXED d2[3], MIXED m0, MIXED m1, MIXED m2, float distAB,
                                       float distBC) {
    /* old bond vectors, from the apex */
    const MIXED xb0 = pos1[0] - pos0[0], yb0 = pos1[1] - pos0[1], zb0 = pos1[2] - pos0[2];
    const MIXED xc0 = pos2[0] - pos0[0], yc0 = pos2[1] - pos0[1], zc0 = pos2[2] - pos0[2];

    const MIXED invTotalMass = (MIXED) 1 / (m0 + m1 + m2);
    const MIXED xcom = S_FMA(xc0 + d2[0], m2, S_FMA(xb0 + d1[0], m1, d0[0] * m0)) * invTotalMass;
    const MIXED ycom = S_FMA(yc0 + d2[1], m2, S_FMA(yb0 + d1[1], m1, d0[1] * m0)) * invTotalMass;
    const MIXED zcom = S_FMA(zc0 + d2[2], m2, S_FMA(zb0 + d1[2], m1, d0[2] * m0)) * invTotalMass;

    /* unconstrained new positions relative to the new centre of mass */
    const MIXED xa1 = d0[0] - xcom, ya1 = d0[1] - ycom, za1 = d0[2] - zcom;
    const MIXED xb1 = xb0 + d1[0] - xcom, yb1 = yb0 + d1[1] - ycom, zb1 = zb0 + d1[2] - zcom;
    const MIXED xc1 = xc0 + d2[0] - xcom, yc1 = yc0 + d2[1] - ycom, zc1 = zc0 + d2[2] - zcom;

    /* orthonormal frame: Z = normal of the old molecular plane, X = a1 x Z, Y = Z x X */
    const MIXED xaksZd = S_FMA(yb0, zc0, -(zb0 * yc0));
    const MIXED yaksZd = S_FMA(zb0, xc0, -(xb0 * zc0));
    const MIXED zaksZd = S_FMA(xb0, yc0, -(yb0 * xc0));
    const MIXED xaksXd = S_FMA(ya1, zaksZd, -(za1 * yaksZd));
    const MIXED yaksXd = S_FMA(za1, xaksZd, -(xa1 * zaksZd));
    const MIXED zaksXd = S_FMA(xa1, yaksZd, -(ya1 * xaksZd));
    const MIXED xaksYd = S_FMA(yaksZd, zaksXd, -(zaksZd * yaksXd));
    const MIXED yaksYd = S_FMA(zaksZd, xaksXd, -(xaksZd * zaksXd));
    const MIXED zaksYd = S_FMA(xaksZd, yaksXd, -(yaksZd * xaksXd));

    /* one correctly rounded square root and one correctly rounded reciprocal per axis */
    const MIXED invX = (MIXED) 1 / M_SQRT(S_FMA(zaksXd, zaksXd, S_FMA(yaksXd, yaksXd, xaksXd * xaksXd)));
    const MIXED invY = (MIXED) 1 / M_SQRT(S_FMA(zaksYd, zaksYd, S_FMA(yaksYd, yaksYd, xaksYd * xaksYd)));
    const MIXED invZ = (MIXED) 1 / M_SQRT(S_FMA(zaksZd, zaksZd, S_FMA(yaksZd, yaksZd, xaksZd * xaksZd)));
    const MIXED trns11 = xaksXd * invX, trns21 = yaksXd * invX, trns31 = zaksXd * invX;
    const MIXED trns12 = xaksYd * invY, trns22 = yaksYd * invY, trns32 = zaksYd * invY;
    const MIXED trns13 = xaksZd * invZ, trns23 = yaksZd * invZ, trns33 = zaksZd * invZ;

#define S_DOT3(a1, b1, a2, b2, a3, b3) S_FMA(a3, b3, S_FMA(a2, b2, (a1) * (b1)))
    const MIXED xb0d = S_DOT3(trns11, xb0, trns21, yb0, trns31, zb0);
    const MIXED yb0d = S_DOT3(trns12, xb0, trns22, yb0, trns32, zb0);
    const MIXED xc0d = S_DOT3(trns11, xc0, trns21, yc0, trns31, zc0);
    const MIXED yc0d = S_DOT3(trns12, xc0, trns22, yc0, trns32, zc0);
    const MIXED za1d = S_DOT3(trns13, xa1, trns23, ya1, trns33, za1);
    const MIXED xb1d = S_DOT3(trns11, xb1, trns21, yb1, trns31, zb1);
    const MIXED yb1d = S_DOT3(trns12, xb1, trns22, yb1, trns32, zb1);
    const MIXED zb1d = S_DOT3(trns13, xb1, trns23, yb1, trns33, zb1);
    const MIXED xc1d = S_DOT3(trns11, xc1, trns21, yc1, trns31, zc1);
    const MIXED yc1d = S_DOT3(trns12, xc1, trns22, yc1, trns32, zc1);
    const MIXED zc1d = S_DOT3(trns13, xc1, trns23, yc1, trns33, zc1);

    /* step 2: the canonical triangle (apex at +ra on Y, b and c at -rb, x = -+rc), tilted by phi and psi.
     * The squared parameters are float arithmetic (float2 parameters in the formulation restated). */
    const float rc = 0.5f * distBC;
    MIXED rb = M_SQRT((MIXED) (distAB * distAB - rc * rc));
    const MIXED ra = rb * (m1 + m2) * invTotalMass;
    rb = rb - ra;
    const MIXED sinphi = za1d / ra;
    const MIXED cosphi = M_SQRT(S_FMA(-sinphi, sinphi, (MIXED) 1));
    const MIXED sinpsi = (zb1d - zc1d) / ((MIXED) 2 * rc * cosphi);
    const MIXED cospsi = M_SQRT(S_FMA(-sinpsi, sinpsi, (MIXED) 1));

    const MIXED ya2d = ra * cosphi;
    MIXED xb2d = -(rc * cospsi);
    const MIXED tilt = rc * sinpsi * sinphi;
    const MIXED yb2d = -(rb * cosphi) - tilt;
    const MIXED yc2d = -(rb * cosphi) + tilt;
    const MIXED xb2d2 = xb2d * xb2d;
    const MIXED hh2 = S_FMA(zb1d - zc1d, zb1d - zc1d, S_FMA(yb2d - yc2d, yb2d - yc2d, (MIXED) 4 * xb2d2));
    const MIXED deltx = S_FMA((MIXED) 2, xb2d, M_SQRT(S_FMA((MIXED) 4, xb2d2, -hh2) + (MIXED) (distBC * distBC)));
    xb2d = S_FMA(-deltx, (MIXED) 0.5, xb2d);

    /* step 3: the in-plane rotation theta that conserves the angular momentum about Z */
    const MIXED alpha = S_FMA(yc0d, yc2d, S_FMA(yb0d, yb2d, xb2d * (xb0d - xc0d)));
    const MIXED beta = S_FMA(xc0d, yc2d, S_FMA(xb0d, yb2d, xb2d * (yc0d - yb0d)));
    const MIXED gamma = S_FMA(xc0d, yc1d, -(xc1d * yc0d)) + S_FMA(xb0d, yb1d, -(xb1d * yb0d));
    const MIXED al2be2 = S_FMA(beta, beta, alpha * alpha);
    const MIXED sintheta = S_FMA(alpha, gamma, -(beta * M_SQRT(S_FMA(-gamma, gamma, al2be2)))) / al2be2;

    /* step 4 */
    const MIXED costheta = M_SQRT(S_FMA(-sintheta, sintheta, (MIXED) 1));
    const MIXED xa3d = -(ya2d * sintheta);
    const MIXED ya3d = ya2d * costheta;
    const MIXED za3d = za1d;
    const MIXED xb3d = S_FMA(xb2d, costheta, -(yb2d * sintheta));
    const MIXED yb3d = S_FMA(xb2d, sintheta, yb2d * costheta);
    const MIXED zb3d = zb1d;
    const MIXED xc3d = -S_FMA(xb2d, costheta, yc2d * sintheta);
    const MIXED yc3d = S_FMA(-xb2d, sintheta, yc2d * costheta);
    const MIXED zc3d = zc1d;

    /* step 5: back to the laboratory frame, and to displacements */
    d0[0] = xcom + S_DOT3(trns11, xa3d, trns12, ya3d, trns13, za3d);
    d0[1] = ycom + S_DOT3(trns21, xa3d, trns22, ya3d, trns23, za3d);
    d0[2] = zcom + S_DOT3(trns31, xa3d, trns32, ya3d, trns33, za3d);
    d1[0] = xcom + S_DOT3(trns11, xb3d, trns12, yb3d, trns13, zb3d) - xb0;
    d1[1] = ycom + S_DOT3(trns21, xb3d, trns22, yb3d, trns23, zb3d) - yb0;
    d1[2] = zcom + S_DOT3(trns31, xb3d, trns32, yb3d, trns33, zb3d) - zb0;
    d2[0] = xcom + S_DOT3(trns11, xc3d, trns12, yc3d, trns13, zc3d) - xc0;
    d2[1] = ycom + S_DOT3(trns21, xc3d, trns22, yc3d, trns23, zc3d) - yc0;
    d2[2] = zcom + S_DOT3(trns31, xc3d, trns32, yc3d, trns33, zc3d) - zc0;
#undef S_DOT3
}

static inline void FN(settle_load_pos)(const REAL* posq, const REAL* corr, size_t a, MIXED out[3]) {
#if MIXED_POS
    for (int k = 0; k < 3; k++) out[k] = (MIXED) posq[4 * a + k] + (MIXED) corr[4 * a + k];
#else
    (void) corr;
    for (int k = 0; k < 3; k++) out[k] = (MIXED) posq[4 * a + k];
#endif
}

/*
 * The constraint step on posDelta for every molecule: clusterAtoms = (apex, b, c) slots, clusterParams =
 * (d(apex-b), d(b-c)) floats.  posq/posqCorrection hold the positions of the previous step.
 */
void FN(oracle_settle_positions)(int numClusters, const REAL* posq, const REAL* posqCorrection, MIXED* posDelta,
                                 const MIXED* velm, const int* clusterAtoms, const float* clusterParams) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < numClusters; i++) {
        const size_t a0 = (size_t) clusterAtoms[3 * i], a1 = (size_t) clusterAtoms[3 * i + 1], a2 = (size_t) clusterAtoms[3 * i + 2];
        MIXED p0[3], p1[3], p2[3];
        FN(settle_load_pos)(posq, posqCorrection, a0, p0);
        FN(settle_load_pos)(posq, posqCorrection, a1, p1);
        FN(settle_load_pos)(posq, posqCorrection, a2, p2);
        const MIXED m0 = (MIXED) 1 / velm[4 * a0 + 3], m1 = (MIXED) 1 / velm[4 * a1 + 3], m2 = (MIXED) 1 / velm[4 * a2 + 3];
        FN(settle_molecule)(p0, p1, p2, posDelta + 4 * a0, posDelta + 4 * a1, posDelta + 4 * a2, m0, m1, m2,
                            clusterParams[2 * i], clusterParams[2 * i + 1]);
    }
}

/*
 * One whole step of a system whose constraints are all rigid 3-site molecules, in the reference's launch
 * order (CudaConstVKernels.cpp:146-166): part 1 -> constraints (SETTLE on posDelta) -> part 2 -> image
 * rewrite; then the explicit image zeroing of the B200 path.
 */
void FN(oracle_settle_step)(int numAtoms, int paddedNumAtoms, int numCells, double zshiftPerCell, REAL* posq,
                            REAL* posqCorrection, MIXED* velm, long long* force, MIXED* posDelta, const MIXED* params,
                            MIXED stepSize, const float* random4, unsigned int randomIndex, const int* invAtomOrder,
                            int zeroVelm, int zeroForce, int numClusters, const int* clusterAtoms,
                            const float* clusterParams) {
    const int numRealAtoms = numAtoms / numCells;
    FN(oracle_part1)(numAtoms, paddedNumAtoms, velm, force, posDelta, params, stepSize, random4, randomIndex);
    FN(oracle_settle_positions)(numClusters, posq, posqCorrection, posDelta, velm, clusterAtoms, clusterParams);
    FN(oracle_part2)(numAtoms, posq, posqCorrection, posDelta, velm, stepSize);
    FN(oracle_mirror)(numRealAtoms, numCells, zshiftPerCell, posq, posqCorrection, invAtomOrder);
    FN(oracle_zero_images)(numAtoms, numRealAtoms, paddedNumAtoms, velm, force, invAtomOrder, zeroVelm, zeroForce);
}


/*
 * Velocity constraints of one rigid molecule, in double whatever the precision model (an on-demand report, not
 * the step): v_i' = v_i + w_i * sum_p s_ip lambda_p r_p over the pairs p0 = (0,1), p1 = (0,2), p2 = (1,2), with
 * (v_c' - v_d') . r_q = 0 for every pair q = (c,d).  A 3x3 linear system in the lambdas, solved by Cramer's
 * rule.  This is what OpenMM's applyVelocityConstraints does for these molecules inside
 * integration.computeKineticEnergy (reached from CudaConstVKernels.cpp:175-177); PARITY UNPINNED like the
 * position solver, pinned on its defining properties (tests/test_oracle_settle.py).  Every operation rounded
 * separately; the CUDA kernel spells the same sequence.
 */
static inline void FN(settle_velocity_molecule)(const double x[3][3], double v[3][3], const double w[3]) {
    double r[3][3];
    for (int k = 0; k < 3; k++) {
        r[0][k] = x[0][k] - x[1][k];
        r[1][k] = x[0][k] - x[2][k];
        r[2][k] = x[1][k] - x[2][k];
    }
    double g[3][3];
    for (int p = 0; p < 3; p++)
        for (int q = 0; q < 3; q++) g[p][q] = r[p][0] * r[q][0] + r[p][1] * r[q][1] + r[p][2] * r[q][2];
    const double a00 = (w[0] + w[1]) * g[0][0], a01 = w[0] * g[1][0], a02 = -w[1] * g[2][0];
    const double a10 = w[0] * g[0][1], a11 = (w[0] + w[2]) * g[1][1], a12 = w[2] * g[2][1];
    const double a20 = -w[1] * g[0][2], a21 = w[2] * g[1][2], a22 = (w[1] + w[2]) * g[2][2];
    double b[3];
    b[0] = -((v[0][0] - v[1][0]) * r[0][0] + (v[0][1] - v[1][1]) * r[0][1] + (v[0][2] - v[1][2]) * r[0][2]);
    b[1] = -((v[0][0] - v[2][0]) * r[1][0] + (v[0][1] - v[2][1]) * r[1][1] + (v[0][2] - v[2][2]) * r[1][2]);
    b[2] = -((v[1][0] - v[2][0]) * r[2][0] + (v[1][1] - v[2][1]) * r[2][1] + (v[1][2] - v[2][2]) * r[2][2]);
    const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    const double l0 = (b[0] * c00 + a01 * (a12 * b[2] - b[1] * a22) + a02 * (b[1] * a21 - a11 * b[2])) / det;
    const double l1 = (a00 * (b[1] * a22 - a12 * b[2]) + b[0] * c01 + a02 * (a10 * b[2] - b[1] * a20)) / det;
    const double l2 = (a00 * (a11 * b[2] - b[1] * a21) + a01 * (b[1] * a20 - a10 * b[2]) + b[0] * c02) / det;
    for (int k = 0; k < 3; k++) {
        v[0][k] = v[0][k] + w[0] * (l0 * r[0][k] + l1 * r[1][k]);
        v[1][k] = v[1][k] + w[1] * (l2 * r[2][k] - l0 * r[0][k]);
        v[2][k] = v[2][k] - w[2] * (l1 * r[1][k] + l2 * r[2][k]);
    }
}

/*
 * Kinetic energy of a system with rigid molecules: the shifted velocities v + timeShift*F/m of
 * oracle_kinetic_energy, projected onto the constraints for the molecule atoms, KE = 0.5 sum m |v'|^2.
 */
double FN(oracle_kinetic_energy_rigid)(int numAtoms, int paddedNumAtoms, const REAL* posq, const REAL* posqCorrection,
                                       const MIXED* velm, const long long* force, double timeShift, int numClusters,
                                       const int* clusterAtoms, double* projectedVelocities) {
    const MIXED scale = (MIXED) (timeShift / 4294967296.0);
    const size_t P = (size_t) paddedNumAtoms;
    unsigned char* inMolecule = (unsigned char*) calloc((size_t) numAtoms + 1, 1);
    double energy = 0.0;
    for (int i = 0; i < numClusters; i++) {
        double x[3][3], v[3][3], w[3];
        for (int a = 0; a < 3; a++) {
            const size_t at = (size_t) clusterAtoms[3 * i + a];
            inMolecule[at] = 1;
            MIXED pos[3];
            FN(settle_load_pos)(posq, posqCorrection, at, pos);
            w[a] = (double) velm[4 * at + 3];
            for (int k = 0; k < 3; k++) {
                x[a][k] = (double) pos[k];
                const MIXED f = (MIXED) force[at + P * k];
                v[a][k] = (double) S_FMA(scale * f, velm[4 * at + 3], velm[4 * at + k]);
            }
        }
        FN(settle_velocity_molecule)(x, v, w);
        for (int a = 0; a < 3; a++) {
            energy += (v[a][0] * v[a][0] + v[a][1] * v[a][1] + v[a][2] * v[a][2]) / w[a];
            if (projectedVelocities)
                for (int k = 0; k < 3; k++) projectedVelocities[3 * (size_t) clusterAtoms[3 * i + a] + k] = v[a][k];
        }
    }
    for (int index = 0; index < numAtoms; index++) {
        const MIXED* vel = velm + 4 * (size_t) index;
        const MIXED w = vel[3];
        if (w != 0 && !inMolecule[index]) {
            double e = 0.0;
            for (int k = 0; k < 3; k++) {
                const MIXED f = (MIXED) force[(size_t) index + P * k];
                const MIXED v = S_FMA(scale * f, w, vel[k]);
                e += (double) v * (double) v;
                if (projectedVelocities) projectedVelocities[3 * (size_t) index + k] = (double) v;
            }
            energy += e / (double) w;
        }
    }
    free(inMolecule);
    return 0.5 * energy;
}

#undef M_SQRT
#undef S_FMA
#undef FN
#undef CAT
#undef CAT_
