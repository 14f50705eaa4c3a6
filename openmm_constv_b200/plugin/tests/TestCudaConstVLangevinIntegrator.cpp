// End-to-end tests of ConstVLangevinIntegrator on the CUDA platform through the plugin classes
// (needs a GPU).  Usage: TestCudaConstVLangevinIntegrator [single|mixed|double]   (as the
// reference's platforms/cuda/tests/CMakeLists.txt:22-24 runs its test three times).
//
// The properties are the ones the reference's TestCudaConstVLangevinIntegrator.cpp:52-268 asserts
// -- damped oscillator against the analytic solution, energy conservation at zero friction,
// equipartition temperature, massless particles never move, seed determinism -- but on systems that
// actually satisfy the integrator's requirements (the reference's systems have no image particles,
// so its initialize() throws), plus what the reference never tests: the image
// rewrite, the neutral-atom skip, validation messages, atom reordering, the constraint hand-off.
#include <cmath>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "CudaConstVKernelFactory.h"
#include "CudaConstVKernels.h"
#include "CudaContext.h"
#include "CudaPlatform.h"
#include "TestAssert.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/Context.h"
#include "openmm/HarmonicBondForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/State.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;
using std::vector;

extern "C" void registerConstVCudaKernelFactories();

static std::map<std::string, std::string> properties;
static std::string precision = "single";

static CudaContext& cudaContextOf(Context& context) {
    return *static_cast<CudaPlatform::PlatformData*>(context.getImpl().getPlatformData())->contexts[0];
}

// A slab of `numReal` real particles followed by their `numReal` massless images; cell size 2 nm.
static void addSlabParticles(System& system, const vector<double>& realMasses) {
    for (double m : realMasses) system.addParticle(m);
    for (size_t i = 0; i < realMasses.size(); i++) system.addParticle(0.0);
    system.setDefaultPeriodicBoxVectors(Vec3(4, 0, 0), Vec3(0, 4, 0), Vec3(0, 0, 4));
}

static float asReal(double v) { return (float) v; }

// what the kernel must write for the image of a charged atom at height z (constVLangevin.cu:151-152)
static double mirroredZ(double z, double zmax) {
    if (precision == "single") return (double) (float) (-(double) (float) z + 2.0 * zmax);
    return -z + 2.0 * zmax;
}

static void testSingleBond() {
    System system;
    addSlabParticles(system, {2.0, 2.0});
    ConstVLangevinIntegrator integrator(0, 0.1, 0.01);
    HarmonicBondForce* bond = new HarmonicBondForce();
    bond->addBond(0, 1, 1.5, 1);
    system.addForce(bond);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    vector<Vec3> positions = {Vec3(-1, 0, 1), Vec3(1, 0, 1), Vec3(-1, 0, -1), Vec3(1, 0, -1)};
    context.setPositions(positions);
    context.setCharges({1.0, -1.0, -1.0, 1.0});

    // a damped harmonic oscillator: compare with the analytic solution
    const double freq = std::sqrt(1 - 0.05 * 0.05);
    for (int i = 0; i < 1000; ++i) {
        State state = context.getState(State::Positions | State::Velocities);
        const double time = state.getTime();
        const double dist = 1.5 + 0.5 * std::exp(-0.05 * time) * std::cos(freq * time);
        const double speed = -0.5 * std::exp(-0.05 * time) * (0.05 * std::cos(freq * time) + freq * std::sin(freq * time));
        CHECK_CLOSE(-0.5 * dist, state.getPositions()[0][0], 0.02);
        CHECK_CLOSE(0.5 * dist, state.getPositions()[1][0], 0.02);
        CHECK_CLOSE(-0.5 * speed, state.getVelocities()[0][0], 0.02);
        CHECK_CLOSE(0.5 * speed, state.getVelocities()[1][0], 0.02);
        for (int p = 0; p < 2; p++) {
            CHECK_CLOSE(0.0, state.getPositions()[p][1], 1e-6);
            CHECK_CLOSE(1.0, state.getPositions()[p][2], 1e-6);
            if (i > 0) {  // the images follow their partners: same x, y; z mirrored through z = zmax
                CHECK_CLOSE(state.getPositions()[p][0], state.getPositions()[p + 2][0], 1e-7);
                CHECK_CLOSE(3.0, state.getPositions()[p + 2][2], 1e-6);
            }
            CHECK_EQ(0.0, state.getVelocities()[p + 2][0]);
        }
        integrator.step(1);
    }
    CHECK_CLOSE(10.0, context.getState(0).getTime(), 1e-9);

    // zero friction: total energy is conserved
    integrator.setFriction(0.0);
    context.setPositions(positions);
    context.setVelocities(vector<Vec3>(4));
    State state = context.getState(State::Energy);
    const double initialEnergy = state.getKineticEnergy() + state.getPotentialEnergy();
    for (int i = 0; i < 1000; ++i) {
        state = context.getState(State::Energy);
        CHECK_CLOSE(initialEnergy, state.getKineticEnergy() + state.getPotentialEnergy(), 0.01);
        integrator.step(1);
    }
}

static void testTemperature() {
    const int numReal = 8;
    const double temp = 100.0;
    System system;
    addSlabParticles(system, vector<double>(numReal, 2.0));
    ConstVLangevinIntegrator integrator(temp, 2.0, 0.01);
    integrator.setRandomNumberSeed(31);
    HarmonicWellForce* wells = new HarmonicWellForce();
    vector<Vec3> positions(2 * numReal);
    vector<double> charges(2 * numReal);
    for (int i = 0; i < numReal; ++i) {
        positions[i] = Vec3((i % 2 == 0 ? 1 : -1), (i % 4 < 2 ? 1 : -1), (i < 4 ? 1.5 : 0.5));
        positions[i + numReal] = Vec3(positions[i][0], positions[i][1], -positions[i][2]);
        charges[i] = (i % 2 == 0 ? 1.0 : -1.0);
        charges[i + numReal] = -charges[i];
        wells->addParticle(i, positions[i], 5.0);
    }
    system.addForce(wells);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    context.setPositions(positions);
    context.setCharges(charges);

    integrator.step(5000);  // equilibrate
    double ke = 0.0;
    const int samples = 10000;
    for (int i = 0; i < samples; ++i) {
        ke += context.getState(State::Energy).getKineticEnergy();
        integrator.step(1);
    }
    ke /= samples;
    const double expected = 0.5 * numReal * 3 * BOLTZ * temp;
    CHECK_CLOSE(expected, ke, 6 / std::sqrt((double) samples));
}

// Massless particles (the images, and neutral "wall" atoms among the reals) never acquire a
// velocity and never move on their own; a neutral atom's image is left where the user put it
// (constVLangevin.cu:143; example/nacl_1m_spce/nacl_constv.py:82-85 relies on this).
static void testMasslessAndNeutralParticles() {
    System system;
    addSlabParticles(system, {2.0, 0.0, 3.0});  // atom 1 is a massless, neutral wall atom
    ConstVLangevinIntegrator integrator(300.0, 2.0, 0.01);
    integrator.setRandomNumberSeed(5);
    HarmonicWellForce* wells = new HarmonicWellForce();
    wells->addParticle(0, Vec3(0.5, 0.5, 1.0), 10.0);
    wells->addParticle(2, Vec3(1.5, 1.0, 0.7), 10.0);
    system.addForce(wells);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    vector<Vec3> positions = {Vec3(0.5, 0.5, 1.0), Vec3(1.0, 2.0, 0.0), Vec3(1.5, 1.0, 0.7),
                              Vec3(0.5, 0.5, -1.0), Vec3(1.0, 2.0, 0.001), Vec3(9, 9, 9)};
    context.setPositions(positions);
    context.setCharges({0.8, 0.0, 0.0, -0.8, 0.0, -0.25});  // atom 2 is massive but neutral
    integrator.step(50);
    State state = context.getState(State::Positions | State::Velocities);
    for (int p : {1, 3, 4, 5})
        for (int k = 0; k < 3; k++) CHECK_EQ(0.0, state.getVelocities()[p][k]);
    for (int k = 0; k < 3; k++) {
        CHECK_EQ((double) asReal(positions[1][k]), (double) asReal(state.getPositions()[1][k]));  // wall atom did not move
        CHECK_EQ((double) asReal(positions[4][k]), (double) asReal(state.getPositions()[4][k]));  // its image was not rewritten
        CHECK_EQ((double) asReal(positions[5][k]), (double) asReal(state.getPositions()[5][k]));  // neutral massive atom: image untouched
    }
    CHECK(state.getPositions()[0][0] != positions[0][0]);  // the massive atoms did move
    CHECK(state.getPositions()[2][0] != positions[2][0]);
    // charged atom 0: image = (x, y, -z + 2 zmax), bit for bit
    CHECK_EQ(bitsOf(state.getPositions()[0][0]), bitsOf(state.getPositions()[3][0]));
    CHECK_EQ(bitsOf(state.getPositions()[0][1]), bitsOf(state.getPositions()[3][1]));
    if (precision != "mixed")
        CHECK_EQ(bitsOf(mirroredZ(state.getPositions()[0][2], 2.0)), bitsOf(state.getPositions()[3][2]));
    else
        CHECK_CLOSE(-state.getPositions()[0][2] + 4.0, state.getPositions()[3][2], 1e-12);
}

static void buildNoisySlab(System& system, vector<Vec3>& positions, vector<double>& charges, int numReal, bool constraints) {
    std::mt19937 gen(99);
    std::uniform_real_distribution<double> u(0.1, 1.9);
    vector<double> masses(numReal);
    for (int i = 0; i < numReal; i++) masses[i] = (i % 7 == 3) ? 0.0 : 1.0 + (i % 5);
    addSlabParticles(system, masses);
    HarmonicWellForce* wells = new HarmonicWellForce();
    positions.resize(2 * numReal);
    charges.resize(2 * numReal);
    for (int i = 0; i < numReal; i++) {
        positions[i] = Vec3(u(gen) * 2, u(gen) * 2, u(gen));
        positions[i + numReal] = Vec3(positions[i][0], positions[i][1], -positions[i][2]);
        charges[i] = masses[i] == 0 ? 0.0 : (i % 3 == 0 ? -0.8476 : 0.4238);
        charges[i + numReal] = -charges[i];
        if (masses[i] != 0) wells->addParticle(i, Vec3(positions[i][0] + 0.05, positions[i][1], positions[i][2] - 0.02), 50.0);
    }
    system.addForce(wells);
    if (constraints) {  // massive pairs only; the shim has no solver, the hand-off is what is observed
        system.addConstraint(0, 1, 1.0);
        system.addConstraint(4, 5, 1.0);
    }
}

static void runNoisy(int seed, int steps, bool constraints, const vector<int>* reorder, vector<Vec3>& pos, vector<Vec3>& vel,
                     int* constraintCalls = NULL) {
    const int numReal = 100;
    System system;
    vector<Vec3> positions;
    vector<double> charges;
    buildNoisySlab(system, positions, charges, numReal, constraints);
    ConstVLangevinIntegrator integrator(300.0, 5.0, 0.002);
    integrator.setRandomNumberSeed(seed);
    integrator.setConstraintTolerance(3e-6);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    context.setPositions(positions);
    context.setCharges(charges);
    context.setVelocitiesToTemperature(300.0, 4);
    integrator.step(steps / 2);
    if (reorder != NULL) cudaContextOf(context).setPendingReorder(*reorder);  // applied by cu.reorderAtoms() at the end of the next step
    integrator.step(steps - steps / 2);
    State state = context.getState(State::Positions | State::Velocities);
    pos = state.getPositions();
    vel = state.getVelocities();
    if (constraintCalls != NULL) {
        *constraintCalls = cudaContextOf(context).getIntegrationUtilities().constraintCalls;
        CHECK_EQ(3e-6, cudaContextOf(context).getIntegrationUtilities().lastConstraintTol);
    }
}

static bool identical(const vector<Vec3>& a, const vector<Vec3>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
        for (int k = 0; k < 3; k++)
            if (bitsOf(a[i][k]) != bitsOf(b[i][k])) return false;
    return true;
}

// Same seed -> the same trajectory (bit for bit: the Philox stream is a pure function of seed, atom
// and step); another seed -> a different one (TestCudaConstVLangevinIntegrator.cpp:212-268).
static void testRandomSeed() {
    vector<Vec3> p1, v1, p2, v2, p3, v3;
    runNoisy(5, 20, false, NULL, p1, v1);
    runNoisy(5, 20, false, NULL, p2, v2);
    runNoisy(10, 20, false, NULL, p3, v3);
    CHECK(identical(p1, p2));
    CHECK(identical(v1, v2));
    CHECK(!identical(p1, p3));
    CHECK(!identical(v1, v3));
}

// OpenMM's periodic re-sort of the atoms must not change the trajectory: after cu.reorderAtoms()
// the kernels address slots through atomIndex / invAtomOrder (CudaConstVKernels.cpp:139-142,
// constVLangevin.cu:142-158) and the noise follows the ORIGINAL atom index.
static void testReorderedAtoms() {
    const int n = 200;
    vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = (i * 77 + 13) % n;  // 77 is coprime with 200: a permutation mixing reals and images
    vector<Vec3> p1, v1, p2, v2;
    runNoisy(8, 30, false, NULL, p1, v1);
    runNoisy(8, 30, false, &order, p2, v2);
    CHECK(identical(p1, p2));
    CHECK(identical(v1, v2));
}

// A re-sort that happens BETWEEN two steps (OpenMM re-sorts inside setPositions, loadCheckpoint and the
// minimizer) leaves cu.getAtomsWereReordered() false by the time execute() runs, because every later
// reorderAtoms() call that does not re-sort clears the flag.  The kernel must still notice (it listens for
// re-sorts instead of polling the flag as the reference does, CudaConstVKernels.cpp:139-142).
static void testReorderBetweenSteps() {
    const int numReal = 100, n = 2 * numReal;
    vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = (i * 77 + 13) % n;
    vector<Vec3> ref, refVel;
    runNoisy(21, 12, false, NULL, ref, refVel);

    System system;
    vector<Vec3> positions;
    vector<double> charges;
    buildNoisySlab(system, positions, charges, numReal, false);
    ConstVLangevinIntegrator integrator(300.0, 5.0, 0.002);
    integrator.setRandomNumberSeed(21);
    integrator.setConstraintTolerance(3e-6);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    context.setPositions(positions);
    context.setCharges(charges);
    context.setVelocitiesToTemperature(300.0, 4);
    CudaContext& cu = cudaContextOf(context);
    integrator.step(6);
    cu.setPendingReorder(order);
    cu.reorderAtoms();  // the re-sort, outside any step
    cu.reorderAtoms();  // a later call that re-sorts nothing: the flag is false again
    CHECK(!cu.getAtomsWereReordered());
    integrator.step(6);
    State state = context.getState(State::Positions | State::Velocities);
    CHECK(identical(ref, state.getPositions()));
    CHECK(identical(refVel, state.getVelocities()));
}

// The noise of step k is a function of (seed, atom, k) with k = OpenMM's own step count: a state carried into
// a NEW Context together with its step count (what loadCheckpoint does) continues the trajectory bit for bit
// instead of replaying the noise of steps 0, 1, ...
static void testStepCountDrivesTheNoise() {
    const int numReal = 100;
    vector<Vec3> ref, refVel;
    runNoisy(33, 10, false, NULL, ref, refVel);  // 5 + 5 steps in one Context

    System system;
    vector<Vec3> positions;
    vector<double> charges;
    buildNoisySlab(system, positions, charges, numReal, false);
    Platform& platform = Platform::getPlatformByName("CUDA");
    vector<Vec3> midPos, midVel;
    {
        ConstVLangevinIntegrator integrator(300.0, 5.0, 0.002);
        integrator.setRandomNumberSeed(33);
        Context context(system, integrator, platform, properties);
        context.setPositions(positions);
        context.setCharges(charges);
        context.setVelocitiesToTemperature(300.0, 4);
        integrator.step(4);
        State state = context.getState(State::Positions | State::Velocities);
        midPos = state.getPositions();
        midVel = state.getVelocities();
    }
    ConstVLangevinIntegrator integrator(300.0, 5.0, 0.002);
    integrator.setRandomNumberSeed(33);
    Context context(system, integrator, platform, properties);
    context.setPositions(midPos);
    context.setCharges(charges);
    context.setVelocities(midVel);
    cudaContextOf(context).setStepCount(4);
    cudaContextOf(context).setTime(4 * 0.002);
    integrator.step(6);
    State state = context.getState(State::Positions | State::Velocities);
    CHECK(identical(ref, state.getPositions()));
    CHECK(identical(refVel, state.getVelocities()));
    CHECK_EQ(10, cudaContextOf(context).getStepCount());
}

// A charged virtual site (the M site of a 4-site water): the reference places the sites after part 2 and
// rewrites the images after that (part 2, computeVirtualSites, updateImageParticlePositions:
// CudaConstVKernels.cpp:158-166), so the site's image mirrors the position the site has at the END of the step.
static void testVirtualSiteImage() {
    System system;
    addSlabParticles(system, {16.0, 1.0, 0.0});  // atom 2: massless site halfway between atoms 0 and 1
    system.setVirtualSite(2, new TwoParticleAverageSite(0, 1, 0.5, 0.5));
    ConstVLangevinIntegrator integrator(300.0, 2.0, 0.004);
    integrator.setRandomNumberSeed(9);
    HarmonicWellForce* wells = new HarmonicWellForce();
    wells->addParticle(0, Vec3(1.0, 1.0, 0.8), 20.0);
    wells->addParticle(1, Vec3(1.2, 1.1, 1.1), 20.0);
    system.addForce(wells);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    vector<Vec3> positions = {Vec3(1.0, 1.0, 0.8), Vec3(1.2, 1.1, 1.1), Vec3(1.1, 1.05, 0.95),
                              Vec3(1.0, 1.0, -0.8), Vec3(1.2, 1.1, -1.1), Vec3(0, 0, 0)};
    context.setPositions(positions);
    context.setCharges({0.0, 0.52, -1.04, 0.0, -0.52, 1.04});
    for (int k = 0; k < 25; k++) {
        integrator.step(1);
        State state = context.getState(State::Positions);
        const vector<Vec3>& p = state.getPositions();
        const double tol = precision == "single" ? 1e-6 : 1e-12;
        for (int c = 0; c < 3; c++)  // the shim placed the site from this step's positions
            CHECK_CLOSE(0.5 * (p[0][c] + p[1][c]), p[2][c], tol);
        // ... and the image of the site is the mirror of THAT position, bit for bit, not of last step's
        CHECK_EQ(bitsOf(p[2][0]), bitsOf(p[5][0]));
        CHECK_EQ(bitsOf(p[2][1]), bitsOf(p[5][1]));
        if (precision != "mixed")
            CHECK_EQ(bitsOf(mirroredZ(p[2][2], 2.0)), bitsOf(p[5][2]));
        else
            CHECK_CLOSE(-p[2][2] + 4.0, p[5][2], 1e-12);
        CHECK(p[2][2] != positions[2][2] || k == 0);
    }
    CHECK_EQ(25, cudaContextOf(context).getIntegrationUtilities().virtualSiteCalls);
}

// With constraints in the System the step is split around integration.applyConstraints(tol)
// (CudaConstVKernels.cpp:146-166).  The shim's solver is a no-op, so the split path must reproduce
// the fused path exactly, and the solver must have been handed posDelta once per step.
static void testConstraintHandOff() {
    vector<Vec3> p1, v1, p2, v2;
    int calls = 0;
    runNoisy(3, 24, false, NULL, p1, v1);
    runNoisy(3, 24, true, NULL, p2, v2, &calls);
    CHECK_EQ(24, calls);
    CHECK(identical(p1, p2));
    CHECK(identical(v1, v2));
}

// getState(Energy) goes through computeKineticEnergy (CudaConstVKernels.cpp:175-177): the
// half-step-shifted kinetic energy of the massive atoms.
// Rigid 3-site water (the reference's example: constraints=AllBonds on SPC/E): every constraint is part of a
// triangle, so the kernel solves them inside the step (SETTLE) and nothing is handed to applyConstraints --
// the shim has no solver at all, yet the molecules stay rigid.
static void testRigidWater() {
    const int numMolecules = 150;
    const double dOH = 0.1, dHH = 0.16330;
    System system;
    vector<double> masses;
    for (int k = 0; k < numMolecules; k++) {
        masses.push_back(15.9994);
        masses.push_back(1.008);
        masses.push_back(1.008);
    }
    addSlabParticles(system, masses);
    const int numReal = 3 * numMolecules;
    HarmonicWellForce* wells = new HarmonicWellForce();
    vector<Vec3> positions(2 * numReal);
    vector<double> charges(2 * numReal);
    const double c = std::sqrt(dOH * dOH - 0.25 * dHH * dHH), h = 0.5 * dHH;
    for (int k = 0; k < numMolecules; k++) {
        const Vec3 o(0.35 * (k % 10), 0.35 * ((k / 10) % 10), 0.4 + 1.2 * ((k * 7) % 15) / 15.0);
        // three different orientations of the same rigid triangle
        const Vec3 u = k % 3 == 0 ? Vec3(1, 0, 0) : (k % 3 == 1 ? Vec3(0, 1, 0) : Vec3(0, 0, 1));
        const Vec3 w = k % 3 == 0 ? Vec3(0, 1, 0) : (k % 3 == 1 ? Vec3(0, 0, 1) : Vec3(1, 0, 0));
        positions[3 * k] = o;
        positions[3 * k + 1] = o + u * c + w * h;
        positions[3 * k + 2] = o + u * c - w * h;
        charges[3 * k] = -0.8476;
        charges[3 * k + 1] = charges[3 * k + 2] = 0.4238;
        wells->addParticle(3 * k, o, 300.0);
        system.addConstraint(3 * k, 3 * k + 1, dOH);
        system.addConstraint(3 * k, 3 * k + 2, dOH);
        system.addConstraint(3 * k + 1, 3 * k + 2, dHH);
    }
    for (int i = 0; i < numReal; i++) {
        positions[i + numReal] = Vec3(positions[i][0], positions[i][1], -positions[i][2]);
        charges[i + numReal] = -charges[i];
    }
    system.addForce(wells);
    const double temp = 300.0;
    ConstVLangevinIntegrator integrator(temp, 5.0, 0.002);
    integrator.setRandomNumberSeed(11);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    context.setPositions(positions);
    context.setCharges(charges);
    Kernel kernel = platform.createKernel(IntegrateConstVLangevinStepKernel::Name(), context.getImpl());
    CudaContext& cu = cudaContextOf(context);

    integrator.step(500);
    // OpenMM re-sorts identical molecules now and then: real molecule k trades slots with real molecule
    // numMolecules-1-k, the image atoms are reversed one by one.  From the next step on the kernel addresses
    // the images through atomIndex / invAtomOrder and the molecules stay rigid all the same.
    vector<int> order(2 * numReal);
    for (int k = 0; k < numMolecules; k++)
        for (int a = 0; a < 3; a++) order[3 * k + a] = 3 * (numMolecules - 1 - k) + a;
    for (int i = 0; i < numReal; i++) order[numReal + i] = 2 * numReal - 1 - i;
    cu.setPendingReorder(order);
    integrator.step(500);
    CHECK(cu.getAtomIndex()[0] == 3 * (numMolecules - 1));
    double ke = 0;
    const int samples = 400;
    for (int n = 0; n < samples; n++) {
        integrator.step(5);
        const State state = context.getState(State::Positions | State::Velocities);
        const vector<Vec3>& p = state.getPositions();
        const vector<Vec3>& v = state.getVelocities();
        const double tol = precision == "single" ? 2e-5 : 2e-7;
        for (int k = 0; k < numMolecules; k++) {
            const Vec3 a = p[3 * k + 1] - p[3 * k], b = p[3 * k + 2] - p[3 * k], bc = p[3 * k + 2] - p[3 * k + 1];
            CHECK_CLOSE(dOH, std::sqrt(a.dot(a)), tol);
            CHECK_CLOSE(dOH, std::sqrt(b.dot(b)), tol);
            CHECK_CLOSE(dHH, std::sqrt(bc.dot(bc)), tol);
        }
        for (int i = 0; i < numReal; i++) ke += 0.5 * masses[i] * v[i].dot(v[i]);
    }
    ke /= samples;
    // a rigid triangle has 6 degrees of freedom; v = delta/dt lies on the constraint surface
    const double expected = 0.5 * 6 * numMolecules * BOLTZ * temp;
    CHECK_CLOSE(expected, ke, 0.05);
    CHECK_EQ(0, cu.getIntegrationUtilities().constraintCalls);  // nothing was handed over
    // every (charged) atom's image followed it
    const vector<Vec3> p = context.getState(State::Positions).getPositions();
    for (int i = 0; i < numReal; i++)
        CHECK_CLOSE(mirroredZ(p[i][2], 2.0), p[i + numReal][2], precision == "single" ? 1e-6 : 1e-12);
}

static void testKineticEnergy() {
    System system;
    vector<Vec3> positions;
    vector<double> charges;
    buildNoisySlab(system, positions, charges, 100, false);
    ConstVLangevinIntegrator integrator(300.0, 1.0, 0.002);
    integrator.setRandomNumberSeed(1);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(system, integrator, platform, properties);
    context.setPositions(positions);
    context.setCharges(charges);
    context.setVelocitiesToTemperature(300.0, 12);
    integrator.step(10);
    const double ke = context.getState(State::Energy).getKineticEnergy();
    const double expected = cudaContextOf(context).getIntegrationUtilities().computeKineticEnergy(0.5 * integrator.getStepSize());
    CHECK(expected > 0);
    CHECK_CLOSE(expected, ke, precision == "single" ? 1e-6 : 1e-12);
}

static void testValidation() {
    Platform& platform = Platform::getPlatformByName("CUDA");
    {
        System system;
        addSlabParticles(system, {1.0, 1.0, 1.0});
        ConstVLangevinIntegrator integrator(300, 1, 0.001, 3);
        CHECK_THROWS(Context(system, integrator, platform, properties), "Not even number of cells");
    }
    {
        System system;
        addSlabParticles(system, {1.0, 1.0, 1.0});
        ConstVLangevinIntegrator integrator(300, 1, 0.001, 4);
        CHECK_THROWS(Context(system, integrator, platform, properties), "Number of particles is not multiple of number of cells");
    }
    {
        System system;
        addSlabParticles(system, {1.0, 1.0});
        system.setParticleMass(3, 12.0);
        ConstVLangevinIntegrator integrator(300, 1, 0.001);
        CHECK_THROWS(Context(system, integrator, platform, properties), "Image Particle has nonzero mass");
    }
    {
        System system;
        addSlabParticles(system, {1.0, 1.0});
        ConstVLangevinIntegrator integrator(300, 1, 0.001, 2, 1.9);  // box z = 4 != 2 * 1.9
        CHECK_THROWS(Context(system, integrator, platform, properties), "Unit cell dimension does not match with the provided zmax value");
    }
    {
        System system;
        addSlabParticles(system, {1.0, 1.0});
        ConstVLangevinIntegrator integrator(300, 1, 0.001, 2, 2.0);  // explicit and consistent
        Context context(system, integrator, platform, properties);
        CHECK_THROWS(Context(system, integrator, platform, properties), "This Integrator is already bound to a context");
        // unknown kernel names are refused by the factory with the reference's text
        CudaConstVKernelFactory factory;
        CHECK_THROWS(factory.createKernelImpl("NoSuchKernel", platform, context.getImpl()),
                     "Tried to create kernel with illegal kernel name 'NoSuchKernel'");
        // the kernel resolved the cell size from the integrator
        Kernel kernel = platform.createKernel(IntegrateConstVLangevinStepKernel::Name(), context.getImpl());
        CHECK(dynamic_cast<CudaIntegrateConstVLangevinStepKernel*>(&kernel.getImpl()) != NULL);
    }
}

int main(int argc, char* argv[]) {
    try {
        registerConstVCudaKernelFactories();
        if (argc > 1) precision = argv[1];
        properties[CudaPlatform::CudaPrecision()] = precision;
        testSingleBond();
        testTemperature();
        testMasslessAndNeutralParticles();
        testRandomSeed();
        testReorderedAtoms();
        testReorderBetweenSteps();
        testStepCountDrivesTheNoise();
        testVirtualSiteImage();
        testConstraintHandOff();
        testRigidWater();
        testKineticEnergy();
        testValidation();
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "Done (" << precision << ")" << std::endl;
    return 0;
}
