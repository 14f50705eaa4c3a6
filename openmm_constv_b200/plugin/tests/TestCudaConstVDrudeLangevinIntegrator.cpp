// End-to-end tests of ConstVDrudeLangevinIntegrator on the CUDA platform through the plugin classes
// (needs a GPU).  Usage: TestCudaConstVDrudeLangevinIntegrator [single|mixed|double].
//
// The reference ships no test for this integrator.  The properties checked here are the ones the
// algorithm (platforms/cuda/src/kernels/constVDrudeLangevin.cu) promises: the dual thermostat brings
// each pair's centre of mass to `temperature` and its internal motion to `drudeTemperature`
// (:31-63), the hard wall bounds the Drude displacement (:108-211), images follow their real partners
// (constVLangevin.cu:138-164), equal seeds reproduce the trajectory, and the platform-independent
// class raises the reference's errors (openmmapi/src/ConstVDrudeLangevinIntegrator.cpp:61-80).
#include <cmath>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "CudaConstVKernelFactory.h"
#include "CudaConstVKernels.h"
#include "CudaContext.h"
#include "CudaPlatform.h"
#include "TestAssert.h"
#include "openmm/ConstVDrudeLangevinIntegrator.h"
#include "openmm/ConstVKernels.h"
#include "openmm/Context.h"
#include "openmm/DrudeForce.h"
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

namespace {

const double kCellSize = 2.0;

class ExposedDrudeIntegrator : public ConstVDrudeLangevinIntegrator {
public:
    using ConstVDrudeLangevinIntegrator::ConstVDrudeLangevinIntegrator;
    vector<std::string> kernelNames() { return getKernelNames(); }
};

// `numPairs` (Drude particle, parent) pairs -- slots 2k (Drude, mass mDrude) and 2k+1 (parent) -- then
// `numPlain` ordinary particles, then the massless images of all of them.
struct DrudeSlab {
    System system;
    DrudeForce* drude;
    vector<Vec3> positions;
    vector<double> charges;
    int numReal;
};

void buildSlab(DrudeSlab& s, int numPairs, int numPlain, double mDrude, double mParent, double polarizability) {
    s.numReal = 2 * numPairs + numPlain;
    for (int k = 0; k < numPairs; k++) {
        s.system.addParticle(mDrude);
        s.system.addParticle(mParent);
    }
    for (int k = 0; k < numPlain; k++) s.system.addParticle(12.0);
    for (int k = 0; k < s.numReal; k++) s.system.addParticle(0.0);
    s.system.setDefaultPeriodicBoxVectors(Vec3(4, 0, 0), Vec3(0, 4, 0), Vec3(0, 0, 2 * kCellSize));
    s.drude = new DrudeForce();
    HarmonicWellForce* wells = new HarmonicWellForce();  // keeps every molecule near its site
    s.positions.assign(2 * s.numReal, Vec3());
    s.charges.assign(2 * s.numReal, 0.0);
    for (int i = 0; i < s.numReal; i++) {
        const int site = i < 2 * numPairs ? i / 2 : i;
        Vec3 p(0.37 * (site % 10), 0.41 * ((site / 10) % 10), 0.3 + 1.4 * ((site * 7) % 13) / 13.0);
        s.positions[i] = p;
        const bool isDrude = i < 2 * numPairs && i % 2 == 0;
        const bool isParent = i < 2 * numPairs && i % 2 == 1;
        s.charges[i] = isDrude ? -1.5 : (isParent ? 1.5 : (i % 2 ? 1.0 : -1.0));
        if (isDrude) s.drude->addParticle(i, i + 1, -1, -1, -1, -1.5, polarizability, 1, 1);
        if (!isDrude) wells->addParticle(i, p, 400.0);
    }
    for (int i = 0; i < s.numReal; i++) {
        s.positions[i + s.numReal] = Vec3(s.positions[i][0], s.positions[i][1], -s.positions[i][2]);
        s.charges[i + s.numReal] = -s.charges[i];
    }
    s.system.addForce(s.drude);
    s.system.addForce(wells);
}

double mirroredZ(double z) {
    if (precision == "single") return (double) (float) (-(double) (float) z + 2.0 * kCellSize);
    return -z + 2.0 * kCellSize;
}

void testDefaultsAndErrors() {
    ExposedDrudeIntegrator integ(300.0, 5.0, 1.0, 20.0, 0.001);
    CHECK_EQ(300.0, integ.getTemperature());
    CHECK_EQ(5.0, integ.getFriction());
    CHECK_EQ(1.0, integ.getDrudeTemperature());
    CHECK_EQ(20.0, integ.getDrudeFriction());
    CHECK_EQ(0.001, integ.getStepSize());
    CHECK_EQ(0.0, integ.getMaxDrudeDistance());
    CHECK_EQ(2, integ.getNumCells());
    CHECK_EQ(-1.0, integ.getCellSize());
    CHECK_EQ(1e-5, integ.getConstraintTolerance());
    CHECK_EQ(0, integ.getRandomNumberSeed());
    integ.setMaxDrudeDistance(0.02);
    CHECK_EQ(0.02, integ.getMaxDrudeDistance());
    CHECK_THROWS(integ.setMaxDrudeDistance(-0.1), "setMaxDrudeDistance: Distance cannot be negative");
    CHECK_EQ((size_t) 1, integ.kernelNames().size());
    CHECK_EQ(std::string("IntegrateConstVDrudeLangevinStep"), integ.kernelNames()[0]);
    CHECK_THROWS(integ.step(1), "This Integrator is not bound to a context!");

    Platform& platform = Platform::getPlatformByName("CUDA");
    {
        System system;
        for (int i = 0; i < 4; i++) system.addParticle(i < 2 ? 1.0 : 0.0);
        system.setDefaultPeriodicBoxVectors(Vec3(4, 0, 0), Vec3(0, 4, 0), Vec3(0, 0, 4));
        ConstVDrudeLangevinIntegrator a(300, 5, 1, 20, 0.001);
        CHECK_THROWS(Context(system, a, platform, properties), "The System does not contain a DrudeForce");
        system.addForce(new DrudeForce());
        system.addForce(new DrudeForce());
        ConstVDrudeLangevinIntegrator b(300, 5, 1, 20, 0.001);
        CHECK_THROWS(Context(system, b, platform, properties), "The System contains multiple DrudeForces");
    }
    {
        DrudeSlab s;
        buildSlab(s, 2, 1, 0.4, 15.6, 0.001);
        s.system.setParticleMass(s.numReal + 1, 3.0);
        ConstVDrudeLangevinIntegrator integrator(300, 5, 1, 20, 0.001);
        CHECK_THROWS(Context(s.system, integrator, platform, properties), "Image Particle has nonzero mass");
    }
    {
        DrudeSlab s;
        buildSlab(s, 2, 1, 0.4, 15.6, 0.001);
        ConstVDrudeLangevinIntegrator integrator(300, 5, 1, 20, 0.001, 2, 1.7);
        CHECK_THROWS(Context(s.system, integrator, platform, properties), "Unit cell dimension does not match with the provided zmax value");
    }
    {
        DrudeSlab s;
        buildSlab(s, 2, 1, 0.4, 15.6, 0.001);
        ConstVDrudeLangevinIntegrator integrator(300, 5, 1, 20, 0.001);
        Context context(s.system, integrator, platform, properties);
        CHECK_THROWS(Context(s.system, integrator, platform, properties), "This Integrator is already bound to a context");
        Kernel kernel = platform.createKernel(IntegrateConstVDrudeLangevinStepKernel::Name(), context.getImpl());
        CHECK(dynamic_cast<CudaIntegrateConstVDrudeLangevinStepKernel*>(&kernel.getImpl()) != NULL);
    }
}

// Pair centres of mass equilibrate at `temperature`, the internal (relative) motion at `drudeTemperature`.
void testDualThermostat() {
    const int numPairs = 400, numPlain = 100;
    const double temp = 300.0, drudeTemp = 10.0, mDrude = 0.4, mParent = 15.6;
    DrudeSlab s;
    buildSlab(s, numPairs, numPlain, mDrude, mParent, 0.001);  // k = 138.9*2.25/0.001: period ~ 7 fs
    ConstVDrudeLangevinIntegrator integrator(temp, 20.0, drudeTemp, 20.0, 0.0005);
    integrator.setRandomNumberSeed(17);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(s.system, integrator, platform, properties);
    context.setPositions(s.positions);
    context.setCharges(s.charges);
    integrator.step(2000);  // 20 relaxation times
    double keCm = 0, keRel = 0, kePlain = 0;
    const int samples = 200;
    const double mu = mDrude * mParent / (mDrude + mParent), total = mDrude + mParent;
    for (int n = 0; n < samples; n++) {
        integrator.step(10);
        const vector<Vec3> v = context.getState(State::Velocities).getVelocities();
        for (int k = 0; k < numPairs; k++) {
            const Vec3 vcm = (v[2 * k] * mDrude + v[2 * k + 1] * mParent) * (1.0 / total);
            const Vec3 vrel = v[2 * k + 1] - v[2 * k];
            keCm += 0.5 * total * vcm.dot(vcm);
            keRel += 0.5 * mu * vrel.dot(vrel);
        }
        for (int k = 2 * numPairs; k < s.numReal; k++) kePlain += 0.5 * 12.0 * v[k].dot(v[k]);
        for (int k = s.numReal; k < 2 * s.numReal; k++) CHECK(v[k].dot(v[k]) == 0.0);  // images never move on their own
    }
    keCm /= samples * numPairs;
    keRel /= samples * numPairs;
    kePlain /= samples * numPlain;
    CHECK_CLOSE(1.5 * BOLTZ * temp, keCm, 0.05 * 1.5 * BOLTZ * temp);
    CHECK_CLOSE(1.5 * BOLTZ * drudeTemp, keRel, 0.05);  // CHECK_CLOSE scales by max(1, |expected|): 0.125 +- 0.05 absolute
    CHECK(std::fabs(keRel / (1.5 * BOLTZ * drudeTemp) - 1.0) < 0.08);
    CHECK_CLOSE(1.5 * BOLTZ * temp, kePlain, 0.08 * 1.5 * BOLTZ * temp);
}

// With a hot Drude bath and a soft spring the displacement would exceed the wall; it never does.
void testHardWallAndImages() {
    const int numPairs = 200;
    const double wall = 0.02;
    DrudeSlab s;
    buildSlab(s, numPairs, 10, 0.4, 15.6, 0.02);  // soft spring
    ConstVDrudeLangevinIntegrator integrator(300.0, 10.0, 300.0, 10.0, 0.001);
    integrator.setMaxDrudeDistance(wall);
    integrator.setRandomNumberSeed(5);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(s.system, integrator, platform, properties);
    context.setPositions(s.positions);
    context.setCharges(s.charges);
    double largest = 0;
    int nearWall = 0;
    for (int n = 0; n < 300; n++) {
        integrator.step(1);
        const vector<Vec3> p = context.getState(State::Positions).getPositions();
        for (int k = 0; k < numPairs; k++) {
            const Vec3 d = p[2 * k] - p[2 * k + 1];
            const double r = std::sqrt(d.dot(d));
            largest = std::max(largest, r);
            if (r > 0.9 * wall) nearWall++;
        }
        if (n % 50 == 49)
            for (int i = 0; i < s.numReal; i++) {  // every (charged) atom's image is its mirror in the electrode plane
                CHECK_EQ(p[i][0], p[i + s.numReal][0]);
                CHECK_EQ(p[i][1], p[i + s.numReal][1]);
                CHECK_CLOSE(mirroredZ(p[i][2]), p[i + s.numReal][2], precision == "single" ? 1e-6 : 1e-12);
            }
    }
    CHECK(largest <= wall * (1 + 1e-3));
    CHECK(nearWall > 100);  // the wall was really reached

    // without the wall the same system strays beyond it
    DrudeSlab t;
    buildSlab(t, numPairs, 10, 0.4, 15.6, 0.02);
    ConstVDrudeLangevinIntegrator free(300.0, 10.0, 300.0, 10.0, 0.001);
    free.setRandomNumberSeed(5);
    Context other(t.system, free, platform, properties);
    other.setPositions(t.positions);
    other.setCharges(t.charges);
    free.step(300);
    const vector<Vec3> p = other.getState(State::Positions).getPositions();
    double beyond = 0;
    for (int k = 0; k < numPairs; k++) {
        const Vec3 d = p[2 * k] - p[2 * k + 1];
        beyond = std::max(beyond, std::sqrt(d.dot(d)));
    }
    CHECK(beyond > wall);
}

void testRandomSeed() {
    Platform& platform = Platform::getPlatformByName("CUDA");
    vector<vector<Vec3> > results;
    const int seeds[] = {5, 5, 6};
    for (int seed : seeds) {
        DrudeSlab s;
        buildSlab(s, 20, 5, 0.4, 15.6, 0.001);
        ConstVDrudeLangevinIntegrator integrator(300.0, 5.0, 1.0, 20.0, 0.001);
        integrator.setRandomNumberSeed(seed);
        integrator.setMaxDrudeDistance(0.02);
        Context context(s.system, integrator, platform, properties);
        context.setPositions(s.positions);
        context.setCharges(s.charges);
        integrator.step(25);
        results.push_back(context.getState(State::Positions).getPositions());
        CHECK_CLOSE(25 * 0.001, context.getState(State::Positions).getTime(), 1e-12);
    }
    bool differs = false;
    for (size_t i = 0; i < results[0].size(); i++)
        for (int k = 0; k < 3; k++) {
            CHECK_EQ(bitsOf(results[0][i][k]), bitsOf(results[1][i][k]));
            differs = differs || results[0][i][k] != results[2][i][k];
        }
    CHECK(differs);
}

// A System with constraints takes the split path: part 1 | applyConstraints | part 2 + hard wall + images.
void testConstraintHandOff() {
    DrudeSlab s;
    buildSlab(s, 10, 4, 0.4, 15.6, 0.001);
    s.system.addConstraint(20, 21, 0.1);
    ConstVDrudeLangevinIntegrator integrator(300.0, 5.0, 1.0, 20.0, 0.001);
    integrator.setConstraintTolerance(3e-6);
    Platform& platform = Platform::getPlatformByName("CUDA");
    Context context(s.system, integrator, platform, properties);
    context.setPositions(s.positions);
    context.setCharges(s.charges);
    integrator.step(7);
    CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getImpl().getPlatformData())->contexts[0];
    CHECK_EQ(7, cu.getIntegrationUtilities().constraintCalls);
    CHECK_EQ(3e-6, cu.getIntegrationUtilities().lastConstraintTol);
    CHECK_EQ(7, cu.getStepCount());
    const double ke = context.getState(State::Energy).getKineticEnergy();
    CHECK(ke > 0 && std::isfinite(ke));
}

}  // namespace

int main(int argc, char* argv[]) {
    try {
        registerConstVCudaKernelFactories();
        if (argc > 1) precision = argv[1];
        properties[CudaPlatform::CudaPrecision()] = precision;
        testDefaultsAndErrors();
        testDualThermostat();
        testHardWallAndImages();
        testRandomSeed();
        testConstraintHandOff();
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "Done (" << precision << ")" << std::endl;
    return 0;
}
