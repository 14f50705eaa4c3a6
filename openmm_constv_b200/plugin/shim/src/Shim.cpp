// Implementation of the OpenMM shim (see openmm/ShimCore.h): just enough System / Context /
// Platform / CudaContext / serialization machinery to compile, load and exercise the ConstV plugin's
// host C++ without OpenMM.  Test scaffolding only.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <sstream>
#include <type_traits>

#include "CudaContext.h"
#include "openmm/serialization/XmlSerializer.h"

using namespace OpenMM;
using std::string;
using std::vector;

#define CUCHECK(x)                                                                                  \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) throw OpenMMException(string(#x) + ": " + cudaGetErrorString(e_));   \
    } while (0)

// ---- forces ---------------------------------------------------------------------------------------

int HarmonicBondForce::addBond(int p1, int p2, double length, double k) {
    bonds.push_back({p1, p2, length, k});
    return (int) bonds.size() - 1;
}

double HarmonicBondForce::compute(const vector<Vec3>& pos, vector<Vec3>& f) const {
    double energy = 0;
    for (const Bond& b : bonds) {
        Vec3 d = pos[b.p2] - pos[b.p1];
        double r = std::sqrt(d.dot(d));
        double dr = r - b.length;
        energy += 0.5 * b.k * dr * dr;
        Vec3 g = d * (b.k * dr / r);
        f[b.p1] = f[b.p1] + g;
        f[b.p2] = f[b.p2] - g;
    }
    return energy;
}

int HarmonicWellForce::addParticle(int particle, const Vec3& centre, double k) {
    wells.push_back({particle, centre, k});
    return (int) wells.size() - 1;
}

double HarmonicWellForce::compute(const vector<Vec3>& pos, vector<Vec3>& f) const {
    double energy = 0;
    for (const Well& w : wells) {
        Vec3 d = pos[w.p] - w.c;
        energy += 0.5 * w.k * d.dot(d);
        f[w.p] = f[w.p] - d * w.k;
    }
    return energy;
}

int DrudeForce::addParticle(int particle, int particle1, int particle2, int particle3, int particle4, double charge,
                            double polarizability, double aniso12, double aniso34) {
    particles.push_back({particle, particle1, particle2, particle3, particle4, charge, polarizability, aniso12, aniso34});
    return (int) particles.size() - 1;
}

void DrudeForce::getParticleParameters(int index, int& particle, int& particle1, int& particle2, int& particle3, int& particle4,
                                       double& charge, double& polarizability, double& aniso12, double& aniso34) const {
    const Entry& e = particles.at(index);
    particle = e.p; particle1 = e.p1; particle2 = e.p2; particle3 = e.p3; particle4 = e.p4;
    charge = e.charge; polarizability = e.polarizability; aniso12 = e.aniso12; aniso34 = e.aniso34;
}

double DrudeForce::compute(const vector<Vec3>& pos, vector<Vec3>& f) const {
    const double ONE_4PI_EPS0 = 138.935456;
    double energy = 0;
    for (const Entry& e : particles) {
        if (!(e.polarizability > 0)) continue;
        const double k = ONE_4PI_EPS0 * e.charge * e.charge / e.polarizability;
        Vec3 d = pos[e.p] - pos[e.p1];
        energy += 0.5 * k * d.dot(d);
        f[e.p] = f[e.p] - d * k;
        f[e.p1] = f[e.p1] + d * k;
    }
    return energy;
}

// ---- System ---------------------------------------------------------------------------------------

System::System() {
    box[0] = Vec3(2, 0, 0);
    box[1] = Vec3(0, 2, 0);
    box[2] = Vec3(0, 0, 2);
}

System::~System() {
    for (Force* f : forces) delete f;
    for (VirtualSite* v : virtualSites) delete v;
}

void System::setVirtualSite(int index, VirtualSite* virtualSite) {
    if ((int) virtualSites.size() <= index) virtualSites.resize(index + 1, NULL);
    delete virtualSites[index];
    virtualSites[index] = virtualSite;
}

const VirtualSite& System::getVirtualSite(int index) const {
    if (!isVirtualSite(index)) throw OpenMMException("This particle is not a virtual site");
    return *virtualSites[index];
}

int System::addConstraint(int p1, int p2, double distance) {
    constraints.push_back({p1, p2, distance});
    return (int) constraints.size() - 1;
}

void System::getConstraintParameters(int index, int& p1, int& p2, double& distance) const {
    const Constraint& c = constraints.at(index);
    p1 = c.p1;
    p2 = c.p2;
    distance = c.distance;
}

// ---- Platform registry ----------------------------------------------------------------------------

static vector<Platform*>& platforms() {
    static vector<Platform*> list;
    return list;
}

void Platform::registerPlatform(Platform* platform) { platforms().push_back(platform); }
int Platform::getNumPlatforms() { return (int) platforms().size(); }

Platform& Platform::getPlatformByName(const string& name) {
    for (Platform* p : platforms())
        if (p->getName() == name) return *p;
    throw OpenMMException("There is no registered Platform called \"" + name + "\"");
}

Kernel Platform::createKernel(const string& name, ContextImpl& context) const {
    auto it = kernelFactories.find(name);
    if (it == kernelFactories.end())
        throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
    return Kernel(it->second->createKernelImpl(name, *this, context));
}

const string& Platform::getPropertyDefaultValue(const string& property) const {
    static const string empty;
    auto it = defaults.find(property);
    return it == defaults.end() ? empty : it->second;
}

// ---- CudaArray ------------------------------------------------------------------------------------

CudaArray::CudaArray(CudaContext& context, int size, int elementSize, const string& name)
    : pointer(0), size(size), elementSize(elementSize), name(name) {
    void* p = NULL;
    CUCHECK(cudaMalloc(&p, (size_t) size * elementSize));
    CUCHECK(cudaMemset(p, 0, (size_t) size * elementSize));
    pointer = (CUdeviceptr) p;
}

CudaArray::~CudaArray() { cudaFree((void*) pointer); }
void CudaArray::uploadBytes(const void* data) { CUCHECK(cudaMemcpy((void*) pointer, data, (size_t) size * elementSize, cudaMemcpyHostToDevice)); }
void CudaArray::downloadBytes(void* data) const { CUCHECK(cudaMemcpy(data, (void*) pointer, (size_t) size * elementSize, cudaMemcpyDeviceToHost)); }

// ---- CudaIntegrationUtilities -----------------------------------------------------------------------

CudaIntegrationUtilities::CudaIntegrationUtilities(CudaContext& context, const System& system)
    : context(context), system(system), lastStepSize(0), randomSeed(0) {
    const int P = context.getPaddedNumAtoms();
    const int mixed = (context.getUseDoublePrecision() || context.getUseMixedPrecision()) ? 8 : 4;
    posDelta = new CudaArray(context, P, 4 * mixed, "posDelta");
    random = new CudaArray(context, 4 * P, 16, "random");
    stepSize = new CudaArray(context, 1, 2 * mixed, "stepSize");
}

CudaIntegrationUtilities::~CudaIntegrationUtilities() {
    delete posDelta;
    delete random;
    delete stepSize;
}

void CudaIntegrationUtilities::setNextStepSize(double size) {
    if (size == lastStepSize) return;
    lastStepSize = size;
    if (stepSize->getElementSize() == 16) {
        double v[2] = {0, size};
        stepSize->uploadBytes(v);
    } else {
        float v[2] = {0, (float) size};
        stepSize->uploadBytes(v);
    }
}

void CudaIntegrationUtilities::applyConstraints(double tol) {
    constraintCalls++;
    lastConstraintTol = tol;
}

int CudaIntegrationUtilities::prepareRandomNumbers(int numValues) { return 0; }

void CudaIntegrationUtilities::computeVirtualSites() {
    virtualSiteCalls++;
    const int n = context.getNumAtoms(), P = context.getPaddedNumAtoms();
    if (hasVirtualSites < 0) {  // scanned once: OpenMM builds its site lists at context creation
        hasVirtualSites = 0;
        for (int i = 0; i < n && !hasVirtualSites; i++) hasVirtualSites = system.isVirtualSite(i) ? 1 : 0;
    }
    if (!hasVirtualSites) return;
    const vector<int>& order = context.getAtomIndex();
    vector<int> slotOf(P);
    for (int s = 0; s < P; s++) slotOf[order[s]] = s;
    auto place = [&](auto* p, auto* c) {
        for (int i = 0; i < n; i++) {
            if (!system.isVirtualSite(i)) continue;
            const TwoParticleAverageSite* site = dynamic_cast<const TwoParticleAverageSite*>(&system.getVirtualSite(i));
            if (site == NULL) throw OpenMMException("the shim implements TwoParticleAverageSite only");
            const int s0 = slotOf[i], s1 = slotOf[site->getParticle(0)], s2 = slotOf[site->getParticle(1)];
            for (int k = 0; k < 3; k++) {
                const double a = (double) p[4 * s1 + k] + (c ? (double) c[4 * s1 + k] : 0.0);
                const double b = (double) p[4 * s2 + k] + (c ? (double) c[4 * s2 + k] : 0.0);
                const double v = site->getWeight(0) * a + site->getWeight(1) * b;
                p[4 * s0 + k] = (std::remove_reference_t<decltype(p[0])>) v;
                if (c) c[4 * s0 + k] = (std::remove_reference_t<decltype(c[0])>) (v - (double) p[4 * s0 + k]);
            }
        }
    };
    if (context.getUseDoublePrecision()) {
        vector<double> p(4 * (size_t) P);
        context.getPosq().downloadBytes(p.data());
        place(p.data(), (double*) NULL);
        context.getPosq().uploadBytes(p.data());
    } else {
        vector<float> p(4 * (size_t) P), c(4 * (size_t) P, 0.f);
        context.getPosq().downloadBytes(p.data());
        if (context.getUseMixedPrecision()) context.getPosqCorrection().downloadBytes(c.data());
        place(p.data(), context.getUseMixedPrecision() ? c.data() : (float*) NULL);
        context.getPosq().uploadBytes(p.data());
        if (context.getUseMixedPrecision()) context.getPosqCorrection().uploadBytes(c.data());
    }
}

double CudaIntegrationUtilities::computeKineticEnergy(double timeShift) {
    const int P = context.getPaddedNumAtoms(), n = context.getNumAtoms();
    const bool mixedDouble = context.getUseDoublePrecision() || context.getUseMixedPrecision();
    vector<long long> f(3 * (size_t) P);
    context.getForce().downloadBytes(f.data());
    vector<double> v(4 * (size_t) P);
    if (mixedDouble) {
        context.getVelm().downloadBytes(v.data());
    } else {
        vector<float> vf(4 * (size_t) P);
        context.getVelm().downloadBytes(vf.data());
        for (size_t i = 0; i < vf.size(); i++) v[i] = vf[i];
    }
    double energy = 0;
    for (int i = 0; i < n; i++) {
        double w = v[4 * i + 3];
        if (w == 0) continue;
        for (int k = 0; k < 3; k++) {
            double vk = v[4 * i + k] + timeShift * w * (double) f[i + (size_t) P * k] / 4294967296.0;
            energy += vk * vk / w;
        }
    }
    return 0.5 * energy;
}

// ---- CudaContext / CudaPlatform ---------------------------------------------------------------------

CudaContext::CudaContext(const System& system, int deviceIndex, const string& precision, CudaPlatform::PlatformData& platformData)
    : platformData(platformData), deviceIndex(deviceIndex), stepCount(0), atomsWereReordered(false), time(0) {
    useDouble = precision == "double";
    useMixed = precision == "mixed";
    if (!useDouble && !useMixed && precision != "single")
        throw OpenMMException("Illegal value for Precision: " + precision);
    CUCHECK(cudaSetDevice(deviceIndex));
    numAtoms = system.getNumParticles();
    paddedNumAtoms = (numAtoms + 31) / 32 * 32;
    const int real = useDouble ? 8 : 4, mixed = (useDouble || useMixed) ? 8 : 4;
    posq = new CudaArray(*this, paddedNumAtoms, 4 * real, "posq");
    posqCorrection = new CudaArray(*this, paddedNumAtoms, 4 * real, "posqCorrection");
    velm = new CudaArray(*this, paddedNumAtoms, 4 * mixed, "velm");
    force = new CudaArray(*this, 3 * paddedNumAtoms, 8, "force");
    atomIndex = CudaArray::create<int>(*this, paddedNumAtoms, "atomIndex");
    atomIndexHost.resize(paddedNumAtoms);
    for (int i = 0; i < paddedNumAtoms; i++) atomIndexHost[i] = i;
    atomIndex->upload(atomIndexHost);
    // inverse masses
    vector<double> v(4 * (size_t) paddedNumAtoms, 0.0);
    for (int i = 0; i < numAtoms; i++) {
        double m = system.getParticleMass(i);
        v[4 * i + 3] = m == 0 ? 0.0 : 1.0 / m;
    }
    if (mixed == 8) {
        velm->uploadBytes(v.data());
    } else {
        vector<float> vf(v.begin(), v.end());
        velm->uploadBytes(vf.data());
    }
    integration = new CudaIntegrationUtilities(*this, system);
}

CudaContext::~CudaContext() {
    for (ReorderListener* l : reorderListeners) delete l;
    delete integration;
    delete posq;
    delete posqCorrection;
    delete velm;
    delete force;
    delete atomIndex;
}

void CudaContext::setAsCurrent() { cudaSetDevice(deviceIndex); }

template <typename T>
static void permuteRows(CudaArray& array, int paddedNumAtoms, int numAtoms, const vector<int>& oldIndex, const vector<int>& newIndex) {
    vector<T> data(4 * (size_t) paddedNumAtoms), out(4 * (size_t) paddedNumAtoms);
    array.downloadBytes(data.data());
    out = data;
    vector<int> slotOf(paddedNumAtoms);
    for (int s = 0; s < paddedNumAtoms; s++) slotOf[oldIndex[s]] = s;
    for (int s = 0; s < numAtoms; s++) {
        const int from = slotOf[newIndex[s]];
        for (int k = 0; k < 4; k++) out[4 * (size_t) s + k] = data[4 * (size_t) from + k];
    }
    array.uploadBytes(out.data());
}

void CudaContext::reorderAtoms() {
    atomsWereReordered = false;
    if (pendingOrder.empty()) return;
    vector<int> next = atomIndexHost;
    for (int i = 0; i < numAtoms; i++) next[i] = pendingOrder[i];
    if (useDouble) permuteRows<double>(*posq, paddedNumAtoms, numAtoms, atomIndexHost, next);
    else permuteRows<float>(*posq, paddedNumAtoms, numAtoms, atomIndexHost, next);
    if (useMixed) permuteRows<float>(*posqCorrection, paddedNumAtoms, numAtoms, atomIndexHost, next);
    if (useDouble || useMixed) permuteRows<double>(*velm, paddedNumAtoms, numAtoms, atomIndexHost, next);
    else permuteRows<float>(*velm, paddedNumAtoms, numAtoms, atomIndexHost, next);
    atomIndexHost = next;
    atomIndex->upload(atomIndexHost);
    pendingOrder.clear();
    atomsWereReordered = true;
    for (ReorderListener* l : reorderListeners) l->execute();
}

CudaPlatform::CudaPlatform() {
    setPropertyDefaultValue(CudaPrecision(), "single");
    setPropertyDefaultValue(CudaDeviceIndex(), "0");
}

CudaPlatform::PlatformData::PlatformData(ContextImpl* context, const System& system, const string& deviceIndex, const string& precision)
    : context(context), contextsInitialized(true) {
    contexts.push_back(new CudaContext(system, std::atoi(deviceIndex.c_str()), precision, *this));
}

CudaPlatform::PlatformData::~PlatformData() {
    for (CudaContext* c : contexts) delete c;
}

void CudaPlatform::contextCreated(ContextImpl& context, const std::map<string, string>& properties) const {
    auto get = [&](const string& key) {
        auto it = properties.find(key);
        if (it != properties.end()) return it->second;
        it = properties.find("Cuda" + key);  // OpenMM accepts the deprecated CudaPrecision spelling
        return it != properties.end() ? it->second : getPropertyDefaultValue(key);
    };
    context.setPlatformData(new PlatformData(&context, context.getSystem(), get(CudaDeviceIndex()), get(CudaPrecision())));
}

void CudaPlatform::contextDestroyed(ContextImpl& context) const {
    delete static_cast<PlatformData*>(context.getPlatformData());
    context.setPlatformData(NULL);
}

// ---- ContextImpl / Context --------------------------------------------------------------------------

static CudaContext& cudaOf(ContextImpl& impl) {
    return *static_cast<CudaPlatform::PlatformData*>(impl.getPlatformData())->contexts[0];
}

ContextImpl::ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform& platform,
                         const std::map<string, string>& properties)
    : owner(owner), system(system), integrator(integrator), platform(&platform), platformData(NULL), lastEnergy(0) {
    platform.contextCreated(*this, properties);
    try {
        integrator.initialize(*this);
    } catch (...) {
        platform.contextDestroyed(*this);
        throw;
    }
}

ContextImpl::~ContextImpl() {
    integrator.cleanup();
    platform->contextDestroyed(*this);
}

static void downloadPositions(CudaContext& cu, vector<Vec3>& pos, vector<double>* charges = NULL) {
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    const vector<int>& order = cu.getAtomIndex();
    pos.assign(n, Vec3());
    if (charges) charges->assign(n, 0.0);
    if (cu.getUseDoublePrecision()) {
        vector<double> p(4 * (size_t) P);
        cu.getPosq().downloadBytes(p.data());
        for (int s = 0; s < n; s++) {
            pos[order[s]] = Vec3(p[4 * s], p[4 * s + 1], p[4 * s + 2]);
            if (charges) (*charges)[order[s]] = p[4 * s + 3];
        }
    } else {
        vector<float> p(4 * (size_t) P), c(4 * (size_t) P, 0.f);
        cu.getPosq().downloadBytes(p.data());
        if (cu.getUseMixedPrecision()) cu.getPosqCorrection().downloadBytes(c.data());
        for (int s = 0; s < n; s++) {
            pos[order[s]] = Vec3((double) p[4 * s] + c[4 * s], (double) p[4 * s + 1] + c[4 * s + 1], (double) p[4 * s + 2] + c[4 * s + 2]);
            if (charges) (*charges)[order[s]] = p[4 * s + 3];
        }
    }
}

double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    CudaContext& cu = cudaOf(*this);
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    vector<Vec3> pos, f(n);
    downloadPositions(cu, pos);
    double energy = 0;
    for (int i = 0; i < system.getNumForces(); i++) energy += system.getForce(i).compute(pos, f);
    vector<long long> fixed(3 * (size_t) P, 0);
    const vector<int>& order = cu.getAtomIndex();
    for (int s = 0; s < n; s++)
        for (int k = 0; k < 3; k++) fixed[s + (size_t) P * k] = (long long) std::llround(f[order[s]][k] * 4294967296.0);
    cu.getForce().uploadBytes(fixed.data());
    lastEnergy = energy;
    return energy;
}

Context::Context(const System& system, Integrator& integrator, Platform& platform) : impl(NULL) {
    impl = new ContextImpl(*this, system, integrator, platform, properties);
}

Context::Context(const System& system, Integrator& integrator, Platform& platform, const std::map<string, string>& props)
    : impl(NULL), properties(props) {
    impl = new ContextImpl(*this, system, integrator, platform, properties);
}

Context::~Context() { delete impl; }

void Context::reinitialize() {
    const System& system = impl->getSystem();
    Integrator& integrator = impl->getIntegrator();
    Platform& platform = impl->getPlatform();
    delete impl;
    impl = NULL;
    impl = new ContextImpl(*this, system, integrator, platform, properties);
}

void Context::setTime(double time) { cudaOf(*impl).setTime(time); }

void Context::setPositions(const vector<Vec3>& positions) {
    CudaContext& cu = cudaOf(*impl);
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    if ((int) positions.size() != n) throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
    const vector<int>& order = cu.getAtomIndex();
    if (cu.getUseDoublePrecision()) {
        vector<double> p(4 * (size_t) P);
        cu.getPosq().downloadBytes(p.data());
        for (int s = 0; s < n; s++)
            for (int k = 0; k < 3; k++) p[4 * s + k] = positions[order[s]][k];
        cu.getPosq().uploadBytes(p.data());
    } else {
        vector<float> p(4 * (size_t) P), c(4 * (size_t) P, 0.f);
        cu.getPosq().downloadBytes(p.data());
        for (int s = 0; s < n; s++)
            for (int k = 0; k < 3; k++) {
                p[4 * s + k] = (float) positions[order[s]][k];
                c[4 * s + k] = (float) (positions[order[s]][k] - (double) p[4 * s + k]);
            }
        cu.getPosq().uploadBytes(p.data());
        if (cu.getUseMixedPrecision()) cu.getPosqCorrection().uploadBytes(c.data());
    }
}

void Context::setCharges(const vector<double>& charges) {
    CudaContext& cu = cudaOf(*impl);
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    const vector<int>& order = cu.getAtomIndex();
    if (cu.getUseDoublePrecision()) {
        vector<double> p(4 * (size_t) P);
        cu.getPosq().downloadBytes(p.data());
        for (int s = 0; s < n; s++) p[4 * s + 3] = charges[order[s]];
        cu.getPosq().uploadBytes(p.data());
    } else {
        vector<float> p(4 * (size_t) P);
        cu.getPosq().downloadBytes(p.data());
        for (int s = 0; s < n; s++) p[4 * s + 3] = (float) charges[order[s]];
        cu.getPosq().uploadBytes(p.data());
    }
}

void Context::setVelocities(const vector<Vec3>& velocities) {
    CudaContext& cu = cudaOf(*impl);
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    const vector<int>& order = cu.getAtomIndex();
    if (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) {
        vector<double> v(4 * (size_t) P);
        cu.getVelm().downloadBytes(v.data());
        for (int s = 0; s < n; s++)
            for (int k = 0; k < 3; k++) v[4 * s + k] = velocities[order[s]][k];
        cu.getVelm().uploadBytes(v.data());
    } else {
        vector<float> v(4 * (size_t) P);
        cu.getVelm().downloadBytes(v.data());
        for (int s = 0; s < n; s++)
            for (int k = 0; k < 3; k++) v[4 * s + k] = (float) velocities[order[s]][k];
        cu.getVelm().uploadBytes(v.data());
    }
}

void Context::setVelocitiesToTemperature(double temperature, int randomSeed) {
    const System& system = impl->getSystem();
    std::mt19937_64 gen(randomSeed);
    std::normal_distribution<double> normal(0.0, 1.0);
    vector<Vec3> v(system.getNumParticles());
    for (int i = 0; i < system.getNumParticles(); i++) {
        double m = system.getParticleMass(i);
        if (m == 0) continue;
        double s = std::sqrt(BOLTZ * temperature / m);
        v[i] = Vec3(s * normal(gen), s * normal(gen), s * normal(gen));
    }
    setVelocities(v);
}

State Context::getState(int types) {
    CudaContext& cu = cudaOf(*impl);
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    const vector<int>& order = cu.getAtomIndex();
    State state;
    state.time = cu.getTime();
    if (types & State::Positions) downloadPositions(cu, state.positions);
    if (types & State::Velocities) {
        state.velocities.assign(n, Vec3());
        if (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) {
            vector<double> v(4 * (size_t) P);
            cu.getVelm().downloadBytes(v.data());
            for (int s = 0; s < n; s++) state.velocities[order[s]] = Vec3(v[4 * s], v[4 * s + 1], v[4 * s + 2]);
        } else {
            vector<float> v(4 * (size_t) P);
            cu.getVelm().downloadBytes(v.data());
            for (int s = 0; s < n; s++) state.velocities[order[s]] = Vec3(v[4 * s], v[4 * s + 1], v[4 * s + 2]);
        }
    }
    if (types & (State::Energy | State::Forces)) {
        state.pe = impl->calcForcesAndEnergy(true, true);
        if (types & State::Energy) state.ke = impl->getIntegrator().computeKineticEnergy();
        if (types & State::Forces) {
            vector<long long> f(3 * (size_t) P);
            cu.getForce().downloadBytes(f.data());
            state.forces.assign(n, Vec3());
            for (int s = 0; s < n; s++)
                state.forces[order[s]] = Vec3(f[s] / 4294967296.0, f[s + (size_t) P] / 4294967296.0, f[s + 2 * (size_t) P] / 4294967296.0);
        }
    }
    return state;
}

// ---- serialization ----------------------------------------------------------------------------------

const string& SerializationNode::getStringProperty(const string& key) const {
    auto it = properties.find(key);
    if (it == properties.end()) throw OpenMMException("Unknown property '" + key + "' in node '" + name + "'");
    return it->second;
}

const string& SerializationNode::getStringProperty(const string& key, const string& defaultValue) const {
    auto it = properties.find(key);
    return it == properties.end() ? defaultValue : it->second;
}

SerializationNode& SerializationNode::setStringProperty(const string& key, const string& value) {
    properties[key] = value;
    return *this;
}

int SerializationNode::getIntProperty(const string& key) const { return std::atoi(getStringProperty(key).c_str()); }
int SerializationNode::getIntProperty(const string& key, int d) const { return hasProperty(key) ? getIntProperty(key) : d; }

SerializationNode& SerializationNode::setIntProperty(const string& key, int value) {
    std::stringstream s;
    s << value;
    return setStringProperty(key, s.str());
}

double SerializationNode::getDoubleProperty(const string& key) const { return std::strtod(getStringProperty(key).c_str(), NULL); }
double SerializationNode::getDoubleProperty(const string& key, double d) const { return hasProperty(key) ? getDoubleProperty(key) : d; }

SerializationNode& SerializationNode::setDoubleProperty(const string& key, double value) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.17g", value);  // round-trips exactly
    return setStringProperty(key, buf);
}

SerializationNode& SerializationNode::createChildNode(const string& childName) {
    children.push_back(SerializationNode());
    children.back().setName(childName);
    return children.back();
}

static std::map<string, const SerializationProxy*>& proxiesByType() {
    static std::map<string, const SerializationProxy*> m;
    return m;
}
static std::map<string, const SerializationProxy*>& proxiesByName() {
    static std::map<string, const SerializationProxy*> m;
    return m;
}

void SerializationProxy::registerProxy(const std::type_info& type, const SerializationProxy* proxy) {
    proxiesByType()[type.name()] = proxy;
    proxiesByName()[proxy->getTypeName()] = proxy;
}

const SerializationProxy& SerializationProxy::getProxy(const string& typeName) {
    auto it = proxiesByName().find(typeName);
    if (it == proxiesByName().end()) throw OpenMMException("There is no serialization proxy registered for type " + typeName);
    return *it->second;
}

const SerializationProxy& SerializationProxy::getProxy(const std::type_info& type) {
    auto it = proxiesByType().find(type.name());
    if (it == proxiesByType().end()) throw OpenMMException(string("There is no serialization proxy registered for type ") + type.name());
    return *it->second;
}

void XmlSerializer::write(const SerializationNode& node, std::ostream& stream, int depth) {
    if (depth == 0) stream << "<?xml version=\"1.0\" ?>\n";
    stream << string(depth, '\t') << '<' << node.getName();
    for (const auto& kv : node.getProperties()) stream << ' ' << kv.first << "=\"" << kv.second << '"';
    if (node.getChildren().empty()) {
        stream << "/>\n";
        return;
    }
    stream << ">\n";
    for (const SerializationNode& child : node.getChildren()) write(child, stream, depth + 1);
    stream << string(depth, '\t') << "</" << node.getName() << ">\n";
}

namespace {
struct Parser {
    string text;
    size_t pos = 0;
    void skipSpace() { while (pos < text.size() && isspace((unsigned char) text[pos])) pos++; }
    bool startsWith(const char* s) const { return text.compare(pos, strlen(s), s) == 0; }
    string readName() {
        size_t start = pos;
        while (pos < text.size() && (isalnum((unsigned char) text[pos]) || text[pos] == '_' || text[pos] == ':' || text[pos] == '.' || text[pos] == '-')) pos++;
        return text.substr(start, pos - start);
    }
    void parseElement(SerializationNode& node) {
        skipSpace();
        while (startsWith("<?") || startsWith("<!--")) {
            pos = text.find('>', pos) + 1;
            skipSpace();
        }
        if (text[pos] != '<') throw OpenMMException("XML parse error: expected '<'");
        pos++;
        node.setName(readName());
        for (;;) {
            skipSpace();
            if (startsWith("/>")) { pos += 2; return; }
            if (text[pos] == '>') { pos++; break; }
            string key = readName();
            skipSpace();
            if (text[pos] != '=') throw OpenMMException("XML parse error: expected '='");
            pos++;
            skipSpace();
            char quote = text[pos++];
            size_t end = text.find(quote, pos);
            node.setStringProperty(key, text.substr(pos, end - pos));
            pos = end + 1;
        }
        for (;;) {
            skipSpace();
            if (startsWith("</")) { pos = text.find('>', pos) + 1; return; }
            SerializationNode& child = node.createChildNode("");
            parseElement(child);
        }
    }
};
}  // namespace

void XmlSerializer::read(std::istream& stream, SerializationNode& node) {
    Parser p;
    std::stringstream buffer;
    buffer << stream.rdbuf();
    p.text = buffer.str();
    p.parseElement(node);
}
