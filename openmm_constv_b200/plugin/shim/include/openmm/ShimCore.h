// Minimal stand-in for the part of OpenMM 7.x that the ConstV Langevin plugin touches.
//
// OpenMM is not available in the build container, so the plugin's host C++ is compiled
// and tested against this shim; with a real OpenMM, pass OPENMM_DIR to build.sh and none of these
// headers is used.  Only the symbols listed in SURVEY.md 8b "OpenMM symbols the host code touches"
// are declared, with the signatures the reference's call sites use.  This is test scaffolding, not
// a reimplementation of OpenMM: no force fields, no constraint solver, no reordering heuristics.
#ifndef OPENMM_SHIM_CORE_H_
#define OPENMM_SHIM_CORE_H_

#include <cmath>
#include <exception>
#include <functional>
#include <map>
#include <string>
#include <vector>

#define OPENMM_EXPORT __attribute__((visibility("default")))

namespace OpenMM {

class OPENMM_EXPORT OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& message) : message(message) {}
    ~OpenMMException() throw() {}
    const char* what() const throw() { return message.c_str(); }
private:
    std::string message;
};

class Vec3 {
public:
    Vec3() : data{0, 0, 0} {}
    Vec3(double x, double y, double z) : data{x, y, z} {}
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    Vec3 operator+(const Vec3& o) const { return Vec3(data[0] + o[0], data[1] + o[1], data[2] + o[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(data[0] - o[0], data[1] - o[1], data[2] - o[2]); }
    Vec3 operator*(double s) const { return Vec3(data[0] * s, data[1] * s, data[2] * s); }
    double dot(const Vec3& o) const { return data[0] * o[0] + data[1] * o[1] + data[2] * o[2]; }
private:
    double data[3];
};

// A force in the shim is a host callback: positions (nm) in, forces (kJ/mol/nm) accumulated, energy out.
class OPENMM_EXPORT Force {
public:
    virtual ~Force() {}
    virtual double compute(const std::vector<Vec3>& positions, std::vector<Vec3>& forces) const = 0;
};

class OPENMM_EXPORT HarmonicBondForce : public Force {
public:
    int addBond(int p1, int p2, double length, double k);
    double compute(const std::vector<Vec3>& positions, std::vector<Vec3>& forces) const;
private:
    struct Bond { int p1, p2; double length, k; };
    std::vector<Bond> bonds;
};

// Independent harmonic wells 0.5*k*|x - x0|^2 per particle (stands in for any bounded force field).
class OPENMM_EXPORT HarmonicWellForce : public Force {
public:
    int addParticle(int particle, const Vec3& centre, double k);
    double compute(const std::vector<Vec3>& positions, std::vector<Vec3>& forces) const;
private:
    struct Well { int p; Vec3 c; double k; };
    std::vector<Well> wells;
};

// The slice of OpenMM's DrudeForce (Drude plugin) the Drude integrator reads: one entry per Drude
// particle with its parent `particle1` (CudaConstVKernels.cpp:219-224).  In the shim the force itself is
// the isotropic spring 0.5*k*|r - r1|^2 with k = ONE_4PI_EPS0*charge^2/polarizability; anisotropy and
// Thole screening are not modelled.
class OPENMM_EXPORT DrudeForce : public Force {
public:
    int addParticle(int particle, int particle1, int particle2, int particle3, int particle4, double charge,
                    double polarizability, double aniso12, double aniso34);
    int getNumParticles() const { return (int) particles.size(); }
    void getParticleParameters(int index, int& particle, int& particle1, int& particle2, int& particle3, int& particle4,
                               double& charge, double& polarizability, double& aniso12, double& aniso34) const;
    double compute(const std::vector<Vec3>& positions, std::vector<Vec3>& forces) const;
private:
    struct Entry { int p, p1, p2, p3, p4; double charge, polarizability, aniso12, aniso34; };
    std::vector<Entry> particles;
};

// Virtual sites as OpenMM declares them (openmm/VirtualSite.h); the shim implements the two-particle average.
class OPENMM_EXPORT VirtualSite {
public:
    virtual ~VirtualSite() {}
    int getNumParticles() const { return (int) particles.size(); }
    int getParticle(int particle) const { return particles.at(particle); }
protected:
    void setParticles(const std::vector<int>& p) { particles = p; }
private:
    std::vector<int> particles;
};

class OPENMM_EXPORT TwoParticleAverageSite : public VirtualSite {
public:
    TwoParticleAverageSite(int particle1, int particle2, double weight1, double weight2) : weight1(weight1), weight2(weight2) {
        setParticles({particle1, particle2});
    }
    double getWeight(int particle) const { return particle == 0 ? weight1 : weight2; }
private:
    double weight1, weight2;
};

class OPENMM_EXPORT System {
public:
    System();
    ~System();
    void setVirtualSite(int index, VirtualSite* virtualSite);  // takes ownership
    bool isVirtualSite(int index) const { return index < (int) virtualSites.size() && virtualSites[index] != NULL; }
    const VirtualSite& getVirtualSite(int index) const;
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size() - 1; }
    int getNumParticles() const { return (int) masses.size(); }
    double getParticleMass(int index) const { return masses.at(index); }
    void setParticleMass(int index, double mass) { masses.at(index) = mass; }
    int addConstraint(int p1, int p2, double distance);
    int getNumConstraints() const { return (int) constraints.size(); }
    void getConstraintParameters(int index, int& p1, int& p2, double& distance) const;
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size() - 1; }  // takes ownership
    int getNumForces() const { return (int) forces.size(); }
    const Force& getForce(int index) const { return *forces.at(index); }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
private:
    struct Constraint { int p1, p2; double distance; };
    std::vector<double> masses;
    std::vector<Constraint> constraints;
    std::vector<Force*> forces;
    std::vector<VirtualSite*> virtualSites;
    Vec3 box[3];
};

class Platform;
class ContextImpl;
class Context;

class OPENMM_EXPORT KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), referenceCount(1) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int referenceCount;
};

// Reference-counted handle, as in OpenMM: the integrator drops its kernel with `kernel = Kernel()`.
class OPENMM_EXPORT Kernel {
public:
    Kernel() : impl(NULL) {}
    Kernel(KernelImpl* impl) : impl(impl) {}
    Kernel(const Kernel& copy) : impl(copy.impl) { if (impl) impl->referenceCount++; }
    ~Kernel() { release(); }
    Kernel& operator=(const Kernel& copy) {
        if (copy.impl) copy.impl->referenceCount++;
        release();
        impl = copy.impl;
        return *this;
    }
    std::string getName() const { return impl->getName(); }
    const KernelImpl& getImpl() const { return *impl; }
    KernelImpl& getImpl() { return *impl; }
    template <class T> const T& getAs() const { return dynamic_cast<const T&>(*impl); }
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
private:
    void release() { if (impl && --impl->referenceCount == 0) delete impl; impl = NULL; }
    KernelImpl* impl;
};

class OPENMM_EXPORT KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
};

class OPENMM_EXPORT Platform {
public:
    virtual ~Platform() {}
    virtual const std::string& getName() const = 0;
    void registerKernelFactory(const std::string& name, KernelFactory* factory) { kernelFactories[name] = factory; }
    Kernel createKernel(const std::string& name, ContextImpl& context) const;
    void setPropertyDefaultValue(const std::string& property, const std::string& value) { defaults[property] = value; }
    const std::string& getPropertyDefaultValue(const std::string& property) const;
    virtual void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const = 0;
    virtual void contextDestroyed(ContextImpl& context) const = 0;
    static void registerPlatform(Platform* platform);
    static Platform& getPlatformByName(const std::string& name);
    static int getNumPlatforms();
private:
    std::map<std::string, KernelFactory*> kernelFactories;
    std::map<std::string, std::string> defaults;
};

class OPENMM_EXPORT State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16 };
    double getTime() const { return time; }
    const std::vector<Vec3>& getPositions() const { return positions; }
    const std::vector<Vec3>& getVelocities() const { return velocities; }
    const std::vector<Vec3>& getForces() const { return forces; }
    double getKineticEnergy() const { return ke; }
    double getPotentialEnergy() const { return pe; }
private:
    friend class Context;
    double time = 0, ke = 0, pe = 0;
    std::vector<Vec3> positions, velocities, forces;
};

class OPENMM_EXPORT Integrator {
public:
    Integrator() : context(NULL), owner(NULL), stepSize(0), constraintTol(1e-5) {}
    virtual ~Integrator() {}
    virtual double getStepSize() const { return stepSize; }
    virtual void setStepSize(double size) { stepSize = size; }
    virtual double getConstraintTolerance() const { return constraintTol; }
    virtual void setConstraintTolerance(double tol) { constraintTol = tol; }
    virtual void step(int steps) = 0;
protected:
    friend class Context;
    friend class ContextImpl;
    ContextImpl* context;
    Context* owner;
    virtual void initialize(ContextImpl& context) = 0;
    virtual void cleanup() {}
    virtual std::vector<std::string> getKernelNames() = 0;
    virtual double computeKineticEnergy() = 0;
private:
    double stepSize, constraintTol;
};

class OPENMM_EXPORT ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform& platform,
                const std::map<std::string, std::string>& properties);
    ~ContextImpl();
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    void setPlatformData(void* data) { platformData = data; }
    void updateContextState() {}
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    double getLastPotentialEnergy() const { return lastEnergy; }
private:
    friend class Context;
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    void* platformData;
    double lastEnergy;
};

class OPENMM_EXPORT Context {
public:
    Context(const System& system, Integrator& integrator, Platform& platform);
    Context(const System& system, Integrator& integrator, Platform& platform,
            const std::map<std::string, std::string>& properties);
    ~Context();
    const System& getSystem() const { return impl->getSystem(); }
    Integrator& getIntegrator() { return impl->getIntegrator(); }
    Platform& getPlatform() { return impl->getPlatform(); }
    State getState(int types);
    void setTime(double time);
    void setPositions(const std::vector<Vec3>& positions);
    void setVelocities(const std::vector<Vec3>& velocities);
    void setVelocitiesToTemperature(double temperature, int randomSeed = 1234);
    // shim extension: per-particle charges live in posq.w (OpenMM sets them from NonbondedForce)
    void setCharges(const std::vector<double>& charges);
    void reinitialize();
    ContextImpl& getImpl() { return *impl; }
private:
    ContextImpl* impl;
    std::map<std::string, std::string> properties;
};

}  // namespace OpenMM

#define BOLTZ 8.31446261815324e-3 /* kJ/mol/K, as SimTKOpenMMRealType.h (OpenMM >= 7.4) */

#endif  // OPENMM_SHIM_CORE_H_
