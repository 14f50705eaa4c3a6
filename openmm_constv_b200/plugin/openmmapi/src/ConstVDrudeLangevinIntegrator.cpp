// Platform-independent half of the Drude integrator.  Behaviour follows scychon/openmm_constV
// openmmapi/src/ConstVDrudeLangevinIntegrator.cpp:44-129 (defaults :49-54, setMaxDrudeDistance
// :61-65, the DrudeForce lookup and its two messages :70-80, per-step sequence :120-128).
#include "openmm/ConstVDrudeLangevinIntegrator.h"

#include "openmm/ConstVKernels.h"
#include "openmm/Context.h"
#include "openmm/DrudeForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

ConstVDrudeLangevinIntegrator::ConstVDrudeLangevinIntegrator(double temperature, double frictionCoeff, double drudeTemperature,
                                                             double drudeFrictionCoeff, double stepSize, int numCells, double zmax)
    : temperature(temperature), friction(frictionCoeff), drudeTemperature(drudeTemperature), drudeFriction(drudeFrictionCoeff),
      maxDrudeDistance(0), cellsize(zmax), randomNumberSeed(0), numCellCopies(numCells) {
    setStepSize(stepSize);
    setConstraintTolerance(1e-5);
}

void ConstVDrudeLangevinIntegrator::setMaxDrudeDistance(double distance) {
    if (distance < 0)
        throw OpenMMException("setMaxDrudeDistance: Distance cannot be negative");
    maxDrudeDistance = distance;
}

void ConstVDrudeLangevinIntegrator::initialize(ContextImpl& contextRef) {
    if (owner != NULL && &contextRef.getOwner() != owner)
        throw OpenMMException("This Integrator is already bound to a context");
    const System& system = contextRef.getSystem();
    const DrudeForce* drude = NULL;
    for (int i = 0; i < system.getNumForces(); i++) {
        const DrudeForce* candidate = dynamic_cast<const DrudeForce*>(&system.getForce(i));
        if (candidate == NULL)
            continue;
        if (drude != NULL)
            throw OpenMMException("The System contains multiple DrudeForces");
        drude = candidate;
    }
    if (drude == NULL)
        throw OpenMMException("The System does not contain a DrudeForce");
    context = &contextRef;
    owner = &contextRef.getOwner();
    kernel = context->getPlatform().createKernel(IntegrateConstVDrudeLangevinStepKernel::Name(), contextRef);
    try {
        kernel.getAs<IntegrateConstVDrudeLangevinStepKernel>().initialize(system, *this, *drude);
    } catch (...) {
        kernel = Kernel();  // release it while the platform data is alive
        throw;
    }
}

void ConstVDrudeLangevinIntegrator::cleanup() {
    kernel = Kernel();
}

std::vector<std::string> ConstVDrudeLangevinIntegrator::getKernelNames() {
    return std::vector<std::string>(1, IntegrateConstVDrudeLangevinStepKernel::Name());
}

double ConstVDrudeLangevinIntegrator::computeKineticEnergy() {
    return kernel.getAs<IntegrateConstVDrudeLangevinStepKernel>().computeKineticEnergy(*context, *this);
}

void ConstVDrudeLangevinIntegrator::step(int steps) {
    if (context == NULL)
        throw OpenMMException("This Integrator is not bound to a context!");
    IntegrateConstVDrudeLangevinStepKernel& stepKernel = kernel.getAs<IntegrateConstVDrudeLangevinStepKernel>();
    for (int i = 0; i < steps; ++i) {
        context->updateContextState();
        context->calcForcesAndEnergy(true, false);
        stepKernel.execute(*context, *this);
    }
}
