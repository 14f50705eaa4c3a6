// Platform-independent half of the integrator: owns the parameters and drives the kernel.
// Behaviour follows scychon/openmm_constV openmmapi/src/ConstVLangevinIntegrator.cpp:44-85
// (defaults :50-51, binding errors :56,79, per-step sequence :80-83).
#include "openmm/ConstVLangevinIntegrator.h"

#include "openmm/ConstVKernels.h"
#include "openmm/Context.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

ConstVLangevinIntegrator::ConstVLangevinIntegrator(double temperature, double frictionCoeff, double stepSize, int numCells, double zmax)
    : temperature(temperature), friction(frictionCoeff), cellsize(zmax), randomNumberSeed(0), numCellCopies(numCells) {
    setStepSize(stepSize);
    setConstraintTolerance(1e-5);
}

void ConstVLangevinIntegrator::initialize(ContextImpl& contextRef) {
    if (owner != NULL && &contextRef.getOwner() != owner)
        throw OpenMMException("This Integrator is already bound to a context");
    context = &contextRef;
    owner = &contextRef.getOwner();
    kernel = context->getPlatform().createKernel(IntegrateConstVLangevinStepKernel::Name(), contextRef);
    try {
        kernel.getAs<IntegrateConstVLangevinStepKernel>().initialize(contextRef.getSystem(), *this);
    } catch (...) {
        // the Context is about to be torn down: release the kernel while its platform data is alive
        kernel = Kernel();
        throw;
    }
}

void ConstVLangevinIntegrator::cleanup() {
    kernel = Kernel();
}

std::vector<std::string> ConstVLangevinIntegrator::getKernelNames() {
    return std::vector<std::string>(1, IntegrateConstVLangevinStepKernel::Name());
}

double ConstVLangevinIntegrator::computeKineticEnergy() {
    return kernel.getAs<IntegrateConstVLangevinStepKernel>().computeKineticEnergy(*context, *this);
}

void ConstVLangevinIntegrator::step(int steps) {
    if (context == NULL)
        throw OpenMMException("This Integrator is not bound to a context!");
    IntegrateConstVLangevinStepKernel& stepKernel = kernel.getAs<IntegrateConstVLangevinStepKernel>();
    for (int i = 0; i < steps; ++i) {
        context->updateContextState();
        context->calcForcesAndEnergy(true, false);
        stepKernel.execute(*context, *this);
    }
}
