#ifndef CONSTV_KERNELS_H_
#define CONSTV_KERNELS_H_

// Platform-independent kernel interfaces of the ConstV Langevin steps; same names and virtuals as
// scychon/openmm_constV openmmapi/include/openmm/ConstVKernels.h:49-77 (Langevin) and :82-108
// (Drude Langevin) so that Platform::createKernel("IntegrateConstVLangevinStep") /
// ("IntegrateConstVDrudeLangevinStep") find the B200 implementations.

#include <string>

#include "openmm/ConstVDrudeLangevinIntegrator.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/DrudeForce.h"
#include "openmm/Platform.h"
#include "openmm/System.h"

namespace OpenMM {

class IntegrateConstVLangevinStepKernel : public KernelImpl {
public:
    static std::string Name() { return "IntegrateConstVLangevinStep"; }
    IntegrateConstVLangevinStepKernel(std::string name, const Platform& platform) : KernelImpl(name, platform) {}
    /** Validate the System (image particles massless, cell size consistent) and allocate state. */
    virtual void initialize(const System& system, const ConstVLangevinIntegrator& integrator) = 0;
    /** Take one time step. */
    virtual void execute(ContextImpl& context, const ConstVLangevinIntegrator& integrator) = 0;
    /** Kinetic energy at the current time (velocities shifted by half a step). */
    virtual double computeKineticEnergy(ContextImpl& context, const ConstVLangevinIntegrator& integrator) = 0;
};

class IntegrateConstVDrudeLangevinStepKernel : public KernelImpl {
public:
    static std::string Name() { return "IntegrateConstVDrudeLangevinStep"; }
    IntegrateConstVDrudeLangevinStepKernel(std::string name, const Platform& platform) : KernelImpl(name, platform) {}
    /** As above, plus the (Drude particle, parent) pairs read from `force`. */
    virtual void initialize(const System& system, const ConstVDrudeLangevinIntegrator& integrator, const DrudeForce& force) = 0;
    virtual void execute(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator) = 0;
    virtual double computeKineticEnergy(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator) = 0;
};

}  // namespace OpenMM

#endif /* CONSTV_KERNELS_H_ */
