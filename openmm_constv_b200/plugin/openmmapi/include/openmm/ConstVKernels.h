#ifndef CONSTV_KERNELS_H_
#define CONSTV_KERNELS_H_

// Platform-independent kernel interface of the ConstV Langevin step; same name and virtuals as
// scychon/openmm_constV openmmapi/include/openmm/ConstVKernels.h:49-77 so that
// Platform::createKernel("IntegrateConstVLangevinStep") finds the B200 implementation.

#include <string>

#include "openmm/ConstVLangevinIntegrator.h"
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

}  // namespace OpenMM

#endif /* CONSTV_KERNELS_H_ */
