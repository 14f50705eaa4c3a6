#ifndef CUDA_CONSTV_KERNELS_H_
#define CUDA_CONSTV_KERNELS_H_

// CUDA-platform implementation of IntegrateConstVLangevinStepKernel on top of the B200 C ABI
// (include/constv_b200.h).  Replaces the reference's class of the same name
// (platforms/cuda/src/CudaConstVKernels.h:46-79, CudaConstVKernels.cpp:53-177): instead of four
// runtime-compiled kernels it binds OpenMM's device arrays to a constv_slab and calls the
// precompiled sm_100a kernels.

#include "CudaArray.h"
#include "CudaContext.h"
#include "constv_b200.h"
#include "openmm/ConstVKernels.h"

namespace OpenMM {

class CudaIntegrateConstVLangevinStepKernel : public IntegrateConstVLangevinStepKernel {
public:
    CudaIntegrateConstVLangevinStepKernel(std::string name, const Platform& platform, CudaContext& cu)
        : IntegrateConstVLangevinStepKernel(name, platform), cu(cu), slab(NULL), zmax(0), hasConstraints(false),
          fusedConstraints(false), numRigidMolecules(0), slotsPermuted(false) {
    }
    ~CudaIntegrateConstVLangevinStepKernel();
    void initialize(const System& system, const ConstVLangevinIntegrator& integrator);
    void execute(ContextImpl& context, const ConstVLangevinIntegrator& integrator);
    double computeKineticEnergy(ContextImpl& context, const ConstVLangevinIntegrator& integrator);
    /** The C-ABI handle (tests and tools). */
    constv_slab* getSlab() { return slab; }
    double getCellSize() const { return zmax; }
    /** True when every constraint of the System is part of a rigid 3-site molecule: the constraints are then
     *  solved inside the step kernel instead of being handed to integration.applyConstraints. */
    bool getConstraintsAreFused() const { return fusedConstraints; }
    int getNumRigidMolecules() const { return numRigidMolecules; }

private:
    void bindArrays();
    CudaContext& cu;
    constv_slab* slab;
    double zmax;
    bool hasConstraints;
    bool fusedConstraints;
    int numRigidMolecules;
    bool slotsPermuted;  // true once OpenMM has re-sorted the atoms at least once
};

// CUDA-platform implementation of IntegrateConstVDrudeLangevinStepKernel; replaces the reference's
// class of the same name (CudaConstVKernels.h:84-118, CudaConstVKernels.cpp:179-357): its four
// runtime-compiled kernels per step (part 1, part 2, hard wall, image rewrite) become one precompiled
// kernel, or two around the constraint solver.
class CudaIntegrateConstVDrudeLangevinStepKernel : public IntegrateConstVDrudeLangevinStepKernel {
public:
    CudaIntegrateConstVDrudeLangevinStepKernel(std::string name, const Platform& platform, CudaContext& cu)
        : IntegrateConstVDrudeLangevinStepKernel(name, platform), cu(cu), slab(NULL), zmax(0), hasConstraints(false), slotsPermuted(false) {
    }
    ~CudaIntegrateConstVDrudeLangevinStepKernel();
    void initialize(const System& system, const ConstVDrudeLangevinIntegrator& integrator, const DrudeForce& force);
    void execute(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator);
    double computeKineticEnergy(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator);
    constv_slab* getSlab() { return slab; }
    double getCellSize() const { return zmax; }

private:
    void bindArrays();
    CudaContext& cu;
    constv_slab* slab;
    double zmax;
    bool hasConstraints;
    bool slotsPermuted;
};

}  // namespace OpenMM

#endif /* CUDA_CONSTV_KERNELS_H_ */
