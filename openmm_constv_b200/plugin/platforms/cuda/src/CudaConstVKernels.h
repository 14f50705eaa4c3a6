#ifndef CUDA_CONSTV_KERNELS_H_
#define CUDA_CONSTV_KERNELS_H_

// CUDA-platform implementation of IntegrateConstVLangevinStepKernel on top of the B200 C ABI
// (include/constv_b200.h).  Replaces the reference's class of the same name
// (platforms/cuda/src/CudaConstVKernels.h:46-79, CudaConstVKernels.cpp:53-177): instead of four
// runtime-compiled kernels it binds OpenMM's device arrays to a constv_slab and calls the
// precompiled sm_100a kernels.
//
// Targets the OpenMM the reference was written against (7.3 / 7.4: CudaArray::create<T>, CUfunction kernels,
// everything on the context's default stream).  Build with -DCONSTV_OPENMM_HAS_CURRENT_STREAM against
// OpenMM >= 7.5, whose CudaContext has getCurrentStream().

#include <memory>

#include "CudaArray.h"
#include "CudaContext.h"
#include "constv_b200.h"
#include "openmm/ConstVKernels.h"

namespace OpenMM {

// What both kernel classes share: the C-ABI slab bound to a CudaContext's arrays, kept consistent with
// OpenMM's atom re-sorting, step counter and virtual sites.
class ConstVSlabBinding {
public:
    explicit ConstVSlabBinding(CudaContext& cu) : cu(cu), slab(NULL), zmax(0), slotsPermuted(false), hasVirtualSites(false),
                                                  lastStepCount(-1), stepsSinceChargeRefresh(0), chargeRefreshInterval(64) {}
    ~ConstVSlabBinding();
    /** Validation + constv_create on OpenMM's current CUDA context (reference :76-96 / :184-205). */
    void create(const System& system, int numCells, double cellSize, int randomNumberSeed);
    /** First thing in execute(): follow OpenMM's slot order, step counter and image-charge snapshot. */
    void beforeStep(void* stream);
    /** After the step kernels: virtual sites, then the image rewrite the reference runs after them (:161-166, :342-347). */
    void finishStep(void* stream);
    void* stream();
    constv_slab* get() { return slab; }
    double getCellSize() const { return zmax; }
    bool getSlotsPermuted() const { return slotsPermuted; }
    bool getHasVirtualSites() const { return hasVirtualSites; }

private:
    struct OrderWatch { bool dirty = true; };  // shared with the ReorderListener OpenMM's context owns
    class Listener;
    void bindArrays();
    CudaContext& cu;
    constv_slab* slab;
    double zmax;
    bool slotsPermuted;      // the slab addresses slots through atomIndex / invAtomOrder
    bool hasVirtualSites;
    long long lastStepCount;  // cu.getStepCount() as this binding left it (the device Philox counter follows it)
    int stepsSinceChargeRefresh, chargeRefreshInterval;
    std::shared_ptr<OrderWatch> watch;
};

class CudaIntegrateConstVLangevinStepKernel : public IntegrateConstVLangevinStepKernel {
public:
    CudaIntegrateConstVLangevinStepKernel(std::string name, const Platform& platform, CudaContext& cu)
        : IntegrateConstVLangevinStepKernel(name, platform), cu(cu), binding(cu), hasConstraints(false),
          fusedConstraints(false), numRigidMolecules(0) {
    }
    ~CudaIntegrateConstVLangevinStepKernel();
    void initialize(const System& system, const ConstVLangevinIntegrator& integrator);
    void execute(ContextImpl& context, const ConstVLangevinIntegrator& integrator);
    double computeKineticEnergy(ContextImpl& context, const ConstVLangevinIntegrator& integrator);
    /** The C-ABI handle (tests and tools). */
    constv_slab* getSlab() { return binding.get(); }
    double getCellSize() const { return binding.getCellSize(); }
    /** True when every constraint of the System is part of a rigid 3-site molecule: the constraints are then
     *  solved inside the step kernel instead of being handed to integration.applyConstraints. */
    bool getConstraintsAreFused() const { return fusedConstraints; }
    int getNumRigidMolecules() const { return numRigidMolecules; }
    bool getSlotsPermuted() const { return binding.getSlotsPermuted(); }

private:
    CudaContext& cu;
    ConstVSlabBinding binding;
    bool hasConstraints;
    bool fusedConstraints;
    int numRigidMolecules;
};

// CUDA-platform implementation of IntegrateConstVDrudeLangevinStepKernel; replaces the reference's
// class of the same name (CudaConstVKernels.h:84-118, CudaConstVKernels.cpp:179-357): its four
// runtime-compiled kernels per step (part 1, part 2, hard wall, image rewrite) become one precompiled
// kernel, or two around the constraint solver.
class CudaIntegrateConstVDrudeLangevinStepKernel : public IntegrateConstVDrudeLangevinStepKernel {
public:
    CudaIntegrateConstVDrudeLangevinStepKernel(std::string name, const Platform& platform, CudaContext& cu)
        : IntegrateConstVDrudeLangevinStepKernel(name, platform), cu(cu), binding(cu), hasConstraints(false) {
    }
    ~CudaIntegrateConstVDrudeLangevinStepKernel();
    void initialize(const System& system, const ConstVDrudeLangevinIntegrator& integrator, const DrudeForce& force);
    void execute(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator);
    double computeKineticEnergy(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator);
    constv_slab* getSlab() { return binding.get(); }
    double getCellSize() const { return binding.getCellSize(); }
    bool getSlotsPermuted() const { return binding.getSlotsPermuted(); }

private:
    CudaContext& cu;
    ConstVSlabBinding binding;
    bool hasConstraints;
};

}  // namespace OpenMM

#endif /* CUDA_CONSTV_KERNELS_H_ */
