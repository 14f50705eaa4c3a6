#ifndef OPENMM_CUDACONSTVKERNELFACTORY_H_
#define OPENMM_CUDACONSTVKERNELFACTORY_H_

// Creates the B200 implementation of IntegrateConstVLangevinStepKernel for the CUDA platform.
// Same class name and exported registration symbols as the reference
// (platforms/cuda/include/CudaConstVKernelFactory.h:44-47, src/CudaConstVKernelFactory.cpp:38-70),
// so OpenMM's plugin loader picks this library up from lib/plugins unchanged.

#include "openmm/KernelFactory.h"

namespace OpenMM {

class CudaConstVKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const;
};

}  // namespace OpenMM

#endif /* OPENMM_CUDACONSTVKERNELFACTORY_H_ */
