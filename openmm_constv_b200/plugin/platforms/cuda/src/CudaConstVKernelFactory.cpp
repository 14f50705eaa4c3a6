#include "CudaConstVKernelFactory.h"

#include <exception>

#include "CudaConstVKernels.h"
#include "CudaPlatform.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/windowsExport.h"

using namespace OpenMM;

// Called by OpenMM's plugin loader right after dlopen; this plugin adds no platform of its own.
extern "C" OPENMM_EXPORT void registerPlatforms() {
}

// Called by the plugin loader after registerPlatforms(): hook the kernel name onto the CUDA
// platform.  A machine without the CUDA platform is not an error for a plugin, so it is ignored.
extern "C" OPENMM_EXPORT void registerKernelFactories() {
    try {
        Platform& cuda = Platform::getPlatformByName("CUDA");
        CudaConstVKernelFactory* factory = new CudaConstVKernelFactory();  // one factory, both kernel names (reference :44-46)
        cuda.registerKernelFactory(IntegrateConstVLangevinStepKernel::Name(), factory);
        cuda.registerKernelFactory(IntegrateConstVDrudeLangevinStepKernel::Name(), factory);
    } catch (const std::exception&) {
    }
}

// For programs that link the plugin directly (the tests): make sure a CUDA platform exists first.
extern "C" OPENMM_EXPORT void registerConstVCudaKernelFactories() {
    try {
        Platform::getPlatformByName("CUDA");
    } catch (...) {
        Platform::registerPlatform(new CudaPlatform());
    }
    registerKernelFactories();
}

KernelImpl* CudaConstVKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    CudaPlatform::PlatformData& data = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData());
    CudaContext& cu = *data.contexts[0];  // the integrator runs on the first device, as in the reference
    if (name == IntegrateConstVLangevinStepKernel::Name())
        return new CudaIntegrateConstVLangevinStepKernel(name, platform, cu);
    if (name == IntegrateConstVDrudeLangevinStepKernel::Name())
        return new CudaIntegrateConstVDrudeLangevinStepKernel(name, platform, cu);
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
}
