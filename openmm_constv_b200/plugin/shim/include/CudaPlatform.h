// Shim of OpenMM's CudaPlatform and its PlatformData (CudaConstVKernelFactory.cpp:53-64).
#ifndef OPENMM_SHIM_CUDAPLATFORM_H_
#define OPENMM_SHIM_CUDAPLATFORM_H_
#include "openmm/ShimCore.h"
namespace OpenMM {
class CudaContext;
class OPENMM_EXPORT CudaPlatform : public Platform {
public:
    class PlatformData;
    CudaPlatform();
    const std::string& getName() const { static const std::string name = "CUDA"; return name; }
    static const std::string& CudaPrecision() { static const std::string key = "Precision"; return key; }
    static const std::string& CudaDeviceIndex() { static const std::string key = "DeviceIndex"; return key; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const;
    void contextDestroyed(ContextImpl& context) const;
};
class OPENMM_EXPORT CudaPlatform::PlatformData {
public:
    PlatformData(ContextImpl* context, const System& system, const std::string& deviceIndex, const std::string& precision);
    ~PlatformData();
    void initializeContexts(const System& system) {}
    ContextImpl* context;
    std::vector<CudaContext*> contexts;
    bool contextsInitialized;
};
}
#endif
