// Drop-in surface of the plugin libraries, checked without a GPU: the exported registration
// symbols OpenMM's plugin loader and the reference's tests look up
// (platforms/cuda/src/CudaConstVKernelFactory.cpp:38-61,
// serialization/src/ConstVSerializationProxyRegistration.cpp:63-65), their behaviour on a machine
// with no CUDA platform, and the integrator's platform-independent defaults and errors
// (openmmapi/src/ConstVLangevinIntegrator.cpp:44-85).
#include <dlfcn.h>
#include <libgen.h>
#include <unistd.h>

#include <iostream>
#include <string>

#include "TestAssert.h"
#include "openmm/ConstVKernels.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"

using namespace OpenMM;

namespace {

class ExposedIntegrator : public ConstVLangevinIntegrator {
public:
    using ConstVLangevinIntegrator::ConstVLangevinIntegrator;
    std::vector<std::string> kernelNames() { return getKernelNames(); }
};

std::string directoryOfThisBinary() {
    char path[4096];
    ssize_t n = readlink("/proc/self/exe", path, sizeof path - 1);
    CHECK(n > 0);
    path[n] = 0;
    return dirname(path);
}

typedef void (*RegisterFn)();

RegisterFn lookup(void* handle, const char* name) {
    void* sym = dlsym(handle, name);
    if (sym == NULL)
        throw std::runtime_error(std::string("missing exported symbol ") + name);
    return reinterpret_cast<RegisterFn>(sym);
}

void testIntegratorDefaults() {
    ExposedIntegrator integ(300.0, 1.5, 0.002);
    CHECK_EQ(300.0, integ.getTemperature());
    CHECK_EQ(1.5, integ.getFriction());
    CHECK_EQ(0.002, integ.getStepSize());
    CHECK_EQ(2, integ.getNumCells());
    CHECK_EQ(-1.0, integ.getCellSize());
    CHECK_EQ(1e-5, integ.getConstraintTolerance());
    CHECK_EQ(0, integ.getRandomNumberSeed());
    integ.setNumCells(4);
    integ.setCellSize(2.5);
    integ.setTemperature(10);
    integ.setFriction(0);
    integ.setRandomNumberSeed(-3);
    CHECK_EQ(4, integ.getNumCells());
    CHECK_EQ(2.5, integ.getCellSize());
    CHECK_EQ(10.0, integ.getTemperature());
    CHECK_EQ(0.0, integ.getFriction());
    CHECK_EQ(-3, integ.getRandomNumberSeed());
    CHECK_EQ((size_t) 1, integ.kernelNames().size());
    CHECK_EQ(std::string("IntegrateConstVLangevinStep"), integ.kernelNames()[0]);
    CHECK_EQ(std::string("IntegrateConstVLangevinStep"), IntegrateConstVLangevinStepKernel::Name());
    CHECK_THROWS(integ.step(1), "This Integrator is not bound to a context!");
}

void testExportedSymbols() {
    const std::string dir = directoryOfThisBinary();
    void* plugin = dlopen((dir + "/../lib/plugins/libConstVPluginCUDA.so").c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (plugin == NULL)
        throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
    RegisterFn registerPlatforms = lookup(plugin, "registerPlatforms");
    RegisterFn registerKernelFactories = lookup(plugin, "registerKernelFactories");
    RegisterFn registerConstVCuda = lookup(plugin, "registerConstVCudaKernelFactories");
    void* api = dlopen((dir + "/../lib/libOpenMMConstV.so").c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (api == NULL)
        throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
    lookup(api, "registerConstVSerializationProxies")();

    // the plugin loader's sequence on a machine without a CUDA platform: must not throw
    const int before = Platform::getNumPlatforms();
    registerPlatforms();
    CHECK_EQ(before, Platform::getNumPlatforms());  // this plugin adds no platform
    bool hadCuda = true;
    try {
        Platform::getPlatformByName("CUDA");
    } catch (const std::exception&) {
        hadCuda = false;
    }
    registerKernelFactories();  // swallowed when there is no CUDA platform
    // the direct-link entry point creates the CUDA platform if needed, then registers the factory
    registerConstVCuda();
    Platform& cuda = Platform::getPlatformByName("CUDA");
    CHECK_EQ(std::string("CUDA"), cuda.getName());
    if (!hadCuda)
        CHECK_EQ(before + 1, Platform::getNumPlatforms());
}

}  // namespace

int main() {
    try {
        testIntegratorDefaults();
        testExportedSymbols();
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "Done" << std::endl;
    return 0;
}
