// Registers the XML proxies when libOpenMMConstV is loaded.  The reference declares the constructor
// attribute on a misspelt name (serialization/src/ConstVSerializationProxyRegistration.cpp:58 vs
// :63), so its proxy never auto-registers on Linux; here the attribute sits on the function that
// exists.  The exported symbol keeps the reference's name.
#include <typeinfo>

#include "openmm/ConstVDrudeLangevinIntegrator.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/serialization/ConstVDrudeLangevinIntegratorProxy.h"
#include "openmm/serialization/ConstVLangevinIntegratorProxy.h"
#include "openmm/serialization/SerializationProxy.h"

using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerConstVSerializationProxies();

extern "C" void __attribute__((constructor)) constvRegisterSerializationProxiesOnLoad() {
    registerConstVSerializationProxies();
}

extern "C" OPENMM_EXPORT void registerConstVSerializationProxies() {
    static bool done = false;
    if (done) return;
    done = true;
    SerializationProxy::registerProxy(typeid(ConstVLangevinIntegrator), new ConstVLangevinIntegratorProxy());
    SerializationProxy::registerProxy(typeid(ConstVDrudeLangevinIntegrator), new ConstVDrudeLangevinIntegratorProxy());
}
