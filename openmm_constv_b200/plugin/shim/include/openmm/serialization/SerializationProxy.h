// Shim of OpenMM's SerializationProxy registry.
#ifndef OPENMM_SHIM_SERIALIZATIONPROXY_H_
#define OPENMM_SHIM_SERIALIZATIONPROXY_H_
#include <map>
#include <string>
#include <typeinfo>
#include "openmm/serialization/SerializationNode.h"
namespace OpenMM {
class OPENMM_EXPORT SerializationProxy {
public:
    explicit SerializationProxy(const std::string& typeName) : typeName(typeName) {}
    virtual ~SerializationProxy() {}
    const std::string& getTypeName() const { return typeName; }
    virtual void serialize(const void* object, SerializationNode& node) const = 0;
    virtual void* deserialize(const SerializationNode& node) const = 0;
    static void registerProxy(const std::type_info& type, const SerializationProxy* proxy);
    static const SerializationProxy& getProxy(const std::string& typeName);
    static const SerializationProxy& getProxy(const std::type_info& type);
private:
    std::string typeName;
};
}
#endif
