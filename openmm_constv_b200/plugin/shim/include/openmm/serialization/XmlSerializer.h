// Shim of OpenMM's XmlSerializer: <root type="..." prop="..."> <child .../> </root>.
#ifndef OPENMM_SHIM_XMLSERIALIZER_H_
#define OPENMM_SHIM_XMLSERIALIZER_H_
#include <iosfwd>
#include <typeinfo>
#include "openmm/serialization/SerializationProxy.h"
namespace OpenMM {
class OPENMM_EXPORT XmlSerializer {
public:
    template <class T> static void serialize(const T* object, const std::string& rootName, std::ostream& stream) {
        const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(*object));
        SerializationNode node;
        node.setName(rootName);
        proxy.serialize(object, node);
        node.setStringProperty("type", proxy.getTypeName());
        write(node, stream);
    }
    template <class T> static T* deserialize(std::istream& stream) {
        SerializationNode node;
        read(stream, node);
        const SerializationProxy& proxy = SerializationProxy::getProxy(node.getStringProperty("type"));
        return reinterpret_cast<T*>(proxy.deserialize(node));
    }
    static void write(const SerializationNode& node, std::ostream& stream, int depth = 0);
    static void read(std::istream& stream, SerializationNode& node);
};
}
#endif
