// Shim of OpenMM's SerializationNode: a named node with string-typed properties and children.
#ifndef OPENMM_SHIM_SERIALIZATIONNODE_H_
#define OPENMM_SHIM_SERIALIZATIONNODE_H_
#include <map>
#include <string>
#include <vector>
#include "openmm/ShimCore.h"
namespace OpenMM {
class OPENMM_EXPORT SerializationNode {
public:
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    const std::map<std::string, std::string>& getProperties() const { return properties; }
    bool hasProperty(const std::string& key) const { return properties.count(key) != 0; }
    const std::string& getStringProperty(const std::string& key) const;
    const std::string& getStringProperty(const std::string& key, const std::string& defaultValue) const;
    SerializationNode& setStringProperty(const std::string& key, const std::string& value);
    int getIntProperty(const std::string& key) const;
    int getIntProperty(const std::string& key, int defaultValue) const;
    SerializationNode& setIntProperty(const std::string& key, int value);
    double getDoubleProperty(const std::string& key) const;
    double getDoubleProperty(const std::string& key, double defaultValue) const;
    SerializationNode& setDoubleProperty(const std::string& key, double value);
    const std::vector<SerializationNode>& getChildren() const { return children; }
    std::vector<SerializationNode>& getChildren() { return children; }
    SerializationNode& createChildNode(const std::string& childName);
private:
    std::string name;
    std::map<std::string, std::string> properties;
    std::vector<SerializationNode> children;
};
}
#endif
