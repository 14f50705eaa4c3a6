#ifndef OPENMM_CONSTVLANGEVIN_INTEGRATOR_PROXY_H_
#define OPENMM_CONSTVLANGEVIN_INTEGRATOR_PROXY_H_

// XML proxy for ConstVLangevinIntegrator.  Node type "ConstVLangevinIntegrator" as in the reference
// (serialization/src/ConstVLangevinIntegratorProxy.cpp:41-63).  Writes version 2, which adds the
// numCells and cellSize properties the reference's version 1 silently dropped; reads both.

#include "openmm/serialization/XmlSerializer.h"

namespace OpenMM {

class ConstVLangevinIntegratorProxy : public SerializationProxy {
public:
    ConstVLangevinIntegratorProxy();
    void serialize(const void* object, SerializationNode& node) const;
    void* deserialize(const SerializationNode& node) const;
};

}  // namespace OpenMM

#endif /* OPENMM_CONSTVLANGEVIN_INTEGRATOR_PROXY_H_ */
