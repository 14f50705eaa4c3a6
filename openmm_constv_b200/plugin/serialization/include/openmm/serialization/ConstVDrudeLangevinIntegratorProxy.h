#pragma once

// XML (de)serialization of ConstVDrudeLangevinIntegrator.  The reference ships no proxy for this integrator
// (serialization/src holds ConstVLangevinIntegratorProxy only), so a Context that uses it cannot be written to
// XML there; this one follows the Langevin proxy's conventions: node type = class name, an integer version,
// every constructor argument and setter as a property.

#include "openmm/serialization/SerializationNode.h"
#include "openmm/serialization/SerializationProxy.h"

namespace OpenMM {

class ConstVDrudeLangevinIntegratorProxy final : public SerializationProxy {
public:
    static constexpr int kWritten = 1;

    ConstVDrudeLangevinIntegratorProxy();
    /** Returns a new ConstVDrudeLangevinIntegrator owned by the caller; throws "Unsupported version number". */
    void* deserialize(const SerializationNode& node) const override;
    void serialize(const void* object, SerializationNode& node) const override;
};

}  // namespace OpenMM
