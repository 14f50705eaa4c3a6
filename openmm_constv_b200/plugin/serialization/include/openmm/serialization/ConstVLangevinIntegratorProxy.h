#pragma once

// XML (de)serialization of ConstVLangevinIntegrator.  The node type is "ConstVLangevinIntegrator" and the
// version-1 property names are the reference's (serialization/src/ConstVLangevinIntegratorProxy.cpp:41-63), so
// files written by the reference load here.  Version 2, which this proxy writes, adds the two properties the
// reference's version 1 silently dropped: numCells and cellSize.

#include "openmm/serialization/SerializationNode.h"
#include "openmm/serialization/SerializationProxy.h"

namespace OpenMM {

class ConstVLangevinIntegratorProxy final : public SerializationProxy {
public:
    /** Version written; files of every version from kOldestReadable on are read. */
    static constexpr int kWritten = 2;
    static constexpr int kOldestReadable = 1;

    ConstVLangevinIntegratorProxy();
    /** Returns a new ConstVLangevinIntegrator owned by the caller; throws "Unsupported version number". */
    void* deserialize(const SerializationNode& node) const override;
    void serialize(const void* object, SerializationNode& node) const override;
};

}  // namespace OpenMM
