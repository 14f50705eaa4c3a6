#pragma once

// XML (de)serialization of ConstVLangevinIntegrator.  The node type is "ConstVLangevinIntegrator" and the
// version-1 property names are the reference's (serialization/src/ConstVLangevinIntegratorProxy.cpp:41-63), so
// files written by the reference load here AND files written here load in the reference: the proxy writes
// version 1 (the only one the reference's reader accepts, :55-56) and adds the two properties the reference
// silently drops, numCells and cellSize, as extra attributes -- the reference's reader looks properties up by
// name and ignores the rest; this reader takes them when present.  Version-2 files (what round 1 of this
// plugin wrote: the same attributes under version="2") are still read.

#include "openmm/serialization/SerializationNode.h"
#include "openmm/serialization/SerializationProxy.h"

namespace OpenMM {

class ConstVLangevinIntegratorProxy final : public SerializationProxy {
public:
    /** Version written, and the range of versions read. */
    static constexpr int kWritten = 1;
    static constexpr int kOldestReadable = 1;
    static constexpr int kNewestReadable = 2;

    ConstVLangevinIntegratorProxy();
    /** Returns a new ConstVLangevinIntegrator owned by the caller; throws "Unsupported version number". */
    void* deserialize(const SerializationNode& node) const override;
    void serialize(const void* object, SerializationNode& node) const override;
};

}  // namespace OpenMM
