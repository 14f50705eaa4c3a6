#include "openmm/serialization/ConstVLangevinIntegratorProxy.h"

#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/OpenMMException.h"
#include "openmm/serialization/SerializationNode.h"

using namespace OpenMM;

ConstVLangevinIntegratorProxy::ConstVLangevinIntegratorProxy() : SerializationProxy("ConstVLangevinIntegrator") {
}

void ConstVLangevinIntegratorProxy::serialize(const void* object, SerializationNode& node) const {
    const ConstVLangevinIntegrator& integrator = *reinterpret_cast<const ConstVLangevinIntegrator*>(object);
    node.setIntProperty("version", kWritten);
    // version 1 properties (same names as the reference)
    node.setDoubleProperty("stepSize", integrator.getStepSize());
    node.setDoubleProperty("constraintTolerance", integrator.getConstraintTolerance());
    node.setDoubleProperty("temperature", integrator.getTemperature());
    node.setDoubleProperty("friction", integrator.getFriction());
    node.setIntProperty("randomSeed", integrator.getRandomNumberSeed());
    // what the reference drops: extra attributes its reader ignores (it rejects any version but 1)
    node.setIntProperty("numCells", integrator.getNumCells());
    node.setDoubleProperty("cellSize", integrator.getCellSize());
}

void* ConstVLangevinIntegratorProxy::deserialize(const SerializationNode& node) const {
    const int version = node.getIntProperty("version");
    if (version < kOldestReadable || version > kNewestReadable)
        throw OpenMMException("Unsupported version number");
    // a file written by the reference carries neither numCells nor cellSize: the reference's defaults apply
    ConstVLangevinIntegrator* integrator = new ConstVLangevinIntegrator(
        node.getDoubleProperty("temperature"), node.getDoubleProperty("friction"), node.getDoubleProperty("stepSize"),
        node.getIntProperty("numCells", 2), node.getDoubleProperty("cellSize", -1.0));
    integrator->setConstraintTolerance(node.getDoubleProperty("constraintTolerance"));
    integrator->setRandomNumberSeed(node.getIntProperty("randomSeed"));
    return integrator;
}
