#include "openmm/serialization/ConstVDrudeLangevinIntegratorProxy.h"

#include "openmm/ConstVDrudeLangevinIntegrator.h"
#include "openmm/OpenMMException.h"

using namespace OpenMM;

ConstVDrudeLangevinIntegratorProxy::ConstVDrudeLangevinIntegratorProxy() : SerializationProxy("ConstVDrudeLangevinIntegrator") {
}

void ConstVDrudeLangevinIntegratorProxy::serialize(const void* object, SerializationNode& node) const {
    const ConstVDrudeLangevinIntegrator& integrator = *reinterpret_cast<const ConstVDrudeLangevinIntegrator*>(object);
    node.setIntProperty("version", kWritten);
    node.setDoubleProperty("stepSize", integrator.getStepSize());
    node.setDoubleProperty("constraintTolerance", integrator.getConstraintTolerance());
    node.setDoubleProperty("temperature", integrator.getTemperature());
    node.setDoubleProperty("friction", integrator.getFriction());
    node.setDoubleProperty("drudeTemperature", integrator.getDrudeTemperature());
    node.setDoubleProperty("drudeFriction", integrator.getDrudeFriction());
    node.setDoubleProperty("maxDrudeDistance", integrator.getMaxDrudeDistance());
    node.setIntProperty("randomSeed", integrator.getRandomNumberSeed());
    node.setIntProperty("numCells", integrator.getNumCells());
    node.setDoubleProperty("cellSize", integrator.getCellSize());
}

void* ConstVDrudeLangevinIntegratorProxy::deserialize(const SerializationNode& node) const {
    if (node.getIntProperty("version") != kWritten)
        throw OpenMMException("Unsupported version number");
    ConstVDrudeLangevinIntegrator* integrator = new ConstVDrudeLangevinIntegrator(
        node.getDoubleProperty("temperature"), node.getDoubleProperty("friction"), node.getDoubleProperty("drudeTemperature"),
        node.getDoubleProperty("drudeFriction"), node.getDoubleProperty("stepSize"), node.getIntProperty("numCells"),
        node.getDoubleProperty("cellSize"));
    try {
        integrator->setMaxDrudeDistance(node.getDoubleProperty("maxDrudeDistance"));
    } catch (...) {
        delete integrator;
        throw;
    }
    integrator->setConstraintTolerance(node.getDoubleProperty("constraintTolerance"));
    integrator->setRandomNumberSeed(node.getIntProperty("randomSeed"));
    return integrator;
}
