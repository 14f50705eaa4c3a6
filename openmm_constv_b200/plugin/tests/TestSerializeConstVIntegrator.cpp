// XML round trip of ConstVLangevinIntegrator (CPU only).  Pins what the reference's
// serialization/tests/TestSerializeConstVIntegrator.cpp:47-59 pins (constraintTolerance, stepSize,
// temperature, friction, seed survive), plus what the reference loses: numCells and cellSize
// (version 2), reading a version-1 document written by the reference, rejecting other versions
// with the reference's message (ConstVLangevinIntegratorProxy.cpp:56-57), and that the proxy is
// registered by merely loading libOpenMMConstV (the reference's constructor attribute sits on a
// misspelt symbol, ConstVSerializationProxyRegistration.cpp:58).
#include <iostream>
#include <sstream>

#include "TestAssert.h"
#include "openmm/ConstVDrudeLangevinIntegrator.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/serialization/XmlSerializer.h"

using namespace OpenMM;

extern "C" void registerConstVSerializationProxies();

static void testRoundTrip() {
    ConstVLangevinIntegrator original(301.25, 0.95, 0.0015, 4, 4.2183);
    original.setConstraintTolerance(2.5e-6);
    original.setRandomNumberSeed(10);
    std::stringstream buffer;
    XmlSerializer::serialize<ConstVLangevinIntegrator>(&original, "Integrator", buffer);
    ConstVLangevinIntegrator* copy = XmlSerializer::deserialize<ConstVLangevinIntegrator>(buffer);
    CHECK_EQ(original.getConstraintTolerance(), copy->getConstraintTolerance());
    CHECK_EQ(original.getStepSize(), copy->getStepSize());
    CHECK_EQ(original.getTemperature(), copy->getTemperature());
    CHECK_EQ(original.getFriction(), copy->getFriction());
    CHECK_EQ(original.getRandomNumberSeed(), copy->getRandomNumberSeed());
    CHECK_EQ(4, copy->getNumCells());
    CHECK_EQ(4.2183, copy->getCellSize());
    delete copy;
}

static void testReadsVersion1() {
    // what the reference writes: no numCells / cellSize
    std::stringstream v1(
        "<?xml version=\"1.0\" ?>\n"
        "<Integrator type=\"ConstVLangevinIntegrator\" version=\"1\" stepSize=\".002\" constraintTolerance=\"1e-05\" "
        "temperature=\"300\" friction=\"1\" randomSeed=\"7\"/>\n");
    ConstVLangevinIntegrator* integ = XmlSerializer::deserialize<ConstVLangevinIntegrator>(v1);
    CHECK_EQ(0.002, integ->getStepSize());
    CHECK_EQ(1e-5, integ->getConstraintTolerance());
    CHECK_EQ(300.0, integ->getTemperature());
    CHECK_EQ(1.0, integ->getFriction());
    CHECK_EQ(7, integ->getRandomNumberSeed());
    CHECK_EQ(2, integ->getNumCells());      // the reference's defaults after a round trip
    CHECK_EQ(-1.0, integ->getCellSize());
    delete integ;
}

static void testReadsVersion2() {
    // what round 1 of this plugin wrote
    std::stringstream v2(
        "<Integrator type=\"ConstVLangevinIntegrator\" version=\"2\" stepSize=\".002\" constraintTolerance=\"1e-05\" "
        "temperature=\"300\" friction=\"1\" randomSeed=\"7\" numCells=\"4\" cellSize=\"2.5\"/>\n");
    ConstVLangevinIntegrator* integ = XmlSerializer::deserialize<ConstVLangevinIntegrator>(v2);
    CHECK_EQ(4, integ->getNumCells());
    CHECK_EQ(2.5, integ->getCellSize());
    delete integ;
}

static void testRejectsUnknownVersion() {
    std::stringstream v3(
        "<Integrator type=\"ConstVLangevinIntegrator\" version=\"3\" stepSize=\".002\" constraintTolerance=\"1e-05\" "
        "temperature=\"300\" friction=\"1\" randomSeed=\"7\"/>\n");
    CHECK_THROWS(XmlSerializer::deserialize<ConstVLangevinIntegrator>(v3), "Unsupported version number");
}

static void testWrittenDocument() {
    ConstVLangevinIntegrator integ(300, 1, 0.001);
    std::stringstream buffer;
    XmlSerializer::serialize<ConstVLangevinIntegrator>(&integ, "Integrator", buffer);
    const std::string text = buffer.str();
    CHECK(text.find("type=\"ConstVLangevinIntegrator\"") != std::string::npos);
    // version 1: the only one the reference's reader accepts (its ConstVLangevinIntegratorProxy.cpp:55-56), so a
    // document written here loads in the reference plugin; numCells / cellSize ride along as extra attributes
    CHECK(text.find("version=\"1\"") != std::string::npos);
    for (const char* key : {"stepSize=", "constraintTolerance=", "temperature=", "friction=", "randomSeed=", "numCells=", "cellSize="})
        CHECK(text.find(key) != std::string::npos);
}

// The reference has no proxy for the Drude integrator; this one carries every constructor argument and setter.
static void testDrudeRoundTrip() {
    ConstVDrudeLangevinIntegrator original(299.5, 4.5, 1.25, 19.0, 0.00075, 4, 3.5);
    original.setMaxDrudeDistance(0.021);
    original.setConstraintTolerance(3.5e-6);
    original.setRandomNumberSeed(-12);
    std::stringstream buffer;
    XmlSerializer::serialize<ConstVDrudeLangevinIntegrator>(&original, "Integrator", buffer);
    CHECK(buffer.str().find("type=\"ConstVDrudeLangevinIntegrator\"") != std::string::npos);
    ConstVDrudeLangevinIntegrator* copy = XmlSerializer::deserialize<ConstVDrudeLangevinIntegrator>(buffer);
    CHECK_EQ(original.getTemperature(), copy->getTemperature());
    CHECK_EQ(original.getFriction(), copy->getFriction());
    CHECK_EQ(original.getDrudeTemperature(), copy->getDrudeTemperature());
    CHECK_EQ(original.getDrudeFriction(), copy->getDrudeFriction());
    CHECK_EQ(original.getMaxDrudeDistance(), copy->getMaxDrudeDistance());
    CHECK_EQ(original.getStepSize(), copy->getStepSize());
    CHECK_EQ(original.getConstraintTolerance(), copy->getConstraintTolerance());
    CHECK_EQ(original.getRandomNumberSeed(), copy->getRandomNumberSeed());
    CHECK_EQ(4, copy->getNumCells());
    CHECK_EQ(3.5, copy->getCellSize());
    delete copy;
    std::stringstream v2(
        "<Integrator type=\"ConstVDrudeLangevinIntegrator\" version=\"2\" stepSize=\".001\" constraintTolerance=\"1e-05\" "
        "temperature=\"300\" friction=\"5\" drudeTemperature=\"1\" drudeFriction=\"20\" maxDrudeDistance=\"0\" randomSeed=\"0\" "
        "numCells=\"2\" cellSize=\"-1\"/>\n");
    CHECK_THROWS(XmlSerializer::deserialize<ConstVDrudeLangevinIntegrator>(v2), "Unsupported version number");
}

int main() {
    try {
        // no explicit registration first: loading the library must have registered the proxy
        testRoundTrip();
        registerConstVSerializationProxies();  // and calling it again is harmless
        testRoundTrip();
        testReadsVersion1();
        testReadsVersion2();
        testRejectsUnknownVersion();
        testWrittenDocument();
        testDrudeRoundTrip();
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "Done" << std::endl;
    return 0;
}
