// Shim of the slice of OpenMM's CudaIntegrationUtilities the ConstV kernel calls
// (CudaConstVKernels.cpp:64,112,146-161,176).
#ifndef OPENMM_SHIM_CUDAINTEGRATIONUTILITIES_H_
#define OPENMM_SHIM_CUDAINTEGRATIONUTILITIES_H_
#include "CudaArray.h"
namespace OpenMM {
class OPENMM_EXPORT CudaIntegrationUtilities {
public:
    CudaIntegrationUtilities(CudaContext& context, const System& system);
    ~CudaIntegrationUtilities();
    CudaArray& getPosDelta() { return *posDelta; }
    CudaArray& getRandom() { return *random; }
    CudaArray& getStepSize() { return *stepSize; }
    void setNextStepSize(double size);
    double getLastStepSize() const { return lastStepSize; }
    // The shim has no SHAKE/SETTLE/CCMA: applyConstraints only counts calls (systems under test
    // carry no constraints); a hook lets tests observe the hand-off.
    void applyConstraints(double tol);
    // two-particle-average sites only (host round trip: the shim is test scaffolding)
    void computeVirtualSites();
    int virtualSiteCalls = 0;
    int hasVirtualSites = -1;
    void initRandomNumberGenerator(unsigned int seed) { randomSeed = seed; }
    int prepareRandomNumbers(int numValues);
    double computeKineticEnergy(double timeShift);
    int constraintCalls = 0;
    double lastConstraintTol = 0;
private:
    CudaContext& context;
    const System& system;
    CudaArray* posDelta;
    CudaArray* random;
    CudaArray* stepSize;
    double lastStepSize;
    unsigned int randomSeed;
};
}
#endif
