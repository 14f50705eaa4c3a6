// Shim of the slice of OpenMM's CudaContext the ConstV kernel calls.
#ifndef OPENMM_SHIM_CUDACONTEXT_H_
#define OPENMM_SHIM_CUDACONTEXT_H_
#include "CudaArray.h"
#include "CudaIntegrationUtilities.h"
#include "CudaPlatform.h"
namespace OpenMM {
class OPENMM_EXPORT CudaContext {
public:
    CudaContext(const System& system, int deviceIndex, const std::string& precision, CudaPlatform::PlatformData& platformData);
    ~CudaContext();
    // OpenMM's hook for code that caches slot-dependent state: execute() runs after every re-sort.  The context
    // owns (and deletes) the listeners, as OpenMM's does.
    class ReorderListener {
    public:
        virtual void execute() = 0;
        virtual ~ReorderListener() {}
    };
    void addReorderListener(ReorderListener* listener) { reorderListeners.push_back(listener); }
    void setAsCurrent();
    CudaPlatform::PlatformData& getPlatformData() { return platformData; }
    int getDeviceIndex() const { return deviceIndex; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return useMixed; }
    CudaArray& getPosq() { return *posq; }
    CudaArray& getPosqCorrection() { return *posqCorrection; }
    CudaArray& getVelm() { return *velm; }
    CudaArray& getForce() { return *force; }
    CudaArray& getAtomIndexArray() { return *atomIndex; }
    const std::vector<int>& getAtomIndex() const { return atomIndexHost; }
    bool getAtomsWereReordered() const { return atomsWereReordered; }
    // OpenMM re-sorts atoms spatially every few hundred steps; the shim applies a caller-provided
    // permutation once (setPendingReorder) so that the invAtomOrder path is exercised.
    void reorderAtoms();
    void setPendingReorder(const std::vector<int>& newAtomIndex) { pendingOrder = newAtomIndex; }
    void clearReorderedFlag() { atomsWereReordered = false; }
    double getTime() const { return time; }
    void setTime(double t) { time = t; }
    int getStepCount() const { return stepCount; }
    void setStepCount(int c) { stepCount = c; }
    CUstream getCurrentStream() { return 0; }
    CudaIntegrationUtilities& getIntegrationUtilities() { return *integration; }
private:
    CudaPlatform::PlatformData& platformData;
    int deviceIndex, numAtoms, paddedNumAtoms, stepCount;
    bool useDouble, useMixed, atomsWereReordered;
    double time;
    CudaArray *posq, *posqCorrection, *velm, *force, *atomIndex;
    std::vector<int> atomIndexHost, pendingOrder;
    std::vector<ReorderListener*> reorderListeners;
    CudaIntegrationUtilities* integration;
};
}
#endif
