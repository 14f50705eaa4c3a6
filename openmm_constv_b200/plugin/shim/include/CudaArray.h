// Shim of OpenMM's CudaArray: a typed device allocation exposing a CUdeviceptr.
#ifndef OPENMM_SHIM_CUDAARRAY_H_
#define OPENMM_SHIM_CUDAARRAY_H_
#include <cuda.h>
#include <string>
#include <vector>
#include "openmm/ShimCore.h"
namespace OpenMM {
class CudaContext;
class OPENMM_EXPORT CudaArray {
public:
    CudaArray(CudaContext& context, int size, int elementSize, const std::string& name);
    ~CudaArray();
    template <class T> static CudaArray* create(CudaContext& context, int size, const std::string& name) {
        return new CudaArray(context, size, sizeof(T), name);
    }
    int getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    CUdeviceptr& getDevicePointer() { return pointer; }
    void uploadBytes(const void* data);
    void downloadBytes(void* data) const;
    template <class T> void upload(const std::vector<T>& data) { uploadBytes(data.data()); }
    template <class T> void download(std::vector<T>& data) const { data.resize(size); downloadBytes(data.data()); }
private:
    CUdeviceptr pointer;
    int size, elementSize;
    std::string name;
};
}
#endif
