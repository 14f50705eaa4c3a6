// A read-only sweep for BenchPluginPath: reads `bytes` of device memory and keeps a checksum, so that the L2 ends
// up full of CLEAN lines of something else (a memset would leave 126 MB of dirty lines whose write-back the
// next kernel pays for).
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void benchSweepReadKernel(const uint4* __restrict__ p, size_t n, unsigned int* __restrict__ sink) {
    unsigned int acc = 0;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const uint4 v = p[i];
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9e3779b9u) *sink = acc;  // never true for a zero-filled buffer; keeps the loads alive
}

// A copy KERNEL (not a memcpy node): what sits in front of execute() in a real Context is a force kernel, and the
// step kernel's programmatic early launch only engages behind a kernel.
__global__ void benchSweepCopyKernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" void benchSweepCopy(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    benchSweepCopyKernel<<<148 * 8, 256, 0, stream>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), bytes / 16);
}

extern "C" void benchSweepRead(const void* p, size_t bytes, unsigned int* sink, cudaStream_t stream) {
    benchSweepReadKernel<<<148 * 8, 256, 0, stream>>>(static_cast<const uint4*>(p), bytes / 16, sink);
}
