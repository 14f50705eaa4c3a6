// Microbenchmark: what HBM bandwidth does the fused step's ACCESS PATTERN reach with no arithmetic?
// Same 11 address streams as constv_fused_step_kernel (single precision, numCells = 2):
//   read  velm[i] 16, posq[i] 16, fx[i] fy[i] fz[i] 8 each, q[i] 4
//   write velm[i] 16, posq[i] 16, posq[R+i] 16, velm[R+i] 16, fx/fy/fz[R+i] 8 each
// Variants differ in traversal order and cache operators.  Rotates over several replicas so that
// every pass streams from HBM (footprint > 2x L2).  Build: nvcc -O3 -arch=sm_100a stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Slab { float4* posq; float4* velm; long long* force; float* q; int R; int P; };

enum { LD_DEFAULT = 0, LD_CS = 1, LD_NC_NOALLOC = 2 };
enum { ST_DEFAULT = 0, ST_CS = 1, ST_WT = 2 };

template <int LD> __device__ __forceinline__ float4 ld4(const float4* p) {
    if (LD == LD_CS) return __ldcs(p);
    if (LD == LD_NC_NOALLOC) { float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
    return *p;
}
template <int LD> __device__ __forceinline__ long long ld8(const long long* p) {
    if (LD == LD_CS) return __ldcs(p);
    if (LD == LD_NC_NOALLOC) { long long r; asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p)); return r; }
    return *p;
}
template <int ST> __device__ __forceinline__ void st4(float4* p, float4 v) {
    if (ST == ST_CS) __stcs(p, v); else if (ST == ST_WT) __stwt(p, v); else *p = v;
}
template <int ST> __device__ __forceinline__ void st8(long long* p, long long v) {
    if (ST == ST_CS) __stcs(p, v); else if (ST == ST_WT) __stwt(p, v); else *p = v;
}

// mode bit0: do reads, bit1: do writes.  CONTIG: contiguous range per CTA instead of grid-stride.
// MODE bit2: read the image charge from posq[R+i].w (the line that is overwritten afterwards) instead
// of the side array; bit3: additionally touch (read) the image velm line before overwriting it.
template <int LD, int ST, bool CONTIG, int MODE>
__global__ void __launch_bounds__(256) pattern(Slab s) {
    const int R = s.R; const size_t P = s.P;
    int lo, hi, stride;
    if (CONTIG) {
        const long long units = ((long long) R + 31) >> 5;
        lo = (int) min((long long) R, (units * blockIdx.x / gridDim.x) << 5) + threadIdx.x;
        hi = (int) min((long long) R, (units * (blockIdx.x + 1) / gridDim.x) << 5);
        stride = blockDim.x;
    } else { lo = blockIdx.x * blockDim.x + threadIdx.x; hi = R; stride = gridDim.x * blockDim.x; }
    for (int i = lo; i < hi; i += stride) {
        float4 v = make_float4(1, 2, 3, 4), p = make_float4(5, 6, 7, 8); long long fx = 1, fy = 2, fz = 3; float q = 1;
        if (MODE & 1) {
            v = ld4<LD>(s.velm + i); p = ld4<LD>(s.posq + i);
            fx = ld8<LD>(s.force + i); fy = ld8<LD>(s.force + P + i); fz = ld8<LD>(s.force + 2 * P + i);
            if (MODE & 4) q = reinterpret_cast<const float*>(s.posq + R + i)[3]; else q = __ldg(s.q + i);
            if (MODE & 8) q += reinterpret_cast<const float*>(s.velm + R + i)[3];
        }
        if (MODE & 2) {
            v.x += (float) fx; v.y += (float) fy; v.z += (float) fz; p.x += v.x; p.y += v.y; p.z += v.z;
            st4<ST>(s.velm + i, v); st4<ST>(s.posq + i, p);
            st4<ST>(s.posq + R + i, make_float4(p.x, p.y, -p.z, q));
            st4<ST>(s.velm + R + i, make_float4(0, 0, 0, 0));
            st8<ST>(s.force + R + i, 0); st8<ST>(s.force + P + R + i, 0); st8<ST>(s.force + 2 * P + R + i, 0);
        } else if (v.x + p.x + (float) (fx + fy + fz) + q == 123.456f) s.q[i] = 0;  // keep the loads alive
    }
}

__global__ void copy4(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename F> float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) f(i);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f(i);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(int argc, char** argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 500000;
    const int P = 2 * R;
    const int NREP = (int) (700e6 / (116.0 * R)) + 2;
    std::vector<Slab> slabs(NREP);
    for (auto& s : slabs) {
        s.R = R; s.P = P;
        CK(cudaMalloc(&s.posq, 16 * (size_t) P)); CK(cudaMalloc(&s.velm, 16 * (size_t) P));
        CK(cudaMalloc(&s.force, 24 * (size_t) P)); CK(cudaMalloc(&s.q, 4 * (size_t) R));
        CK(cudaMemset(s.posq, 0, 16 * (size_t) P)); CK(cudaMemset(s.velm, 0, 16 * (size_t) P));
        CK(cudaMemset(s.force, 0, 24 * (size_t) P)); CK(cudaMemset(s.q, 0, 4 * (size_t) R));
    }
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double rd = 60.0 * R, wr = 88.0 * R;
    printf("R=%d replicas=%d SMs=%d  bytes/pass: read %.1f MB write %.1f MB\n", R, NREP, sms, rd / 1e6, wr / 1e6);
    {   // calibration: plain copy of 1 GiB
        size_t n = (size_t) 1 << 26; float4 *a, *b; CK(cudaMalloc(&a, 16 * n)); CK(cudaMalloc(&b, 16 * n));
        float ms = timeit([&](int) { copy4<<<sms * 16, 256>>>(a, b, n); }, 10);
        printf("%-44s %8.2f us  %7.1f GB/s\n", "copy float4 1GiB (calibration)", ms * 1e3, 2.0 * 16 * n / ms / 1e6);
        cudaFree(a); cudaFree(b);
    }
    const int reps = 400;
#define RUN(name, LD, ST, CONTIG, MODE, grid, bytes) { \
        float ms = timeit([&](int i) { pattern<LD, ST, CONTIG, MODE><<<grid, 256>>>(slabs[i % NREP]); }, reps); \
        printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, (bytes) / ms / 1e6); }
    const int gs = (R + 255) / 256;
    for (int c : {4, 8}) {
        const int g = sms * c < gs ? sms * c : gs;
        printf("-- grid = %d CTAs (%d per SM)\n", g, c);
        RUN("R+W grid-stride default", LD_DEFAULT, ST_DEFAULT, false, 3, g, rd + wr);
        RUN("R+W grid-stride ld.cs st.cs", LD_CS, ST_CS, false, 3, g, rd + wr);
        RUN("R+W grid-stride ld.nc.noalloc st.default", LD_NC_NOALLOC, ST_DEFAULT, false, 3, g, rd + wr);
        RUN("R+W grid-stride ld.cs st.wt", LD_CS, ST_WT, false, 3, g, rd + wr);
        RUN("R+W contiguous default", LD_DEFAULT, ST_DEFAULT, true, 3, g, rd + wr);
        RUN("R+W contiguous ld.cs st.cs", LD_CS, ST_CS, true, 3, g, rd + wr);
        RUN("R only grid-stride default", LD_DEFAULT, ST_DEFAULT, false, 1, g, rd);
        RUN("W only grid-stride default", LD_DEFAULT, ST_DEFAULT, false, 2, g, wr);
        RUN("W only grid-stride st.cs", LD_DEFAULT, ST_CS, false, 2, g, wr);
    }
    printf("-- grid = one CTA per 256 pairs (%d CTAs)\n", gs);
    RUN("R+W one-shot default", LD_DEFAULT, ST_DEFAULT, false, 3, gs, rd + wr);
    RUN("R+W one-shot ld.cs st.cs", LD_CS, ST_CS, false, 3, gs, rd + wr);
    RUN("R+W one-shot default, q from posq[R+i].w", LD_DEFAULT, ST_DEFAULT, false, 7, gs, rd + wr);
    RUN("R+W one-shot ld.cs st.cs, q from posq[R+i].w", LD_CS, ST_CS, false, 7, gs, rd + wr);
    RUN("R+W one-shot cs, q from posq + touch img velm", LD_CS, ST_CS, false, 15, gs, rd + wr);
    { const int g = sms * 8 < gs ? sms * 8 : gs;
      RUN("R+W persistent8 cs, q from posq[R+i].w", LD_CS, ST_CS, false, 7, g, rd + wr); }
    return 0;
}
