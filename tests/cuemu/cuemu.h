// cuemu.h — single-threaded, fiber-based emulation of the small CUDA subset used by
// medical-vision-textural-bias_b200/mvtb/csrc/*.cu.
//
// DEBUG SCAFFOLDING ONLY.  The build container has nvcc but no GPU; this header lets the
// *same kernel sources* be compiled by g++ (-DMVTB_EMU) into tests/cuemu/_build/libmvtb_emu.so
// so that index arithmetic, barrier placement and the C-ABI argument handling can be
// debugged on tiny shapes before GPU time is spent.  It is never loaded by the product
// package (which loads only the nvcc-built libmvtb.so and refuses to run without CUDA);
// only tests/test_emu_*.py load it.
//
// Model: a launch runs blocks one after another; the threads of a block are ucontext
// fibers; __syncthreads() and the warp primitives yield to a scheduler that releases a
// barrier when every live thread of the block (or warp) has arrived.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace cuemu {
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void block_barrier();
void warp_barrier();
int lane_id();
int warp_id();
void* warp_slot();   // 32 x 16-byte exchange slots of the calling thread's warp
}  // namespace cuemu

#define threadIdx (cuemu::g_threadIdx)
#define blockIdx (cuemu::g_blockIdx)
#define blockDim (cuemu::g_blockDim)
#define gridDim (cuemu::g_gridDim)
static const int warpSize = 32;

static inline void __threadfence() {}
static inline void __nanosleep(unsigned) {}
static inline void __syncthreads() { cuemu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::warp_barrier(); }

template <typename T>
static inline T cuemu_shfl(T v, int src) {
    static_assert(sizeof(T) <= 16, "shuffle payload too large");
    char* slots = (char*)cuemu::warp_slot();
    std::memcpy(slots + 16 * cuemu::lane_id(), &v, sizeof(T));
    cuemu::warp_barrier();
    T r;
    std::memcpy(&r, slots + 16 * (src & 31), sizeof(T));
    cuemu::warp_barrier();
    return r;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int l = cuemu::lane_id();
    return cuemu_shfl(v, (l / width) * width + (src % width));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    return cuemu_shfl(v, cuemu::lane_id() ^ m);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int l = cuemu::lane_id();
    int s = l - (int)d;
    if (s < 0 || (s / width) != (l / width)) s = l;
    return cuemu_shfl(v, s);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int l = cuemu::lane_id();
    int s = l + (int)d;
    if ((s / width) != (l / width)) s = l;
    return cuemu_shfl(v, s);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (cuemu_shfl(pred ? 1 : 0, i) ? 1u : 0u) << i;
    return r;
}

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
static inline float expf_(float x) { return std::exp(x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __int2float_rn(int i) { return (float)i; }
static inline float __uint2float_rn(unsigned i) { return (float)i; }

template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; *p = std::min(o, v); return o; }
static inline int atomicMax(int* p, int v) { int o = *p; *p = std::max(o, v); return o; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; *p = std::min(o, v); return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; *p = std::max(o, v); return o; }

// ---- runtime API subset (device memory == host memory)
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cuemu"; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
typedef void* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; p->minor = 0; return cudaSuccess; }
