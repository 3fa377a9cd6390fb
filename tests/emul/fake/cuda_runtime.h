// TEST INFRASTRUCTURE: minimal host stand-in for the CUDA runtime so that the product's kernel
// sources (myochallenge_b200/csrc/*.cu, *.cuh) compile with g++ and run single-lane on the CPU.
// Used only by tests/emul (CPU-side debugging of the kernel logic); never part of the product library.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__

struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
struct dim3e { unsigned x = 1, y = 1, z = 1; };
extern thread_local dim3e blockIdx, threadIdx, blockDim, gridDim;
extern thread_local float4* emul_smem;

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaFuncAttributes { int numRegs = 0; };
struct cudaDeviceProp { size_t sharedMemPerBlockOptin = 232448; int multiProcessorCount = 2; };

inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
template <class F> cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->numRegs = 0; return 0; }
template <class F> cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
template <class F> cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, int) { *n = 1; return 0; }

inline void __syncthreads() {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
using std::isfinite;
using std::max;
using std::min;
