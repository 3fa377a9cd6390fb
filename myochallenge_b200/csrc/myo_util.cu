// Measurement helpers exported through the C ABI (bench.py's compute roofline).
#include <cuda_runtime.h>

#include "../../include/myo_b200.h"
#include "myo_model.hpp"

namespace {

// 16 independent FMA chains per thread: enough to cover the FMA pipe's latency with 8 resident warps per scheduler
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; k++) s += x[k];
  if (s == 123.456f) out[0] = s;      // keeps the chains alive
}

}  // namespace

extern "C" int myo_fp32_fma_peak(int device, double* tflops) {
  if (!tflops) { myo::set_error("null argument"); return MYO_E_ARG; }
  if (cudaSetDevice(device) != cudaSuccess) { myo::set_error("cudaSetDevice failed"); return MYO_E_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { myo::set_error("cudaGetDeviceProperties failed"); return MYO_E_CUDA; }
  float* out = nullptr;
  if (cudaMalloc(&out, 16) != cudaSuccess) { myo::set_error("cudaMalloc failed"); return MYO_E_CUDA; }
  const int grid = prop.multiProcessorCount * 8, iters = 8192;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {      // first repetition warms up
    cudaEventRecord(e0);
    fma_peak_kernel<<<grid, 256>>>(out, iters, 1.000001f, 1e-7f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); myo::set_error("fma peak kernel failed"); return MYO_E_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * (double)iters * 256.0 * (double)grid;
    if (rep > 0 && ms > 0.f) best = best > flops / (ms * 1e-3) ? best : flops / (ms * 1e-3);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best * 1e-12;
  return MYO_OK;
}
