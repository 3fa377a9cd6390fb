// TEST INFRASTRUCTURE: builds the product's kernel translation unit for the host, one lane per world,
// behind the same C ABI (include/myo_b200.h), so the kernel *logic* can be diffed against the oracle
// on a machine without a GPU. Loaded only by tests/test_emul_*.py; the package never loads it.
#include "fake/cuda_runtime.h"

thread_local dim3e blockIdx, threadIdx, blockDim, gridDim;
thread_local float4* emul_smem = nullptr;

#include <vector>
template <class K, class... A>
void emul_launch(K kernel, unsigned grid, unsigned block, size_t smem, A... args) {
  std::vector<float4> buf(smem / sizeof(float4) + 1);
  emul_smem = buf.data();
  gridDim.x = grid; blockDim.x = block;
  for (unsigned b = 0; b < grid; b++)
    for (unsigned t = 0; t < block; t++) { blockIdx.x = b; threadIdx.x = t; kernel(args...); }
}
#define MYO_LAUNCH(kernel, grid, block, smem, stream, ...) emul_launch(kernel, (unsigned)(grid), (unsigned)(block), (size_t)(smem), __VA_ARGS__)
#define MYO_LANES_CASES(b, FN, ...) { (b)->pm.lanes = 1; rc = FN<1>(__VA_ARGS__); }
#define MYO_EMUL 1
#include "../../myochallenge_b200/csrc/myo_kernels.cu"
