// DDIM update around the UNet (sampler.cu): shared between the single-GPU C entry and the frame-sharded one in unet.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace dawn {

// cross-rank reductions of the exact radix-select (in place, stream-ordered); ctx is the caller's communicator
struct DdimReduce {
  void* ctx;
  int (*sum_u32)(void* ctx, unsigned int* buf, size_t n, cudaStream_t st);
  int (*sum_u64)(void* ctx, unsigned long long* buf, size_t n, cudaStream_t st);
  int (*min_u32)(void* ctx, unsigned int* buf, size_t n, cudaStream_t st);
};

int ddim_step_impl(float* x, const float* eps, const float* noise, int64_t n_local, int64_t n_global, float ca, float cb,
                   float sqrt_an, float c, float sigma, float q, void* scratch, cudaStream_t st, const DdimReduce* red);

}  // namespace dawn
