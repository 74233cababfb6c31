// torch.optim.Adam(lr) (code/training/volsdf_train.py:178, 374) for all parameter tensors in ONE launch.  The reference
// (and torch's own fused=True path) walks the 64 parameter tensors of the model with multi-tensor kernels (2 x 76 us
// at this model size); here a table of (param, grad, exp_avg, exp_avg_sq, numel) is walked by one grid, and the 1/world
// scale of the data-parallel gradient mean rides along (no separate flat.mul_ launch after the all-reduce).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/neat_b200.h"

namespace neat {

constexpr int ADAM_MAX_TENSORS = 128;
constexpr int ADAM_BLOCK_ELEMS = 2048;  // 256 threads x 8
struct AdamTable {
  neat_adam_tensor t[ADAM_MAX_TENSORS];
  int blk_start[ADAM_MAX_TENSORS + 1];
  int n;
};

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamTable* __restrict__ tp, float lr, float beta1, float beta2,
                                                        float eps, float weight_decay, float bc1, float rsqrt_bc2,
                                                        float grad_scale) {
  const AdamTable& T = *tp;
  int lo = 0, hi = T.n;  // tensor of this block: last i with blk_start[i] <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (T.blk_start[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid;
  }
  const neat_adam_tensor a = T.t[lo];
  const long long base = static_cast<long long>(blockIdx.x - T.blk_start[lo]) * ADAM_BLOCK_ELEMS;
  const float step_size = lr / bc1;
#pragma unroll
  for (int k = 0; k < ADAM_BLOCK_ELEMS / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < a.numel) {
      float g = a.grad[i] * grad_scale;
      const float p = a.param[i];
      if (weight_decay != 0.f) g += weight_decay * p;
      const float m = a.exp_avg[i] + (g - a.exp_avg[i]) * (1.0f - beta1);   // lerp, as torch
      const float v = a.exp_avg_sq[i] * beta2 + (1.0f - beta2) * g * g;
      a.exp_avg[i] = m;
      a.exp_avg_sq[i] = v;
      const float denom = sqrtf(v) * rsqrt_bc2 + eps;
      a.param[i] = p - step_size * (m / denom);
    }
  }
}

}  // namespace neat
