// nn.utils.weight_norm (dim=0) for all Linear layers of the three MLPs in ONE launch each way
// (neat_wfr_rend_a.py:71-72, 167-168, 227-228):  W[r,:] = g[r] v[r,:] / |v[r,:]|,  and its adjoint
//   g_bar[r] = <W_bar[r,:], v[r,:]> / |v[r,:]|,   v_bar[r,:] = g[r]/|v| (W_bar[r,:] - <W_bar[r,:], v[r,:]> v[r,:] / |v|^2).
// One warp per output row; the effective weights / their gradients live in the flat buffer of neat_param_count().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/neat_b200.h"

namespace neat {

constexpr int WN_MAX_LAYERS = 32;
struct WnTable {
  neat_wn_layer l[WN_MAX_LAYERS];
  int row_start[WN_MAX_LAYERS + 1];
  int n;
};

__device__ __forceinline__ float wn_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(const WnTable* __restrict__ tp, float* __restrict__ flat) {
  const WnTable& t = *tp;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= t.row_start[t.n]) return;
  int li = 0;
  while (row >= t.row_start[li + 1]) ++li;
  const neat_wn_layer L = t.l[li];
  const int r = row - t.row_start[li];
  const float* v = L.v + static_cast<size_t>(r) * L.cols;
  float* w = flat + L.w_off + static_cast<size_t>(r) * L.cols;
  if (L.cols <= 10 * 32) {   // the row in registers: all loads in flight at once (see weight_norm_bwd_kernel)
    float vk[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) vk[j] = lane + 32 * j < L.cols ? v[lane + 32 * j] : 0.f;
    float scale = 1.f;
    if (L.g) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 10; ++j) ss += vk[j] * vk[j];
      scale = L.g[r] / sqrtf(wn_warp_sum(ss));
    }
#pragma unroll
    for (int j = 0; j < 10; ++j)
      if (lane + 32 * j < L.cols) w[lane + 32 * j] = L.g ? vk[j] * scale : vk[j];
  } else {
    float scale = 1.f;
    if (L.g) {
      float ss = 0.f;
      for (int k = lane; k < L.cols; k += 32) ss += v[k] * v[k];
      scale = L.g[r] / sqrtf(wn_warp_sum(ss));
    }
    for (int k = lane; k < L.cols; k += 32) w[k] = L.g ? v[k] * scale : v[k];
  }
  if (lane == 0) flat[L.b_off + r] = L.b[r];
}

// Rows are at most WN_REG * 32 = 320 wide in every supported net (<= 256 hidden + 48 aux inputs), so a lane keeps its
// <= 10 elements of v and W_bar in registers: every load of the row is issued up front (the plain `for k += 32` loop
// took one dependent round trip per element and pass: 27 us for 17 MB) and the second pass re-reads nothing.
constexpr int WN_REG = 10;

__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const WnTable* __restrict__ tp, const float* __restrict__ flat_grad,
                                                              int accumulate) {
  const WnTable& t = *tp;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= t.row_start[t.n]) return;
  int li = 0;
  while (row >= t.row_start[li + 1]) ++li;
  const neat_wn_layer L = t.l[li];
  const int r = row - t.row_start[li];
  const float* __restrict__ v = L.v + static_cast<size_t>(r) * L.cols;
  const float* __restrict__ wb = flat_grad + L.w_off + static_cast<size_t>(r) * L.cols;
  float* gv = L.gv + static_cast<size_t>(r) * L.cols;
  if (L.cols <= WN_REG * 32) {
    float vk[WN_REG], wk[WN_REG], old[WN_REG];
#pragma unroll
    for (int j = 0; j < WN_REG; ++j) {
      const int k = lane + 32 * j;
      const bool in = k < L.cols;
      vk[j] = (in && L.g) ? v[k] : 0.f;
      wk[j] = in ? wb[k] : 0.f;
      old[j] = (in && accumulate) ? gv[k] : 0.f;
    }
    float gi = 1.f, c = 0.f;
    if (L.g) {
      float ss = 0.f, dot = 0.f;
#pragma unroll
      for (int j = 0; j < WN_REG; ++j) { ss += vk[j] * vk[j]; dot += wk[j] * vk[j]; }
      ss = wn_warp_sum(ss);
      dot = wn_warp_sum(dot);
      const float inv = rsqrtf(ss);
      gi = L.g[r] * inv;
      c = dot / ss;
      if (lane == 0) L.gg[r] = (accumulate ? L.gg[r] : 0.f) + dot * inv;
    }
#pragma unroll
    for (int j = 0; j < WN_REG; ++j) {
      const int k = lane + 32 * j;
      if (k < L.cols) gv[k] = old[j] + (L.g ? gi * (wk[j] - c * vk[j]) : wk[j]);
    }
  } else if (L.g) {
    float ss = 0.f, dot = 0.f;
    for (int k = lane; k < L.cols; k += 32) { ss += v[k] * v[k]; dot += wb[k] * v[k]; }
    ss = wn_warp_sum(ss);
    dot = wn_warp_sum(dot);
    const float inv = rsqrtf(ss);
    const float gi = L.g[r] * inv;
    const float c = dot / ss;
    for (int k = lane; k < L.cols; k += 32) gv[k] = (accumulate ? gv[k] : 0.f) + gi * (wb[k] - c * v[k]);
    if (lane == 0) L.gg[r] = (accumulate ? L.gg[r] : 0.f) + dot * inv;
  } else {
    for (int k = lane; k < L.cols; k += 32) gv[k] = (accumulate ? gv[k] : 0.f) + wb[k];
  }
  if (lane == 0) L.gb[r] = (accumulate ? L.gb[r] : 0.f) + flat_grad[L.b_off + r];
}

}  // namespace neat
