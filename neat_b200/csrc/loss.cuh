// VolSDFLoss (code/model/networks/loss_wfr.py:34-79) fused with its own backward: rgb L1, eikonal, and the two
// endpoint-order-invariant 2D line losses (uncalibrated for gating/statistics, calibrated for the gradient).
// Two tiny kernels: (1) per-ray terms + the global sums (counts are denominators), (2) the gradients w.r.t.
// rgb_values, lines2d_calib and grad_theta.  project2D's adjoint (neat_wfr_rend_a.py:317-331, 442) is a third kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace neat {

struct LossParams {
  int R, n_eik;
  const float* rgb_values;     // [R,3]
  const float* rgb_gt;         // [R,3]
  const float* lines2d;        // [R,4]  (uncalibrated, no gradient)
  const float* lines2d_calib;  // [R,4]
  const float* lines_gt;       // [R,5]  x1 y1 x2 y2 weight
  const float* labels;         // [R] or nullptr (multiplies the weight)
  const float* K3;             // [3,3] row-major (ld = k_ld)
  int k_ld;
  const float* grad_theta;     // [n_eik,3] or nullptr
  float eikonal_weight, line_weight;
  // sums[0] rgb abs sum, [1] eik sum, [2] uncal numerator, [3] uncal count, [4] calib numerator, [5] calib count
  float* sums;                 // [8] zeroed by the caller
  float* per_uncal;            // [R] scratch: uncalibrated per-ray loss (gates the calibrated one)
  // outputs of pass 2
  float* out;                  // [8]: loss_core, rgb_loss, eikonal_loss, line_loss, l2d_loss, count
  float* g_rgb;                // [R,3]
  float* g_calib;              // [R,4]
  float* g_theta;              // [n_eik,3]
};

__device__ __forceinline__ void inv3x3(const float* m, int ld, float* o) {
  const float a = m[0], b = m[1], c = m[2], d = m[ld], e = m[ld + 1], f = m[ld + 2], g = m[2 * ld], h = m[2 * ld + 1], i = m[2 * ld + 2];
  const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const float id = 1.0f / (a * A + b * B + c * C);
  o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
  o[3] = B * id; o[4] = (a * i - c * g) * id;  o[5] = -(a * f - c * d) * id;
  o[6] = C * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}

// get_line_loss for one ray: returns mean |l - tgt| and writes sign(l - tgt) (for the gradient)
__device__ __forceinline__ float line_term(const float l[4], const float gt[4], float sgn[4]) {
  const float sw[4] = {gt[2], gt[3], gt[0], gt[1]};
  float d1 = 0.f, d2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { d1 += (l[i] - gt[i]) * (l[i] - gt[i]); d2 += (l[i] - sw[i]) * (l[i] - sw[i]); }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float df = l[i] - (d1 < d2 ? gt[i] : sw[i]);
    s += fabsf(df);
    sgn[i] = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
  }
  return s * 0.25f;
}

__device__ __forceinline__ void calib_gt(const float* Kinv, const float gt[4], float out[4]) {
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float x = gt[2 * e], y = gt[2 * e + 1];
    const float hx = Kinv[0] * x + Kinv[1] * y + Kinv[2], hy = Kinv[3] * x + Kinv[4] * y + Kinv[5],
                hz = Kinv[6] * x + Kinv[7] * y + Kinv[8];
    out[2 * e] = hx / hz; out[2 * e + 1] = hy / hz;
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) loss_terms_kernel(LossParams p) {
  __shared__ float red[8];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float s_rgb = 0.f, s_eik = 0.f, s_un = 0.f, c_un = 0.f;
  if (r < p.R) {
#pragma unroll
    for (int c = 0; c < 3; ++c) s_rgb += fabsf(p.rgb_values[3 * r + c] - p.rgb_gt[3 * r + c]);
    float l[4], gt[4], sg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { l[i] = p.lines2d[4 * r + i]; gt[i] = p.lines_gt[5 * r + i]; }
    const float w = p.lines_gt[5 * r + 4] * (p.labels ? p.labels[r] : 1.f);
    const float per = line_term(l, gt, sg);
    p.per_uncal[r] = per;
    if (per < 100.f) { s_un = per * w; c_un = 1.f; }
  }
  if (p.grad_theta && r < p.n_eik) {
    const float gx = p.grad_theta[3 * r], gy = p.grad_theta[3 * r + 1], gz = p.grad_theta[3 * r + 2];
    const float n = sqrtf(gx * gx + gy * gy + gz * gz);
    s_eik = (n - 1.f) * (n - 1.f);
  }
  float t;
  t = block_sum(s_rgb, red); if (threadIdx.x == 0) atomicAdd(p.sums + 0, t);
  t = block_sum(s_eik, red); if (threadIdx.x == 0) atomicAdd(p.sums + 1, t);
  t = block_sum(s_un, red);  if (threadIdx.x == 0) atomicAdd(p.sums + 2, t);
  t = block_sum(c_un, red);  if (threadIdx.x == 0) atomicAdd(p.sums + 3, t);
}

// calibrated line loss needs the uncalibrated gate of every ray (already complete), so it is a second launch
__global__ void __launch_bounds__(256) loss_calib_kernel(LossParams p) {
  __shared__ float red[8];
  __shared__ float Kinv[9];
  if (threadIdx.x == 0) inv3x3(p.K3, p.k_ld, Kinv);
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float s_c = 0.f, c_c = 0.f;
  if (r < p.R) {
    float l[4], gt[4], gc[4], sg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { l[i] = p.lines2d_calib[4 * r + i]; gt[i] = p.lines_gt[5 * r + i]; }
    calib_gt(Kinv, gt, gc);
    const float w = p.lines_gt[5 * r + 4] * (p.labels ? p.labels[r] : 1.f) * (p.per_uncal[r] < 100.f ? 1.f : 0.f);
    const float per = line_term(l, gc, sg);
    if (per < 100.f) { s_c = per * w; c_c = 1.f; }
  }
  float t;
  t = block_sum(s_c, red); if (threadIdx.x == 0) atomicAdd(p.sums + 4, t);
  t = block_sum(c_c, red); if (threadIdx.x == 0) atomicAdd(p.sums + 5, t);
}

__global__ void __launch_bounds__(256) loss_grads_kernel(LossParams p) {
  __shared__ float Kinv[9];
  if (threadIdx.x == 0) inv3x3(p.K3, p.k_ld, Kinv);
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const float n_un = fmaxf(p.sums[3], 1.f), n_c = fmaxf(p.sums[5], 1.f);
  if (r == 0) {
    const float rgb_loss = p.sums[0] / (3.f * p.R);
    const float eik = p.grad_theta ? p.sums[1] / p.n_eik : 0.f;
    const float line = p.sums[4] / n_c, l2d = p.sums[2] / n_un;
    p.out[0] = rgb_loss + p.eikonal_weight * eik + p.line_weight * line;
    p.out[1] = rgb_loss; p.out[2] = eik; p.out[3] = line; p.out[4] = l2d; p.out[5] = p.sums[3];
  }
  if (r < p.R) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = p.rgb_values[3 * r + c] - p.rgb_gt[3 * r + c];
      p.g_rgb[3 * r + c] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / (3.f * p.R);
    }
    float l[4], gt[4], gc[4], sg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { l[i] = p.lines2d_calib[4 * r + i]; gt[i] = p.lines_gt[5 * r + i]; }
    calib_gt(Kinv, gt, gc);
    const float w = p.lines_gt[5 * r + 4] * (p.labels ? p.labels[r] : 1.f) * (p.per_uncal[r] < 100.f ? 1.f : 0.f);
    const float per = line_term(l, gc, sg);
    const float k = (per < 100.f ? 1.f : 0.f) * w * 0.25f * p.line_weight / n_c;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.g_calib[4 * r + i] = k * sg[i];
  }
  if (p.grad_theta && r < p.n_eik) {
    const float gx = p.grad_theta[3 * r], gy = p.grad_theta[3 * r + 1], gz = p.grad_theta[3 * r + 2];
    const float n = sqrtf(gx * gx + gy * gy + gz * gz);
    // |grad_theta| = 0: torch's norm backward yields the zero subgradient there (no 0/0)
    const float k = n > 0.f ? p.eikonal_weight * 2.f * (n - 1.f) / (n * p.n_eik) : 0.f;
    p.g_theta[3 * r] = k * gx; p.g_theta[3 * r + 1] = k * gy; p.g_theta[3 * r + 2] = k * gz;
  }
}

// adjoint of lines2d_calib = project2D(I, R, T, lines3d): g_calib [R,4] -> g_lines3d [R,6]
__global__ void project_calib_bwd_kernel(int R, const float* __restrict__ pose_inv, const float* __restrict__ lines3d,
                                         const float* __restrict__ g_calib, float* __restrict__ g_lines3d) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* RT = pose_inv;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float X[3] = {lines3d[6 * r + 3 * e], lines3d[6 * r + 3 * e + 1], lines3d[6 * r + 3 * e + 2]};
    float c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) c[i] = RT[4 * i] * X[0] + RT[4 * i + 1] * X[1] + RT[4 * i + 2] * X[2] + RT[4 * i + 3];
    const float sign = c[2] >= 0.f ? 1.f : -1.f;
    const float eps = fabsf(c[2]) < 1e-8f ? 1e-8f : 0.f;
    const float dd = c[2] + eps * sign;
    const float g0 = g_calib[4 * r + 2 * e], g1 = g_calib[4 * r + 2 * e + 1];
    const float gc[3] = {g0 / dd, g1 / dd, -(g0 * c[0] + g1 * c[1]) / (dd * dd)};
#pragma unroll
    for (int j = 0; j < 3; ++j) g_lines3d[6 * r + 3 * e + j] = RT[j] * gc[0] + RT[4 + j] * gc[1] + RT[8 + j] * gc[2];
  }
}

}  // namespace neat
