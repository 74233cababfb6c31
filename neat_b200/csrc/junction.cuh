// Device side of the junction terms: the projections of the global junctions (VolSDFNetwork.forward,
// neat_wfr_rend_a.py:484-486: project2D with K and with the identity) and the Hungarian-matched junction losses
// (VolSDFLoss.forward, loss_wfr.py:110-125) with their adjoints.  Each of them is a dozen-odd elementwise / index /
// reduce launches in eager PyTorch; on ~1000 points that is pure launch latency, so each is ONE small kernel here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "composite.cuh"

namespace neat {

// out_pix = project2D(K, R, T, X), out_cal = project2D(I, R, T, X);  RT = rows 0..2 of pose^-1 (row-major [4,4])
__global__ void project_points_kernel(int N, const float* __restrict__ pose_inv, const float* __restrict__ K, int k_ld,
                                      const float* __restrict__ X, float* __restrict__ out_pix,
                                      float* __restrict__ out_cal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float x[3] = {X[3 * i], X[3 * i + 1], X[3 * i + 2]};
  const float I3[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  float K3[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) K3[3 * r + c] = K[r * k_ld + c];
  float o[2];
  if (out_pix) {
    project2d(K3, pose_inv, x, o);
    out_pix[2 * i] = o[0]; out_pix[2 * i + 1] = o[1];
  }
  if (out_cal) {
    project2d(I3, pose_inv, x, o);
    out_cal[2 * i] = o[0]; out_cal[2 * i + 1] = o[1];
  }
}

// adjoint of one projection u = (M c)_{0,1} / ((M c)_2 + eps sign), c = R X + T:  g_u [2] -> g_X [3] (added)
__device__ __forceinline__ void project2d_adjoint(const float* M3, const float* RT, const float X[3], const float g[2],
                                                  float gX[3]) {
  float c[3], x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = RT[4 * i] * X[0] + RT[4 * i + 1] * X[1] + RT[4 * i + 2] * X[2] + RT[4 * i + 3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = M3[3 * i] * c[0] + M3[3 * i + 1] * c[1] + M3[3 * i + 2] * c[2];
  const float sign = x[2] >= 0.f ? 1.f : -1.f;
  const float eps = fabsf(x[2]) < 1e-8f ? 1e-8f : 0.f;
  const float dd = x[2] + eps * sign;
  const float gx[3] = {g[0] / dd, g[1] / dd, -(g[0] * x[0] + g[1] * x[1]) / (dd * dd)};
  float gc[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gc[j] = M3[j] * gx[0] + M3[3 + j] * gx[1] + M3[6 + j] * gx[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) gX[j] += RT[j] * gc[0] + RT[4 + j] * gc[1] + RT[8 + j] * gc[2];
}

__global__ void project_points_bwd_kernel(int N, const float* __restrict__ pose_inv, const float* __restrict__ K, int k_ld,
                                          const float* __restrict__ X, const float* __restrict__ g_pix,
                                          const float* __restrict__ g_cal, float* __restrict__ gX) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float x[3] = {X[3 * i], X[3 * i + 1], X[3 * i + 2]};
  float acc[3] = {0.f, 0.f, 0.f};
  if (g_pix) {
    float K3[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) K3[3 * r + c] = K[r * k_ld + c];
    const float g[2] = {g_pix[2 * i], g_pix[2 * i + 1]};
    project2d_adjoint(K3, pose_inv, x, g, acc);
  }
  if (g_cal) {
    const float I3[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    const float g[2] = {g_cal[2 * i], g_cal[2 * i + 1]};
    project2d_adjoint(I3, pose_inv, x, g, acc);
  }
  gX[3 * i] = acc[0]; gX[3 * i + 1] = acc[1]; gX[3 * i + 2] = acc[2];
}

// loss_wfr.py:110-121 for the matched pairs (rows[i], cols[i]), i < n:
//   out[0] = mean_i |j3l[r] - j3g[c]|_1,  out[1] = mean_i |j2lc[r] - j2gc[c]|_1,  out[2] = mean_i |j2l[r] - j2g[c]|_1
// One block; n is at most the number of ground-truth junctions.
__global__ void __launch_bounds__(256) junction_terms_kernel(int n, const float* __restrict__ j3l, const float* __restrict__ j3g,
                                                             const float* __restrict__ j2lc, const float* __restrict__ j2gc,
                                                             const float* __restrict__ j2l, const float* __restrict__ j2g,
                                                             const int* __restrict__ rows, const int* __restrict__ cols,
                                                             float* __restrict__ out) {
  __shared__ float red[3][8];
  float s3 = 0.f, s2 = 0.f, su = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = rows[i], c = cols[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) s3 += fabsf(j3l[3 * r + k] - j3g[3 * c + k]);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      s2 += fabsf(j2lc[2 * r + k] - j2gc[2 * c + k]);
      su += fabsf(j2l[2 * r + k] - j2g[2 * c + k]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    su += __shfl_xor_sync(0xffffffffu, su, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = s3; red[1][w] = s2; red[2][w] = su; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) t += red[threadIdx.x][k];
    out[threadIdx.x] = t / static_cast<float>(n);
  }
}

// adjoint w.r.t. the GLOBAL junctions (the local ones are detached cluster centroids): g_out [2] = d L / d out[0..1];
// g_j3g [G,3] and g_j2gc [G,2] must be zero-filled (unmatched rows get no gradient); the assignment is injective.
__global__ void junction_terms_bwd_kernel(int n, const float* __restrict__ j3l, const float* __restrict__ j3g,
                                          const float* __restrict__ j2lc, const float* __restrict__ j2gc,
                                          const int* __restrict__ rows, const int* __restrict__ cols,
                                          const float* __restrict__ g_out, float* __restrict__ g_j3g,
                                          float* __restrict__ g_j2gc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i], c = cols[i];
  const float k3 = g_out[0] / static_cast<float>(n), k2 = g_out[1] / static_cast<float>(n);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float d = j3l[3 * r + k] - j3g[3 * c + k];
    g_j3g[3 * c + k] = -k3 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float d = j2lc[2 * r + k] - j2gc[2 * c + k];
    g_j2gc[2 * c + k] = -k2 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
  }
}

// use_l3d junction candidates (neat_wfr_rend_a.py:454-455, 461-465): score_i = |(l3d_i - a_i) x (l3d_i - b_i)| / |a_i - b_i|
// for the 3D line (a_i, b_i) of ray i; the rays with score < max(median(score), 0.01) hand over both end points (ray
// order), followed by their l3d points.  The reference does this with median / boolean-mask / cat launches and a host
// sync for the mask; here: ONE block -- scores and a bitonic sort of their copy in shared memory (torch.median = the
// lower median, element (R-1)/2 of the sorted scores), then an ordered compaction.  out [3R,3]; *n_out = 3 n_selected.
constexpr int L3D_MAX_R = 4096;
__global__ void __launch_bounds__(1024) l3d_candidates_kernel(int R, const float* __restrict__ lines3d,
                                                              const float* __restrict__ l3d, float* __restrict__ out,
                                                              int* __restrict__ n_out, float* __restrict__ score_out) {
  __shared__ float score[L3D_MAX_R];
  __shared__ float key[L3D_MAX_R];
  __shared__ int wsum[32];
  __shared__ int base;
  const int tid = threadIdx.x;
  int P2 = 1;
  while (P2 < R) P2 <<= 1;
  for (int i = tid; i < P2; i += blockDim.x) {
    float sc = INFINITY;
    if (i < R) {
      const float* a = lines3d + 6 * i;
      const float* q = l3d + 3 * i;
      const float u[3] = {q[0] - a[0], q[1] - a[1], q[2] - a[2]}, v[3] = {q[0] - a[3], q[1] - a[4], q[2] - a[5]};
      const float cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
      const float dx = a[0] - a[3], dy = a[1] - a[4], dz = a[2] - a[5];
      sc = sqrtf(cx * cx + cy * cy + cz * cz) / sqrtf(dx * dx + dy * dy + dz * dz);
      score[i] = sc;
      if (score_out) score_out[i] = sc;
    }
    key[i] = sc == sc ? sc : INFINITY;   // a NaN score (degenerate line) sorts last and is never selected
  }
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float x = key[i], y = key[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { key[i] = y; key[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  const float thr = fmaxf(key[(R - 1) / 2], 0.01f);
  // pass 1: number selected; pass 2: ordered positions
  int cnt = 0;
  for (int i = tid; i < R; i += blockDim.x) cnt += score[i] < thr ? 1 : 0;
  if (tid == 0) base = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) wsum[tid >> 5] = cnt;
  __syncthreads();
  int total = 0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) total += wsum[w];
  __syncthreads();
  for (int start = 0; start < R; start += blockDim.x) {
    const int i = start + tid;
    const int flag = (i < R && score[i] < thr) ? 1 : 0;
    int v = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if ((tid & 31) >= o) v += t;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      int sct = tid < (blockDim.x >> 5) ? wsum[tid] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sct, o);
        if (tid >= o) sct += t;
      }
      wsum[tid] = sct;
    }
    __syncthreads();
    const int woff = (tid >> 5) ? wsum[(tid >> 5) - 1] : 0;
    if (flag) {
      const int pos = base + woff + v - 1;
#pragma unroll
      for (int c = 0; c < 6; ++c) out[6 * pos + c] = lines3d[6 * i + c];
#pragma unroll
      for (int c = 0; c < 3; ++c) out[3 * (2 * total + pos) + c] = l3d[3 * i + c];
    }
    __syncthreads();
    if (tid == 0) base += wsum[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  if (tid == 0) *n_out = 3 * total;
}

}  // namespace neat
