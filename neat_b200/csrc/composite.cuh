// Per-ray HBM-bound kernels around the MLPs: pixel -> ray (rend_util.py:55-108), Laplace density + alpha
// compositing (density.py:21-26, neat_wfr_rend_a.py:404-429, 540-554) and the 2D line geometry
// (neat_wfr_rend_a.py:317-331, 433-456).  One warp per ray, lane-strided over the samples (coalesced).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sampler.cuh"

namespace neat {

// rend_util.get_camera_params + lift: uv [R,2], pose [4,4] (camera-to-world), K [4,4] -> dirs [R,3], cam [3]
__global__ void camera_rays_kernel(const float* __restrict__ uv, const float* __restrict__ pose,
                                   const float* __restrict__ K, int R, float* __restrict__ dirs,
                                   float* __restrict__ cam) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) { cam[0] = pose[3]; cam[1] = pose[7]; cam[2] = pose[11]; }
  if (r >= R) return;
  const float fx = K[0], sk = K[1], cx = K[2], fy = K[5], cy = K[6];
  const float x = uv[2 * r], y = uv[2 * r + 1], z = 1.0f;
  const float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z;
  const float yl = (y - cy) / fy * z;
  float w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float wi = pose[4 * i] * xl + pose[4 * i + 1] * yl + pose[4 * i + 2] * z + pose[4 * i + 3];
    w[i] = wi - pose[4 * i + 3];
  }
  const float n = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);  // F.normalize
  dirs[3 * r] = w[0] / n; dirs[3 * r + 1] = w[1] / n; dirs[3 * r + 2] = w[2] / n;
}

struct CompositeParams {
  int R, S;
  const float* z;        // [R,S]
  const float* sdf;      // [R,S]
  const float* rgb;      // [R,S,3]
  const float* lines;    // [R,S,6]
  const float* normals;  // [R,S,3] (eval: normal map) or nullptr
  const float* rays_o;   // [3]
  const float* rays_d;   // [R,3]
  const float* beta_param;
  float beta_min;
  float* weights;     // [R,S]
  float* rgb_values;  // [R,3]
  float* lines3d;     // [R,6]
  float* depth;       // [R]
  float* points3d;    // [R,3]
  float* normal_map;  // [R,3] or nullptr
  const float* bg;    // [3] or nullptr: white_bkgd, rgb_values += (1 - sum_i w_i) bg  (neat_wfr_rend_a.py:411-413)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// weights_i = (1 - exp(-delta_i sigma_i)) exp(-sum_{j<i} delta_j sigma_j), delta_last = 1e10
__global__ void __launch_bounds__(128) composite_fwd_kernel(CompositeParams p) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= p.R) return;
  const int S = p.S;
  const float beta = fabsf(*p.beta_param) + p.beta_min;
  const float* z = p.z + static_cast<size_t>(r) * S;
  const float* s = p.sdf + static_cast<size_t>(r) * S;
  const float d0 = p.rays_d[3 * r], d1 = p.rays_d[3 * r + 1], d2 = p.rays_d[3 * r + 2];
  const float o0 = p.rays_o[0], o1 = p.rays_o[1], o2 = p.rays_o[2];
  float acc[3 + 6 + 1 + 3 + 3];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float carry = 0.f;  // sum of free energy of all previous 32-sample blocks
  float wsum = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    float fe = 0.f, zi = 0.f;
    if (i < S) {
      zi = z[i];
      const float delta = i + 1 < S ? z[i + 1] - zi : 1e10f;
      fe = delta * laplace_density(s[i], beta);
    }
    float tot;
    const float ex = warp_excl_scan(fe, tot);
    if (i < S) {
      const float w = (1.0f - expf(-fe)) * expf(-(carry + ex));
      if (p.weights) p.weights[static_cast<size_t>(r) * S + i] = w;
      wsum += w;
      const size_t q = static_cast<size_t>(r) * S + i;
      if (p.rgb) { acc[0] += w * p.rgb[3 * q]; acc[1] += w * p.rgb[3 * q + 1]; acc[2] += w * p.rgb[3 * q + 2]; }
      if (p.lines) {
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[3 + c] += w * p.lines[6 * q + c];
      }
      const float v0 = zi * d0, v1 = zi * d1, v2 = zi * d2;
      acc[9] += w * sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
      acc[10] += w * (o0 + v0); acc[11] += w * (o1 + v1); acc[12] += w * (o2 + v2);
      if (p.normals) {
        const float n0 = p.normals[3 * q], n1 = p.normals[3 * q + 1], n2 = p.normals[3 * q + 2];
        const float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
        acc[13] += w * n0 / nn; acc[14] += w * n1 / nn; acc[15] += w * n2 / nn;
      }
    }
    carry += tot;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = warp_sum(acc[i]);
  if (p.bg) wsum = warp_sum(wsum);
  if (lane == 0) {
    if (p.rgb) for (int c = 0; c < 3; ++c) p.rgb_values[3 * r + c] = p.bg ? acc[c] + (1.0f - wsum) * p.bg[c] : acc[c];
    if (p.lines) for (int c = 0; c < 6; ++c) p.lines3d[6 * r + c] = acc[3 + c];
    if (p.depth) p.depth[r] = acc[9];
    if (p.points3d) for (int c = 0; c < 3; ++c) p.points3d[3 * r + c] = acc[10 + c];
    if (p.normal_map) for (int c = 0; c < 3; ++c) p.normal_map[3 * r + c] = acc[13 + c];
  }
}

// general 4x4 inverse (pose[0].inverse(), neat_wfr_rend_a.py:433) by cofactors, one thread
__device__ inline void inverse4x4(const float* m, float* inv) {
  float t[16];
  t[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  t[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  t[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  t[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  t[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  t[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  t[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  t[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  t[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  t[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  t[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  t[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  t[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  t[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  t[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  t[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const float det = m[0] * t[0] + m[1] * t[4] + m[2] * t[8] + m[3] * t[12];
  const float id = 1.0f / det;
  for (int i = 0; i < 16; ++i) inv[i] = t[i] * id;
}

// VolSDFNetwork.project2D (neat_wfr_rend_a.py:317-331) of one 3D point with world-to-camera [R|T] and K3
__device__ __forceinline__ void project2d(const float* K3, const float* RT /*[3][4]*/, const float X[3], float out[2]) {
  float c[3], x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = RT[4 * i] * X[0] + RT[4 * i + 1] * X[1] + RT[4 * i + 2] * X[2] + RT[4 * i + 3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = K3[3 * i] * c[0] + K3[3 * i + 1] * c[1] + K3[3 * i + 2] * c[2];
  const float den = x[2];
  const float sign = den >= 0.f ? 1.f : -1.f;
  const float eps = fabsf(den) < 1e-8f ? 1e-8f : 0.f;
  const float dd = den + eps * sign;
  out[0] = x[0] / dd; out[1] = x[1] / dd;
}

struct GeometryParams {
  int R;
  const float* pose;      // [4,4] camera-to-world
  const float* K;         // [4,4]
  const float* uv_proj;   // [R,2]
  const float* points3d;  // [R,3]
  const float* grad3d;    // [R,3] sdf gradient at points3d
  const float* lines3d;   // [R,6]
  float* lines2d;         // [R,4]
  float* lines2d_calib;   // [R,4]
  float* l3d;             // [R,3]
  float* pose_inv;        // [16] scratch/output: world-to-camera
};

__global__ void pose_inverse_kernel(const float* __restrict__ pose, float* __restrict__ inv) {
  if (threadIdx.x == 0 && blockIdx.x == 0) inverse4x4(pose, inv);
}

// lines2d / lines2d_calib (:439-442) and l3d = ray(uv_proj) ^ tangent plane at points3d (:444-456)
__global__ void line_geometry_kernel(GeometryParams p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  float K3[9];
  const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) K3[3 * i + j] = p.K[4 * i + j];
  const float* RT = p.pose_inv;  // rows 0..2 of the inverse pose
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float X[3] = {p.lines3d[6 * r + 3 * e], p.lines3d[6 * r + 3 * e + 1], p.lines3d[6 * r + 3 * e + 2]};
    float a[2], b[2];
    project2d(K3, RT, X, a);
    project2d(I3, RT, X, b);
    p.lines2d[4 * r + 2 * e] = a[0]; p.lines2d[4 * r + 2 * e + 1] = a[1];
    p.lines2d_calib[4 * r + 2 * e] = b[0]; p.lines2d_calib[4 * r + 2 * e + 1] = b[1];
  }
  // ray through uv_proj (same lift as camera_rays_kernel)
  const float* pose = p.pose;
  const float fx = p.K[0], sk = p.K[1], cx = p.K[2], fy = p.K[5], cy = p.K[6];
  const float x = p.uv_proj[2 * r], y = p.uv_proj[2 * r + 1];
  const float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx;
  const float yl = (y - cy) / fy;
  float w[3], ro[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ro[i] = pose[4 * i + 3];
    w[i] = (pose[4 * i] * xl + pose[4 * i + 1] * yl + pose[4 * i + 2] + pose[4 * i + 3]) - ro[i];
  }
  const float n = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);
  float rd[3] = {w[0] / n, w[1] / n, w[2] / n};
  const float g[3] = {p.grad3d[3 * r], p.grad3d[3 * r + 1], p.grad3d[3 * r + 2]};
  const float den = rd[0] * g[0] + rd[1] * g[1] + rd[2] * g[2];
  const float de = den >= 0.f ? 1e-6f : -1e-6f;
  const float num = (p.points3d[3 * r] - ro[0]) * g[0] + (p.points3d[3 * r + 1] - ro[1]) * g[1] + (p.points3d[3 * r + 2] - ro[2]) * g[2];
  const float tt = num / (den + de);
#pragma unroll
  for (int i = 0; i < 3; ++i) p.l3d[3 * r + i] = ro[i] + rd[i] * tt;
}

}  // namespace neat

namespace neat {

// Adjoint of composite_fwd_kernel w.r.t. the per-point head outputs, the sdf and beta.  Only the outputs the
// reference losses consume carry gradients: rgb_values (through weights and rgb) and lines3d (through the
// per-point lines only: the weights are detached there, neat_wfr_rend_a.py:410).
struct CompositeBwdParams {
  int R, S;
  const float* z;            // [R,S]
  const float* sdf;          // [R,S]
  const float* weights;      // [R,S] (forward output)
  const float* rgb;          // [R,S,3] (sigmoid output)
  const float* act;          // [R,S] or nullptr: 1 where the sdf comes from the network
  const float* rgb_values_bar;  // [R,3]
  const float* lines3d_bar;     // [R,6]
  const float* beta_param;
  float beta_min;
  float* rgb_pre_bar;   // [R,S,3]  dL/d(pre-sigmoid rgb) = w * rgb_values_bar * rgb (1 - rgb)
  float* lines_bar;     // [R,S,6]  = w * lines3d_bar
  float* sdf_bar;       // [R,S]    dL/d(raw network sdf) (masked by act)
  float* beta_bar;      // [1]      accumulated with atomicAdd: dL/d(density.beta parameter)
  const float* bg;      // [3] or nullptr: white_bkgd -- d rgb_values / d w_i = rgb_i - bg
};

__global__ void __launch_bounds__(128) composite_bwd_kernel(CompositeBwdParams p) {
  __shared__ float red[4];
  const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + wq;
  float bsum = 0.f;
  if (r < p.R) {
    const int S = p.S;
    const float bp = *p.beta_param;
    const float beta = fabsf(bp) + p.beta_min;
    const float* z = p.z + static_cast<size_t>(r) * S;
    const float* s = p.sdf + static_cast<size_t>(r) * S;
    const float* w = p.weights + static_cast<size_t>(r) * S;
    float gb[3], lb[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) gb[c] = p.rgb_values_bar[3 * r + c];
    const float gbg = p.bg ? gb[0] * p.bg[0] + gb[1] * p.bg[1] + gb[2] * p.bg[2] : 0.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) lb[c] = p.lines3d_bar[6 * r + c];
    // pass 1 (forward order): exclusive prefix of the free energy -> T_i ; pass 2 (reverse): suffix of w_bar w
    // Both are done block-wise; T is recomputed in the reverse pass from the stored prefix per block.
    const int nblk = (S + 31) / 32;
    float blk_prefix[8];  // S <= 256
    float carry = 0.f;
    for (int b = 0; b < nblk; ++b) {
      const int i = b * 32 + lane;
      float fe = 0.f;
      if (i < S) {
        const float delta = i + 1 < S ? z[i + 1] - z[i] : 1e10f;
        fe = delta * laplace_density(s[i], beta);
      }
      blk_prefix[b] = carry;
      carry += warp_sum(fe);
    }
    float suffix_carry = 0.f;  // sum_{k in later blocks} w_bar_k w_k
    for (int b = nblk - 1; b >= 0; --b) {
      const int i = b * 32 + lane;
      const size_t q = static_cast<size_t>(r) * S + i;
      float fe = 0.f, delta = 0.f, sig = 0.f, wi = 0.f, wbar = 0.f, si = 0.f;
      float rgbv[3] = {0.f, 0.f, 0.f};
      if (i < S) {
        si = s[i];
        delta = i + 1 < S ? z[i + 1] - z[i] : 1e10f;
        sig = laplace_density(si, beta);
        fe = delta * sig;
        wi = w[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) { rgbv[c] = p.rgb[3 * q + c]; wbar += gb[c] * rgbv[c]; }
        wbar -= gbg;
      }
      float tot;
      const float ex = warp_excl_scan(fe, tot);
      const float ww = wbar * wi;
      // suffix_i = sum_{k>i} ww_k by a reverse scan.  It must be EXACTLY 0 for the last sample (its sigma_bar is
      // multiplied by delta = 1e10), so it is built from shuffles, never as total - prefix - own.
      float inc = ww;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float tdn = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += tdn;
      }
      const float tot2 = __shfl_sync(0xffffffffu, inc, 0);
      const float nxt = __shfl_down_sync(0xffffffffu, inc, 1);
      const float suffix = (lane == 31 ? 0.f : nxt) + suffix_carry;
      if (i < S) {
        const float T = expf(-(blk_prefix[b] + ex));
        const float em = expf(-fe);
        const float fe_bar = wbar * em * T - suffix;
        const float sigma_bar = fe_bar * delta;
        const float e = expf(-fabsf(si) / beta);
        const float sg2 = si != 0.f ? 1.f : 0.f;
        const float dsig_ds = -e / (2.0f * beta * beta) * sg2;
        const float dsig_db = -sig / beta + si * e / (2.0f * beta * beta * beta);
        const float a = p.act ? p.act[q] : 1.f;
        p.sdf_bar[q] = sigma_bar * dsig_ds * a;
        bsum += sigma_bar * dsig_db;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.rgb_pre_bar[3 * q + c] = wi * gb[c] * rgbv[c] * (1.0f - rgbv[c]);
#pragma unroll
        for (int c = 0; c < 6; ++c) p.lines_bar[6 * q + c] = wi * lb[c];
      }
      suffix_carry += tot2;
    }
    bsum = warp_sum(bsum) * (bp >= 0.f ? 1.f : -1.f);
  }
  if (lane == 0) red[wq] = bsum;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(p.beta_bar, red[0] + red[1] + red[2] + red[3]);
}

}  // namespace neat
