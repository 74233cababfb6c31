// Dataset-side attraction precompute (runs once per image, feeds the step its uv_proj / labels):
//   encodels_kernel          == hawp.base._C.encodels, `encode_kernel` of the reference's only CUDA file
//                               (third-party/hawp/hawp/base/csrc/linesegment.cu:23-103)
//   point_line_attraction    == SceneDataset.compute_point_line_attraction (code/datasets/scene_hawp_dataset.py:92-146)
//                               fused: no [num_lines, H, W] one-hot tensor is ever materialised.
// One thread per pixel, line segments staged through shared memory, every result written exactly once (the reference
// kernel rewrites its six output planes each time a closer segment is found).  Arithmetic is spelled out with explicit
// round-to-nearest intrinsics in exactly the form nvcc gives the reference kernel at its default flags (-fmad=true):
// every `a*a + b*b` is fma(a, a, rn(b*b)), `x1 + t*dx` is fma(t, dx, x1), the division runs in double.  That pattern was
// read off the SASS of the reference kernel built for sm_100a (oracle/build_ref.py -> oracle/_ref/), and
// tests/test_hawp_oracle.py::test_gpu_encodels_vs_reference_kernel checks bit-equality against that binary on the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace neat {

constexpr int ATT_LINES_PER_PASS = 1024;

struct NearestLine {
  float ax, ay, ux, uy, vx, vy, t, dis;
  int idx;
  bool inside;
};

// the body of encode_kernel's loop for one pixel; lines staged in shared memory by the whole block
__device__ __forceinline__ NearestLine nearest_line(const float* __restrict__ lines, int num, float xs, float ys, float px,
                                                    float py, float4* sl) {
  NearestLine b;
  b.dis = 1e30f; b.idx = -1; b.inside = true;
  b.ax = b.ay = b.ux = b.uy = b.vx = b.vy = b.t = 0.f;
  for (int base = 0; base < num; base += ATT_LINES_PER_PASS) {
    const int cnt = min(ATT_LINES_PER_PASS, num - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float4 l = reinterpret_cast<const float4*>(lines)[base + i];
      sl[i] = make_float4(__fmul_rn(l.x, xs), __fmul_rn(l.y, ys), __fmul_rn(l.z, xs), __fmul_rn(l.w, ys));
    }
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      const float x1 = sl[i].x, y1 = sl[i].y, x2 = sl[i].z, y2 = sl[i].w;
      const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1);
      const float norm2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
      const float num_ = __fmaf_rn(__fsub_rn(px, x1), dx, __fmul_rn(__fsub_rn(py, y1), dy));
      float t = static_cast<float>(static_cast<double>(num_) / (static_cast<double>(norm2) + 1e-6));  // double, as the reference
      const bool flag = t <= 1.f && t >= 0.f;
      t = t < 0.f ? 0.f : t;
      t = t > 1.f ? 1.f : t;
      const float ax = __fsub_rn(__fmaf_rn(t, dx, x1), px);
      const float ay = __fsub_rn(__fmaf_rn(t, dy, y1), py);
      const float dis = __fmaf_rn(ax, ax, __fmul_rn(ay, ay));
      if (dis < b.dis) {
        b.dis = dis; b.ax = ax; b.ay = ay; b.t = t; b.idx = base + i; b.inside = flag;
        const float ux = __fsub_rn(x1, px), uy = __fsub_rn(y1, py), vx = __fsub_rn(x2, px), vy = __fsub_rn(y2, py);
        const bool first = __fmaf_rn(ux, ux, __fmul_rn(uy, uy)) < __fmaf_rn(vx, vx, __fmul_rn(vy, vy));
        b.ux = first ? ux : vx; b.uy = first ? uy : vy; b.vx = first ? vx : ux; b.vy = first ? vy : uy;
      }
    }
  }
  return b;
}

// map [6,H,W], label [num,H,W] (bool, zero-initialised by the caller), tmap [1,H,W]
__global__ void __launch_bounds__(256) encodels_kernel(const float* __restrict__ lines, int input_height, int input_width,
                                                       int num, int height, int width, float* __restrict__ map,
                                                       uint8_t* __restrict__ label, float* __restrict__ tmap) {
  __shared__ float4 sl[ATT_LINES_PER_PASS];
  const int index = blockIdx.x * blockDim.x + threadIdx.x;
  const int hw = height * width;
  const int w = index % width, h = index / width;
  const float xs = __fdiv_rn(static_cast<float>(width), static_cast<float>(input_width));
  const float ys = __fdiv_rn(static_cast<float>(height), static_cast<float>(input_height));
  const NearestLine b = nearest_line(lines, num, xs, ys, static_cast<float>(w), static_cast<float>(h), sl);
  if (index >= hw || b.idx < 0) return;
  map[index] = b.ax; map[hw + index] = b.ay;
  map[2 * hw + index] = b.ux; map[3 * hw + index] = b.uy;
  map[4 * hw + index] = b.vx; map[5 * hw + index] = b.vy;
  tmap[index] = b.t;
  label[static_cast<size_t>(b.idx) * hw + index] = b.inside ? 1 : 0;
}

// mask [HW] (bool), labels [HW] (int64), proj_points [HW,2] (x = column + ax, y = row + ay; zero where !mask).
// The reference's further conditions (pos_angle > 0, neg_angle < 0, scene_hawp_dataset.py:131-140) hold identically:
// after the clamps pos = (>= 1e-9, >= 1e-9) and neg = (>= 1e-9, <= -1e-9), so atan2 is in (0, pi/2) resp. (-pi/2, 0).
__global__ void __launch_bounds__(256) point_line_attraction_kernel(const float* __restrict__ lines, int num, int height,
                                                                    int width, float distance, uint8_t* __restrict__ mask,
                                                                    long long* __restrict__ labels,
                                                                    float* __restrict__ proj) {
  __shared__ float4 sl[ATT_LINES_PER_PASS];
  const int index = blockIdx.x * blockDim.x + threadIdx.x;
  const int hw = height * width;
  const int w = index % width, h = index / width;
  const NearestLine b = nearest_line(lines, num, 1.0f, 1.0f, static_cast<float>(w), static_cast<float>(h), sl);
  if (index >= hw) return;
  const bool on = b.idx >= 0 && b.inside;
  const float dismap = sqrtf(__fadd_rn(__fmul_rn(b.ax, b.ax), __fmul_rn(b.ay, b.ay)));
  const bool m = on && dismap <= distance;
  mask[index] = m ? 1 : 0;
  labels[index] = on ? b.idx : 0;  // labels_onehot.max(dim=0): index of the single True, 0 for an all-False column
  proj[2 * index] = m ? __fadd_rn(b.ax, static_cast<float>(w)) : 0.f;
  proj[2 * index + 1] = m ? __fadd_rn(b.ay, static_cast<float>(h)) : 0.f;
}

}  // namespace neat
