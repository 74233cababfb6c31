// Weight-gradient GEMMs of the hand-written backward:   dW[n][k] += scale * sum_pt X[pt][n] * Y[pt][k]
// where X (dL/dz of a layer, or the normal-pass vector a_l) and Y (the layer input u_l, or the tangent p_l)
// are operand tiles saved by the forward / backward kernels (bf16 hi/lo planes, chunk-major, see engine.cuh).
// The reduction runs over points, so both operands are MN-major for tcgen05 and the very bytes written as a
// K-major A tile are reused unchanged.  One CTA owns (job, 128-row half of dW, split of the tile range),
// accumulates in TMEM over all its tiles, then flushes once with vector atomics.  A 16-column "ones" operand
// appended to Y yields the bias gradient db[n] = sum_pt X[pt][n] from the same MMAs.
#pragma once
#include "engine.cuh"

namespace neat {

struct WJob {
  const uint8_t* x_base;  // per tile: x_base + tile * x_stride (+ x_hi / x_lo) = plane of the X segment
  const uint8_t* y_base;
  uint64_t x_stride, y_stride;
  uint32_t x_hi, x_lo, y_hi, y_lo;
  int m0;        // first X column handled (multiple of 8); this CTA covers columns [m0, m0+128)
  int x_cols;    // columns available in the X plane from m0 (multiple of 8, <= 128); the rest is zero-filled
  int n_cols;    // Y columns fed to the MMA (multiple of 16, <= 256)
  int x_valid;   // rows of dW written:   i < x_valid
  int y_valid;   // columns of dW written: j < y_valid
  float* out;    // out[(row0 + i) * ld + col0 + j] += scale * D[i][j]
  float* bias;   // bias[row0 + i] += scale * sum_pt X[pt][m0 + i]   (nullptr: none)
  int ld, row0, col0;
  float scale;
  int n_tiles;
  int split, n_split;
  int x_f16, y_f16;  // operand formats (umma.cuh): tiles written by the forward kernels hold fp16 pairs, by the backward
                     // kernels bf16 pairs; the instruction descriptor carries one format per operand
};

// 16-byte vector reduction into global memory (sm_90+): one L2 transaction instead of four
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// scalar fire-and-forget reduction; a plain atomicAdd on a generic pointer compiles to a RETURNING ATOM plus a
// shared-memory CAS fallback loop
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
// adds scale * v[0..31] to row[0..31] (columns c0.. of a dW row), skipping columns >= n_valid.  MIS = float index of
// row[0] modulo 4 (the same for every row of a job whose leading dimension is a multiple of 4): the first (4 - MIS) % 4
// elements and the ragged tail go out as scalar REDs, everything else as 16-byte vector REDs.  MIS < 0: scalar only.
template <int MIS>
__device__ __forceinline__ void flush_row32(float* row, const float* v, int n_valid, float scale) {
  constexpr int HEAD = MIS < 0 ? 32 : (4 - MIS) % 4;
#pragma unroll
  for (int j = 0; j < HEAD; ++j)
    if (j < n_valid) red_add_f32(row + j, scale * v[j]);
#pragma unroll
  for (int j = HEAD; j + 3 < 32; j += 4) {
    if (j + 3 < n_valid) {
      red_add_v4(row + j, scale * v[j], scale * v[j + 1], scale * v[j + 2], scale * v[j + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (j + k < n_valid) red_add_f32(row + j + k, scale * v[j + k]);
    }
  }
  constexpr int TAIL = HEAD + (32 - HEAD) / 4 * 4;
#pragma unroll
  for (int j = TAIL; j < 32; ++j)
    if (j < n_valid) red_add_f32(row + j, scale * v[j]);
}

constexpr int WG_X_PLANE = 128 / 8 * A_CHUNK_BYTES;  // 32768
constexpr int WG_Y_PLANE = 256 / 8 * A_CHUNK_BYTES;    // 65536
constexpr int WG_ONES_BYTES = 2 * A_CHUNK_BYTES;  // 16 columns
constexpr int WG_THREADS = 192;                   // warps 0-3 flush, warp 4 loads, warp 5 issues MMAs

struct alignas(1024) WgradSmem {
  uint8_t x_hi[WG_X_PLANE], x_lo[WG_X_PLANE];
  uint8_t y_hi[WG_Y_PLANE], y_lo[WG_Y_PLANE];
  uint8_t ones[WG_ONES_BYTES];
  uint64_t full, empty, d_ready;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const WJob* __restrict__ jobs) {
  extern __shared__ uint8_t smem_raw[];
  WgradSmem& sm = *reinterpret_cast<WgradSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const WJob job = jobs[blockIdx.x];
  const int warp = warp_idx_uniform();
  const int my_tiles = job.n_tiles > job.split ? (job.n_tiles - job.split + job.n_split - 1) / job.n_split : 0;

  if (threadIdx.x == 0) {
    mbar_init(&sm.full, 1);
    mbar_init(&sm.empty, 1);
    mbar_init(&sm.d_ready, 1);
    fence_mbar_init();
  }
  // ones operand: element (pt, col) -> chunk(col/8) * 2048 + pt * 16 + (col % 8) * 2 ; column 0 = 1.0
  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) {
    const int byte = i * 4;
    const bool first = byte < A_CHUNK_BYTES && (byte % 16) == 0;
    reinterpret_cast<uint32_t*>(sm.ones)[i] = first ? 0x00003F80u : 0u;
  }
  // zero the part of the X planes that is never loaded (x_cols < 128)
  if (job.x_cols < 128) {
    const int from = job.x_cols / 8 * A_CHUNK_BYTES;
    for (int i = from / 16 + threadIdx.x; i < WG_X_PLANE / 16; i += blockDim.x) {
      reinterpret_cast<uint4*>(sm.x_hi)[i] = make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(sm.x_lo)[i] = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t xb = static_cast<uint32_t>(job.x_cols / 8) * A_CHUNK_BYTES;
  const uint32_t yb = static_cast<uint32_t>(job.n_cols / 8) * A_CHUNK_BYTES;
  if (warp == 4) {  // load warp (converged; one elected lane issues the bulk copies)
    for (int t = 0; t < my_tiles; ++t) {
      const uint64_t tile = static_cast<uint64_t>(job.split) + static_cast<uint64_t>(t) * job.n_split;
      const uint8_t* xs = job.x_base + tile * job.x_stride + static_cast<uint64_t>(job.m0 / 8) * A_CHUNK_BYTES;
      const uint8_t* ys = job.y_base + tile * job.y_stride;
      mbar_wait(&sm.empty, (t & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sm.full, 2 * xb + 2 * yb);
        bulk_g2s(sm.x_hi, xs + job.x_hi, xb, &sm.full);
        bulk_g2s(sm.x_lo, xs + job.x_lo, xb, &sm.full);
        bulk_g2s(sm.y_hi, ys + job.y_hi, yb, &sm.full);
        bulk_g2s(sm.y_lo, ys + job.y_lo, yb, &sm.full);
        // the tile `dist` ahead streams from HBM into the L2 while this one is multiplied (g_l2_prefetch = dist; the
        // first iteration also requests the tiles in between)
        const int dist = g_l2_prefetch;
        for (int a = (t == 0 ? 1 : dist); a <= dist; ++a) {
          if (t + a >= my_tiles) break;
          const uint8_t* xn = xs + static_cast<uint64_t>(a) * job.n_split * job.x_stride;
          const uint8_t* yn = ys + static_cast<uint64_t>(a) * job.n_split * job.y_stride;
          bulk_prefetch_l2(xn + job.x_hi, xb);
          bulk_prefetch_l2(xn + job.x_lo, xb);
          bulk_prefetch_l2(yn + job.y_hi, yb);
          bulk_prefetch_l2(yn + job.y_lo, yb);
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {  // MMA warp (converged; one elected lane issues)
    // MN-major operands: core matrix = 8 points (K) x 8 columns (16 B); K-direction stride 128 B,
    // MN-direction stride = one chunk (2048 B)
    const uint32_t idesc = make_idesc(128, job.n_cols, 1, 1, job.x_f16, job.y_f16);
    const uint32_t idesc1 = make_idesc(128, 16, 1, 1, job.x_f16, 0);  // the in-kernel ones operand is bf16
    const uint32_t d = __shfl_sync(0xffffffffu, sm.tmem_base, 0), d1 = d + 256;
    const uint64_t dxh0 = make_desc_k(smem_u32(sm.x_hi), 128, A_CHUNK_BYTES), dxl0 = make_desc_k(smem_u32(sm.x_lo), 128, A_CHUNK_BYTES);
    const uint64_t dyh0 = make_desc_k(smem_u32(sm.y_hi), 128, A_CHUNK_BYTES), dyl0 = make_desc_k(smem_u32(sm.y_lo), 128, A_CHUNK_BYTES);
    const uint64_t don0 = make_desc_k(smem_u32(sm.ones), 128, A_CHUNK_BYTES);
    const bool has_bias = job.bias != nullptr;
    const bool t_hl = !(g_dbg & 1u), t_lh = !(g_dbg & 2u);
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(&sm.full, t & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < TILE_M / 16; ++ks) {
          const uint32_t ko = ks * 256;  // 16 points = 2 core matrices of 128 B
          const uint64_t dxh = desc_advance(dxh0, ko), dxl = desc_advance(dxl0, ko);
          const uint64_t dyh = desc_advance(dyh0, ko), dyl = desc_advance(dyl0, ko);
          const uint32_t acc = (t | ks) ? 1u : 0u;
          umma_bf16(d, dxh, dyh, idesc, acc);
          if (t_hl) umma_bf16(d, dxh, dyl, idesc, 1u);
          if (t_lh) umma_bf16(d, dxl, dyh, idesc, 1u);
          if (has_bias) {
            const uint64_t don = desc_advance(don0, ko);
            umma_bf16(d1, dxh, don, idesc1, acc);
            if (t_lh) umma_bf16(d1, dxl, don, idesc1, 1u);
          }
        }
        umma_commit(&sm.empty);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&sm.d_ready);
    __syncwarp();
  } else if (warp < 4) {
    if (my_tiles > 0) {
      mbar_wait(&sm.d_ready, 0);
      tc_fence_after();
      const int i = threadIdx.x;  // row of this dW block
      const uint32_t tm = sm.tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
      const bool row_ok = i < job.x_valid;
      float* orow = job.out + static_cast<size_t>(job.row0 + i) * job.ld + job.col0;
      // alignment class of the job's rows (uniform when ld % 4 == 0; the 32-column blocks keep it)
      const int mis = (job.ld & 3) == 0 ? static_cast<int>((reinterpret_cast<uintptr_t>(job.out + job.col0) >> 2) & 3) : -1;
      for (int c0 = 0; c0 < job.n_cols; c0 += 32) {
        float v[32];
        tmem_ld32(tm + c0, v);   // n_cols is a multiple of 16: the upper half of the last load may be stale
        tmem_ld_wait();
        if (row_ok) {
          const int nv = job.y_valid - c0;  // valid columns of this 32-column block
          switch (mis) {
            case 0: flush_row32<0>(orow + c0, v, nv, job.scale); break;
            case 1: flush_row32<1>(orow + c0, v, nv, job.scale); break;
            case 2: flush_row32<2>(orow + c0, v, nv, job.scale); break;
            case 3: flush_row32<3>(orow + c0, v, nv, job.scale); break;
            default: flush_row32<-1>(orow + c0, v, nv, job.scale); break;
          }
        }
      }
      if (job.bias) {
        float v[32];
        tmem_ld32(tm + 256, v);
        tmem_ld_wait();
        if (row_ok) red_add_f32(job.bias + job.row0 + i, job.scale * v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(sm.tmem_base, TMEM_COLS);
}

}  // namespace neat
