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
  int x_main, y_main;  // 1: MAIN operand tile in the half-tile-contiguous global layout (engine.cuh: gtile_off), one bulk
                       // copy per plane and half; 0: aux tile in shared-memory order ([chunk][128 rows]): one 1 KB copy
                       // per chunk, plane and half (<= 6 chunks)
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

// Shared memory: a TWO-STAGE ring of HALF tiles (64 points): while the MMAs of one half run, the next half lands.
// A stage holds the hi / lo planes of X (128 columns) and Y (256 columns) in chunk-major order over 64 rows:
// element (pt, col) -> plane + (col / 8) * 1024 + pt * 16 + (col % 8) * 2 -- the bytes of the global layout, so a plane of
// a main operand is ONE bulk copy.  (Round 1 measured this pipeline with [chunk][128 rows] global tiles: 96 copies of
// 1 KB per stage, 1.66 -> 2.35 ms.  The single-stage kernel it replaces ran load and MMA back to back: 1.20 ms at 1024
// rays with the MMAs exposed, profiles/r02_whatif_timing.log.)
constexpr int WG_HALF = TILE_M / 2;                       // points per stage
constexpr int WG_CHUNK = WG_HALF * 16;                    // 1024
constexpr int WG_X_PLANE = 128 / 8 * WG_CHUNK;            // 16384
constexpr int WG_Y_PLANE = 256 / 8 * WG_CHUNK;            // 32768
constexpr int WG_ONES_BYTES = 2 * WG_CHUNK;               // 16 columns x 64 points
constexpr int WG_STAGES = 2;
constexpr int WG_THREADS = 192;                           // warps 0-3 flush, warp 4 loads, warp 5 issues MMAs

struct alignas(1024) WgradSmem {
  uint8_t x_hi[WG_STAGES][WG_X_PLANE], x_lo[WG_STAGES][WG_X_PLANE];
  uint8_t y_hi[WG_STAGES][WG_Y_PLANE], y_lo[WG_STAGES][WG_Y_PLANE];
  uint8_t ones[WG_ONES_BYTES];
  uint64_t full[WG_STAGES], empty[WG_STAGES], d_ready;
  uint32_t tmem_base;
};

// copies `n_chunks` chunks of one plane of one half tile into a stage plane; main layout: contiguous, aux: per chunk
__device__ __forceinline__ void wg_load_plane(uint8_t* dst, const uint8_t* tile_plane, int first_chunk, int n_chunks, int half,
                                              int is_main, uint64_t* bar) {
  if (is_main) {
    bulk_g2s(dst, tile_plane + half * G_HALF_PLANE_BYTES + first_chunk * G_CHUNK_BYTES, n_chunks * WG_CHUNK, bar);
  } else {
    for (int c = 0; c < n_chunks; ++c)
      bulk_g2s(dst + c * WG_CHUNK, tile_plane + (first_chunk + c) * A_CHUNK_BYTES + half * WG_CHUNK, WG_CHUNK, bar);
  }
}
__device__ __forceinline__ void wg_prefetch_plane(const uint8_t* tile_plane, int first_chunk, int n_chunks, int is_main) {
  if (is_main) {
    bulk_prefetch_l2(tile_plane + first_chunk * G_CHUNK_BYTES, n_chunks * WG_CHUNK);
    bulk_prefetch_l2(tile_plane + G_HALF_PLANE_BYTES + first_chunk * G_CHUNK_BYTES, n_chunks * WG_CHUNK);
  } else {
    bulk_prefetch_l2(tile_plane + first_chunk * A_CHUNK_BYTES, n_chunks * A_CHUNK_BYTES);
  }
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const WJob* __restrict__ jobs) {
  extern __shared__ uint8_t smem_raw[];
  WgradSmem& sm = *reinterpret_cast<WgradSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const WJob job = jobs[blockIdx.x];
  const int warp = warp_idx_uniform();
  const int my_tiles = job.n_tiles > job.split ? (job.n_tiles - job.split + job.n_split - 1) / job.n_split : 0;
  const int n_halves = 2 * my_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.d_ready, 1);
    fence_mbar_init();
  }
  // ones operand: element (pt, col) -> chunk(col/8) * 1024 + pt * 16 + (col % 8) * 2 ; column 0 = 1.0
  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) {
    const int byte = i * 4;
    const bool first = byte < WG_CHUNK && (byte % 16) == 0;
    reinterpret_cast<uint32_t*>(sm.ones)[i] = first ? 0x00003F80u : 0u;
  }
  // zero the part of the X planes that is never loaded (x_cols < 128)
  if (job.x_cols < 128) {
    const int from = job.x_cols / 8 * WG_CHUNK;
    for (int s = 0; s < WG_STAGES; ++s)
      for (int i = from / 16 + threadIdx.x; i < WG_X_PLANE / 16; i += blockDim.x) {
        reinterpret_cast<uint4*>(sm.x_hi[s])[i] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(sm.x_lo[s])[i] = make_uint4(0, 0, 0, 0);
      }
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int xc = job.x_cols / 8, yc = job.n_cols / 8;   // chunks per plane
  const uint32_t stage_bytes = static_cast<uint32_t>(2 * (xc + yc)) * WG_CHUNK;
  if (warp == 4) {  // load warp (converged; one elected lane issues the bulk copies)
    for (int h = 0; h < n_halves; ++h) {
      const int t = h >> 1, half = h & 1, s = h & (WG_STAGES - 1);
      const uint64_t tile = static_cast<uint64_t>(job.split) + static_cast<uint64_t>(t) * job.n_split;
      const uint8_t* xs = job.x_base + tile * job.x_stride;
      const uint8_t* ys = job.y_base + tile * job.y_stride;
      mbar_wait(&sm.empty[s], ((h / WG_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sm.full[s], stage_bytes);
        wg_load_plane(sm.x_hi[s], xs + job.x_hi, job.m0 / 8, xc, half, job.x_main, &sm.full[s]);
        wg_load_plane(sm.x_lo[s], xs + job.x_lo, job.m0 / 8, xc, half, job.x_main, &sm.full[s]);
        wg_load_plane(sm.y_hi[s], ys + job.y_hi, 0, yc, half, job.y_main, &sm.full[s]);
        wg_load_plane(sm.y_lo[s], ys + job.y_lo, 0, yc, half, job.y_main, &sm.full[s]);
        // the next tile is requested from HBM into the L2 while this one is multiplied.  With the two-stage ring the
        // distance that paid for the single-stage kernel (2 tiles) only evicts what is about to be used: measured at
        // 1024 rays 0 / 1 / 2 / 3 / 4 tiles ahead = 1.00 / 1.00 / 1.08 / 1.31 / 1.46 ms
        if (half == 0) {
          const int dist = g_l2_prefetch ? 1 : 0;
          for (int a = (t == 0 ? 1 : dist); a <= dist; ++a) {
            if (t + a >= my_tiles) break;
            const uint8_t* xn = xs + static_cast<uint64_t>(a) * job.n_split * job.x_stride;
            const uint8_t* yn = ys + static_cast<uint64_t>(a) * job.n_split * job.y_stride;
            wg_prefetch_plane(xn + job.x_hi, job.m0 / 8, xc, job.x_main);
            wg_prefetch_plane(xn + job.x_lo, job.m0 / 8, xc, job.x_main);
            wg_prefetch_plane(yn + job.y_hi, 0, yc, job.y_main);
            wg_prefetch_plane(yn + job.y_lo, 0, yc, job.y_main);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {  // MMA warp (converged; one elected lane issues)
    // MN-major operands: core matrix = 8 points (K) x 8 columns (16 B); K-direction stride 128 B,
    // MN-direction stride = one chunk (1024 B)
    const uint32_t idesc = make_idesc(128, job.n_cols, 1, 1, job.x_f16, job.y_f16);
    const uint32_t idesc1 = make_idesc(128, 16, 1, 1, job.x_f16, 0);  // the in-kernel ones operand is bf16
    const uint32_t d = __shfl_sync(0xffffffffu, sm.tmem_base, 0), d1 = d + 256;
    const uint64_t don0 = make_desc_k(smem_u32(sm.ones), 128, WG_CHUNK);
    const bool has_bias = job.bias != nullptr;
    const bool t_hl = !(g_dbg & 1u), t_lh = !(g_dbg & 2u);
    for (int h = 0; h < n_halves; ++h) {
      const int s = h & (WG_STAGES - 1);
      mbar_wait(&sm.full[s], (h / WG_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dxh0 = make_desc_k(smem_u32(sm.x_hi[s]), 128, WG_CHUNK), dxl0 = make_desc_k(smem_u32(sm.x_lo[s]), 128, WG_CHUNK);
        const uint64_t dyh0 = make_desc_k(smem_u32(sm.y_hi[s]), 128, WG_CHUNK), dyl0 = make_desc_k(smem_u32(sm.y_lo[s]), 128, WG_CHUNK);
#pragma unroll
        for (int ks = 0; ks < WG_HALF / 16; ++ks) {
          const uint32_t ko = ks * 256;  // 16 points = 2 core matrices of 128 B
          const uint64_t dxh = desc_advance(dxh0, ko), dxl = desc_advance(dxl0, ko);
          const uint64_t dyh = desc_advance(dyh0, ko), dyl = desc_advance(dyl0, ko);
          const uint32_t acc = (h | ks) ? 1u : 0u;
          umma_bf16(d, dxh, dyh, idesc, acc);
          if (t_hl) umma_bf16(d, dxh, dyl, idesc, 1u);
          if (t_lh) umma_bf16(d, dxl, dyh, idesc, 1u);
          if (has_bias) {
            const uint64_t don = desc_advance(don0, ko);
            umma_bf16(d1, dxh, don, idesc1, acc);
            if (t_lh) umma_bf16(d1, dxl, don, idesc1, 1u);
          }
        }
        umma_commit(&sm.empty[s]);
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
