// Tile-MLP engine shared by all tensor-core kernels.
//
// One CTA owns a tile of 128 points (TMEM lane i == tile row i == thread i of the 4 epilogue warps).
// The activation tile ("A") stays resident in shared memory as two bf16 planes (hi / lo split of the
// fp32 value) in the UMMA no-swizzle canonical layout, chunk-major:
//     element (row r, column k)  ->  plane + (k / 8) * 2048 + r * 16 + (k % 8) * 2      [bytes]
// i.e. an 8-column chunk of all 128 rows is one contiguous 2 KB block.  The same bytes are a valid
// K-major operand (M = rows, K = columns: layer GEMMs) and a valid MN-major operand
// (M/N = columns, K = rows: weight-gradient GEMMs).
// Columns [0,256) are the "main" segment, columns [256,304) the "aux" segment (positional encoding,
// view dirs, normals ... whatever the first layer of a net concatenates to its main input).
//
// Weights are pre-packed (pack.cu) into per-k-step slabs in exactly the B-operand layout, so a slab is
// one contiguous cp.async.bulk (1-D TMA) into a ring of shared-memory stages:
//     slab(k-step) = hi plane [2 chunks][npad rows][8] bf16, then lo plane (same shape)
// Warp roles: warps 0-3 epilogue (TMEM -> registers -> activation -> A tile), warp 4 lane 0 weight
// producer (TMA), warp 5 lane 0 MMA issuer (tcgen05.mma, 3 MMAs per k-step: hi*hi + hi*lo + lo*hi).
#pragma once
#include "layout.h"
#include "umma.cuh"

namespace neat {

// Debug knob (neat_debug_set_desc_swap): exchanges the LBO / SBO fields of every matrix descriptor.
__constant__ int g_desc_swap = 0;
// k_stride: bytes between core matrices adjacent along K; mn_stride: along M/N  (K-major operands)
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr, uint32_t k_stride, uint32_t mn_stride) {
  return g_desc_swap ? make_desc(saddr, mn_stride, k_stride) : make_desc(saddr, k_stride, mn_stride);
}

template <int STAGES>
struct alignas(1024) EngineSmem {
  uint8_t a_hi[A_PLANE_BYTES];
  uint8_t a_lo[A_PLANE_BYTES];
  uint8_t w[STAGES][W_STAGE_BYTES];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t a_ready;
  uint64_t d_ready;
  uint64_t in_ready;  // bulk loads of input operand tiles (heads / backward kernels)
  uint32_t tmem_base;
};

template <int STAGES>
__device__ __forceinline__ void engine_init(EngineSmem<STAGES>& sm) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    mbar_init(&sm.a_ready, TILE_M);
    mbar_init(&sm.d_ready, 1);
    mbar_init(&sm.in_ready, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

template <int STAGES>
__device__ __forceinline__ void engine_fini(EngineSmem<STAGES>& sm) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 4) tmem_dealloc(sm.tmem_base, TMEM_COLS);
}

// warp 4, lane 0: stream every slab of every step, for every tile this CTA owns
template <int STAGES>
__device__ __forceinline__ void producer_loop(EngineSmem<STAGES>& sm, const Program& prog, const uint8_t* packed,
                                              int n_tiles) {
  uint32_t stage = 0, phase = 0;
  for (int t = 0; t < n_tiles; ++t) {
    for (int i = 0; i < prog.n; ++i) {
      const PLayer w = prog.s[i].w;
      const uint32_t slab = static_cast<uint32_t>(w.npad) * 64u;
      const int nk = w.nk_main + w.nk_aux;
      const uint8_t* src = packed + w.off;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(&sm.empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&sm.full[stage], slab);
        bulk_g2s(sm.w[stage], src + static_cast<size_t>(ks) * slab, slab, &sm.full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
}

// warp 5, lane 0: issue the MMAs
template <int STAGES>
__device__ __forceinline__ void mma_loop(EngineSmem<STAGES>& sm, const Program& prog, int n_tiles) {
  uint32_t stage = 0, phase = 0, a_phase = 0;
  const uint32_t tmem = sm.tmem_base;
  const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
  for (int t = 0; t < n_tiles; ++t) {
    for (int i = 0; i < prog.n; ++i) {
      const Step st = prog.s[i];
      if (st.wait_a) {
        mbar_wait(&sm.a_ready, a_phase);
        a_phase ^= 1;
        tc_fence_after();
      }
      const uint32_t npad = st.w.npad;
      const uint32_t idesc = make_idesc(TILE_M, npad, 0, 0);
      const uint32_t d = tmem + st.d_col;
      const int nk = st.w.nk_main + st.w.nk_aux;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(&sm.full[stage], phase);
        tc_fence_after();
        const uint32_t col0 = ks < st.w.nk_main ? 16u * ks : A_MAIN_COLS + 16u * (ks - st.w.nk_main);
        const uint32_t a_off = (col0 >> 3) * A_CHUNK_BYTES;
        const uint64_t da_hi = make_desc_k(a_hi + a_off, A_CHUNK_BYTES, 128);
        const uint64_t da_lo = make_desc_k(a_lo + a_off, A_CHUNK_BYTES, 128);
        const uint32_t wb = smem_u32(sm.w[stage]);
        const uint64_t db_hi = make_desc_k(wb, npad * 16, 128);
        const uint64_t db_lo = make_desc_k(wb + npad * 32, npad * 16, 128);
        umma_bf16(d, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(d, da_hi, db_lo, idesc, 1u);
        umma_bf16(d, da_lo, db_hi, idesc, 1u);
        umma_commit(&sm.empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (st.commit_d) umma_commit(&sm.d_ready);
    }
  }
}

// ---------------------------------------------------------------- epilogue-side helpers (warps 0-3)
struct EpiState {
  uint32_t d_phase = 0;
};

template <int STAGES>
__device__ __forceinline__ void epi_wait_d(EngineSmem<STAGES>& sm, EpiState& es) {
  mbar_wait(&sm.d_ready, es.d_phase);
  es.d_phase ^= 1;
  tc_fence_after();
}
// call after the thread finished writing its row of the A tile (and reading the accumulator)
template <int STAGES>
__device__ __forceinline__ void epi_publish_a(EngineSmem<STAGES>& sm) {
  fence_proxy_async();
  tc_fence_before();
  mbar_arrive(&sm.a_ready);
}

// write 32 consecutive columns [c0, c0+32) of row `row` (c0 % 8 == 0) into the A tile
__device__ __forceinline__ void store_a32(uint8_t* a_hi, uint8_t* a_lo, int row, int c0, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 hi, lo;
    split8(v + 8 * j, hi, lo);
    const int off = ((c0 >> 3) + j) * A_CHUNK_BYTES + row * 16;
    *reinterpret_cast<uint4*>(a_hi + off) = hi;
    *reinterpret_cast<uint4*>(a_lo + off) = lo;
  }
}
__device__ __forceinline__ void store_a8(uint8_t* a_hi, uint8_t* a_lo, int row, int c0, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  const int off = (c0 >> 3) * A_CHUNK_BYTES + row * 16;
  *reinterpret_cast<uint4*>(a_hi + off) = hi;
  *reinterpret_cast<uint4*>(a_lo + off) = lo;
}

// ---------------------------------------------------------------- activations
constexpr float SP_BETA = 100.0f;
constexpr float SP_THRESH = 20.0f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// nn.Softplus(beta=100, threshold=20):  z if 100 z > 20 else log1p(exp(100 z)) / 100
__device__ __forceinline__ float softplus100(float z) {
  const float bz = z * SP_BETA;
  const float e = exp2f(fminf(bz, SP_THRESH) * LOG2E);
  const float sp = __log2f(1.0f + e) * (LN2 / SP_BETA);
  return bz > SP_THRESH ? z : sp;
}
// softplus and its derivative sigmoid(100 z) (1 above the threshold, as autograd computes it)
__device__ __forceinline__ void softplus100_d1(float z, float& h, float& d1) {
  const float bz = z * SP_BETA;
  const float e = exp2f(fminf(bz, SP_THRESH) * LOG2E);
  const float ope = 1.0f + e;
  const float sp = __log2f(ope) * (LN2 / SP_BETA);
  const float sg = __fdividef(e, ope);
  const bool lin = bz > SP_THRESH;
  h = lin ? z : sp;
  d1 = lin ? 1.0f : sg;
}

// NeRF positional encoding of a 3-vector: [x, sin(2^j x), cos(2^j x)]_{j<L}  (embedder.py:5-36)
// out must hold 3 + 6 L floats
__device__ __forceinline__ void embed3(const float x[3], int L, float* out) {
  out[0] = x[0]; out[1] = x[1]; out[2] = x[2];
  float f = 1.0f;
  for (int j = 0; j < L; ++j) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(x[c] * f, &s, &co);
      out[3 + 6 * j + c] = s;
      out[3 + 6 * j + 3 + c] = co;
    }
    f *= 2.0f;
  }
}

}  // namespace neat
