// Hand-written backward of the three MLPs (no autograd graph), SURVEY.md Appendix A.
//
// head_bwd_kernel : reverse sweep of a ReLU head.  dL/d(out) -> z_bar_l tiles (saved for wgrad.cuh),
//                   dL/d(feature) (fp32, accumulated over the two heads) and dL/d(normal).
// sdf_bwd_kernel  : ImplicitNetwork double backward.  Given n_bar = dL/d(normal) and o_bar = dL/d(raw output):
//     tangent sweep  p_0 = J_pe (act * n_bar);  q_l = W_l p_l;  zhat_l = sigma''_l g_{l+1} q_l;  p_{l+1} = sigma'_l q_l
//     reverse sweep  z_bar_{L-1} = o_bar;  g = W_l^T z_bar_l;  z_bar_{l-1} = sigma'_{l-1} g + zhat_{l-1}
//   and saves p_l / z_bar_l tiles; the weight gradients dW_l = z_bar_l^T u_l + a_l^T p_l are then plain GEMMs
//   over points (wgrad.cuh).  sigma'' g_{l+1} is recovered from the saved a_l = sigma'_l g_{l+1}:
//   sigma''_l g_{l+1} = 100 (1 - sigma'_l) a_l.
#pragma once
#include "heads.cuh"

namespace neat {

// read 8 consecutive columns (one chunk) of a saved operand tile row as fp32 (hi + lo)
__device__ __forceinline__ void load_tile8(const uint8_t* hi_plane, const uint8_t* lo_plane, int chunk, int row, float* v) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi_plane + chunk * A_CHUNK_BYTES + row * 16);
  const uint4 l = *reinterpret_cast<const uint4*>(lo_plane + chunk * A_CHUNK_BYTES + row * 16);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    v[2 * i + 1] = __uint_as_float(hw[i] & 0xFFFF0000u) + __uint_as_float(lw[i] & 0xFFFF0000u);
  }
}
// ReLU mask of 8 columns from the hi plane of the saved ReLU output (bf16(u) != 0  <=>  u > 0)
__device__ __forceinline__ uint32_t load_mask8(const uint8_t* hi_plane, int chunk, int row) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi_plane + chunk * A_CHUNK_BYTES + row * 16);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((hw[i] & 0x7FFFu) != 0u ? 1u : 0u) << (2 * i);
    m |= ((hw[i] & 0x7FFF0000u) != 0u ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}

struct HeadBwdSaveLayout {
  uint32_t zb_aux;  // TILE_AUX_BYTES: dL/d(out) in aux columns (X operand of the last layer's dW)
  uint32_t zb;      // [HL-1] x TILE_MAIN_BYTES: z_bar_0 .. z_bar_{HL-2}
  uint32_t total;
};
__host__ __device__ inline HeadBwdSaveLayout head_bwd_layout(int HL) {
  HeadBwdSaveLayout s;
  s.zb_aux = 0;
  s.zb = TILE_AUX_BYTES;
  s.total = s.zb + (HL - 1) * TILE_MAIN_BYTES;
  return s;
}

struct HeadBwdParams {
  Program prog;  // HT_{HL-1}, ..., HT_1, HT_0 (feature part), HT_0^aux (normal part)
  const uint8_t* packed;
  int M, HL, out_dim;
  const float* out_bar;     // [M, out_dim]  dL/d(pre-activation output of the head)
  const uint8_t* fwd_save;  // head forward save records
  uint8_t* bwd_save;        // [n_tiles][head_bwd_layout.total]
  float* feat_bar;          // [n_tiles][256][128] fp32
  float* n_bar;             // [M,3]
  int accumulate;           // 0: overwrite feat_bar / n_bar, 1: add
};

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) head_bwd_kernel(const __grid_constant__ HeadBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  EngineSmem<STAGES>& sm =
      *reinterpret_cast<EngineSmem<STAGES>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  engine_init(sm);
  const int n_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4) {
    if (lane == 0) producer_loop(sm, p.prog, p.packed, my_tiles);
  } else if (warp == 5) {
    if (lane == 0) mma_loop(sm, p.prog, my_tiles);
  } else {
    const int row = threadIdx.x;
    const uint32_t tm = sm.tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const HeadSaveLayout fl = head_save_layout(p.HL);
    const HeadBwdSaveLayout bl = head_bwd_layout(p.HL);
    EpiState es;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int pt = tile * TILE_M + row;
      const bool valid = pt < p.M;
      const uint8_t* frec = p.fwd_save + static_cast<size_t>(tile) * fl.total;
      uint8_t* brec = p.bwd_save + static_cast<size_t>(tile) * bl.total;
      // ---- dL/d(out) -> aux columns 0..15
      if (row == 0) bulk_wait_read0();
      epi_bar();
      {
        float e[A_AUX_COLS];
#pragma unroll
        for (int i = 0; i < A_AUX_COLS; ++i) e[i] = 0.f;
        if (valid) {
#pragma unroll
          for (int c = 0; c < 6; ++c)
            if (c < p.out_dim) e[c] = p.out_bar[static_cast<size_t>(pt) * p.out_dim + c];
        }
#pragma unroll
        for (int i = 0; i < A_AUX_COLS / 8; ++i) store_a8(sm.a_hi, sm.a_lo, row, A_MAIN_COLS + 8 * i, e + 8 * i);
      }
      fence_proxy_async();
      epi_bar();
      if (row == 0) {
        bulk_s2g(brec + bl.zb_aux, sm.a_hi + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_s2g(brec + bl.zb_aux + PLANE_AUX_BYTES, sm.a_lo + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_commit();
      }
      epi_publish_a(sm);
      // ---- layers HL-1 .. 1: D = dL/d(u_l);  z_bar_{l-1} = D * [u_l > 0]
      for (int l = p.HL - 1; l >= 1; --l) {
        const int npad = p.prog.s[p.HL - 1 - l].w.npad;
        const uint8_t* u_hi = frec + fl.u + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES;
        epi_wait_d(sm, es);
        if (row == 0) bulk_wait_read0();
        epi_bar();
        for (int c0 = 0; c0 < npad; c0 += 32) {
          float acc[32];
          tmem_ld32(tm + c0, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t m = load_mask8(u_hi, (c0 >> 3) + j, row);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (!((m >> q) & 1u)) acc[8 * j + q] = 0.f;
          }
          store_a32(sm.a_hi, sm.a_lo, row, c0, acc);
        }
        fence_proxy_async();
        epi_bar();
        if (row == 0) {
          uint8_t* dst = brec + bl.zb + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES;
          bulk_s2g(dst, sm.a_hi, PLANE_MAIN_BYTES);
          bulk_s2g(dst + PLANE_MAIN_BYTES, sm.a_lo, PLANE_MAIN_BYTES);
          bulk_commit();
        }
        epi_publish_a(sm);
      }
      // ---- first layer: dL/d(feature) (D cols 0..F) and dL/d(normal) (D cols 256..258)
      {
        const int npad = p.prog.s[p.HL - 1].w.npad;
        epi_wait_d(sm, es);
        float* fb = p.feat_bar + static_cast<size_t>(tile) * (256 * TILE_M);
        for (int c0 = 0; c0 < npad; c0 += 32) {
          float acc[32];
          tmem_ld32(tm + c0, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float* dst = fb + (c0 + j) * TILE_M + row;
            *dst = p.accumulate ? *dst + acc[j] : acc[j];
          }
        }
        float acc[32];
        tmem_ld32(tm + 256, acc);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float* dst = p.n_bar + 3 * static_cast<size_t>(pt) + c;
            *dst = p.accumulate ? *dst + acc[c] : acc[c];
          }
        }
      }
    }
    if (row == 0) bulk_wait0();
  }
  engine_fini(sm);
}

// ------------------------------------------------------------------------------------------------
struct SdfBwdSaveLayout {
  uint32_t p_aux;   // TILE_AUX_BYTES   : p_0 (tangent seed, E columns)
  uint32_t p;       // [L-1] x TILE_MAIN_BYTES : p_1 .. p_{L-1}
  uint32_t zb_aux;  // TILE_AUX_BYTES   : sdf column of z_bar_{L-1}
  uint32_t zb;      // [L] x TILE_MAIN_BYTES   : z_bar_0 .. z_bar_{L-1} (the last one: feature columns)
  uint32_t total;
};
__host__ __device__ inline SdfBwdSaveLayout sdf_bwd_layout(int L) {
  SdfBwdSaveLayout s;
  s.p_aux = 0;
  s.p = TILE_AUX_BYTES;
  s.zb_aux = s.p + (L - 1) * TILE_MAIN_BYTES;
  s.zb = s.zb_aux + TILE_AUX_BYTES;
  s.total = s.zb + L * TILE_MAIN_BYTES;
  return s;
}

struct SdfBwdParams {
  Program prog;  // F_0 .. F_{L-2} (tangent), T_{L-1} .. T_1 (reverse)
  const uint8_t* packed;
  SdfQueryParams pts;
  int L, skip, H, E, F;
  const float* n_bar;     // [M,3]
  const float* s_bar;     // [M] or nullptr (eikonal points)
  const float* feat_bar;  // [n_tiles][256][128] or nullptr
  const float* act;       // [M] or nullptr (= 1)
  const uint8_t* fwd_save;  // sdf_render training records
  uint8_t* bwd_save;        // [n_tiles][sdf_bwd_layout.total]
  float* zhat;              // scratch [gridDim.x][L-1][256][128] fp32
};

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) sdf_bwd_kernel(const __grid_constant__ SdfBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  EngineSmem<STAGES>& sm =
      *reinterpret_cast<EngineSmem<STAGES>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  engine_init(sm);
  const int n_tiles = (p.pts.M + TILE_M - 1) / TILE_M;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;

  if (warp == 4) {
    if (lane == 0) producer_loop(sm, p.prog, p.packed, my_tiles);
  } else if (warp == 5) {
    if (lane == 0) mma_loop(sm, p.prog, my_tiles);
  } else {
    const int row = threadIdx.x;
    const uint32_t tm = sm.tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const SdfSaveLayout fl = sdf_save_layout(L, true);
    const SdfBwdSaveLayout bl = sdf_bwd_layout(L);
    float* zhat_base = p.zhat + static_cast<size_t>(blockIdx.x) * (L - 1) * (256 * TILE_M);
    EpiState es;
    auto store_tile = [&](uint8_t* dst) {  // all 128 threads: publish the main planes of A to global
      fence_proxy_async();
      epi_bar();
      if (row == 0) {
        bulk_s2g(dst, sm.a_hi, PLANE_MAIN_BYTES);
        bulk_s2g(dst + PLANE_MAIN_BYTES, sm.a_lo, PLANE_MAIN_BYTES);
        bulk_commit();
      }
    };
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int pt = tile * TILE_M + row;
      const bool valid = pt < p.pts.M;
      const uint8_t* frec = p.fwd_save + static_cast<size_t>(tile) * fl.total;
      uint8_t* brec = p.bwd_save + static_cast<size_t>(tile) * bl.total;
      const float* d1_base = reinterpret_cast<const float*>(frec + fl.d1);
      float x[3] = {0.f, 0.f, 0.f}, nb[3] = {0.f, 0.f, 0.f};
      float actv = 0.f;
      if (valid) {
        load_point(p.pts, pt, x);
        actv = p.act ? p.act[pt] : 1.f;
        nb[0] = actv * p.n_bar[3 * pt]; nb[1] = actv * p.n_bar[3 * pt + 1]; nb[2] = actv * p.n_bar[3 * pt + 2];
      }
      // ---------------------------------------------------------------- tangent seed p_0 = J (act * n_bar)
      if (row == 0) bulk_wait_read0();
      epi_bar();
      {
        float e[A_AUX_COLS];
#pragma unroll
        for (int i = 0; i < A_AUX_COLS; ++i) e[i] = 0.f;
        e[0] = nb[0]; e[1] = nb[1]; e[2] = nb[2];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          if (j < p.pts.multires) {
            const float f = static_cast<float>(1 << j);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float s, co;
              sincosf(x[c] * f, &s, &co);
              e[3 + 6 * j + c] = f * co * nb[c];
              e[3 + 6 * j + 3 + c] = -f * s * nb[c];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < A_AUX_COLS / 8; ++i) store_a8(sm.a_hi, sm.a_lo, row, A_MAIN_COLS + 8 * i, e + 8 * i);
      }
      fence_proxy_async();
      epi_bar();
      if (row == 0) {
        bulk_s2g(brec + bl.p_aux, sm.a_hi + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_s2g(brec + bl.p_aux + PLANE_AUX_BYTES, sm.a_lo + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_commit();
      }
      epi_publish_a(sm);
      // ---------------------------------------------------------------- tangent sweep, layers 0 .. L-2
      for (int l = 0; l < L - 1; ++l) {
        const int npad = p.prog.s[l].w.npad;
        const float* d1 = d1_base + static_cast<size_t>(l) * (256 * TILE_M);
        const uint8_t* a_hi = frec + fl.a + static_cast<size_t>(l) * TILE_MAIN_BYTES;
        const uint8_t* a_lo = a_hi + PLANE_MAIN_BYTES;
        float* zh = zhat_base + static_cast<size_t>(l) * (256 * TILE_M);
        epi_wait_d(sm, es);  // D = q_l = W_l p_l
        if (row == 0) bulk_wait_read0();
        epi_bar();
        for (int c0 = 0; c0 < npad; c0 += 32) {
          float q[32];
          tmem_ld32(tm + c0, q);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a[8];
            load_tile8(a_hi, a_lo, (c0 >> 3) + j, row, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = c0 + 8 * j + k;
              const float s1 = d1[c * TILE_M + row];
              zh[c * TILE_M + row] = SP_BETA * (1.0f - s1) * a[k] * q[8 * j + k];  // sigma'' g_{l+1} q_l
              q[8 * j + k] *= s1;                                                   // p_{l+1}
            }
          }
          store_a32(sm.a_hi, sm.a_lo, row, c0, q);
        }
        store_tile(brec + bl.p + static_cast<size_t>(l) * TILE_MAIN_BYTES);  // p_{l+1}
        if (l < L - 2) epi_publish_a(sm);                                     // -> F_{l+1}
      }
      // ---------------------------------------------------------------- z_bar_{L-1} = o_bar = [s_bar | feat_bar]
      if (row == 0) bulk_wait_read0();
      epi_bar();
      {
        const int npadF = p.prog.s[L - 1].w.nk_main * 16;  // feature columns read by T_{L-1}
        const float* fb = p.feat_bar ? p.feat_bar + static_cast<size_t>(tile) * (256 * TILE_M) : nullptr;
        for (int c0 = 0; c0 < npadF; c0 += 32) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (fb && valid) ? fb[(c0 + j) * TILE_M + row] : 0.f;
          store_a32(sm.a_hi, sm.a_lo, row, c0, v);
        }
        float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid && p.s_bar) e[0] = p.s_bar[pt];  // already masked by act (composite_bwd)
        store_a8(sm.a_hi, sm.a_lo, row, A_MAIN_COLS, e);
        const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_a8(sm.a_hi, sm.a_lo, row, A_MAIN_COLS + 8, zero8);
      }
      fence_proxy_async();
      epi_bar();
      if (row == 0) {
        uint8_t* dst = brec + bl.zb + static_cast<size_t>(L - 1) * TILE_MAIN_BYTES;
        bulk_s2g(dst, sm.a_hi, PLANE_MAIN_BYTES);
        bulk_s2g(dst + PLANE_MAIN_BYTES, sm.a_lo, PLANE_MAIN_BYTES);
        bulk_s2g(brec + bl.zb_aux, sm.a_hi + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_s2g(brec + bl.zb_aux + PLANE_AUX_BYTES, sm.a_lo + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_commit();
      }
      epi_publish_a(sm);  // -> T_{L-1}
      // ---------------------------------------------------------------- reverse sweep: layers L-1 .. 1
      for (int l = L - 1; l >= 1; --l) {
        const int ncols = p.prog.s[l - 1].w.npad;  // width of z_bar_{l-1}
        const float* d1 = d1_base + static_cast<size_t>(l - 1) * (256 * TILE_M);
        const float* zh = zhat_base + static_cast<size_t>(l - 1) * (256 * TILE_M);
        epi_wait_d(sm, es);  // D = W_l^T z_bar_l  (gradient w.r.t. the input of layer l)
        if (row == 0) bulk_wait_read0();
        epi_bar();
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          float g[32];
          tmem_ld32(tm + c0, g);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            g[j] = d1[c * TILE_M + row] * g[j] + zh[c * TILE_M + row];
          }
          store_a32(sm.a_hi, sm.a_lo, row, c0, g);
        }
        store_tile(brec + bl.zb + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES);  // z_bar_{l-1}
        if (l >= 2) epi_publish_a(sm);                                            // -> T_{l-1}
      }
    }
    if (row == 0) bulk_wait0();
  }
  engine_fini(sm);
}

}  // namespace neat
