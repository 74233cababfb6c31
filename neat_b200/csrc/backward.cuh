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
__device__ __forceinline__ uint32_t mask8(const uint4& h) {
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
  float* feat_bar;          // [n_tiles][64][128][4] fp32 (engine.cuh: f4_at)
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
  const int warp = warp_idx_uniform();

  if (warp == EPI_WARPS) {
    reg_dec<AUX_REGS>();
    producer_loop(sm, p.prog, p.packed, my_tiles);
  } else if (warp == EPI_WARPS + 1) {
    reg_dec<AUX_REGS>();
    mma_loop(sm, p.prog, my_tiles);
  } else if (warp > EPI_WARPS + 1) {
    reg_dec<AUX_REGS>();   // the two idle warps of the fifth warpgroup (layout.h)
  } else {
    reg_inc<EPI_REGS>();
    Epi e = epi_make(sm);
    const HeadSaveLayout fl = head_save_layout(p.HL);
    const HeadBwdSaveLayout bl = head_bwd_layout(p.HL);
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int pt = tile * TILE_M + e.row;
      const bool valid = pt < p.M;
      const uint8_t* frec = p.fwd_save + static_cast<size_t>(tile) * fl.total;
      uint8_t* brec = p.bwd_save + static_cast<size_t>(tile) * bl.total;
      // ---- stage 0: dL/d(out) -> aux columns 0..15
      epi_planes_free(sm, e);
      if (e.j == 0) {
        float a[A_AUX_COLS];
#pragma unroll
        for (int i = 0; i < A_AUX_COLS; ++i) a[i] = 0.f;
        if (valid) {
#pragma unroll
          for (int c = 0; c < 6; ++c)
            if (c < p.out_dim) a[c] = p.out_bar[static_cast<size_t>(pt) * p.out_dim + c];
        }
#pragma unroll
        for (int i = 0; i < A_AUX_COLS / 8; ++i) store_a8<false>(sm.a_hi, sm.a_lo, e.row, A_MAIN_COLS + 8 * i, a + 8 * i);
        epi_publish_aux(sm);
      }
      epi_publish_all(sm);
      epi_store_main(sm, e, sm.a_hi + PLANE_MAIN_BYTES, sm.a_lo + PLANE_MAIN_BYTES, brec + bl.zb_aux, PLANE_AUX_BYTES);
      // ---- layers HL-1 .. 1: D = dL/d(u_l);  z_bar_{l-1} = D * [u_l > 0]
      for (int l = p.HL - 1; l >= 1; --l) {
        const Step st = p.prog.s[p.HL - 1 - l];
        const uint8_t* __restrict__ u_hi = frec + fl.u + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES;
        // the saved ReLU outputs (masks) of all four groups are requested before waiting for the accumulator
        uint4 mraw[N_GROUPS][2];
#pragma unroll
        for (int g = 0; g < N_GROUPS; ++g) {
          const int c0 = epi_col(e, g);
          if (c0 < st.w.npad) {
#pragma unroll
            for (int j = 0; j < 2; ++j) mraw[g][j] = ldg128(u_hi + gtile_off((c0 >> 3) + j, e.row));
          }
        }
        uint8_t* zsave = brec + bl.zb + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES;  // z_bar_{l-1}
        epi_wait_d(sm, e);
        uint32_t mbits[2] = {0u, 0u};  // 16 mask bits per group (the raw vectors are dead before the accumulator loads)
#pragma unroll
        for (int g = 0; g < N_GROUPS; ++g) {
          if (epi_col(e, g) < st.w.npad)
            mbits[g >> 1] |= (mask8(mraw[g][0]) | (mask8(mraw[g][1]) << 8)) << (16 * (g & 1));
        }
        float nxt[16];
        tmem_ld16(e.tm + st.d_col + epi_col(e, 0), nxt);
#pragma unroll
        for (int g = 0; g < N_GROUPS; ++g) {
          const int c0 = epi_col(e, g);
          float acc[16];
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = nxt[j];
          if (g + 1 < N_GROUPS) tmem_ld16(e.tm + st.d_col + epi_col(e, g + 1), nxt);
          if (c0 < st.w.npad) {
            const uint32_t m = mbits[g >> 1] >> (16 * (g & 1));
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (!((m >> k) & 1u)) acc[k] = 0.f;
            store_a16_save<false>(sm.a_hi, sm.a_lo, zsave, e.row, c0, acc);  // (global stores after the publish: spills here)
          }
          epi_publish_group(sm, g);
        }
      }
      // ---- first layer: dL/d(feature) (step HL-1) and dL/d(normal) (step HL)
      {
        const Step sf = p.prog.s[p.HL - 1], sa = p.prog.s[p.HL];
        epi_wait_d(sm, e);
        float* fb = p.feat_bar + static_cast<size_t>(tile) * (256 * TILE_M);
        for (int g = 0; g < N_GROUPS; ++g) {
          const int c0 = epi_col(e, g);
          if (c0 < sf.w.npad) {
            float acc[16];
            tmem_ld16(e.tm + sf.d_col + c0, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* q = f4_at(fb, c0 + 4 * j, e.row);
              float4 o = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
              if (p.accumulate) {
                const float4 old = *q;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *q = o;
            }
          }
        }
        if (e.j == 0) {
          float acc[16];
          tmem_ld16(e.tm + sa.d_col, acc);
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
    }
    if (e.lead) bulk_wait0();
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
  const float* feat_bar;  // [n_tiles][64][128][4] (f4_at) or nullptr
  const float* act;       // [M] or nullptr (= 1)
  const uint8_t* fwd_save;  // sdf_render training records
  uint8_t* bwd_save;        // [n_tiles][sdf_bwd_layout.total]
  float* zhat;              // scratch [gridDim.x][L-1][256][128] fp32
};

#ifndef NEAT_SDF_BWD_PF
#define NEAT_SDF_BWD_PF 1
#endif
constexpr int SDF_BWD_PF = NEAT_SDF_BWD_PF;  // prefetch distance of the sweeps' saved-tensor loads, in 8-column units

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) sdf_bwd_kernel(const __grid_constant__ SdfBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  EngineSmem<STAGES>& sm =
      *reinterpret_cast<EngineSmem<STAGES>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  engine_init(sm);
  const int n_tiles = (p.pts.M + TILE_M - 1) / TILE_M;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int warp = warp_idx_uniform();
  const int L = p.L;

  if (warp == EPI_WARPS) {
    reg_dec<AUX_REGS>();
    producer_loop(sm, p.prog, p.packed, my_tiles);
  } else if (warp == EPI_WARPS + 1) {
    reg_dec<AUX_REGS>();
    mma_loop(sm, p.prog, my_tiles);
  } else if (warp > EPI_WARPS + 1) {
    reg_dec<AUX_REGS>();   // the two idle warps of the fifth warpgroup (layout.h)
  } else {
    reg_inc<EPI_REGS>();
    Epi e = epi_make(sm);
    const SdfSaveLayout fl = sdf_save_layout(L, true);
    const SdfBwdSaveLayout bl = sdf_bwd_layout(L);
#ifdef NEAT_ZHAT_HINT
    const uint64_t zh_pol = l2_policy_evict_last();
#endif
    uint8_t* __restrict__ zhat_base = reinterpret_cast<uint8_t*>(p.zhat) + static_cast<size_t>(blockIdx.x) * (L - 1) * (2 * ZH_PLANE_BYTES);
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int pt = tile * TILE_M + e.row;
      const bool valid = pt < p.pts.M;
      const uint8_t* __restrict__ frec = p.fwd_save + static_cast<size_t>(tile) * fl.total;
      uint8_t* brec = p.bwd_save + static_cast<size_t>(tile) * bl.total;
      const uint8_t* __restrict__ d1_base = frec + fl.d1;   // sigma' records: 16-bit fixed point (engine.cuh)
      // ---------------------------------------------------------------- stage 0: tangent seed p_0 = J (act * n_bar)
      epi_planes_free(sm, e);
      if (e.j == 0) {
        float x[3] = {0.f, 0.f, 0.f}, nb[3] = {0.f, 0.f, 0.f};
        if (valid) {
          load_point(p.pts, pt, x);
          const float actv = p.act ? p.act[pt] : 1.f;
          nb[0] = actv * p.n_bar[3 * pt]; nb[1] = actv * p.n_bar[3 * pt + 1]; nb[2] = actv * p.n_bar[3 * pt + 2];
        }
        float a[A_AUX_COLS];
#pragma unroll
        for (int i = 0; i < A_AUX_COLS; ++i) a[i] = 0.f;
        a[0] = nb[0]; a[1] = nb[1]; a[2] = nb[2];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          if (j < p.pts.multires) {
            const float f = static_cast<float>(1 << j);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float sn, co;
              sincosf(x[c] * f, &sn, &co);
              a[3 + 6 * j + c] = f * co * nb[c];
              a[3 + 6 * j + 3 + c] = -f * sn * nb[c];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < A_AUX_COLS / 8; ++i) store_a8<false>(sm.a_hi, sm.a_lo, e.row, A_MAIN_COLS + 8 * i, a + 8 * i);
        epi_publish_aux(sm);
      }
      epi_publish_all(sm);
      epi_store_main(sm, e, sm.a_hi + PLANE_MAIN_BYTES, sm.a_lo + PLANE_MAIN_BYTES, brec + bl.p_aux, PLANE_AUX_BYTES);
      // ---------------------------------------------------------------- tangent sweep, layers 0 .. L-2
      for (int l = 0; l < L - 1; ++l) {
        const Step st = p.prog.s[l];
        const uint8_t* __restrict__ d1 = d1_base + static_cast<size_t>(l) * D1_BYTES;
        const uint8_t* __restrict__ a_hi = frec + fl.a + static_cast<size_t>(l) * TILE_MAIN_BYTES;
        const uint8_t* __restrict__ a_lo = a_hi + PLANE_MAIN_BYTES;
        uint8_t* __restrict__ zh = zhat_base + static_cast<size_t>(l) * (2 * ZH_PLANE_BYTES);
        const int npad = st.w.npad;
        // Loads of the saved tensors run SDF_BWD_PF units ahead of their use (the first ones before the accumulator wait):
        // every one is an HBM round trip, and taken one unit ahead they formed a chain of eight per layer.  After full
        // unrolling the arrays below are plain registers with short live ranges (3 x 12 in flight).
        uint4 sp[SDF_BWD_PF + 1], ahv[SDF_BWD_PF + 1], alv[SDF_BWD_PF + 1];   // slot = unit % (PF + 1)
        auto issue = [&](int u) {
          const int c = epi_unit_col(e, u), k = u % (SDF_BWD_PF + 1);
          if (c < npad) {
            sp[k] = __ldg(s1_at(d1, e, u));
            ahv[k] = ldg128(a_hi + gunit_off(e, u));
            alv[k] = ldg128(a_lo + gunit_off(e, u));
          }
        };
        uint8_t* psave = brec + bl.p + static_cast<size_t>(l) * TILE_MAIN_BYTES;  // p_{l+1}
#pragma unroll
        for (int u = 0; u < SDF_BWD_PF; ++u) issue(u);
        epi_wait_d(sm, e);  // D = q_l = W_l p_l
#pragma unroll
        for (int u = 0; u < N_UNITS; ++u) {
          const int c = epi_unit_col(e, u);
          constexpr int PFS = SDF_BWD_PF + 1;
          const uint4 spu = sp[u % PFS], ahu = ahv[u % PFS], alu = alv[u % PFS];
          if (u + SDF_BWD_PF < N_UNITS) issue(u + SDF_BWD_PF);
          if (c < npad) {
            float a[8], q[8], s1[8];
            tmem_ld8(e.tm + st.d_col + c, q);
            unpack_hilo8<false>(ahu, alu, a);
            s1_unpack8(spu, s1);
            tmem_ld_wait();
            float zv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              zv[i] = SP_BETA * (1.0f - s1[i]) * a[i] * q[i];  // zhat_l = sigma'' g_{l+1} q_l
              q[i] *= s1[i];                                     // p_{l+1}
            }
#ifdef NEAT_ZHAT_HINT
            stg128_hint(zh_at(zh, 0, e, u), make_float4(zv[0], zv[1], zv[2], zv[3]), zh_pol);
            stg128_hint(zh_at(zh, 1, e, u), make_float4(zv[4], zv[5], zv[6], zv[7]), zh_pol);
#else
            *zh_at(zh, 0, e, u) = make_float4(zv[0], zv[1], zv[2], zv[3]);
#ifndef NEAT_WHATIF_ZHAT_HALF
            *zh_at(zh, 1, e, u) = make_float4(zv[4], zv[5], zv[6], zv[7]);
#endif
#endif
            store_a8_save<false>(sm.a_hi, sm.a_lo, psave, e.row, c, q);
          }
          if ((u & 1) && l < L - 2) epi_publish_group(sm, u >> 1);  // -> F_{l+1}
        }
      }
      // ---------------------------------------------------------------- z_bar_{L-1} = o_bar = [s_bar | feat_bar]
      epi_planes_free(sm, e);
      {
        const int npadF = p.prog.s[L - 1].w.nk_main * 16;  // feature columns read by T_{L-1}
        const float* __restrict__ fb = p.feat_bar ? p.feat_bar + static_cast<size_t>(tile) * (256 * TILE_M) : nullptr;
        for (int g = 0; g < N_GROUPS; ++g) {
          const int c0 = epi_col(e, g);
          if (c0 < npadF) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t = fb ? __ldg(f4_at(fb, c0 + 4 * j, e.row)) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * j] = valid ? t.x : 0.f; v[4 * j + 1] = valid ? t.y : 0.f;
              v[4 * j + 2] = valid ? t.z : 0.f; v[4 * j + 3] = valid ? t.w : 0.f;
            }
            store_a16_save<false>(sm.a_hi, sm.a_lo, brec + bl.zb + static_cast<size_t>(L - 1) * TILE_MAIN_BYTES, e.row, c0, v);
          }
        }
        if (e.j == 0) {
          float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (valid && p.s_bar) a8[0] = p.s_bar[pt];  // already masked by act (composite_bwd)
          store_a8<false>(sm.a_hi, sm.a_lo, e.row, A_MAIN_COLS, a8);
          store_a8<false>(sm.a_hi, sm.a_lo, e.row, A_MAIN_COLS + 8, zero8);
          epi_publish_aux(sm);
        }
      }
      epi_publish_all(sm);  // -> T_{L-1}
      // the main columns of z_bar_{L-1} went out from registers above (global operand-tile layout); the aux columns
      // (the sdf column) leave as one bulk store out of the aux planes
      epi_wrote(sm);
      if (e.lead) {
        mbar_wait(&sm.wr_done, e.wr_phase);
        bulk_s2g(brec + bl.zb_aux, sm.a_hi + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_s2g(brec + bl.zb_aux + PLANE_AUX_BYTES, sm.a_lo + PLANE_MAIN_BYTES, PLANE_AUX_BYTES);
        bulk_commit();
      }
      e.wr_phase ^= 1;
      // ---------------------------------------------------------------- reverse sweep: layers L-1 .. 1
      for (int l = L - 1; l >= 1; --l) {
        const Step st = p.prog.s[(L - 1) + (L - 1 - l)];
        const int ncols = p.prog.s[l - 1].w.npad;  // width of z_bar_{l-1}
        const uint8_t* __restrict__ d1 = d1_base + static_cast<size_t>(l - 1) * D1_BYTES;
        uint8_t* zh = zhat_base + static_cast<size_t>(l - 1) * (2 * ZH_PLANE_BYTES);
        uint4 sp[SDF_BWD_PF + 1];
        float4 z0[SDF_BWD_PF + 1], z1[SDF_BWD_PF + 1];
        auto issue = [&](int u) {
          const int c = epi_unit_col(e, u), k = u % (SDF_BWD_PF + 1);
          if (c < ncols) {
            sp[k] = __ldg(s1_at(d1, e, u));
#ifdef NEAT_ZHAT_HINT
            z0[k] = ldg128_hint(zh_at(zh, 0, e, u), zh_pol);
            z1[k] = ldg128_hint(zh_at(zh, 1, e, u), zh_pol);
#else
            z0[k] = *zh_at(zh, 0, e, u);  // written by this very thread in the tangent sweep
#ifndef NEAT_WHATIF_ZHAT_HALF
            z1[k] = *zh_at(zh, 1, e, u);
#else
            z1[k] = z0[k];
#endif
#endif
          }
        };
        uint8_t* zsave = brec + bl.zb + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES;  // z_bar_{l-1}
#pragma unroll
        for (int u = 0; u < SDF_BWD_PF; ++u) issue(u);
        epi_wait_d(sm, e);  // D = W_l^T z_bar_l  (gradient w.r.t. the input of layer l)
        if (l == L - 1) epi_planes_free(sm, e);  // the bulk store of z_bar_{L-1} out of the A planes (above)
#pragma unroll
        for (int u = 0; u < N_UNITS; ++u) {
          const int c = epi_unit_col(e, u);
          constexpr int PFS = SDF_BWD_PF + 1;
          const uint4 spu = sp[u % PFS];
          const float4 z0u = z0[u % PFS], z1u = z1[u % PFS];
          if (u + SDF_BWD_PF < N_UNITS) issue(u + SDF_BWD_PF);
          if (c < ncols) {
            float s1[8], zz[8];
            s1_unpack8(spu, s1);
            f4_unpack(z0u, zz);
            f4_unpack(z1u, zz + 4);
            float gq[8];
            tmem_ld8(e.tm + st.d_col + c, gq);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) gq[j] = s1[j] * gq[j] + zz[j];
            store_a8_save<false>(sm.a_hi, sm.a_lo, zsave, e.row, c, gq);
          }
          if ((u & 1) && l >= 2) epi_publish_group(sm, u >> 1);  // -> T_{l-1}
        }
      }
    }
    if (e.lead) bulk_wait0();
  }
  engine_fini(sm);
}

}  // namespace neat
