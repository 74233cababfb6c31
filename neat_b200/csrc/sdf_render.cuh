// ImplicitNetwork.get_outputs / .gradient as one fused kernel (neat_wfr_rend_a.py:98-129):
//   forward pass  F : PE -> L Linear layers (+Softplus(100)), keeping sigma'(z_l) per layer
//   normal pass   N : d sdf / d x by the hand-derived reverse recurrence
//                     a_{L-2} = sigma'_{L-2} * W_{L-1}[0,:],  v_l = a_l W_l,  a_{l-1} = sigma'_{l-1} * v_l,
//                     n = J_pe^T (v_0 + skip part)
// Both passes run on the tensor cores with the point tile resident in shared memory; no autograd graph.
// Outputs per point: sdf (sphere-clamped), normal (unnormalised), feature tile (bf16 hi/lo, tile layout).
// In training mode the kernel also writes what the hand-written backward needs (per 128-point tile):
//   sigma'_l (16-bit fixed point), the layer inputs u_l and the normal-pass vectors a_l (bf16 hi/lo operand tiles).
#pragma once
#include "engine.cuh"
#include "sdf_query.cuh"

namespace neat {

constexpr int PLANE_MAIN_BYTES = A_MAIN_COLS / 8 * A_CHUNK_BYTES;  // 65536
constexpr int PLANE_AUX_BYTES = A_AUX_COLS / 8 * A_CHUNK_BYTES;    // 12288
constexpr int TILE_MAIN_BYTES = 2 * PLANE_MAIN_BYTES;              // hi + lo
constexpr int TILE_AUX_BYTES = 2 * PLANE_AUX_BYTES;
constexpr int D1_BYTES = S1_LAYER_BYTES;                           // sigma' of one layer, 16-bit fixed point (engine.cuh)
constexpr int RSKIP_BYTES = A_AUX_COLS * TILE_M * 4;

// Byte layout of one tile's save record (training) / one CTA's scratch (inference)
struct SdfSaveLayout {
  uint32_t d1;     // [L-1] x D1_BYTES
  uint32_t rskip;  // RSKIP_BYTES
  uint32_t pe;     // TILE_AUX_BYTES   (u_0 and the aux part of u_skip)
  uint32_t u;      // [L-1] x TILE_MAIN_BYTES : u_1 .. u_{L-1}   (training only)
  uint32_t a;      // [L-1] x TILE_MAIN_BYTES : a_0 .. a_{L-2}   (training only)
  uint32_t feat;   // TILE_MAIN_BYTES : the feature vectors as bf16 pairs for wgrad (training only; the heads read the
                   // fp16 tile in SdfRenderParams::feat_tiles)
  uint32_t total;  // bytes per tile
};
__host__ __device__ inline SdfSaveLayout sdf_save_layout(int L, bool training) {
  SdfSaveLayout s;
  s.d1 = 0;
  s.rskip = s.d1 + (L - 1) * D1_BYTES;
  s.pe = s.rskip + RSKIP_BYTES;
  s.u = s.pe + TILE_AUX_BYTES;
  s.a = s.u + (training ? (L - 1) * TILE_MAIN_BYTES : 0);
  s.feat = s.a + (training ? (L - 1) * TILE_MAIN_BYTES : 0);
  s.total = s.feat + (training ? TILE_MAIN_BYTES : 0);
  return s;
}

struct SdfRenderParams {
  Program prog;  // F_0..F_{L-2}, F_{L-1}^feat, F_{L-1}^sdf, T_{L-2}, ..., T_0
  const uint8_t* packed;
  SdfQueryParams pts;  // point source (x | rays_o, rays_d, z), M, multires, sphere_*; prog/packed unused
  int L, skip, H, E, F;
  int clamp;     // 1: get_outputs (sphere clamp), 0: gradient() (eikonal points)
  int training;  // 1: write the full save record per tile
  uint32_t w_last_row_off;
  float* sdf;          // [M] or nullptr
  float* grad;         // [M,3]
  float* act;          // [M] 1 where the network branch of min() is active (training; may be nullptr)
  uint8_t* feat_tiles; // [n_tiles][TILE_MAIN_BYTES] or nullptr
  uint8_t* save;       // training: [n_tiles][layout.total]; inference: [gridDim.x][layout.total]
};

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) sdf_render_kernel(const __grid_constant__ SdfRenderParams p) {
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
    const SdfSaveLayout lay = sdf_save_layout(L, p.training != 0);
    const float* __restrict__ w_row = reinterpret_cast<const float*>(p.packed + p.w_last_row_off);
    const bool tr = p.training != 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int pt = tile * TILE_M + e.row;
      const bool valid = pt < p.pts.M;
      uint8_t* rec = p.save + static_cast<size_t>(tr ? tile : static_cast<int>(blockIdx.x)) * lay.total;
      uint8_t* d1_base = rec + lay.d1;
      float* rskip = reinterpret_cast<float*>(rec + lay.rskip);
      float x[3] = {0.f, 0.f, 0.f};
      if (e.j == 0 && valid) load_point(p.pts, pt, x);

      // ------------------------------------------------------------ stage 0: positional encoding
      epi_planes_free(sm, e);
      if (e.j == 0) {
        pe_to_aux(sm.a_hi, sm.a_lo, e.row, x, p.pts.multires, tr ? rec + lay.pe : nullptr);
        epi_publish_aux(sm);
      }
      epi_publish_all(sm);

      // ------------------------------------------------------------ forward pass, hidden layers
      for (int l = 0; l < L - 1; ++l) {
        const Step st = p.prog.s[l];
        const float4* bias = reinterpret_cast<const float4*>(p.packed + st.w.bias_off);
        uint8_t* d1 = d1_base + static_cast<size_t>(l) * D1_BYTES;
        uint8_t* usave = tr ? rec + lay.u + static_cast<size_t>(l) * TILE_MAIN_BYTES : nullptr;  // u_{l+1}
        epi_wait_d(sm, e);
        // (TMEM columns past npad are allocated but hold stale data: loaded unconditionally, never used)
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
          const bool has = c0 < st.w.npad;
          uint4 s1p[2];
          if (has) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float dd[8];
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const float4 b = __ldg(bias + (c0 >> 2) + 2 * h + j);
                const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float hv;
                  softplus100_d1(fmaf(acc[8 * h + 4 * j + k], st.comp, bb[k]), hv, dd[4 * j + k]);   // comp: layout.h (RZ accumulation)
                  acc[8 * h + 4 * j + k] = hv;
                }
              }
              s1p[h] = s1_pack8(dd);
            }
            store_a16<true>(sm.a_hi, sm.a_lo, e.row, c0, acc);
          }
          epi_publish_group(sm, g);
          if (has) {  // after the publish: nothing waits for the records (sigma': read back by this thread's normal pass)
            *s1_at(d1, e, 2 * g) = s1p[0];
            *s1_at(d1, e, 2 * g + 1) = s1p[1];
            if (usave) {  // bf16 pairs, see engine.cuh
              stg_bf16_pairs8(usave, PLANE_MAIN_BYTES, e.row, c0 >> 3, acc);
              stg_bf16_pairs8(usave, PLANE_MAIN_BYTES, e.row, (c0 >> 3) + 1, acc + 8);
            }
          }
        }
      }

      // ------------------------------------------------------------ last layer: sdf (step L-1) + features (step L)
      float s_raw = 0.f;
      {
        const Step ss = p.prog.s[L - 1], sf = p.prog.s[L];
        const float4* bias = reinterpret_cast<const float4*>(p.packed + sf.w.bias_off);
        epi_wait_d(sm, e);
        if (e.j == 0) {
          float acc[16];
          tmem_ld16(e.tm + ss.d_col, acc);
          tmem_ld_wait();
          s_raw = fmaf(acc[0], ss.comp, __ldg(reinterpret_cast<const float*>(p.packed + ss.w.bias_off)));
        }
        for (int g = 0; g < N_GROUPS; ++g) {
          const int c0 = epi_col(e, g);
          if (c0 < sf.w.npad) {
            float acc[16];
            tmem_ld16(e.tm + sf.d_col + c0, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b = __ldg(bias + (c0 >> 2) + j);
              acc[4 * j + 0] = fmaf(acc[4 * j + 0], sf.comp, b.x); acc[4 * j + 1] = fmaf(acc[4 * j + 1], sf.comp, b.y);
              acc[4 * j + 2] = fmaf(acc[4 * j + 2], sf.comp, b.z); acc[4 * j + 3] = fmaf(acc[4 * j + 3], sf.comp, b.w);
            }
            store_a16<true>(sm.a_hi, sm.a_lo, e.row, c0, acc);
            if (tr) {
              uint8_t* fsave = rec + lay.feat;
              stg_bf16_pairs8(fsave, PLANE_MAIN_BYTES, e.row, c0 >> 3, acc);
              stg_bf16_pairs8(fsave, PLANE_MAIN_BYTES, e.row, (c0 >> 3) + 1, acc + 8);
            }
          }
        }
        fence_proxy_async();
        tc_fence_before();
        epi_bar();
        if (e.lead && p.feat_tiles) {
          uint8_t* dst = p.feat_tiles + static_cast<size_t>(tile) * TILE_MAIN_BYTES;
          bulk_s2g(dst, sm.a_hi, PLANE_MAIN_BYTES);
          bulk_s2g(dst + PLANE_MAIN_BYTES, sm.a_lo, PLANE_MAIN_BYTES);
          bulk_commit();
          bulk_wait_read0();
        }
        epi_bar();
        tc_fence_after();
      }

      // ------------------------------------------------------------ normal pass seed: a_{L-2} = sigma'_{L-2} * W_{L-1}[0,:]
      {
        const uint8_t* d1 = d1_base + static_cast<size_t>(L - 2) * D1_BYTES;
        const int npad = p.prog.s[L - 2].w.npad;
        for (int g = 0; g < N_GROUPS; ++g) {
          const int c0 = epi_col(e, g);
          if (c0 < npad) {
            float a[16];
            s1_unpack8(*s1_at(d1, e, 2 * g), a);
            s1_unpack8(*s1_at(d1, e, 2 * g + 1), a + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] *= __ldg(w_row + c0 + j);
            store_a16<true>(sm.a_hi, sm.a_lo, e.row, c0, a);
            if (tr) {
              uint8_t* asave = rec + lay.a + static_cast<size_t>(L - 2) * TILE_MAIN_BYTES;
              stg_bf16_pairs8(asave, PLANE_MAIN_BYTES, e.row, c0 >> 3, a);
              stg_bf16_pairs8(asave, PLANE_MAIN_BYTES, e.row, (c0 >> 3) + 1, a + 8);
            }
          }
          epi_publish_group(sm, g);
        }
      }
      // ------------------------------------------------------------ transposed layers l = L-2 .. 1: D = v_l -> a_{l-1}
      for (int l = L - 2; l >= 1; --l) {
        const Step st = p.prog.s[(L + 1) + (L - 2 - l)];
        const int npad = st.w.npad;                            // width of the input of layer l
        const uint8_t* d1 = d1_base + static_cast<size_t>(l - 1) * D1_BYTES;
        const int n_main = (l == p.skip) ? p.H - p.E : npad;   // columns that feed a_{l-1}
        // sigma'_{l-1} (16 bytes per unit, written by this very thread in the forward pass) is requested PF units ahead,
        // the first PF before waiting for the accumulator: the round trips (L2 or HBM, ~1 us each) overlap the MMAs
        // instead of forming a chain of eight as they did when taken one unit ahead (whole layer ahead: spills)
#ifndef NEAT_RENDER_PF
#define NEAT_RENDER_PF 4
#endif
        constexpr int PF = NEAT_RENDER_PF;
        uint4 s1p[PF];
        // (at the skip layer the transposed GEMM is wider than the layer below: its extra columns are the skip part,
        // which needs no sigma' and whose record was never written)
        const int n_s1 = min(npad, static_cast<int>(p.prog.s[l - 1].w.npad));
        auto issue = [&](int u) {
          s1p[u % PF] = epi_unit_col(e, u) < n_s1 ? *s1_at(d1, e, u) : make_uint4(0u, 0u, 0u, 0u);
        };
#pragma unroll
        for (int u = 0; u < PF; ++u) issue(u);
        uint8_t* asave = tr ? rec + lay.a + static_cast<size_t>(l - 1) * TILE_MAIN_BYTES : nullptr;  // a_{l-1}
        epi_wait_d(sm, e);
#pragma unroll
        for (int u = 0; u < N_UNITS; ++u) {
          const int c = epi_unit_col(e, u);
          const uint4 s1u = s1p[u % PF];
          if (u + PF < N_UNITS) issue(u + PF);
          if (c < npad) {
            float s1[8];
            s1_unpack8(s1u, s1);
            float acc[8];
            tmem_ld8(e.tm + st.d_col + c, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= st.comp;
            if (l == p.skip) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int cc = c + j;
                if (cc >= n_main) rskip[(cc - n_main) * TILE_M + e.row] = acc[j];
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= s1[j];
            store_a8<true>(sm.a_hi, sm.a_lo, e.row, c, acc);
            if (asave) stg_bf16_pairs8(asave, PLANE_MAIN_BYTES, e.row, c >> 3, acc);
          }
          if (u & 1) epi_publish_group(sm, u >> 1);
        }
      }
      // ------------------------------------------------------------ layer 0: D = v_0 [E]; n = J^T (v_0 + r)
      {
        const Step st = p.prog.s[2 * L - 1];
        epi_wait_d(sm, e);
        if (p.skip >= 0) {  // the skip part was written by other warps (last column group): make it visible
          __threadfence_block();
          epi_bar();
        }
        if (e.j == 0) {
          float n[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            float v[16];
            tmem_ld16(e.tm + st.d_col + 16 * part, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = 16 * part + i;  // compile-time
              float val = v[i] * st.comp;
              if (p.skip >= 0 && c < p.E) val += rskip[c * TILE_M + e.row];
              if (c < 3) {
                n[c] += val;
              } else {
                const int j = (c - 3) / 6, r6 = (c - 3) % 6, cc = r6 % 3;
                if (j < p.pts.multires) {
                  const float f = static_cast<float>(1 << j);
                  float sn, co;
                  sincosf(x[cc] * f, &sn, &co);
                  n[cc] += (r6 < 3 ? f * co : -f * sn) * val;
                }
              }
            }
          }
          float sdf = s_raw, actv = 1.f;
          if (p.clamp && p.pts.sphere_r > 0.f) {
            const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            const float sph = p.pts.sphere_scale * (p.pts.sphere_r - nrm);
            if (!(s_raw <= sph)) {  // torch.minimum routes the gradient to `self` on ties
              actv = 0.f;
              sdf = sph;
              const float k = -p.pts.sphere_scale / nrm;
              n[0] = k * x[0]; n[1] = k * x[1]; n[2] = k * x[2];
            }
          }
          if (valid) {
            if (p.sdf) p.sdf[pt] = sdf;
            p.grad[3 * pt + 0] = n[0]; p.grad[3 * pt + 1] = n[1]; p.grad[3 * pt + 2] = n[2];
            if (p.act) p.act[pt] = actv;
          }
        }
      }
    }
    if (e.lead) bulk_wait0();
  }
  engine_fini(sm);
}

}  // namespace neat
