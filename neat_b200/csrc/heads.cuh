// RenderingNetwork.forward and AttractionFieldNetwork.forward (neat_wfr_rend_a.py:175-197, 235-255) as one
// fused tensor-core kernel per head:  input = [x, view (PE4 for the rendering head), normal | feature] where
// the feature tile (bf16 hi/lo operand tile written by sdf_render) is bulk-loaded straight into the A tile,
// 4 x (Linear + ReLU) + Linear, then sigmoid (rgb) or x + offsets (3D line end points).
#pragma once
#include "sdf_render.cuh"

namespace neat {

struct HeadSaveLayout {
  uint32_t aux;    // TILE_AUX_BYTES : aux part of u_0
  uint32_t u;      // [HL-1] x TILE_MAIN_BYTES : u_1 .. u_{HL-1}
  uint32_t total;
};
__host__ __device__ inline HeadSaveLayout head_save_layout(int HL) {
  HeadSaveLayout s;
  s.aux = 0;
  s.u = TILE_AUX_BYTES;
  s.total = s.u + (HL - 1) * TILE_MAIN_BYTES;
  return s;
}

struct HeadParams {
  Program prog;  // H_0 .. H_{HL-1}
  const uint8_t* packed;
  SdfQueryParams pts;        // point source; view dir of point pt is rays_d[pt / n_per_ray] unless `dirs` is set
  const float* dirs;         // optional explicit view dirs [M,3]
  const float* normals;      // [M,3]
  const uint8_t* feat_tiles; // [n_tiles][TILE_MAIN_BYTES]
  int head;                  // 0: rendering (sigmoid), 1: attraction (x + offsets)
  int HL, multires_view, out_dim;
  int training;
  uint8_t* save;             // training: [n_tiles][layout.total]
  float* out;                // [M, out_dim]
};

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) head_fwd_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  EngineSmem<STAGES>& sm =
      *reinterpret_cast<EngineSmem<STAGES>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  engine_init(sm);
  const int n_tiles = (p.pts.M + TILE_M - 1) / TILE_M;
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
    const HeadSaveLayout lay = head_save_layout(p.HL);
    const bool tr = p.training != 0;
    uint32_t in_phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int pt = tile * TILE_M + e.row;
      const bool valid = pt < p.pts.M;
      uint8_t* rec = tr ? p.save + static_cast<size_t>(tile) * lay.total : nullptr;
      float x[3] = {0.f, 0.f, 0.f};
      // ---- stage 0: feature tile by bulk copy into the main planes, [x, view, normal] into the aux columns
      epi_planes_free(sm, e);
      // The previous tile's shared-memory writes are already ordered before this bulk load through the mbarrier chain
      // (STS -> a_ready -> MMA -> tcgen05.commit -> d_ready -> here); the barrier below states the same ordering in a
      // form compute-sanitizer's racecheck can see (it does not follow tcgen05.commit arrivals) -- one bar.sync per tile.
      if (t > 0) epi_bar();
      if (e.lead) {
        const uint8_t* src = p.feat_tiles + static_cast<size_t>(tile) * TILE_MAIN_BYTES;
        mbar_arrive_expect_tx(&sm.in_ready, TILE_MAIN_BYTES);
        bulk_g2s(sm.a_hi, src, PLANE_MAIN_BYTES, &sm.in_ready);
        bulk_g2s(sm.a_lo, src + PLANE_MAIN_BYTES, PLANE_MAIN_BYTES, &sm.in_ready);
        // this CTA's NEXT feature tile starts moving from HBM to the L2 now (once per tile: head_fwd 0.515 -> 0.465 ms)
        if (t + 1 < my_tiles) bulk_prefetch_l2(src + static_cast<size_t>(gridDim.x) * TILE_MAIN_BYTES, TILE_MAIN_BYTES);
      }
      if (e.j == 0) {
        float d[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
        if (valid) {
          load_point(p.pts, pt, x);
          const float* dp = p.dirs ? p.dirs + 3 * static_cast<size_t>(pt)
                                   : p.pts.rays_d + 3 * static_cast<size_t>(pt / p.pts.n_per_ray);
          d[0] = dp[0]; d[1] = dp[1]; d[2] = dp[2];
          nrm[0] = p.normals[3 * pt]; nrm[1] = p.normals[3 * pt + 1]; nrm[2] = p.normals[3 * pt + 2];
        }
        float a[A_AUX_COLS];
#pragma unroll
        for (int i = 0; i < A_AUX_COLS; ++i) a[i] = 0.f;
        a[0] = x[0]; a[1] = x[1]; a[2] = x[2];
        a[3] = d[0]; a[4] = d[1]; a[5] = d[2];
        int nv = 3;
        if (p.head == 0 && p.multires_view > 0) {
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (j < p.multires_view) {
              const float f = static_cast<float>(1 << j);
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float sn, co;
                sincosf(d[c] * f, &sn, &co);
                a[3 + 3 + 6 * j + c] = sn;
                a[3 + 3 + 6 * j + 3 + c] = co;
              }
            }
          }
          nv = 3 + 6 * p.multires_view;
        }
#pragma unroll
        for (int i = 6; i < A_AUX_COLS - 2; ++i) {  // normals follow the view block (position known at run time)
          if (i == 3 + nv) { a[i] = nrm[0]; a[i + 1] = nrm[1]; a[i + 2] = nrm[2]; }
        }
#pragma unroll
        for (int i = 0; i < A_AUX_COLS / 8; ++i) store_a8<true>(sm.a_hi, sm.a_lo, e.row, A_MAIN_COLS + 8 * i, a + 8 * i);
        epi_publish_aux(sm);
        if (tr) {  // the aux part of u_0 for wgrad: bf16 pairs (engine.cuh), written from registers
#pragma unroll
          for (int i = 0; i < A_AUX_COLS / 8; ++i) stg_bf16_pairs8(rec + lay.aux, PLANE_AUX_BYTES, e.row, i, a + 8 * i);
        }
      }
      mbar_wait(&sm.in_ready, in_phase);  // the feature planes have landed (every thread observes the TMA completion)
      in_phase ^= 1;
      epi_publish_all(sm);

      for (int l = 0; l < p.HL - 1; ++l) {
        const Step st = p.prog.s[l];
        const float4* bias = reinterpret_cast<const float4*>(p.packed + st.w.bias_off);
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
          if (has) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b = __ldg(bias + (c0 >> 2) + j);
              acc[4 * j + 0] = fmaxf(fmaf(acc[4 * j + 0], st.comp, b.x), 0.f);   // comp: layout.h (RZ accumulation)
              acc[4 * j + 1] = fmaxf(fmaf(acc[4 * j + 1], st.comp, b.y), 0.f);
              acc[4 * j + 2] = fmaxf(fmaf(acc[4 * j + 2], st.comp, b.z), 0.f);
              acc[4 * j + 3] = fmaxf(fmaf(acc[4 * j + 3], st.comp, b.w), 0.f);
            }
            store_a16<true>(sm.a_hi, sm.a_lo, e.row, c0, acc);
          }
          epi_publish_group(sm, g);
          if (has && usave) {  // after the publish: nothing waits for the save record (bf16 pairs, see engine.cuh)
            stg_bf16_pairs8(usave, PLANE_MAIN_BYTES, e.row, c0 >> 3, acc);
            stg_bf16_pairs8(usave, PLANE_MAIN_BYTES, e.row, (c0 >> 3) + 1, acc + 8);
          }
        }
      }
      {
        const Step st = p.prog.s[p.HL - 1];
        const float* bias = reinterpret_cast<const float*>(p.packed + st.w.bias_off);
        epi_wait_d(sm, e);
        if (e.j == 0) {
          float acc[16];
          tmem_ld16(e.tm + st.d_col, acc);
          tmem_ld_wait();
          if (valid) {
            if (p.head == 0) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float v = fmaf(acc[c], st.comp, __ldg(bias + c));
                p.out[3 * static_cast<size_t>(pt) + c] = 1.0f / (1.0f + expf(-v));
              }
            } else {
#pragma unroll
              for (int c = 0; c < 6; ++c) p.out[6 * static_cast<size_t>(pt) + c] = x[c % 3] + fmaf(acc[c], st.comp, __ldg(bias + c));
            }
          }
        }
      }
    }
    if (e.lead) bulk_wait0();
  }
  engine_fini(sm);
}

}  // namespace neat
