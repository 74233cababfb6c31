// ImplicitNetwork.get_sdf_vals as one fused kernel (neat_wfr_rend_a.py:78-96, 131-137):
// point generation (x = o + z d) -> positional encoding -> L weight-normed Linear layers with
// Softplus(100) on the tensor cores (activations never leave the SM) -> sphere clamp -> sdf.
// This is the error-bound sampler's query (ray_sampler.py:146-151), the largest consumer of the step.
#pragma once
#include "engine.cuh"

namespace neat {

struct SdfQueryParams {
  Program prog;  // steps 0 .. L-1 (forward layers; the last one is the sdf row only)
  const uint8_t* packed;
  const float* x;       // explicit points [M,3], or nullptr to generate x = o + z d
  const float* rays_o;  // [R,3] (o_stride 3) or [3] (o_stride 0)
  const float* rays_d;  // [R,3]
  const float* z;       // [R, n_per_ray]
  float* sdf;           // [M]
  const int* skip_flag; // optional device word: non-zero -> the launch is a no-op (sampler already converged)
  int o_stride;
  int n_per_ray;
  int M;
  int multires;
  float sphere_r, sphere_scale;
  // dense-grid mode (grid_n[0] > 0): point idx = (iy * nx + ix) * nz + iz  ->  (X[ix], Y[iy], Z[iz]) with
  // A[i] = float(lo + i * step) in float64, the last one = hi: the raveled np.meshgrid(x, y, z) of np.linspace axes
  // that utils/plots.py:318-324 builds, generated in the kernel instead of materialising [M,3]
  int grid_n[3];
  double grid_lo[3], grid_hi[3], grid_step[3];
};

// positional encoding of x into the aux columns of the A tile (zero padded to 48 columns)
// gsave (training): the same 48 columns as bf16 pairs into the tile's save record (aux-tile layout)
__device__ __forceinline__ void pe_to_aux(uint8_t* a_hi, uint8_t* a_lo, int row, const float x[3], int multires,
                                          uint8_t* gsave = nullptr) {
  float e[A_AUX_COLS];
#pragma unroll
  for (int i = 0; i < A_AUX_COLS; ++i) e[i] = 0.f;
  e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    if (j < multires) {
      const float f = static_cast<float>(1 << j);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s, co;
        sincosf(x[c] * f, &s, &co);
        e[3 + 6 * j + c] = s;
        e[3 + 6 * j + 3 + c] = co;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < A_AUX_COLS / 8; ++i) store_a8<true>(a_hi, a_lo, row, A_MAIN_COLS + 8 * i, e + 8 * i);
  if (gsave) {
#pragma unroll
    for (int i = 0; i < A_AUX_COLS / 8; ++i) stg_bf16_pairs8(gsave, (A_AUX_COLS / 8) * A_CHUNK_BYTES, row, i, e + 8 * i);
  }
}

// the point handled by this thread: explicit, or o + z * d with the reference's rounding (mul, then add)
__device__ __forceinline__ void load_point(const SdfQueryParams& p, int pt, float x[3]) {
  if (p.grid_n[0] > 0) {
    const int iz = pt % p.grid_n[2], t = pt / p.grid_n[2];
    const int idx[3] = {t % p.grid_n[0], t / p.grid_n[0], iz};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      x[c] = static_cast<float>(idx[c] == p.grid_n[c] - 1 ? p.grid_hi[c] : p.grid_lo[c] + idx[c] * p.grid_step[c]);
  } else if (p.x) {
    x[0] = p.x[3 * pt + 0]; x[1] = p.x[3 * pt + 1]; x[2] = p.x[3 * pt + 2];
  } else {
    const int r = pt / p.n_per_ray;
    const float zz = p.z[pt];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      x[c] = __fadd_rn(p.rays_o[static_cast<size_t>(r) * p.o_stride + c], __fmul_rn(zz, p.rays_d[3 * r + c]));
  }
}

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) sdf_query_kernel(const __grid_constant__ SdfQueryParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (p.skip_flag && *p.skip_flag) return;
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
    for (int t = 0; t < my_tiles; ++t) {
      const int pt = (blockIdx.x + t * gridDim.x) * TILE_M + e.row;
      const bool valid = pt < p.M;
      float x[3] = {0.f, 0.f, 0.f};
      // ---- stage 0: the group-0 warps generate the point and its positional encoding (aux columns)
      if (e.j == 0) {
        if (valid) load_point(p, pt, x);
        pe_to_aux(sm.a_hi, sm.a_lo, e.row, x, p.multires);
        epi_publish_aux(sm);
      }
      epi_publish_all(sm);
      // ---- hidden layers: each warp turns its 64 accumulator columns into the next layer's input
      for (int l = 0; l < p.prog.n - 1; ++l) {
        const Step st = p.prog.s[l];
        const float4* bias = reinterpret_cast<const float4*>(p.packed + st.w.bias_off);
        epi_wait_d(sm, e);
        // software pipeline over the column groups: the TMEM load of group g+1 is in flight while group g is computed
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
          if (c0 < st.w.npad) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b = __ldg(bias + (c0 >> 2) + j);
              acc[4 * j + 0] = softplus100(fmaf(acc[4 * j + 0], st.comp, b.x));   // comp: layout.h (RZ accumulation)
              acc[4 * j + 1] = softplus100(fmaf(acc[4 * j + 1], st.comp, b.y));
              acc[4 * j + 2] = softplus100(fmaf(acc[4 * j + 2], st.comp, b.z));
              acc[4 * j + 3] = softplus100(fmaf(acc[4 * j + 3], st.comp, b.w));
            }
            store_a16<true>(sm.a_hi, sm.a_lo, e.row, c0, acc);
          }
          epi_publish_group(sm, g);
        }
      }
      {
        const Step st = p.prog.s[p.prog.n - 1];
        epi_wait_d(sm, e);
        if (e.j == 0) {
          float acc[16];
          tmem_ld16(e.tm + st.d_col, acc);
          tmem_ld_wait();
          float s = fmaf(acc[0], st.comp, __ldg(reinterpret_cast<const float*>(p.packed + st.w.bias_off)));
          if (p.sphere_r > 0.f) {
            const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            s = fminf(s, p.sphere_scale * (p.sphere_r - nrm));
          }
          if (valid) p.sdf[pt] = s;
        }
      }
    }
  }
  engine_fini(sm);
}

}  // namespace neat
