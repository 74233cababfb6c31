// Small kernels that take the host (and eager PyTorch) out of the training step so that the step can be replayed as CUDA
// graphs (neat_b200/trainer.py FusedTrainStep):
//   train_draws_kernel  : every random draw of one training forward in ONE launch, Philox4x32-10 keyed by (seed, a step
//                         counter that lives on the device and is bumped by the kernel itself)
//                         -- replaces torch.rand x3 + topk + randint + uniform_ (code/model/ray_sampler.py:87,234,265,275,
//                         code/model/networks/neat_wfr_rend_a.py:518)
//   gemm_f32_kernel     : fp32 SIMT GEMM with bias / ReLU / ReLU-mask epilogues: the junction `ffn` (3 Linear layers on the
//                         1024 latents, neat_wfr_rend_a.py:274-303, 488) forward and backward -- the last cuBLAS call of the step
//   adam_prepare/adam_graph_kernel : torch.optim.Adam with the step count and the hyper-parameters read from device memory
//   junction_terms with a device-side pair count (the matched-junction count changes every step)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "adam.cuh"
#include "pixels.cuh"

namespace neat {

// ---------------------------------------------------------------------------------------------------- Philox4x32-10
struct Philox {
  uint32_t k0, k1;
};
__device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// 4 x 32 random bits for counter (a, b, c, d) under key (k0, k1)
__device__ __forceinline__ void philox4(uint32_t a, uint32_t b, uint32_t c_, uint32_t d, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {a, b, c_, d};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// torch.rand's float32 recipe: 24 random mantissa bits, [0, 1)
__device__ __forceinline__ float u01(uint32_t x) { return static_cast<float>(x >> 8) * (1.0f / 16777216.0f); }

struct DrawParams {
  int R, n_eval, n_final, n_extra, max_iters, n_out;
  float radius;                 // eikonal points ~ U(-radius, radius)^3
  unsigned long long seed;
  unsigned long long* counter;  // [2] device: [0] = draws made so far (bumped here), [1] = block ticket
  float* t_rand;                // [R, n_eval]   stratified jitter       (ray_sampler.py:87)
  float* u_final;               // [R, n_final]  inverse-CDF uniforms    (:234)
  long long* extra_idx;         // [max_iters, n_extra]: row k-1 = the first n_extra entries of a uniform random
                                //   permutation of [0, n_eval * k)       (:265, randperm(L)[:n_extra] for whichever k)
  long long* eik_idx;           // [R] in [0, n_out)                      (:275)
  float* eik_uniform;           // [R, 3]                                 (neat_wfr_rend_a.py:518)
};

// stream ids keep the sub-draws independent; one thread produces 4 values of a stream
__global__ void __launch_bounds__(256) train_draws_kernel(DrawParams p) {
  const unsigned long long step = p.counter[0];
  const uint32_t k0 = static_cast<uint32_t>(p.seed), k1 = static_cast<uint32_t>(p.seed >> 32);
  const uint32_t s_lo = static_cast<uint32_t>(step), s_hi = static_cast<uint32_t>(step >> 32);
  const long long n_t = static_cast<long long>(p.R) * p.n_eval, n_u = static_cast<long long>(p.R) * p.n_final;
  const long long n_e = 3LL * p.R;
  const long long q_t = (n_t + 3) / 4, q_u = (n_u + 3) / 4, q_e = (n_e + 3) / 4, q_i = (p.R + 3) / 4;
  const long long q_x = static_cast<long long>(p.max_iters) * p.n_extra;
  const long long total = q_t + q_u + q_e + q_i + q_x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t r[4];
    if (i < q_t) {
      philox4(static_cast<uint32_t>(i), 0u, s_lo, s_hi, k0, k1, r);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * i + j < n_t) p.t_rand[4 * i + j] = u01(r[j]);
    } else if (i < q_t + q_u) {
      const long long q = i - q_t;
      philox4(static_cast<uint32_t>(q), 1u, s_lo, s_hi, k0, k1, r);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < n_u) p.u_final[4 * q + j] = u01(r[j]);
    } else if (i < q_t + q_u + q_e) {
      const long long q = i - q_t - q_u;
      philox4(static_cast<uint32_t>(q), 2u, s_lo, s_hi, k0, k1, r);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < n_e) p.eik_uniform[4 * q + j] = (2.0f * u01(r[j]) - 1.0f) * p.radius;
    } else if (i < q_t + q_u + q_e + q_i) {
      const long long q = i - q_t - q_u - q_e;
      philox4(static_cast<uint32_t>(q), 3u, s_lo, s_hi, k0, k1, r);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < p.R)   // 32 random bits onto [0, n_out): the multiply-shift map (bias < 2^-25 for n_out <= 128)
          p.eik_idx[4 * q + j] = static_cast<long long>((static_cast<unsigned long long>(r[j]) * p.n_out) >> 32);
    } else {
      // extra columns: a keyed bijection of [0, L) (4-round Feistel + cycle walking, pixels.cuh) evaluated at 0..n_extra-1
      const long long q = i - q_t - q_u - q_e - q_i;
      const int row = static_cast<int>(q / p.n_extra), j = static_cast<int>(q % p.n_extra);
      philox4(static_cast<uint32_t>(row), 4u, s_lo, s_hi, k0, k1, r);
      PixelPerm P;
      P.n = static_cast<uint32_t>(p.n_eval * (row + 1));
      uint32_t bits = 2;
      while (bits < 32 && (1u << bits) < P.n) ++bits;
      if (bits & 1) ++bits;
      P.half = bits / 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) P.key[k] = r[k];
      p.extra_idx[q] = static_cast<long long>(px_permute(P, static_cast<uint32_t>(j)));
    }
  }
  // the last block to finish bumps the step counter (every block read it before getting here)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(&p.counter[1], 1ULL);
    if (t == gridDim.x - 1) {
      p.counter[1] = 0ULL;
      p.counter[0] = step + 1ULL;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------------- fp32 GEMM
// C[M,N] (+)= op(A)[M,K] * op(B)[K,N], row-major storage with leading dimensions; ta: A is stored [K,M]; tb: B is stored [N,K]
// (a torch Linear weight).  Epilogue: + bias[N], ReLU, or multiply by (mask[m,n] > 0) (the ReLU adjoint).  32 x 64 tiles,
// 128 threads, 4 x 4 outputs per thread, K in slabs of 16 through shared memory with the next slab prefetched into
// registers; split-K over gridDim.z (fire-and-forget atomic adds; linear epilogues only) for the weight-gradient products,
// whose reduction runs over the 1024 latents.  Sizes here are ~1024 x 256 x 256: the first version (64 x 64 tiles, no
// prefetch, no split-K) left most SMs idle -- 84 us per launch, 0.74 ms per training step in the ncu launch list.
struct GemmParams {
  const float *A, *B;
  float* C;
  int M, N, K, lda, ldb, ldc;
  int ta, tb;
  const float* bias;   // [N] or nullptr
  int relu;
  const float* mask;   // [M, ldm] or nullptr: C *= (mask > 0)
  int ldm;
  int accumulate;      // C += instead of C =
  int k_chunk;         // K range per blockIdx.z (multiple of 16); == K: no split
};
constexpr int GEMM_TM = 32, GEMM_TN = 64, GEMM_TK = 16;
__device__ __forceinline__ float gemm_ld_a(const GemmParams& p, int m, int k) {
  return (m < p.M && k < p.K) ? (p.ta ? p.A[static_cast<size_t>(k) * p.lda + m] : p.A[static_cast<size_t>(m) * p.lda + k]) : 0.f;
}
__device__ __forceinline__ float gemm_ld_b(const GemmParams& p, int n, int k) {
  return (n < p.N && k < p.K) ? (p.tb ? p.B[static_cast<size_t>(n) * p.ldb + k] : p.B[static_cast<size_t>(k) * p.ldb + n]) : 0.f;
}
__global__ void __launch_bounds__(128) gemm_f32_kernel(GemmParams p) {
  __shared__ float As[GEMM_TK][GEMM_TM + 4], Bs[GEMM_TK][GEMM_TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;            // 16 x 8 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * GEMM_TM, n0 = blockIdx.x * GEMM_TN;
  const int k_begin = blockIdx.z * p.k_chunk, k_end = min(p.K, k_begin + p.k_chunk);
  // element (kk, mm) of the A slab handled by this thread: 4 per thread (512 / 128), coalesced along the stored row
  int a_kk[4], a_mm[4], b_kk[8], b_nn[8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = tid + 128 * r;
    if (p.ta) { a_kk[r] = i / GEMM_TM; a_mm[r] = i % GEMM_TM; } else { a_mm[r] = i / GEMM_TK; a_kk[r] = i % GEMM_TK; }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = tid + 128 * r;
    if (p.tb) { b_nn[r] = i / GEMM_TK; b_kk[r] = i % GEMM_TK; } else { b_kk[r] = i / GEMM_TN; b_nn[r] = i % GEMM_TN; }
  }
  float ra[4], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) ra[r] = gemm_ld_a(p, m0 + a_mm[r], k0 + a_kk[r] < k_end ? k0 + a_kk[r] : p.K);
#pragma unroll
    for (int r = 0; r < 8; ++r) rb[r] = gemm_ld_b(p, n0 + b_nn[r], k0 + b_kk[r] < k_end ? k0 + b_kk[r] : p.K);
  };
  float acc[4][4] = {};
  fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += GEMM_TK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) As[a_kk[r]][a_mm[r]] = ra[r];
#pragma unroll
    for (int r = 0; r < 8; ++r) Bs[b_kk[r]][b_nn[r]] = rb[r];
    __syncthreads();
    if (k0 + GEMM_TK < k_end) fetch(k0 + GEMM_TK);   // the next slab travels while this one is multiplied
#pragma unroll
    for (int kk = 0; kk < GEMM_TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      float* c = p.C + static_cast<size_t>(m) * p.ldc + n;
      if (split) {   // linear epilogue only (checked on the host); C was zero-filled or holds the value to add to
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(c), "f"(v) : "memory");
        continue;
      }
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.mask) v = p.mask[static_cast<size_t>(m) * p.ldm + n] > 0.f ? v : 0.f;
      *c = p.accumulate ? *c + v : v;
    }
  }
}
// out[n] (+)= sum_m X[m, n]   (bias gradients); one block per 32 columns
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ X, int M, int N, int ldx, float* __restrict__ out,
                                                         int accumulate) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < N)
    for (int m = w; m < M; m += 8) s += X[static_cast<size_t>(m) * ldx + c];
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ---------------------------------------------------------------------------------------------------- graph-safe Adam
// hyper (device): [0] lr, [1] beta1, [2] beta2, [3] eps, [4] weight_decay, [5] grad_scale ; state (device): [0] step count
// (float, as torch keeps it), [1] 1 - beta1^step, [2] 1 / sqrt(1 - beta2^step).  adam_prepare bumps the count and
// evaluates the bias corrections in double, once; the update kernel is adam_step_kernel's arithmetic.
__global__ void adam_prepare_kernel(const float* __restrict__ hyper, float* __restrict__ state) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float step = state[0] + 1.0f;
    state[0] = step;
    state[1] = static_cast<float>(1.0 - pow(static_cast<double>(hyper[1]), static_cast<double>(step)));
    state[2] = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(hyper[2]), static_cast<double>(step))));
  }
}
__global__ void __launch_bounds__(256) adam_graph_kernel(const AdamTable* __restrict__ tp, const float* __restrict__ hyper,
                                                         const float* __restrict__ state) {
  const AdamTable& T = *tp;
  int lo = 0, hi = T.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (T.blk_start[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid;
  }
  const neat_adam_tensor a = T.t[lo];
  const long long base = static_cast<long long>(blockIdx.x - T.blk_start[lo]) * ADAM_BLOCK_ELEMS;
  const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], weight_decay = hyper[4], grad_scale = hyper[5];
  const float step_size = lr / state[1], rsqrt_bc2 = state[2];
#pragma unroll
  for (int k = 0; k < ADAM_BLOCK_ELEMS / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < a.numel) {
      float g = a.grad[i] * grad_scale;
      const float p = a.param[i];
      if (weight_decay != 0.f) g += weight_decay * p;
      const float m = a.exp_avg[i] + (g - a.exp_avg[i]) * (1.0f - beta1);
      const float v = a.exp_avg_sq[i] * beta2 + (1.0f - beta2) * g * g;
      a.exp_avg[i] = m;
      a.exp_avg_sq[i] = v;
      a.param[i] = p - step_size * (m / (sqrtf(v) * rsqrt_bc2 + eps));
    }
  }
}

// ---------------------------------------------------------------------------------------------------- junction terms, device n
// The matched-junction count changes every step; in the replayed graph it is read from device memory.  packed (device):
// [0] n (int), then rows [cap] int, cols [cap] int, local [cap, 7] float (xyz | uv | uv_calib) -- one H2D copy per step.
// out[0..2] = j3d_loss, j2d_loss, j2d_stat (0 when n == 0); w3 / w2 = the loss weights of the two differentiable terms:
// g_j3g [G,3], g_j2gc [G,2] receive d (w3 out[0] + w2 out[1]) / d (global junctions, their calibrated projections).
struct JunctionStepParams {
  const int* packed;
  int cap, G;
  const float *j3g, *j2gc, *j2g;
  float w3, w2;
  float *out, *g_j3g, *g_j2gc;
};
__global__ void __launch_bounds__(256) junction_step_kernel(JunctionStepParams p) {
  __shared__ float red[3][8];
  const int n = min(p.packed[0], p.cap);
  const int* rows = p.packed + 1;
  const int* cols = rows + p.cap;
  const float* local = reinterpret_cast<const float*>(cols + p.cap);
  for (int i = threadIdx.x; i < 3 * p.G; i += blockDim.x) p.g_j3g[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * p.G; i += blockDim.x) p.g_j2gc[i] = 0.f;
  __syncthreads();
  float s3 = 0.f, s2 = 0.f, su = 0.f;
  const float k3 = n > 0 ? p.w3 / static_cast<float>(n) : 0.f, k2 = n > 0 ? p.w2 / static_cast<float>(n) : 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = rows[i], c = cols[i];
    const float* l = local + 7 * r;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float d = l[k] - p.j3g[3 * c + k];
      s3 += fabsf(d);
      p.g_j3g[3 * c + k] = -k3 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float d = l[5 + k] - p.j2gc[2 * c + k];
      s2 += fabsf(d);
      p.g_j2gc[2 * c + k] = -k2 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      su += fabsf(l[3 + k] - p.j2g[2 * c + k]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    su += __shfl_xor_sync(0xffffffffu, su, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = s3; red[1][w] = s2; red[2][w] = su; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
    p.out[threadIdx.x] = n > 0 ? t / static_cast<float>(n) : 0.f;
  }
}

}  // namespace neat
