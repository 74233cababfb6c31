// ErrorBoundSampler.get_z_vals without the MLP (code/model/ray_sampler.py:130-283, 285-293) and
// UniformSampler.get_z_vals (ray_sampler.py:69-95): per-ray HBM/latency-bound kernels, one warp per ray,
// per-ray state staged in shared memory.  The data-dependent loop of the reference (host sync on
// `beta.max() > beta0`, ray_sampler.py:200) is replaced by a device-side state word: every iteration's
// kernels are always launched and return immediately once the sampler has finished.
//
// Iteration i (L = 128 (i+1) samples per ray after the merge):
//   sampler_bounds : merge (samples, sdf_new) into the sorted (z, sdf) state; d* per interval; error bound at
//                    beta0; 10-step bisection on beta; flag |= beta > beta0                  (:152-200)
//   sampler_draw   : more = flag && i+1 < max_iters;
//                    more  -> 128 samples from the bound-opacity pdf (linspace u)             (:202-249)
//                    final -> N_samples from the weight pdf (+1e-5), near/far/extra columns, sort, z_eik
//                                                                                            (:211-276)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace neat {

constexpr int SMP_MAX_L = 640;   // N_samples_eval (128) * max_total_iters (5)
constexpr int SMP_C = SMP_MAX_L / 32;  // elements per lane (blocked ownership)
constexpr int SMP_MAX_NEW = 128;
constexpr int SMP_MAX_OUT = 128;  // N_samples + 2 + N_samples_extra <= 128
constexpr int SMP_WARPS = 2;

struct SamplerState {  // device-resident control block (ints)
  int done;            // 1 once the final samples were drawn
  int n_iters;         // iterations executed (k)
  int flag[8];         // flag[i] = any ray with beta > beta0 after iteration i
};

struct SamplerParams {
  int R;
  int n_eval;       // N_samples_eval (128): samples added per iteration
  int n_final;      // N_samples (64)
  int n_extra;      // N_samples_extra (32)
  int beta_iters;   // 10
  int max_iters;    // 5
  int training;     // 1: u_final / extra_idx / eik_idx are used
  float near, far, eps;
  const float* beta_param;  // density.beta (device scalar); beta0 = |beta| + beta_min
  float beta_min;
  // state
  SamplerState* st;
  float* z;        // [R, SMP_MAX_L] sorted depths
  float* sdf;      // [R, SMP_MAX_L]
  float* samples;  // [R, 128] depths whose sdf is queried next
  float* sdf_new;  // [R, 128]
  float* beta;     // [R]
  // randoms (training) / tables (eval)
  const float* t_rand;      // [R, n_eval] stratified jitter or nullptr (eval)
  const float* u_final;     // [R, n_final] (training)
  const int64_t* extra_idx; // training: [n_extra] columns of z; eval: table [max_iters][n_extra] indexed by k-1
  const int64_t* eik_idx;   // [R] (training)
  // outputs
  float* z_vals;  // [R, n_final + 2 + n_extra] sorted
  float* z_eik;   // [R]
};

// torch.linspace(0, 1, n)[i] in fp32 (symmetric evaluation, as ATen's CPU/CUDA kernels do)
__device__ __forceinline__ float linspace01(int i, int n) {
  const float step = 1.0f / static_cast<float>(n - 1);
  return i < n / 2 ? step * static_cast<float>(i) : 1.0f - step * static_cast<float>(n - 1 - i);
}

// LaplaceDensity.forward (code/model/density.py:21-26)
__device__ __forceinline__ float laplace_density(float s, float beta) {
  const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
  return (1.0f / beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(s) / beta));
}

__device__ __forceinline__ float warp_excl_scan(float v, float& total) {
  const int lane = threadIdx.x & 31;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  // exclusive = inclusive of the previous lane (NOT inc - v: v may be ~1e10 * sigma for the last interval)
  const float prev = __shfl_up_sync(0xffffffffu, inc, 1);
  return lane == 0 ? 0.f : prev;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// UniformSampler.get_z_vals + the initial per-ray beta (ray_sampler.py:69-95, 134-140)
__global__ void sampler_init_kernel(SamplerParams p) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.st->done = 0;
    p.st->n_iters = 0;
    for (int i = 0; i < 8; ++i) p.st->flag[i] = 0;
  }
  if (r >= p.R) return;
  const int n = p.n_eval;
  float* out = p.samples + static_cast<size_t>(r) * SMP_MAX_NEW;
  float ssq = 0.f;
  // lane-strided: consecutive lanes write consecutive samples
  for (int i = lane; i < n; i += 32) {
    auto zi = [&](int k) {
      const float t = linspace01(k, n);
      return p.near * (1.f - t) + p.far * t;
    };
    float zc = zi(i);
    if (p.t_rand) {
      const float lo = i == 0 ? zc : 0.5f * (zc + zi(i - 1));
      const float up = i == n - 1 ? zc : 0.5f * (zi(i + 1) + zc);
      zc = lo + (up - lo) * p.t_rand[static_cast<size_t>(r) * n + i];
    }
    out[i] = zc;
  }
  __syncwarp();
  for (int i = lane; i < n - 1; i += 32) {
    const float d = out[i + 1] - out[i];
    ssq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
  if (lane == 0) p.beta[r] = sqrtf((1.0f / (4.0f * logf(p.eps + 1.0f))) * ssq);
}

struct RaySmem {
  float z[SMP_MAX_L];
  float s[SMP_MAX_L];
  float dist[SMP_MAX_L];
  float dstar[SMP_MAX_L];
  float aux[SMP_MAX_L];
};

// Theorem-1 d* of interval k (ray_sampler.py:161-173)
__device__ __forceinline__ float interval_dstar(float a, float s0, float s1) {
  const float b = fabsf(s0), c = fabsf(s1);
  const bool first = a * a + b * b <= c * c;
  const bool second = a * a + c * c <= b * b;
  float ds = 0.f;
  if (first) ds = b;
  if (second) ds = c;
  const float sp = (a + b + c) / 2.0f;
  const float area = sp * (sp - a) * (sp - b) * (sp - c);
  if (!first && !second && (b + c - a > 0.f)) ds = 2.0f * sqrtf(area) / a;
  const float sg0 = s0 > 0.f ? 1.f : (s0 < 0.f ? -1.f : 0.f);
  const float sg1 = s1 > 0.f ? 1.f : (s1 < 0.f ? -1.f : 0.f);
  return (sg0 * sg1 == 1.f) ? ds : 0.f;
}

// get_error_bound (ray_sampler.py:285-293) for one ray held in shared memory; all lanes return the max
__device__ __forceinline__ float error_bound(const RaySmem& m, int L, float beta_q) {
  const int lane = threadIdx.x & 31;
  // blocked ownership sized to the CURRENT interval count: every lane owns C = ceil((L-1)/32) consecutive intervals, so
  // all 32 lanes work at every iteration (with a fixed block of SMP_C = 20, 60 % of the lanes idle at L = 256)
  const int C = (L - 1 + 31) >> 5;
  const int k0 = lane * C;
  float t_int[SMP_C], t_eps[SMP_C];
  float s_int = 0.f, s_eps = 0.f;
  const float inv4b2 = 1.0f / (4.0f * beta_q * beta_q);
#pragma unroll
  for (int j = 0; j < SMP_C; ++j) {
    const int k = k0 + j;
    float ti = 0.f, te = 0.f;
    if (j < C && k < L - 1) {
      const float d = m.dist[k];
      ti = d * laplace_density(m.s[k], beta_q);
      te = expf(-m.dstar[k] / beta_q) * (d * d) * inv4b2;
    }
    t_int[j] = ti; t_eps[j] = te;
    s_int += ti; s_eps += te;
  }
  float tot;
  float p_int = warp_excl_scan(s_int, tot);  // sum_{k' < k0} dist * sigma  == integral[k0]
  float p_eps = warp_excl_scan(s_eps, tot);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < SMP_C; ++j) {
    const int k = k0 + j;
    p_eps += t_eps[j];  // inclusive
    if (j < C && k < L - 1) mx = fmaxf(mx, (fminf(expf(p_eps), 1.0e6f) - 1.0f) * expf(-p_int));
    p_int += t_int[j];
  }
  return warp_max(mx);
}

__device__ __forceinline__ void load_ray(RaySmem& m, const SamplerParams& p, int r, int L) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < L; i += 32) {
    m.z[i] = p.z[static_cast<size_t>(r) * SMP_MAX_L + i];
    m.s[i] = p.sdf[static_cast<size_t>(r) * SMP_MAX_L + i];
  }
  __syncwarp();
  for (int k = lane; k < L - 1; k += 32) {
    const float a = m.z[k + 1] - m.z[k];
    m.dist[k] = a;
    m.dstar[k] = interval_dstar(a, m.s[k], m.s[k + 1]);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * SMP_WARPS) sampler_bounds_kernel(SamplerParams p, int iter) {
  __shared__ RaySmem sm[SMP_WARPS];
  if (p.st->done) return;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * SMP_WARPS + w;
  if (r >= p.R) return;
  RaySmem& m = sm[w];
  const int n = p.n_eval;
  const int Lold = n * iter, L = Lold + n;
  const float beta0 = fabsf(*p.beta_param) + p.beta_min;

  // ---- merge the new samples into the sorted state (stable: old before new on ties) ----
  float* zg = p.z + static_cast<size_t>(r) * SMP_MAX_L;
  float* sg = p.sdf + static_cast<size_t>(r) * SMP_MAX_L;
  const float* zn = p.samples + static_cast<size_t>(r) * SMP_MAX_NEW;
  const float* sn = p.sdf_new + static_cast<size_t>(r) * SMP_MAX_NEW;
  for (int i = lane; i < Lold; i += 32) { m.dist[i] = zg[i]; m.dstar[i] = sg[i]; }   // old (sorted)
  for (int i = lane; i < n; i += 32) { m.aux[i] = zn[i]; m.aux[SMP_MAX_NEW + i] = sn[i]; }  // new
  __syncwarp();
  // The new samples are non-decreasing for every draw the reference makes (stratified depths, inverse CDF of a sorted
  // u); then both ranks are binary searches.  Checked at run time: a draw that is not sorted (fp32 rounding at an
  // interval boundary) takes the general counting path, so the merge is always the stable sort of the reference.
  bool sorted = true;
  for (int j = lane; j < n - 1; j += 32) sorted = sorted && (m.aux[j] <= m.aux[j + 1]);
  sorted = __all_sync(0xffffffffu, sorted);
  for (int i = lane; i < Lold; i += 32) {  // rank of old i = i + #{new < z_i}
    const float v = m.dist[i];
    int c = 0;
    if (sorted) {
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (m.aux[mid] < v) lo = mid + 1; else hi = mid; }
      c = lo;
    } else {
      for (int j = 0; j < n; ++j) c += m.aux[j] < v;
    }
    m.z[i + c] = v; m.s[i + c] = m.dstar[i];
  }
  for (int j = lane; j < n; j += 32) {  // rank of new j = #{old <= s_j} + #{new k: s_k < s_j or (== and k < j)}
    const float v = m.aux[j];
    int lo = 0, hi = Lold;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (m.dist[mid] <= v) lo = mid + 1; else hi = mid; }
    int c = lo;
    if (sorted) {
      c += j;
    } else {
      for (int k = 0; k < n; ++k) { const float u = m.aux[k]; c += (u < v) || (u == v && k < j); }
    }
    m.z[c] = v; m.s[c] = m.aux[SMP_MAX_NEW + j];
  }
  __syncwarp();
  for (int i = lane; i < L; i += 32) { zg[i] = m.z[i]; sg[i] = m.s[i]; }
  __syncwarp();
  for (int k = lane; k < L - 1; k += 32) {
    const float a = m.z[k + 1] - m.z[k];
    m.dist[k] = a;
    m.dstar[k] = interval_dstar(a, m.s[k], m.s[k + 1]);
  }
  __syncwarp();

  // ---- beta: error bound at beta0, then bisection (ray_sampler.py:177-185) ----
  float beta = p.beta[r];
  const float err0 = error_bound(m, L, beta0);
  if (err0 <= p.eps) beta = beta0;
  float lo = beta0, hi = beta;
  for (int it = 0; it < p.beta_iters; ++it) {
    const float mid = (lo + hi) / 2.0f;
    const float e = error_bound(m, L, mid);
    if (e <= p.eps) hi = mid; else lo = mid;
  }
  beta = hi;
  if (lane == 0) {
    p.beta[r] = beta;
    if (beta > beta0) atomicOr(&p.st->flag[iter], 1);
  }
}

__global__ void __launch_bounds__(32 * SMP_WARPS) sampler_draw_kernel(SamplerParams p, int iter) {
  __shared__ RaySmem sm[SMP_WARPS];
  if (p.st->done) return;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * SMP_WARPS + w;
  const bool more = p.st->flag[iter] != 0 && (iter + 1 < p.max_iters);
  if (r < p.R) {
    RaySmem& m = sm[w];
    const int L = p.n_eval * (iter + 1);
    load_ray(m, p, r, L);
    const float beta = p.beta[r];
    const int C = (L - 1 + 31) >> 5;  // see error_bound
    const int k0 = lane * C;
    // transmittance T[k] = exp(-sum_{j<k} dist_j sigma_j), pdf over the L-1 intervals
    float fe[SMP_C], te[SMP_C];
    float s_fe = 0.f, s_te = 0.f;
    const float inv4b2 = 1.0f / (4.0f * beta * beta);
#pragma unroll
    for (int j = 0; j < SMP_C; ++j) {
      const int k = k0 + j;
      float f = 0.f, e = 0.f;
      if (j < C && k < L - 1) {
        const float d = m.dist[k];
        f = d * laplace_density(m.s[k], beta);
        if (more) e = expf(-m.dstar[k] / beta) * (d * d) * inv4b2;
      }
      fe[j] = f; te[j] = e; s_fe += f; s_te += e;
    }
    float tot;
    float p_fe = warp_excl_scan(s_fe, tot);
    float p_te = warp_excl_scan(s_te, tot);
    float pdf[SMP_C];
    float s_pdf = 0.f;
#pragma unroll
    for (int j = 0; j < SMP_C; ++j) {
      const int k = k0 + j;
      float v = 0.f;
      if (j < C && k < L - 1) {
        const float T = expf(-p_fe);
        if (more) {
          p_te += te[j];
          v = (fminf(expf(p_te), 1.0e6f) - 1.0f) * T;               // bound opacity (:209-211), add_tiny = 0
        } else {
          v = (1.0f - expf(-fe[j])) * T + 1e-5f;                     // weights[:-1] + 1e-5 (:215)
        }
      }
      p_fe += fe[j];
      pdf[j] = v; s_pdf += v;
    }
    float total;
    float p_pdf = warp_excl_scan(s_pdf, total);
    // cdf[0] = 0, cdf[k+1] = cumsum(pdf / total)[k]   (sum of normalised terms, as the reference)
    // NOTE: the reference normalises first and then cumsums; we cumsum pdf/total term by term.
    float s_n = 0.f;
#pragma unroll
    for (int j = 0; j < SMP_C; ++j) { pdf[j] = pdf[j] / total; s_n += pdf[j]; }
    p_pdf = warp_excl_scan(s_n, tot);
    if (lane == 0) m.aux[0] = 0.f;
#pragma unroll
    for (int j = 0; j < SMP_C; ++j) {
      const int k = k0 + j;
      p_pdf += pdf[j];
      if (j < C && k < L - 1) m.aux[k + 1] = p_pdf;
    }
    __syncwarp();
    // inverse-CDF sampling (:231-249)
    const int N = more ? p.n_eval : p.n_final;
    const bool lin = more || !p.training;
    float* dst = p.samples + static_cast<size_t>(r) * SMP_MAX_NEW;  // final draw: the first n_final entries
    for (int i = lane; i < N; i += 32) {
      const float u = lin ? linspace01(i, N) : p.u_final[static_cast<size_t>(r) * N + i];
      int lo = 0, hi = L;  // searchsorted(cdf, u, right=True): first index with cdf > u
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (m.aux[mid] <= u) lo = mid + 1; else hi = mid; }
      const int below = max(lo - 1, 0), above = min(lo, L - 1);
      const float c0 = m.aux[below], c1 = m.aux[above], b0 = m.z[below], b1 = m.z[above];
      float den = c1 - c0;
      if (den < 1e-5f) den = 1.0f;
      dst[i] = b0 + (u - c0) / den * (b1 - b0);
    }
  }
  // the last CTA-independent bookkeeping: every CTA sees the same `more`; one thread records the outcome
  if (!more && blockIdx.x == 0 && threadIdx.x == 0) p.st->n_iters = iter + 1;
}

// marks the sampler finished after the final draw (separate launch: sampler_draw CTAs read st->done on entry)
__global__ void sampler_finish_kernel(SamplerState* st, int iter, int max_iters) {
  if (st->done) return;
  const bool more = st->flag[iter] != 0 && (iter + 1 < max_iters);
  if (!more) st->done = 1;
}

// near, far and the extra columns of z (:259-270), the final sort (:272) and z_eik (:275-276).
// extra_idx is a table [max_iters][n_extra]; row k-1 (k = iterations executed) is used.
__global__ void __launch_bounds__(32 * SMP_WARPS) sampler_final_kernel(SamplerParams p) {
  __shared__ float buf[SMP_WARPS][SMP_MAX_OUT];
  __shared__ float srt[SMP_WARPS][SMP_MAX_OUT];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * SMP_WARPS + w;
  if (r >= p.R) return;
  const int k = p.st->n_iters;
  const int n_out = p.n_final + 2 + p.n_extra;
  const int64_t* eidx = p.extra_idx + static_cast<size_t>(k - 1) * p.n_extra;
  const float* zr = p.z + static_cast<size_t>(r) * SMP_MAX_L;
  for (int i = lane; i < p.n_final; i += 32) buf[w][i] = p.samples[static_cast<size_t>(r) * SMP_MAX_NEW + i];
  if (lane == 0) { buf[w][p.n_final] = p.near; buf[w][p.n_final + 1] = p.far; }
  for (int i = lane; i < p.n_extra; i += 32) buf[w][p.n_final + 2 + i] = zr[eidx[i]];
  __syncwarp();
  for (int i = lane; i < n_out; i += 32) {  // stable rank sort
    const float v = buf[w][i];
    int c = 0;
    for (int j = 0; j < n_out; ++j) { const float u = buf[w][j]; c += (u < v) || (u == v && j < i); }
    srt[w][c] = v;
  }
  __syncwarp();
  float* zo = p.z_vals + static_cast<size_t>(r) * n_out;
  for (int i = lane; i < n_out; i += 32) zo[i] = srt[w][i];
  if (lane == 0) p.z_eik[r] = srt[w][p.training ? static_cast<int>(p.eik_idx[r]) : 0];
}

}  // namespace neat
