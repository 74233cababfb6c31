// Per-step pixel sampling of the dataset (code/datasets/scene_hawp_dataset.py:148-194): the reference
// runs `mask.nonzero()` over the whole image on the CPU, a `randperm` over the masked pixels, five fancy-index gathers and
// then copies the results to the GPU, every step.  Here the per-image tables (rgb, labels, attraction points, the list of
// masked pixels) stay in HBM and one launch of R threads draws the subset and gathers every per-ray input of the step.
//
// The random subset (R distinct masked pixels) is either given (`perm`, the prefix of the reference's CPU `randperm`: same
// seed -> same rays) or drawn on the device by a keyed bijection of [0, n): a 4-round balanced Feistel network on
// ceil(log2 n) bits (rounded up to even) with cycle walking, so thread j gets position perm(j) with no n-sized work and
// no communication; distinct j give distinct positions by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace neat {

__host__ __device__ inline uint32_t px_hash32(uint32_t x) {  // "lowbias32" integer finaliser
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

struct PixelPerm {
  uint32_t n;        // domain size
  uint32_t half;     // bits per Feistel half
  uint32_t key[4];   // round keys
};

__host__ __device__ inline uint32_t px_permute(const PixelPerm& P, uint32_t j) {
  const uint32_t m = (1u << P.half) - 1u;
  uint32_t x = j;
  do {
    uint32_t l = x >> P.half, r = x & m;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t t = l ^ (px_hash32(r ^ P.key[k]) & m);
      l = r;
      r = t;
    }
    x = (l << P.half) | r;
  } while (x >= P.n);  // cycle walking: the 2^(2 half) domain is < 4n, so < 4 rounds expected
  return x;
}

inline PixelPerm make_pixel_perm(uint32_t n, uint64_t seed, uint64_t step) {
  PixelPerm P{};
  P.n = n;
  uint32_t bits = 2;
  while (bits < 32 && (1ull << bits) < n) ++bits;
  if (bits & 1) ++bits;
  P.half = bits / 2;
  const uint32_t s0 = px_hash32(static_cast<uint32_t>(seed) ^ px_hash32(static_cast<uint32_t>(seed >> 32) + 0x9e3779b9U));
  const uint32_t s1 = px_hash32(static_cast<uint32_t>(step) ^ px_hash32(static_cast<uint32_t>(step >> 32) + 0x85ebca6bU));
  for (uint32_t k = 0; k < 4; ++k) P.key[k] = px_hash32(s0 ^ px_hash32(s1 + k * 0x9e3779b9U));
  return P;
}

struct PixelParams {
  int R;                      // pixels to produce
  int W;                      // image width (uv = (pix % W, pix / W))
  long long first;            // full-image mode: pix = first + j
  const int* masked;          // [n] pixel indices of the mask, ascending (= mask.nonzero()); nullptr: full-image mode
  const long long* perm;      // [R] positions into `masked` (reference rng); nullptr: device draw
  PixelPerm draw;
  const float* rgb_image;     // [HW,3]
  const long long* labels;    // [HW]
  const float* att_points;    // [HW,2]
  const float* lines;         // [n_lines,5]
  int n_lines;
  float* uv;                  // [R,2]
  float* uv_proj;             // [R,2]
  float* rgb;                 // [R,3]
  float* lines2d;             // [R,5]
  long long* labels_out;      // [R]
  long long* index_out;       // [R] the sampled pixel index (sampling_idx)
};

__global__ void __launch_bounds__(256) sample_pixels_kernel(PixelParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.R) return;
  long long pix;
  if (p.masked) {
    const uint32_t pos = p.perm ? static_cast<uint32_t>(p.perm[j]) : px_permute(p.draw, static_cast<uint32_t>(j));
    pix = p.masked[pos];
  } else {
    pix = p.first + j;
  }
  const long long row = pix / p.W;
  const int col = static_cast<int>(pix - row * p.W);
  reinterpret_cast<float2*>(p.uv)[j] = make_float2(static_cast<float>(col), static_cast<float>(row));
  reinterpret_cast<float2*>(p.uv_proj)[j] = reinterpret_cast<const float2*>(p.att_points)[pix];
  const float* c = p.rgb_image + 3 * pix;
  float* o = p.rgb + 3 * static_cast<size_t>(j);
  o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
  const long long lab = p.labels[pix];
  p.labels_out[j] = lab;
  if (p.index_out) p.index_out[j] = pix;
  if (p.lines2d) {
    float* l = p.lines2d + 5 * static_cast<size_t>(j);
    if (lab >= 0 && lab < p.n_lines) {
      const float* s = p.lines + 5 * lab;
#pragma unroll
      for (int k = 0; k < 5; ++k) l[k] = s[k];
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) l[k] = 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// mask.nonzero() once per image (ascending pixel order, like torch.nonzero): per-block counts, a one-block scan of the
// counts, and an ordered scatter by ballot / popc ranks.
constexpr int PX_BLOCK = 1024;

__global__ void __launch_bounds__(PX_BLOCK) mask_count_kernel(const uint8_t* __restrict__ mask, long long n, int* __restrict__ block_count) {
  const long long i = static_cast<long long>(blockIdx.x) * PX_BLOCK + threadIdx.x;
  const int c = __syncthreads_count(i < n && mask[i] != 0);
  if (threadIdx.x == 0) block_count[blockIdx.x] = c;
}

// in place: block_count[b] -> exclusive prefix; total -> *n_out
__global__ void __launch_bounds__(PX_BLOCK) mask_scan_kernel(int* __restrict__ block_count, int nb, int* __restrict__ n_out) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nb; b0 += PX_BLOCK) {
    const int b = b0 + threadIdx.x;
    const int v = b < nb ? block_count[b] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t;
      }
      warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int base = carry + (wid ? warp_sum[wid - 1] : 0);
    if (b < nb) block_count[b] = base + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = carry;
}

__global__ void __launch_bounds__(PX_BLOCK) mask_scatter_kernel(const uint8_t* __restrict__ mask, long long n,
                                                                const int* __restrict__ block_off, int* __restrict__ out) {
  __shared__ int warp_sum[32];
  const long long i = static_cast<long long>(blockIdx.x) * PX_BLOCK + threadIdx.x;
  const bool on = i < n && mask[i] != 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, on);
  if (lane == 0) warp_sum[wid] = __popc(bal);
  __syncthreads();
  if (wid == 0) {
    const int v = warp_sum[lane];
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    warp_sum[lane] = inc - v;  // exclusive over warps
  }
  __syncthreads();
  if (on) out[block_off[blockIdx.x] + warp_sum[wid] + __popc(bal & ((1u << lane) - 1u))] = static_cast<int>(i);
}

}  // namespace neat
