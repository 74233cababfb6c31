// cluster_dbscan (neat_wfr_rend_a.py:333-342: sklearn DBSCAN(eps=0.01, min_samples=2) + per-cluster mean) on the GPU.
// With min_samples = 2 every point that has a neighbour within eps is a core point, so the clusters are exactly the
// connected components of the eps-graph with at least two points (hard part 5).  Clusters are
// numbered by their smallest point index, which is the order in which sklearn discovers them.
//   hash grid (cell = eps, open addressing, per-cell linked lists) -> union-find with atomicCAS hooking
//   (larger root under smaller) -> component sizes -> rank of the roots -> double-precision centroid sums.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace neat {

struct DbscanWs {
  unsigned long long* tkey;  // [T] cell key per hash slot
  int* head;                 // [T] first point of the cell's list
  int* next;                 // [N]
  int* parent;               // [N]
  int* count;                // [N] component size (valid at roots)
  int* cid;                  // [N] cluster id of a root (or -1)
  double* sums;              // [N/2 * 3]
  int* csize;                // [N/2]
  int T;
};
constexpr unsigned long long DB_EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long db_key(int cx, int cy, int cz) {
  const unsigned long long o = 1u << 20;
  return ((static_cast<unsigned long long>(cx + o) & 0x1FFFFF) << 42) | ((static_cast<unsigned long long>(cy + o) & 0x1FFFFF) << 21) |
         (static_cast<unsigned long long>(cz + o) & 0x1FFFFF);
}
__device__ __forceinline__ uint32_t db_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return static_cast<uint32_t>(k);
}
__device__ __forceinline__ void db_cell(const float* p, float inv_eps, int c[3]) {
  c[0] = static_cast<int>(floorf(p[0] * inv_eps));
  c[1] = static_cast<int>(floorf(p[1] * inv_eps));
  c[2] = static_cast<int>(floorf(p[2] * inv_eps));
}

__global__ void db_init_kernel(DbscanWs w, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.T) { w.tkey[i] = DB_EMPTY; w.head[i] = -1; }
  if (i < N) { w.parent[i] = i; w.count[i] = 0; w.cid[i] = -1; }
  if (i < (N / 2 + 1)) { w.csize[i] = 0; w.sums[3 * i] = 0.0; w.sums[3 * i + 1] = 0.0; w.sums[3 * i + 2] = 0.0; }
}

__global__ void db_insert_kernel(DbscanWs w, const float* __restrict__ pts, int N, float inv_eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c[3];
  db_cell(pts + 3 * i, inv_eps, c);
  const unsigned long long key = db_key(c[0], c[1], c[2]);
  uint32_t s = db_hash(key) & (w.T - 1);
  while (true) {
    const unsigned long long old = atomicCAS(&w.tkey[s], DB_EMPTY, key);
    if (old == DB_EMPTY || old == key) break;
    s = (s + 1) & (w.T - 1);
  }
  w.next[i] = atomicExch(&w.head[s], i);
}

__device__ __forceinline__ int uf_find(volatile int* parent, int x) {
  while (true) {
    const int p = parent[x];
    if (p == x) return x;
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving; a benign race (any ancestor is a valid parent)
    x = p;
  }
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // hook the larger root under the smaller
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}

__global__ void db_link_kernel(DbscanWs w, const float* __restrict__ pts, int N, float inv_eps, double eps2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
  int c[3];
  db_cell(pts + 3 * i, inv_eps, c);
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const unsigned long long key = db_key(c[0] + dx, c[1] + dy, c[2] + dz);
        uint32_t s = db_hash(key) & (w.T - 1);
        int j = -1;
        while (true) {
          const unsigned long long k = w.tkey[s];
          if (k == key) { j = w.head[s]; break; }
          if (k == DB_EMPTY) break;
          s = (s + 1) & (w.T - 1);
        }
        for (; j >= 0; j = w.next[j]) {
          if (j >= i) continue;  // every pair once
          const double ddx = static_cast<double>(px) - pts[3 * j], ddy = static_cast<double>(py) - pts[3 * j + 1],
                       ddz = static_cast<double>(pz) - pts[3 * j + 2];
          if (ddx * ddx + ddy * ddy + ddz * ddz <= eps2) uf_union(w.parent, i, j);
        }
      }
}

__global__ void db_count_kernel(DbscanWs w, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int r = uf_find(w.parent, i);
  // The root goes to `next` (the cell lists are dead after db_link), NOT back into parent[i]: another thread's path
  // halving (uf_find: read parent[i], read its parent, write parent[i]) could overwrite a flattened entry with an older,
  // non-root ancestor, and db_sum would then drop the point from its cluster's centroid (a rare, run-dependent error:
  // one failure of test_dbscan_vs_sklearn in ~10 full-suite runs led here).
  w.next[i] = r;
  atomicAdd(&w.count[r], 1);
}

// one block: rank the roots of components with >= 2 points by index; n_clusters = their number
__global__ void db_rank_kernel(DbscanWs w, int N, int* n_clusters) {
  __shared__ int wsum[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < N; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const int flag = (i < N && w.parent[i] == i && w.count[i] >= 2) ? 1 : 0;
    int v = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) >= o) v += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      int s = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += t;
      }
      wsum[threadIdx.x] = s;
    }
    __syncthreads();
    const int woff = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    if (flag) {
      const int id = base + woff + v - 1;
      w.cid[i] = id;
      w.csize[id] = w.count[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) base += wsum[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_clusters = base;
}

__global__ void db_sum_kernel(DbscanWs w, const float* __restrict__ pts, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int c = w.cid[w.next[i]];   // next[i] = the root of i (db_count)
  if (c < 0) return;
  atomicAdd(&w.sums[3 * c], static_cast<double>(pts[3 * i]));
  atomicAdd(&w.sums[3 * c + 1], static_cast<double>(pts[3 * i + 1]));
  atomicAdd(&w.sums[3 * c + 2], static_cast<double>(pts[3 * i + 2]));
}

__global__ void db_mean_kernel(DbscanWs w, const int* __restrict__ n_clusters, float* __restrict__ centroids) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= *n_clusters) return;
  const double inv = 1.0 / static_cast<double>(w.csize[c]);
  centroids[3 * c] = static_cast<float>(w.sums[3 * c] * inv);
  centroids[3 * c + 1] = static_cast<float>(w.sums[3 * c + 1] * inv);
  centroids[3 * c + 2] = static_cast<float>(w.sums[3 * c + 2] * inv);
}

}  // namespace neat
