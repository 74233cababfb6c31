// C ABI of neat_b200 (include/neat_b200.h): context, weight packing, kernel launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "plan.h"
#include "attraction.cuh"
#include "backward.cuh"
#include "composite.cuh"
#include "dbscan.cuh"
#include "heads.cuh"
#include "loss.cuh"
#include "sampler.cuh"
#include "sdf_query.cuh"
#include "sdf_render.cuh"
#include "weight_norm.cuh"
#include "wgrad.cuh"
#include "junction.cuh"
#include "adam.cuh"
#include "train_aux.cuh"
#include "parsing.cuh"
#include "pixels.cuh"

using namespace neat;

namespace {
std::atomic<long long> g_launches{0};  // kernels launched by this library (bench.py's gpu_launches)
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e);
}
#define CK(call)                                        \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

template <class T>
cudaError_t upload(const std::vector<T>& h, T** d) {
  cudaError_t e = cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

constexpr int QUERY_STAGES = 4;
constexpr int RENDER_STAGES = 4;
inline size_t engine_smem_bytes(int stages) {
  return (stages == 4 ? sizeof(EngineSmem<4>) : sizeof(EngineSmem<3>)) + 1024;
}

// buf: which 256-column half of TMEM holds the accumulator (consecutive dependent steps alternate, see engine.cuh)
Step mk_step(const PLayer& w, int buf, int wait_a, int wait_aux, int commit_d) {
  Step s{};
  s.w = w;
  s.d_col = static_cast<uint16_t>((buf & 1) * 256);
  s.wait_a = static_cast<uint8_t>(wait_a);
  s.wait_aux = static_cast<uint8_t>(wait_aux);
  s.commit_d = static_cast<uint8_t>(commit_d);
  s.comp = 1.0f + 0.35f * 3.0f * static_cast<float>(w.nk_main + w.nk_aux) * 5.9604645e-08f;  // 3 MMAs per k-step
  return s;
}
}  // namespace

struct neat_ctx {
  Plan plan;
  int device = 0;
  int num_sms = 0;
  uint8_t* packed = nullptr;      // weight slabs as fp16 pairs + fp32 biases: the forward programs
  uint8_t* packed_bf = nullptr;   // the same layout with bf16 pairs: the backward programs
  int32_t* g_src = nullptr;
  uint32_t *g_dst_hi = nullptr, *g_dst_lo = nullptr;
  float* g_scale = nullptr;
  int32_t* g_fsrc = nullptr;
  uint32_t* g_fdst = nullptr;
  Program prog_query{}, prog_render{}, prog_head[2]{}, prog_head_bwd[2]{}, prog_sdf_bwd{};  // zero: pf_base = nullptr
  uint8_t* ones_tile = nullptr;  // X operand with column 0 = 1 (aux-plane sized, hi then lo)
  int epilogue_prefetch = 0;  // NEAT_EPILOGUE_PREFETCH=1: producer-warp L2 hints for the backward epilogues (experiment)
  int grid_cap = 0;           // neat_debug_set_grid_cap: at most this many persistent CTAs per tile-MLP launch (tests force
                              // several tiles per CTA at small sizes with it); 0 = one CTA per SM
  int wgrad_max_split = 32;   // neat_debug_set_wgrad_split / NEAT_WGRAD_SPLIT: cap of the tile-range splits per GEMM
  int wgrad_tiles_per_split = 48;
  WJob* jobs_dev = nullptr;
  std::vector<WJob> jobs_last;  // what jobs_dev holds
  int jobs_cap = 0;
  WnTable* wn_dev[2] = {nullptr, nullptr};
  WnTable wn_host[2]{};
  bool wn_valid[2] = {false, false};
  int wn_rows = 0, wn_slot = 0;
};

// persistent grid of a tile-MLP launch: one CTA per SM (or the debug cap), never more CTAs than tiles
// Persistent grid of a tile-MLP launch: the FEWEST CTAs that still finish in the same number of rounds as one CTA per
// SM would (784 tiles on 148 SMs are 6 rounds; 131 CTAs x 6 tiles do the same work in the same 6 rounds).  The SMs left
// over run the side-stream launches of the step (eikonal / surface points, DBSCAN) concurrently instead of in the tail,
// and the big kernels lose nothing -- measured at 1024 rays, 148 -> 131 CTAs: sdf_render 1.085 -> 1.072 ms, head_bwd
// 0.587 -> 0.518, sdf_bwd 1.125 -> 1.108, head_fwd 0.405 -> 0.399 (less L2 / HBM contention per round).
static inline int grid_for(const neat_ctx* c, int M) {
  const int n_tiles = (M + TILE_M - 1) / TILE_M;
  const int cap = c->grid_cap > 0 && c->grid_cap < c->num_sms ? c->grid_cap : c->num_sms;
  if (n_tiles <= cap) return n_tiles;
  const int rounds = (n_tiles + cap - 1) / cap;
  return (n_tiles + rounds - 1) / rounds;
}

// ---------------------------------------------------------------- packing kernels
// weights -> hi / lo operand slabs: fp16 pairs (packed_f16: ~2^-22 relative, the forward programs) and bf16 pairs
// (packed_bf16: the backward programs, whose activations are bf16 pairs)
__global__ void pack_pairs_kernel(const float* __restrict__ flat, const int32_t* __restrict__ src,
                                  const uint32_t* __restrict__ dst_hi, const uint32_t* __restrict__ dst_lo,
                                  const float* __restrict__ scale, int n, uint16_t* __restrict__ packed_f16,
                                  uint16_t* __restrict__ packed_bf16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = src[i];
  const float v = s >= 0 ? flat[s] * scale[i] : 0.f;
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  packed_bf16[dst_hi[i]] = __bfloat16_as_ushort(hi);
  packed_bf16[dst_lo[i]] = __bfloat16_as_ushort(lo);
  const float vc = fminf(fmaxf(v, -65504.f), 65504.f);
  const __half h = __float2half_rn(vc);
  const __half l = __float2half_rn(vc - __half2float(h));
  packed_f16[dst_hi[i]] = __half_as_ushort(h);
  packed_f16[dst_lo[i]] = __half_as_ushort(l);
}
__global__ void pack_f32_kernel(const float* __restrict__ flat, const int32_t* __restrict__ src,
                                const uint32_t* __restrict__ dst, int n, float* __restrict__ packed, float* __restrict__ packed2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = src[i];
  const float v = s >= 0 ? flat[s] : 0.f;
  packed[dst[i]] = v;
  packed2[dst[i]] = v;
}

extern "C" {

const char* neat_last_error(void) { return g_err.c_str(); }

// fills a fresh context; on any error the caller (neat_create) destroys it, so every early return here is leak-free
static int init_ctx(neat_ctx* c, const neat_net_config* cfg) {
  try {
    build_plan(*cfg, c->plan);
  } catch (const std::exception& e) {
    return fail(NEAT_EUNSUPPORTED, e.what());
  }
  if (const char* e = std::getenv("NEAT_EPILOGUE_PREFETCH")) c->epilogue_prefetch = std::atoi(e);
  if (const char* e = std::getenv("NEAT_WGRAD_SPLIT")) c->wgrad_max_split = std::max(1, std::atoi(e));
  CK(cudaGetDevice(&c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major != 10) return fail(NEAT_ENODEV, "neat_b200 needs an sm_100a device (tcgen05/TMEM)");
  c->num_sms = prop.multiProcessorCount;
  const GatherTables& g = c->plan.g;
  CK(cudaMalloc(&c->packed, c->plan.packed_bytes));
  CK(cudaMemset(c->packed, 0, c->plan.packed_bytes));
  CK(cudaMalloc(&c->packed_bf, c->plan.packed_bytes));
  CK(cudaMemset(c->packed_bf, 0, c->plan.packed_bytes));

  CK(upload(g.src, &c->g_src));
  CK(upload(g.dst_hi, &c->g_dst_hi));
  CK(upload(g.dst_lo, &c->g_dst_lo));
  CK(upload(g.scale, &c->g_scale));
  CK(upload(g.fsrc, &c->g_fsrc));
  CK(upload(g.fdst, &c->g_fdst));

  const Plan& P = c->plan;
  Program& q = c->prog_query;
  q.n = 0;
  for (int l = 0; l < P.cfg.sdf_layers; ++l) q.s[q.n++] = mk_step(P.sdf_f[l], l, 1, l == 0, 1);

  {
    Program& r = c->prog_render;
    const int L = P.cfg.sdf_layers;
    r.n = 0;
    for (int l = 0; l < L - 1; ++l) r.s[r.n++] = mk_step(P.sdf_f[l], l, 1, l == 0, 1);
    const int b = (L - 2) & 1;                                    // accumulator half of F_{L-2}
    r.s[r.n++] = mk_step(P.sdf_f[L - 1], 1 - b, 1, 0, 0);         // sdf row   -> the other half (32 columns)
    r.s[r.n++] = mk_step(P.sdf_f_feat, b, 0, 0, 1);               // features  -> issued after every reader of b is done
    for (int l = L - 2, i = 0; l >= 0; --l, ++i) r.s[r.n++] = mk_step(P.sdf_t[l], (i & 1) ? b : 1 - b, 1, 0, 1);
    for (int h = 0; h < 2; ++h) {
      Program& g2 = c->prog_head[h];
      const std::vector<PLayer>& fw = h == 0 ? P.rend_f : P.att_f;
      g2.n = 0;
      for (const PLayer& w : fw) { g2.s[g2.n] = mk_step(w, g2.n, 1, g2.n == 0, 1); ++g2.n; }
    }
  }
  {
    const int L = P.cfg.sdf_layers, HL = P.cfg.head_layers;
    Program& b = c->prog_sdf_bwd;
    b.n = 0;
    const SdfSaveLayout sfl = sdf_save_layout(L, true);
    for (int l = 0; l < L - 1; ++l, ++b.n) {  // tangent step l: the epilogue reads sigma'_l and a_l
      b.s[b.n] = mk_step(P.sdf_f[l], b.n, 1, b.n == 0, 1);
      b.s[b.n].pf_off[0] = sfl.d1 + l * D1_BYTES;
      b.s[b.n].pf_bytes[0] = D1_BYTES;
      b.s[b.n].pf_off[1] = sfl.a + l * TILE_MAIN_BYTES;
      b.s[b.n].pf_bytes[1] = TILE_MAIN_BYTES;
    }
    for (int l = L - 1; l >= 1; --l, ++b.n) {  // reverse step of layer l: the epilogue reads sigma'_{l-1} (+ its own zhat scratch)
      b.s[b.n] = mk_step(P.sdf_t[l], b.n, 1, l == L - 1, 1);
      b.s[b.n].pf_off[0] = sfl.d1 + (l - 1) * D1_BYTES;
      b.s[b.n].pf_bytes[0] = D1_BYTES;
    }
    for (int h = 0; h < 2; ++h) {
      Program& hb = c->prog_head_bwd[h];
      const std::vector<PLayer>& tr = h == 0 ? P.rend_t : P.att_t;
      hb.n = 0;
      const HeadSaveLayout hsl = head_save_layout(HL);
      for (int l = HL - 1; l >= 1; --l, ++hb.n) {  // the epilogue reads the ReLU masks = hi plane of u_l
        hb.s[hb.n] = mk_step(tr[l], hb.n, 1, hb.n == 0, 1);
        hb.s[hb.n].pf_off[0] = hsl.u + (l - 1) * TILE_MAIN_BYTES;
        hb.s[hb.n].pf_bytes[0] = PLANE_MAIN_BYTES;
      }
      hb.s[hb.n] = mk_step(tr[0], hb.n, 1, 0, 0);
      ++hb.n;
      hb.s[hb.n] = mk_step(h == 0 ? P.rend_t0_aux : P.att_t0_aux, hb.n, 0, 0, 1);
      ++hb.n;
    }
    std::vector<uint16_t> ones(TILE_AUX_BYTES / 2, 0);
    for (int r = 0; r < TILE_M; ++r) ones[r * 8] = 0x3F80;  // hi plane, chunk 0, column 0 = bf16(1.0)
    CK(cudaMalloc(&c->ones_tile, TILE_AUX_BYTES));
    CK(cudaMemcpy(c->ones_tile, ones.data(), TILE_AUX_BYTES, cudaMemcpyHostToDevice));
  }
  // operand formats (umma.cuh): the forward programs multiply fp16 x fp16 pairs, the backward ones bf16 x bf16 (the
  // hardware rejects mixed formats; backward tensors need bf16's range), each from its own copy of the weight slabs
  for (Program* p : {&c->prog_query, &c->prog_render, &c->prog_head[0], &c->prog_head[1]}) p->a_f16 = p->b_f16 = 1;
  for (Program* p : {&c->prog_head_bwd[0], &c->prog_head_bwd[1], &c->prog_sdf_bwd}) p->a_f16 = p->b_f16 = 0;
  CK(cudaFuncSetAttribute(head_bwd_kernel<RENDER_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(engine_smem_bytes(RENDER_STAGES))));
  CK(cudaFuncSetAttribute(sdf_bwd_kernel<RENDER_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(engine_smem_bytes(RENDER_STAGES))));
  CK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(sizeof(WgradSmem) + 1024)));
  CK(cudaFuncSetAttribute(sdf_render_kernel<RENDER_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(engine_smem_bytes(RENDER_STAGES))));
  CK(cudaFuncSetAttribute(head_fwd_kernel<RENDER_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(engine_smem_bytes(RENDER_STAGES))));
  CK(cudaFuncSetAttribute(sdf_query_kernel<QUERY_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(engine_smem_bytes(QUERY_STAGES))));
  return NEAT_OK;
}

int neat_create(const neat_net_config* cfg, neat_ctx** out) {
  if (!cfg || !out) return fail(NEAT_EINVAL, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(NEAT_ENODEV, "no CUDA device");
  neat_ctx* c = new neat_ctx();
  if (int e = init_ctx(c, cfg)) {
    neat_destroy(c);  // frees whatever was allocated before the failure; g_err keeps the message
    return e;
  }
  *out = c;
  return NEAT_OK;
}

void neat_destroy(neat_ctx* c) {
  if (!c) return;
  cudaFree(c->packed);
  cudaFree(c->packed_bf);
  cudaFree(c->g_src);
  cudaFree(c->g_dst_hi);
  cudaFree(c->g_dst_lo);
  cudaFree(c->g_scale);
  cudaFree(c->g_fsrc);
  cudaFree(c->g_fdst);
  cudaFree(c->ones_tile);
  cudaFree(c->jobs_dev);
  cudaFree(c->wn_dev[0]);
  cudaFree(c->wn_dev[1]);
  delete c;
}

size_t neat_param_count(const neat_ctx* c) { return c ? c->plan.n_params : 0; }

static const std::vector<LinearDims>* net_of(const neat_ctx* c, int net) {
  if (!c) return nullptr;
  return net == 0 ? &c->plan.sdf : net == 1 ? &c->plan.rend : net == 2 ? &c->plan.att : nullptr;
}
long neat_param_offset(const neat_ctx* c, int net, int layer, int kind) {
  const auto* v = net_of(c, net);
  if (!v || layer < 0 || layer >= static_cast<int>(v->size())) return -1;
  return static_cast<long>(kind == 0 ? (*v)[layer].w_off : (*v)[layer].b_off);
}
int neat_layer_dims(const neat_ctx* c, int net, int layer, int* in_f, int* out_f) {
  const auto* v = net_of(c, net);
  if (!v || layer < 0 || layer >= static_cast<int>(v->size())) return fail(NEAT_EINVAL, "bad net/layer");
  if (in_f) *in_f = (*v)[layer].in;
  if (out_f) *out_f = (*v)[layer].out;
  return NEAT_OK;
}

namespace {
int upload_wn_table(neat_ctx* c, const neat_wn_layer* layers, int n, cudaStream_t st) {
  const Plan& P = c->plan;
  const int expect = static_cast<int>(P.sdf.size() + P.rend.size() + P.att.size());
  if (!layers || n != expect || n > WN_MAX_LAYERS) return fail(NEAT_EINVAL, "weight_norm: wrong number of layers");
  WnTable t{};
  t.n = n;
  int rows = 0, i = 0;
  for (const std::vector<LinearDims>* net : {&P.sdf, &P.rend, &P.att})
    for (const LinearDims& d : *net) {
      neat_wn_layer L = layers[i];
      if (L.rows != d.out || L.cols != d.in || !L.v || !L.b) return fail(NEAT_EINVAL, "weight_norm: layer shape mismatch");
      L.w_off = static_cast<long>(d.w_off);
      L.b_off = static_cast<long>(d.b_off);
      t.l[i] = L;
      t.row_start[i] = rows;
      rows += d.out;
      ++i;
    }
  t.row_start[n] = rows;
  // two cached device tables: [0] the forward's (no gradient pointers), [1] the backward's -- each is uploaded once, the
  // first time it is seen (alternating one table between the two re-uploaded it twice per step, and an upload from a host
  // struct cannot be captured into a CUDA graph)
  const int slot = layers[0].gv != nullptr ? 1 : 0;
  if (!c->wn_dev[slot]) CK(cudaMalloc(&c->wn_dev[slot], sizeof(WnTable)));
  if (!c->wn_valid[slot] || std::memcmp(&t, &c->wn_host[slot], sizeof(WnTable)) != 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone)
      return fail(NEAT_EINVAL, "weight_norm: a new parameter table cannot be uploaded during stream capture (run one eager step first)");
    CK(cudaStreamSynchronize(st));  // a launch still reading the previous table must have finished
    c->wn_host[slot] = t;
    c->wn_valid[slot] = true;
    CK(cudaMemcpyAsync(c->wn_dev[slot], &c->wn_host[slot], sizeof(WnTable), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  c->wn_rows = rows;
  c->wn_slot = slot;
  return NEAT_OK;
}
}  // namespace

int neat_weight_norm_forward(neat_ctx* c, const neat_wn_layer* layers, int n_layers, float* flat, void* stream) {
  if (!c || !flat) return fail(NEAT_EINVAL, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int e = upload_wn_table(c, layers, n_layers, st)) return e;
  const int rows = c->wn_rows;
  weight_norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(c->wn_dev[c->wn_slot], flat);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_weight_norm_backward(neat_ctx* c, const neat_wn_layer* layers, int n_layers, const float* flat_grad, int accumulate,
                              void* stream) {
  if (!c || !flat_grad) return fail(NEAT_EINVAL, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int e = upload_wn_table(c, layers, n_layers, st)) return e;
  for (int i = 0; i < n_layers; ++i)
    if (!layers[i].gv || !layers[i].gb || (layers[i].g && !layers[i].gg)) return fail(NEAT_EINVAL, "weight_norm: null gradient pointer");
  const int rows = c->wn_rows;
  weight_norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(c->wn_dev[c->wn_slot], flat_grad, accumulate);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_pack_weights(neat_ctx* c, const float* flat, void* stream) {
  if (!c || !flat) return fail(NEAT_EINVAL, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(c->plan.g.src.size());
  const int nf = static_cast<int>(c->plan.g.fsrc.size());
  pack_pairs_kernel<<<(n + 255) / 256, 256, 0, st>>>(flat, c->g_src, c->g_dst_hi, c->g_dst_lo, c->g_scale, n,
                                                    reinterpret_cast<uint16_t*>(c->packed),
                                                    reinterpret_cast<uint16_t*>(c->packed_bf));
  ++g_launches;
  pack_f32_kernel<<<(nf + 255) / 256, 256, 0, st>>>(flat, c->g_fsrc, c->g_fdst, nf, reinterpret_cast<float*>(c->packed),
                                                   reinterpret_cast<float*>(c->packed_bf));
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

static int launch_query(neat_ctx* c, SdfQueryParams& p, void* stream) {
  if (p.M <= 0) return NEAT_OK;
  p.prog = c->prog_query;
  p.packed = c->packed;
  p.multires = c->plan.cfg.multires;
  p.sphere_r = p.sphere_r < 0.f ? 0.f : c->plan.cfg.sphere_radius;  // < 0 on entry: no sphere clamp (raw network sdf)
  p.sphere_scale = c->plan.cfg.sphere_scale;
  const int grid = grid_for(c, p.M);
  sdf_query_kernel<QUERY_STAGES>
      <<<grid, NUM_THREADS, engine_smem_bytes(QUERY_STAGES), static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_sdf_points(neat_ctx* c, const float* x, int M, float* sdf, void* stream) {
  if (!c || !x || !sdf || M < 0) return fail(NEAT_EINVAL, "bad argument");
  SdfQueryParams p{};
  p.x = x;
  p.sdf = sdf;
  p.M = M;
  p.n_per_ray = 1;
  return launch_query(c, p, stream);
}

int neat_sdf_grid(neat_ctx* c, const double* lo, const double* hi, const int* n, int clamp, float* sdf, void* stream) {
  if (!c || !lo || !hi || !n || !sdf) return fail(NEAT_EINVAL, "bad argument");
  long long M = 1;
  SdfQueryParams p{};
  for (int k = 0; k < 3; ++k) {
    if (n[k] < 1) return fail(NEAT_EINVAL, "grid size must be >= 1");
    M *= n[k];
    p.grid_n[k] = n[k];
    p.grid_lo[k] = lo[k];
    p.grid_hi[k] = hi[k];
    p.grid_step[k] = n[k] > 1 ? (hi[k] - lo[k]) / (n[k] - 1) : 0.0;  // np.linspace: step = delta / div
    if (n[k] == 1) p.grid_hi[k] = lo[k];
  }
  if (M > 0x7fffff00LL) return fail(NEAT_EINVAL, "grid too large (more than 2^31 points): evaluate it in slabs");
  p.sdf = sdf;
  p.M = static_cast<int>(M);
  p.n_per_ray = 1;
  p.sphere_r = clamp ? 0.f : -1.f;
  return launch_query(c, p, stream);
}

int neat_sdf_rays(neat_ctx* c, const float* rays_o, int o_stride, const float* rays_d, const float* z, int R, int n,
                  float* sdf, void* stream) {
  if (!c || !rays_o || !rays_d || !z || !sdf || R < 0 || n <= 0 || (o_stride != 0 && o_stride != 3))
    return fail(NEAT_EINVAL, "bad argument");
  if (static_cast<long long>(R) * n > 0x7fffffffLL) return fail(NEAT_EINVAL, "R*n too large");
  SdfQueryParams p{};
  p.rays_o = rays_o;
  p.o_stride = o_stride;
  p.rays_d = rays_d;
  p.z = z;
  p.sdf = sdf;
  p.M = R * n;
  p.n_per_ray = n;
  return launch_query(c, p, stream);
}

// ---------------------------------------------------------------- sampler
namespace {
struct SamplerWs {
  SamplerState* st;
  float *z, *sdf, *samples, *sdf_new, *beta;
};
inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }
SamplerWs carve_sampler_ws(void* ws, int R) {
  uint8_t* b = static_cast<uint8_t*>(ws);
  SamplerWs w;
  w.st = reinterpret_cast<SamplerState*>(b); b += 256;
  w.z = reinterpret_cast<float*>(b); b += al256(sizeof(float) * SMP_MAX_L * R);
  w.sdf = reinterpret_cast<float*>(b); b += al256(sizeof(float) * SMP_MAX_L * R);
  w.samples = reinterpret_cast<float*>(b); b += al256(sizeof(float) * SMP_MAX_NEW * R);
  w.sdf_new = reinterpret_cast<float*>(b); b += al256(sizeof(float) * SMP_MAX_NEW * R);
  w.beta = reinterpret_cast<float*>(b);
  return w;
}
int check_sampler_cfg(const neat_sampler_config* s) {
  if (!s || s->n_eval <= 0 || s->n_eval > SMP_MAX_NEW || s->n_eval * s->max_iters > SMP_MAX_L || s->max_iters < 1 ||
      s->max_iters > 8 || s->n_final <= 0 || s->n_final > SMP_MAX_NEW || s->n_extra < 0 ||
      s->n_final + 2 + s->n_extra > SMP_MAX_OUT)
    return fail(NEAT_EINVAL, "unsupported sampler configuration");
  return NEAT_OK;
}
SamplerParams mk_sampler_params(const neat_sampler_config* s, int R, const SamplerWs& w) {
  SamplerParams p{};
  p.R = R; p.n_eval = s->n_eval; p.n_final = s->n_final; p.n_extra = s->n_extra;
  p.beta_iters = s->beta_iters; p.max_iters = s->max_iters;
  p.near = s->near_; p.far = s->far_; p.eps = s->eps; p.beta_min = s->beta_min;
  p.st = w.st; p.z = w.z; p.sdf = w.sdf; p.samples = w.samples; p.sdf_new = w.sdf_new; p.beta = w.beta;
  return p;
}
}  // namespace

size_t neat_sampler_workspace_bytes(int R) {
  if (R < 0) return 0;
  return 256 + 2 * al256(sizeof(float) * SMP_MAX_L * R) + 2 * al256(sizeof(float) * SMP_MAX_NEW * R) +
         al256(sizeof(float) * R);
}

int neat_sampler_run(neat_ctx* c, const neat_sampler_config* s, const float* rays_o, int o_stride,
                     const float* rays_d, int R, const float* beta_param, const float* t_rand, const float* u_final,
                     void* workspace, int* n_iters_dev, void* stream) {
  if (!c || !rays_o || !rays_d || !beta_param || !workspace || R <= 0 || (o_stride != 0 && o_stride != 3))
    return fail(NEAT_EINVAL, "bad argument");
  if (int e = check_sampler_cfg(s)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const SamplerWs w = carve_sampler_ws(workspace, R);
  SamplerParams p = mk_sampler_params(s, R, w);
  p.beta_param = beta_param;
  p.t_rand = t_rand;
  p.u_final = u_final;
  p.training = u_final != nullptr;
  const int grid = (R + SMP_WARPS - 1) / SMP_WARPS;
  sampler_init_kernel<<<grid, 32 * SMP_WARPS, 0, st>>>(p);
  ++g_launches;
  for (int it = 0; it < s->max_iters; ++it) {
    // SDF of the new samples; the query kernel early-outs through a 0-tile launch guard on `done`
    SdfQueryParams q{};
    q.rays_o = rays_o; q.o_stride = o_stride; q.rays_d = rays_d;
    q.z = w.samples; q.sdf = w.sdf_new; q.M = R * SMP_MAX_NEW; q.n_per_ray = SMP_MAX_NEW;
    q.skip_flag = &w.st->done;
    if (int e = launch_query(c, q, stream)) return e;
    sampler_bounds_kernel<<<grid, 32 * SMP_WARPS, 0, st>>>(p, it);
    ++g_launches;
    sampler_draw_kernel<<<grid, 32 * SMP_WARPS, 0, st>>>(p, it);
    ++g_launches;
    sampler_finish_kernel<<<1, 1, 0, st>>>(w.st, it, s->max_iters);
    ++g_launches;
  }
  if (n_iters_dev) CK(cudaMemcpyAsync(n_iters_dev, &w.st->n_iters, sizeof(int), cudaMemcpyDeviceToDevice, st));
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_sampler_finish(neat_ctx* c, const neat_sampler_config* s, int R, const int64_t* extra_idx,
                        const int64_t* eik_idx, void* workspace, float* z_vals, float* z_eik, void* stream) {
  if (!c || !workspace || !z_vals || !z_eik || R <= 0 || (!extra_idx && s && s->n_extra > 0))
    return fail(NEAT_EINVAL, "bad argument");
  if (int e = check_sampler_cfg(s)) return e;
  const SamplerWs w = carve_sampler_ws(workspace, R);
  SamplerParams p = mk_sampler_params(s, R, w);
  p.extra_idx = extra_idx;
  p.eik_idx = eik_idx;
  p.training = eik_idx != nullptr;
  p.z_vals = z_vals;
  p.z_eik = z_eik;
  sampler_final_kernel<<<(R + SMP_WARPS - 1) / SMP_WARPS, 32 * SMP_WARPS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- render points
namespace {
int fill_points(const neat_ctx* c, const neat_points* pts, SdfQueryParams& q) {
  if (!pts) return fail(NEAT_EINVAL, "null points");
  q = SdfQueryParams{};
  if (pts->x) {
    q.x = pts->x;
    q.M = pts->M;
    q.n_per_ray = 1;
  } else {
    if (!pts->rays_o || !pts->rays_d || !pts->z || (pts->o_stride != 0 && pts->o_stride != 3) || pts->S <= 0)
      return fail(NEAT_EINVAL, "bad ray description");
    q.rays_o = pts->rays_o; q.o_stride = pts->o_stride; q.rays_d = pts->rays_d; q.z = pts->z;
    q.n_per_ray = pts->S;
    q.M = pts->R * pts->S;
  }
  if (q.M <= 0) return fail(NEAT_EINVAL, "empty point batch");
  q.multires = c->plan.cfg.multires;
  q.sphere_r = c->plan.cfg.sphere_radius;
  q.sphere_scale = c->plan.cfg.sphere_scale;
  return NEAT_OK;
}
}  // namespace

size_t neat_feat_tiles_bytes(int M) { return static_cast<size_t>((M + TILE_M - 1) / TILE_M) * TILE_MAIN_BYTES; }

size_t neat_sdf_save_bytes(const neat_ctx* c, int M, int training) {
  if (!c || M <= 0) return 0;
  const SdfSaveLayout lay = sdf_save_layout(c->plan.cfg.sdf_layers, training != 0);
  const size_t n = training ? static_cast<size_t>((M + TILE_M - 1) / TILE_M) : static_cast<size_t>(grid_for(c, M));
  return n * lay.total;
}

int neat_sdf_outputs(neat_ctx* c, const neat_points* pts, int clamp, int training, float* sdf, float* grad,
                     float* act, void* feat_tiles, void* save, void* stream) {
  if (!c || !grad || !save) return fail(NEAT_EINVAL, "bad argument");
  SdfRenderParams p{};
  if (int e = fill_points(c, pts, p.pts)) return e;
  const neat_net_config& g = c->plan.cfg;
  p.prog = c->prog_render;
  p.packed = c->packed;
  p.L = g.sdf_layers; p.skip = g.sdf_skip; p.H = g.sdf_hidden; p.E = c->plan.E; p.F = g.feat;
  p.clamp = clamp; p.training = training;
  p.w_last_row_off = c->plan.w_last_row_off;
  p.sdf = sdf; p.grad = grad; p.act = act;
  p.feat_tiles = static_cast<uint8_t*>(feat_tiles);
  p.save = static_cast<uint8_t*>(save);
  sdf_render_kernel<RENDER_STAGES><<<grid_for(c, p.pts.M), NUM_THREADS, engine_smem_bytes(RENDER_STAGES),
                                     static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

size_t neat_head_save_bytes(const neat_ctx* c, int M) {
  if (!c || M <= 0) return 0;
  return static_cast<size_t>((M + TILE_M - 1) / TILE_M) * head_save_layout(c->plan.cfg.head_layers).total;
}

int neat_head_forward(neat_ctx* c, int head, const neat_points* pts, const float* normals, const void* feat_tiles,
                      int training, void* save, float* out, void* stream) {
  if (!c || (head != 0 && head != 1) || !normals || !feat_tiles || !out || (training && !save))
    return fail(NEAT_EINVAL, "bad argument");
  HeadParams p{};
  if (int e = fill_points(c, pts, p.pts)) return e;
  if (pts->x && !pts->dirs) return fail(NEAT_EINVAL, "explicit points need explicit view dirs");
  p.prog = c->prog_head[head];
  p.packed = c->packed;
  p.dirs = pts->dirs;
  p.normals = normals;
  p.feat_tiles = static_cast<const uint8_t*>(feat_tiles);
  p.head = head;
  p.HL = c->plan.cfg.head_layers;
  p.multires_view = c->plan.cfg.multires_view;
  p.out_dim = head == 0 ? 3 : 6;
  p.training = training;
  p.save = static_cast<uint8_t*>(save);
  p.out = out;
  head_fwd_kernel<RENDER_STAGES><<<grid_for(c, p.pts.M), NUM_THREADS, engine_smem_bytes(RENDER_STAGES),
                                   static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_camera_rays(const float* uv, const float* pose, const float* K, int R, float* dirs, float* cam, void* stream) {
  if (!uv || !pose || !K || !dirs || !cam || R <= 0) return fail(NEAT_EINVAL, "bad argument");
  camera_rays_kernel<<<(R + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(uv, pose, K, R, dirs, cam);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_composite_forward(const neat_composite_args* a, void* stream) {
  if (!a || a->R <= 0 || a->S <= 0 || !a->z || !a->sdf || !a->rays_o || !a->rays_d || !a->beta_param ||
      (a->rgb && !a->rgb_values) || (a->lines && !a->lines3d) || (!a->rgb && !a->lines && !a->weights))
    return fail(NEAT_EINVAL, "bad argument");
  CompositeParams p{};
  p.R = a->R; p.S = a->S; p.z = a->z; p.sdf = a->sdf; p.rgb = a->rgb; p.lines = a->lines; p.normals = a->normals;
  p.rays_o = a->rays_o; p.rays_d = a->rays_d; p.beta_param = a->beta_param; p.beta_min = a->beta_min;
  p.weights = a->weights; p.rgb_values = a->rgb_values; p.lines3d = a->lines3d; p.depth = a->depth;
  p.points3d = a->points3d; p.normal_map = a->normals ? a->normal_map : nullptr;
  p.bg = a->rgb ? a->bg_color : nullptr;
  composite_fwd_kernel<<<(a->R + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_line_geometry(int R, const float* pose, const float* K, const float* uv_proj, const float* points3d,
                       const float* grad3d, const float* lines3d, float* pose_inv, float* lines2d,
                       float* lines2d_calib, float* l3d, void* stream) {
  if (R <= 0 || !pose || !K || !uv_proj || !points3d || !grad3d || !lines3d || !pose_inv || !lines2d ||
      !lines2d_calib || !l3d)
    return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pose_inverse_kernel<<<1, 32, 0, st>>>(pose, pose_inv);
  ++g_launches;
  GeometryParams p{};
  p.R = R; p.pose = pose; p.K = K; p.uv_proj = uv_proj; p.points3d = points3d; p.grad3d = grad3d;
  p.lines3d = lines3d; p.lines2d = lines2d; p.lines2d_calib = lines2d_calib; p.l3d = l3d; p.pose_inv = pose_inv;
  line_geometry_kernel<<<(R + 127) / 128, 128, 0, st>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- dataset-side attraction precompute
int neat_encodels(const float* lines, int input_height, int input_width, int height, int width, int num_lines, float* map,
                  uint8_t* label, float* tmap, void* stream) {
  if (!lines || !map || !label || !tmap || height <= 0 || width <= 0 || num_lines < 0 || input_height <= 0 || input_width <= 0)
    return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t hw = static_cast<size_t>(height) * width;
  CK(cudaMemsetAsync(map, 0, 6 * hw * sizeof(float), st));
  CK(cudaMemsetAsync(tmap, 0, hw * sizeof(float), st));
  CK(cudaMemsetAsync(label, 0, hw * static_cast<size_t>(num_lines), st));
  encodels_kernel<<<static_cast<int>((hw + 255) / 256), 256, 0, st>>>(lines, input_height, input_width, num_lines, height,
                                                                    width, map, label, tmap);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_point_line_attraction(const float* lines, int num_lines, int height, int width, float distance, uint8_t* mask,
                               long long* labels, float* proj_points, void* stream) {
  if (!lines || !mask || !labels || !proj_points || height <= 0 || width <= 0 || num_lines < 0)
    return fail(NEAT_EINVAL, "bad argument");
  const size_t hw = static_cast<size_t>(height) * width;
  point_line_attraction_kernel<<<static_cast<int>((hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      lines, num_lines, height, width, distance, mask, labels, proj_points);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- dataset pixel sampling
size_t neat_mask_compact_workspace_bytes(long long n) {
  if (n <= 0) return 0;
  return al256(sizeof(int) * static_cast<size_t>((n + PX_BLOCK - 1) / PX_BLOCK));
}

int neat_mask_compact(const uint8_t* mask, long long n, void* workspace, int* out_idx, int* n_out, void* stream) {
  if (!mask || n <= 0 || n > 0x7fffffffLL || !workspace || !out_idx || !n_out) return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = static_cast<int>((n + PX_BLOCK - 1) / PX_BLOCK);
  int* block = static_cast<int*>(workspace);
  mask_count_kernel<<<nb, PX_BLOCK, 0, st>>>(mask, n, block);
  ++g_launches;
  mask_scan_kernel<<<1, PX_BLOCK, 0, st>>>(block, nb, n_out);
  ++g_launches;
  mask_scatter_kernel<<<nb, PX_BLOCK, 0, st>>>(mask, n, block, out_idx);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_sample_pixels(const neat_pixel_args* a, void* stream) {
  if (!a || a->R < 0 || a->W <= 0 || !a->rgb_image || !a->labels || !a->att_points || !a->uv || !a->uv_proj || !a->rgb ||
      !a->labels_out || (a->lines2d && (!a->lines || a->n_lines <= 0)) || a->first < 0)
    return fail(NEAT_EINVAL, "bad argument");
  if (a->masked && (a->n_masked <= 0 || (!a->perm && a->R > a->n_masked)))
    return fail(NEAT_EINVAL, "sample_pixels: more rays than masked pixels");
  if (a->R == 0) return NEAT_OK;
  PixelParams p{};
  p.R = a->R; p.W = a->W; p.first = a->first; p.masked = a->masked; p.perm = a->perm;
  if (a->masked && !a->perm) p.draw = make_pixel_perm(static_cast<uint32_t>(a->n_masked), a->seed, a->step);
  p.rgb_image = a->rgb_image; p.labels = a->labels; p.att_points = a->att_points; p.lines = a->lines; p.n_lines = a->n_lines;
  p.uv = a->uv; p.uv_proj = a->uv_proj; p.rgb = a->rgb; p.lines2d = a->lines2d; p.labels_out = a->labels_out;
  p.index_out = a->index_out;
  sample_pixels_kernel<<<(a->R + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_pixel_permutation(unsigned n, unsigned long long seed, unsigned long long step, unsigned first, unsigned count,
                           unsigned* out) {
  if (n == 0 || n > 0x7fffffffu || !out || static_cast<unsigned long long>(first) + count > n)
    return fail(NEAT_EINVAL, "bad argument");
  const PixelPerm P = make_pixel_perm(n, seed, step);
  for (unsigned i = 0; i < count; ++i) out[i] = px_permute(P, first + i);
  return NEAT_OK;
}

// ---------------------------------------------------------------- loss
int neat_loss_forward_backward(const neat_loss_args* a, void* stream) {
  if (!a || a->R <= 0 || !a->rgb_values || !a->rgb_gt || !a->lines2d || !a->lines2d_calib || !a->lines_gt || !a->K3 ||
      !a->scratch || !a->out || !a->g_rgb || !a->g_calib || (a->grad_theta && (!a->g_theta || a->n_eik <= 0)))
    return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LossParams p{};
  p.R = a->R; p.n_eik = a->grad_theta ? a->n_eik : 0;
  p.rgb_values = a->rgb_values; p.rgb_gt = a->rgb_gt; p.lines2d = a->lines2d; p.lines2d_calib = a->lines2d_calib;
  p.lines_gt = a->lines_gt; p.labels = a->labels; p.K3 = a->K3; p.k_ld = a->k_ld; p.grad_theta = a->grad_theta;
  p.eikonal_weight = a->eikonal_weight; p.line_weight = a->line_weight;
  p.sums = a->scratch; p.per_uncal = a->scratch + 8;
  p.out = a->out; p.g_rgb = a->g_rgb; p.g_calib = a->g_calib; p.g_theta = a->g_theta;
  CK(cudaMemsetAsync(p.sums, 0, 8 * sizeof(float), st));
  const int n = std::max(p.R, p.n_eik);
  loss_terms_kernel<<<(n + 255) / 256, 256, 0, st>>>(p);
  ++g_launches;
  loss_calib_kernel<<<(p.R + 255) / 256, 256, 0, st>>>(p);
  ++g_launches;
  loss_grads_kernel<<<(n + 255) / 256, 256, 0, st>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_project_calib_backward(int R, const float* pose_inv, const float* lines3d, const float* g_calib,
                                float* g_lines3d, void* stream) {
  if (R <= 0 || !pose_inv || !lines3d || !g_calib || !g_lines3d) return fail(NEAT_EINVAL, "bad argument");
  project_calib_bwd_kernel<<<(R + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(R, pose_inv, lines3d, g_calib,
                                                                                         g_lines3d);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_project_points(int N, const float* pose_inv, const float* K, int k_ld, const float* X, float* out_pix,
                        float* out_calib, void* stream) {
  if (N < 0 || !pose_inv || !X || (out_pix && !K)) return fail(NEAT_EINVAL, "bad argument");
  if (N == 0) return NEAT_OK;
  project_points_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(N, pose_inv, K, k_ld, X, out_pix,
                                                                                      out_calib);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_project_points_backward(int N, const float* pose_inv, const float* K, int k_ld, const float* X,
                                 const float* g_pix, const float* g_calib, float* g_X, void* stream) {
  if (N < 0 || !pose_inv || !X || !g_X || (g_pix && !K)) return fail(NEAT_EINVAL, "bad argument");
  if (N == 0) return NEAT_OK;
  project_points_bwd_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(N, pose_inv, K, k_ld, X, g_pix,
                                                                                          g_calib, g_X);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_l3d_candidates(int R, const float* lines3d, const float* l3d, float* out, int* n_out, float* score, void* stream) {
  if (R <= 0 || !lines3d || !l3d || !out || !n_out) return fail(NEAT_EINVAL, "bad argument");
  if (R > L3D_MAX_R) return fail(NEAT_EUNSUPPORTED, "use_l3d: at most 4096 rays per call");
  l3d_candidates_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(R, lines3d, l3d, out, n_out, score);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_junction_terms(int n, const float* j3d_local, const float* j3d_global, const float* j2d_local_calib,
                        const float* j2d_global_calib, const float* j2d_local, const float* j2d_global, const int* rows,
                        const int* cols, float* out, void* stream) {
  if (n <= 0 || !j3d_local || !j3d_global || !j2d_local_calib || !j2d_global_calib || !j2d_local || !j2d_global || !rows ||
      !cols || !out)
    return fail(NEAT_EINVAL, "bad argument");
  junction_terms_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, j3d_local, j3d_global, j2d_local_calib,
                                                                        j2d_global_calib, j2d_local, j2d_global, rows, cols,
                                                                        out);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_junction_terms_backward(int n, int n_global, const float* j3d_local, const float* j3d_global,
                                 const float* j2d_local_calib, const float* j2d_global_calib, const int* rows,
                                 const int* cols, const float* g_out, float* g_j3d_global, float* g_j2d_global_calib,
                                 void* stream) {
  if (n <= 0 || n_global <= 0 || !j3d_local || !j3d_global || !j2d_local_calib || !j2d_global_calib || !rows || !cols ||
      !g_out || !g_j3d_global || !g_j2d_global_calib)
    return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaMemsetAsync(g_j3d_global, 0, sizeof(float) * 3 * n_global, st));
  CK(cudaMemsetAsync(g_j2d_global_calib, 0, sizeof(float) * 2 * n_global, st));
  junction_terms_bwd_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, j3d_local, j3d_global, j2d_local_calib, j2d_global_calib,
                                                             rows, cols, g_out, g_j3d_global, g_j2d_global_calib);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- finalisation: per-image line voting
size_t neat_line_vote_workspace_bytes(int N, int G) {
  if (N < 0 || G < 0) return 0;
  return al256(sizeof(int) * 2 * static_cast<size_t>(N)) + al256(sizeof(float) * 8 * static_cast<size_t>(G)) +
         al256(sizeof(int) * 4 * static_cast<size_t>(G));
}

int neat_line_vote(const float* lines2d, const float* lines3d, const float* points3d, int N, const float* gt_lines, int G,
                   float dis_threshold, void* workspace, float* lines3d_mean, float* scores, float* counts, void* stream) {
  if (N <= 0 || G <= 0 || !lines2d || !lines3d || !points3d || !gt_lines || !workspace || !lines3d_mean || !scores || !counts)
    return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* b = static_cast<uint8_t*>(workspace);
  int* assign = reinterpret_cast<int*>(b);
  b += al256(sizeof(int) * 2 * static_cast<size_t>(N));
  float* sums = reinterpret_cast<float*>(b);                 // [G,6]
  float* score_sums = sums + 6 * static_cast<size_t>(G);     // [G]
  b += al256(sizeof(float) * 8 * static_cast<size_t>(G));
  int* slot_count = reinterpret_cast<int*>(b);               // [G]   three-vote lines: their support points
  int* slots = slot_count + G;                               // [G,3]
  CK(cudaMemsetAsync(sums, 0, sizeof(float) * 7 * static_cast<size_t>(G), st));
  CK(cudaMemsetAsync(slot_count, 0, sizeof(int) * 4 * static_cast<size_t>(G), st));
  CK(cudaMemsetAsync(counts, 0, sizeof(float) * static_cast<size_t>(G), st));
  const int blocks = (2 * N + 255) / 256;
  line_vote_assign_kernel<<<blocks, 256, 0, st>>>(lines2d, lines3d, N, gt_lines, G, dis_threshold, assign, sums, counts);
  ++g_launches;
  line_vote_score_kernel<<<blocks, 256, 0, st>>>(points3d, N, assign, sums, counts, score_sums, slot_count, slots);
  ++g_launches;
  line_vote_finish_kernel<<<(G + 127) / 128, 128, 0, st>>>(G, sums, counts, score_sums, points3d, slots, lines3d_mean, scores);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_line_visibility(const float* lines3d, int L, const float* pose_inv, const float* K, int k_ld, const float* gt_lines,
                         int G, float dis_threshold, uint8_t* visible, float* mindis, void* stream) {
  if (L < 0 || G < 0 || !pose_inv || !K || (L && (!lines3d || !visible)) || (G && !gt_lines))
    return fail(NEAT_EINVAL, "bad argument");
  if (L == 0) return NEAT_OK;
  line_visibility_kernel<<<(L + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(lines3d, L, pose_inv, K, k_ld, gt_lines,
                                                                                      G, dis_threshold, visible, mindis);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_line_junction_graph(const float* lines3d, int N, const float* junctions, int J, float rel_threshold, int* midx,
                             uint8_t* matched, float* graph, uint8_t* upper, void* stream) {
  if (N < 0 || J <= 0 || !junctions || !graph || !upper || (N && (!lines3d || !midx || !matched)))
    return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaMemsetAsync(graph, 0, sizeof(float) * static_cast<size_t>(J) * J, st));
  CK(cudaMemsetAsync(upper, 0, static_cast<size_t>(J) * J, st));
  if (N == 0) return NEAT_OK;
  line_junction_graph_kernel<<<(N + 255) / 256, 256, 0, st>>>(lines3d, N, junctions, J, rel_threshold > 0.f ? 1 : 0, midx,
                                                             matched, graph, upper);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- optimizer step
namespace {
struct AdamCache {  // device copies of the tensor tables seen so far on one device (param groups, > 128 tensors: one
  // table per chunk); a table is uploaded once, the first time its tensor list is seen -- no synchronisation afterwards
  static constexpr int MAX = 16;
  AdamTable* host[MAX] = {};
  AdamTable* dev[MAX] = {};
  int n = 0, next = 0;
};
AdamCache g_adam[16];
}  // namespace

int neat_adam_step(const neat_adam_tensor* tensors, int n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, float grad_scale, void* stream) {
  if (n < 0 || n > ADAM_MAX_TENSORS || (n && !tensors) || step < 1) return fail(NEAT_EINVAL, "bad argument");
  if (n == 0) return NEAT_OK;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(NEAT_EINVAL, "device index out of range");
  AdamCache& c = g_adam[dev];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local AdamTable t;
  std::memset(&t, 0, sizeof(t));
  t.n = n;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].param || !tensors[i].grad || !tensors[i].exp_avg || !tensors[i].exp_avg_sq || tensors[i].numel < 0)
      return fail(NEAT_EINVAL, "adam: null tensor");
    t.t[i] = tensors[i];
    t.blk_start[i] = blocks;
    blocks += static_cast<int>((tensors[i].numel + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS);
  }
  for (int i = n; i <= ADAM_MAX_TENSORS; ++i) t.blk_start[i] = blocks;
  int hit = -1;
  for (int i = 0; i < c.n && hit < 0; ++i)
    if (std::memcmp(c.host[i], &t, sizeof(AdamTable)) == 0) hit = i;
  if (hit < 0) {  // first sight of this tensor list: one synchronous upload (the slot's previous table may still be read)
    hit = c.n < AdamCache::MAX ? c.n++ : (c.next++ % AdamCache::MAX);
    if (!c.host[hit]) c.host[hit] = new AdamTable();
    if (!c.dev[hit]) CK(cudaMalloc(&c.dev[hit], sizeof(AdamTable)));
    CK(cudaStreamSynchronize(st));
    *c.host[hit] = t;
    CK(cudaMemcpyAsync(c.dev[hit], c.host[hit], sizeof(AdamTable), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  if (blocks == 0) return NEAT_OK;
  const double bc1 = 1.0 - std::pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - std::pow(static_cast<double>(beta2), step);
  adam_step_kernel<<<blocks, 256, 0, st>>>(c.dev[hit], lr, beta1, beta2, eps, weight_decay, static_cast<float>(bc1),
                                          static_cast<float>(1.0 / std::sqrt(bc2)), grad_scale);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// graph-replayable variant: step count, bias corrections and hyper-parameters live in device memory (train_aux.cuh)
int neat_adam_step_device(const neat_adam_tensor* tensors, int n, const float* hyper_dev, float* state_dev, void* stream) {
  if (n <= 0 || n > ADAM_MAX_TENSORS || !tensors || !hyper_dev || !state_dev) return fail(NEAT_EINVAL, "bad argument");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(NEAT_EINVAL, "device index out of range");
  AdamCache& c = g_adam[dev];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local AdamTable t;
  std::memset(&t, 0, sizeof(t));
  t.n = n;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].param || !tensors[i].grad || !tensors[i].exp_avg || !tensors[i].exp_avg_sq || tensors[i].numel < 0)
      return fail(NEAT_EINVAL, "adam: null tensor");
    t.t[i] = tensors[i];
    t.blk_start[i] = blocks;
    blocks += static_cast<int>((tensors[i].numel + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS);
  }
  for (int i = n; i <= ADAM_MAX_TENSORS; ++i) t.blk_start[i] = blocks;
  int hit = -1;
  for (int i = 0; i < c.n && hit < 0; ++i)
    if (std::memcmp(c.host[i], &t, sizeof(AdamTable)) == 0) hit = i;
  if (hit < 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone)
      return fail(NEAT_EINVAL, "adam: a new tensor list cannot be uploaded during stream capture (run one eager step first)");
    hit = c.n < AdamCache::MAX ? c.n++ : (c.next++ % AdamCache::MAX);
    if (!c.host[hit]) c.host[hit] = new AdamTable();
    if (!c.dev[hit]) CK(cudaMalloc(&c.dev[hit], sizeof(AdamTable)));
    CK(cudaStreamSynchronize(st));
    *c.host[hit] = t;
    CK(cudaMemcpyAsync(c.dev[hit], c.host[hit], sizeof(AdamTable), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  if (blocks == 0) return NEAT_OK;
  adam_prepare_kernel<<<1, 32, 0, st>>>(hyper_dev, state_dev);
  ++g_launches;
  adam_graph_kernel<<<blocks, 256, 0, st>>>(c.dev[hit], hyper_dev, state_dev);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- training-step helpers (train_aux.cuh)
int neat_train_draws(const neat_sampler_config* s, int R, float radius, unsigned long long seed,
                     unsigned long long* counter_dev, float* t_rand, float* u_final, int64_t* extra_idx, int64_t* eik_idx,
                     float* eik_uniform, void* stream) {
  if (!s || R <= 0 || !counter_dev || !t_rand || !u_final || !eik_idx || !eik_uniform || (s->n_extra > 0 && !extra_idx))
    return fail(NEAT_EINVAL, "bad argument");
  if (int e = check_sampler_cfg(s)) return e;
  DrawParams p{};
  p.R = R; p.n_eval = s->n_eval; p.n_final = s->n_final; p.n_extra = s->n_extra; p.max_iters = s->max_iters;
  p.n_out = s->n_final + 2 + s->n_extra;
  p.radius = radius; p.seed = seed; p.counter = counter_dev;
  p.t_rand = t_rand; p.u_final = u_final; p.extra_idx = reinterpret_cast<long long*>(extra_idx);
  p.eik_idx = reinterpret_cast<long long*>(eik_idx); p.eik_uniform = eik_uniform;
  const long long quads = (static_cast<long long>(R) * (s->n_eval + s->n_final + 4) + 3) / 4 + s->max_iters * s->n_extra;
  const int grid = static_cast<int>(std::min<long long>((quads + 255) / 256, 148 * 8));
  train_draws_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_gemm_f32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int ta, int tb,
                  const float* bias, int relu, const float* mask, int ldm, int accumulate, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return fail(NEAT_EINVAL, "bad argument");
  GemmParams p{A, B, C, M, N, K, lda, ldb, ldc, ta, tb, bias, relu, mask, ldm, accumulate, K};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int gx = (N + GEMM_TN - 1) / GEMM_TN, gy = (M + GEMM_TM - 1) / GEMM_TM;
  int gz = 1;
  // split-K for reductions much longer than the output is large (the weight gradients: K = 1024 latents): enough CTAs
  // for every SM; only with a linear epilogue, and the partial sums are ADDED, so a non-accumulating call clears C first
  if (!bias && !relu && !mask && K >= 512 && gx * gy < 148) {
    p.k_chunk = 64;
    gz = (K + p.k_chunk - 1) / p.k_chunk;
    if (!accumulate) {
      if (ldc != N) return fail(NEAT_EINVAL, "gemm: split-K needs a dense C when it overwrites");
      CK(cudaMemsetAsync(C, 0, sizeof(float) * static_cast<size_t>(M) * N, st));
    }
  }
  gemm_f32_kernel<<<dim3(gx, gy, gz), 128, 0, st>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_colsum_f32(const float* X, int M, int N, int ldx, float* out, int accumulate, void* stream) {
  if (!X || !out || M <= 0 || N <= 0) return fail(NEAT_EINVAL, "bad argument");
  colsum_f32_kernel<<<(N + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, M, N, ldx, out, accumulate);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

int neat_junction_step(const int* packed_dev, int cap, int n_global, const float* j3d_global, const float* j2d_global_calib,
                       const float* j2d_global, float w3, float w2, float* out, float* g_j3d_global,
                       float* g_j2d_global_calib, void* stream) {
  if (!packed_dev || cap <= 0 || n_global <= 0 || !j3d_global || !j2d_global_calib || !j2d_global || !out || !g_j3d_global ||
      !g_j2d_global_calib)
    return fail(NEAT_EINVAL, "bad argument");
  JunctionStepParams p{packed_dev, cap, n_global, j3d_global, j2d_global_calib, j2d_global, w3, w2, out, g_j3d_global,
                       g_j2d_global_calib};
  junction_step_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- junction clustering
namespace {
int db_table_size(int N) {
  int T = 64;
  while (T < 2 * N) T <<= 1;
  return T;
}
DbscanWs carve_dbscan_ws(void* ws, int N) {
  uint8_t* b = static_cast<uint8_t*>(ws);
  DbscanWs w;
  w.T = db_table_size(N);
  w.tkey = reinterpret_cast<unsigned long long*>(b); b += al256(sizeof(unsigned long long) * w.T);
  w.sums = reinterpret_cast<double*>(b); b += al256(sizeof(double) * 3 * (N / 2 + 1));
  w.head = reinterpret_cast<int*>(b); b += al256(sizeof(int) * w.T);
  w.next = reinterpret_cast<int*>(b); b += al256(sizeof(int) * N);
  w.parent = reinterpret_cast<int*>(b); b += al256(sizeof(int) * N);
  w.count = reinterpret_cast<int*>(b); b += al256(sizeof(int) * N);
  w.cid = reinterpret_cast<int*>(b); b += al256(sizeof(int) * N);
  w.csize = reinterpret_cast<int*>(b);
  return w;
}
}  // namespace

size_t neat_dbscan_workspace_bytes(int N) {
  if (N <= 0) return 0;
  const int T = db_table_size(N);
  return al256(sizeof(unsigned long long) * T) + al256(sizeof(double) * 3 * (N / 2 + 1)) + al256(sizeof(int) * T) +
         4 * al256(sizeof(int) * N) + al256(sizeof(int) * (N / 2 + 1));
}

int neat_dbscan(const float* points, int N, float eps, void* workspace, float* centroids, int* n_clusters, void* stream) {
  if (!points || N <= 0 || !(eps > 0.f) || !workspace || !centroids || !n_clusters) return fail(NEAT_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DbscanWs w = carve_dbscan_ws(workspace, N);
  const int nb = (std::max(w.T, N) + 255) / 256, nbN = (N + 255) / 256;
  const float inv_eps = 1.0f / eps;
  const double eps2 = static_cast<double>(eps) * static_cast<double>(eps);
  db_init_kernel<<<nb, 256, 0, st>>>(w, N);
  ++g_launches;
  db_insert_kernel<<<nbN, 256, 0, st>>>(w, points, N, inv_eps);
  ++g_launches;
  db_link_kernel<<<nbN, 256, 0, st>>>(w, points, N, inv_eps, eps2);
  ++g_launches;
  db_count_kernel<<<nbN, 256, 0, st>>>(w, N);
  ++g_launches;
  db_rank_kernel<<<1, 1024, 0, st>>>(w, N, n_clusters);
  ++g_launches;
  db_sum_kernel<<<nbN, 256, 0, st>>>(w, points, N);
  ++g_launches;
  db_mean_kernel<<<(N / 2 + 255) / 256, 256, 0, st>>>(w, n_clusters, centroids);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

// ---------------------------------------------------------------- backward
int neat_composite_backward(const neat_composite_bwd_args* a, void* stream) {
  if (!a || a->R <= 0 || a->S <= 0 || a->S > 256 || !a->z || !a->sdf || !a->weights || !a->rgb || !a->rgb_values_bar ||
      !a->lines3d_bar || !a->beta_param || !a->rgb_pre_bar || !a->lines_bar || !a->sdf_bar || !a->beta_bar)
    return fail(NEAT_EINVAL, "bad argument");
  CompositeBwdParams p{};
  p.R = a->R; p.S = a->S; p.z = a->z; p.sdf = a->sdf; p.weights = a->weights; p.rgb = a->rgb; p.act = a->act;
  p.rgb_values_bar = a->rgb_values_bar; p.lines3d_bar = a->lines3d_bar; p.beta_param = a->beta_param;
  p.beta_min = a->beta_min; p.rgb_pre_bar = a->rgb_pre_bar; p.lines_bar = a->lines_bar; p.sdf_bar = a->sdf_bar;
  p.beta_bar = a->beta_bar;
  p.bg = a->bg_color;
  composite_bwd_kernel<<<(a->R + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

size_t neat_head_bwd_save_bytes(const neat_ctx* c, int M) {
  if (!c || M <= 0) return 0;
  return static_cast<size_t>((M + TILE_M - 1) / TILE_M) * head_bwd_layout(c->plan.cfg.head_layers).total;
}
size_t neat_feat_bar_bytes(int M) { return static_cast<size_t>((M + TILE_M - 1) / TILE_M) * 256 * TILE_M * sizeof(float); }

int neat_head_backward(neat_ctx* c, int head, int M, const float* out_bar, const void* fwd_save, void* bwd_save,
                       float* feat_bar, float* n_bar, int accumulate, void* stream) {
  if (!c || (head != 0 && head != 1) || M <= 0 || !out_bar || !fwd_save || !bwd_save || !feat_bar || !n_bar)
    return fail(NEAT_EINVAL, "bad argument");
  HeadBwdParams p{};
  p.prog = c->prog_head_bwd[head];
  if (c->epilogue_prefetch) {
    p.prog.pf_base = static_cast<const uint8_t*>(fwd_save);
    p.prog.pf_stride = head_save_layout(c->plan.cfg.head_layers).total;
  }
  p.packed = c->packed_bf;
  p.M = M; p.HL = c->plan.cfg.head_layers; p.out_dim = head == 0 ? 3 : 6;
  p.out_bar = out_bar;
  p.fwd_save = static_cast<const uint8_t*>(fwd_save);
  p.bwd_save = static_cast<uint8_t*>(bwd_save);
  p.feat_bar = feat_bar; p.n_bar = n_bar; p.accumulate = accumulate;
  head_bwd_kernel<RENDER_STAGES><<<grid_for(c, M), NUM_THREADS, engine_smem_bytes(RENDER_STAGES),
                                   static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

size_t neat_sdf_bwd_save_bytes(const neat_ctx* c, int M) {
  if (!c || M <= 0) return 0;
  return static_cast<size_t>((M + TILE_M - 1) / TILE_M) * sdf_bwd_layout(c->plan.cfg.sdf_layers).total;
}
size_t neat_sdf_bwd_scratch_bytes(const neat_ctx* c, int M) {
  if (!c || M <= 0) return 0;
  return static_cast<size_t>(grid_for(c, M)) * (c->plan.cfg.sdf_layers - 1) * 256 * TILE_M * sizeof(float);
}

int neat_sdf_backward(neat_ctx* c, const neat_points* pts, const float* n_bar, const float* s_bar, const float* feat_bar,
                      const float* act, const void* fwd_save, void* bwd_save, void* scratch, void* stream) {
  if (!c || !n_bar || !fwd_save || !bwd_save || !scratch) return fail(NEAT_EINVAL, "bad argument");
  SdfBwdParams p{};
  if (int e = fill_points(c, pts, p.pts)) return e;
  const neat_net_config& g = c->plan.cfg;
  p.prog = c->prog_sdf_bwd;
  if (c->epilogue_prefetch) {  // off by default: measured slower (sdf_bwd 1.42 -> 1.52 ms), see engine.cuh producer_loop
    p.prog.pf_base = static_cast<const uint8_t*>(fwd_save);
    p.prog.pf_stride = sdf_save_layout(g.sdf_layers, true).total;
  }
  p.packed = c->packed_bf;
  p.L = g.sdf_layers; p.skip = g.sdf_skip; p.H = g.sdf_hidden; p.E = c->plan.E; p.F = g.feat;
  p.n_bar = n_bar; p.s_bar = s_bar; p.feat_bar = feat_bar; p.act = act;
  p.fwd_save = static_cast<const uint8_t*>(fwd_save);
  p.bwd_save = static_cast<uint8_t*>(bwd_save);
  p.zhat = static_cast<float*>(scratch);
  sdf_bwd_kernel<RENDER_STAGES><<<grid_for(c, p.pts.M), NUM_THREADS, engine_smem_bytes(RENDER_STAGES),
                                  static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

namespace {
struct Planes {  // an operand: per-tile base + stride, offsets of the hi / lo planes, columns available, format
  const uint8_t* base;
  uint64_t stride;
  uint32_t hi, lo;
  int cols;
  int f16;  // 1: fp16 pairs (written by a forward kernel), 0: bf16 pairs (written by a backward kernel)
  int main; // 1: main operand tile (half-tile-contiguous global layout, engine.cuh), 0: aux tile (shared-memory order)
};
Planes main_planes(const void* base, uint64_t stride, uint32_t off, int cols, int f16) {
  return Planes{static_cast<const uint8_t*>(base), stride, off, off + PLANE_MAIN_BYTES, cols, f16, 1};
}
Planes aux_planes(const void* base, uint64_t stride, uint32_t off, int cols, int f16) {
  return Planes{static_cast<const uint8_t*>(base), stride, off, off + PLANE_AUX_BYTES, cols, f16, 0};
}
inline int c16(int x) { return (x + 15) / 16 * 16; }
inline int c8(int x) { return (x + 7) / 8 * 8; }

// dW[row0 + i][col0 + j] += scale * sum_pt X[pt][i] Y[pt][j],  i < x_valid, j < y_valid ; optional bias
void add_gemm(std::vector<WJob>& jobs, const Planes& X, int x_valid, const Planes& Y, int y_valid, float* out, int ld,
              int row0, int col0, float scale, float* bias, int n_tiles, int n_split) {
  // the two 128-row halves of one (GEMM, split) read the SAME Y tiles: adjacent block indices run at the same time
  // on different SMs, so the second reader hits the L2 instead of HBM
  for (int sp = 0; sp < n_split; ++sp) {
    for (int m0 = 0; m0 < x_valid; m0 += 128) {
      WJob j{};
      j.x_base = X.base; j.x_stride = X.stride; j.x_hi = X.hi; j.x_lo = X.lo;
      j.y_base = Y.base; j.y_stride = Y.stride; j.y_hi = Y.hi; j.y_lo = Y.lo;
      j.m0 = m0;
      j.x_cols = std::min(128, c8(X.cols - m0));
      j.n_cols = std::min(256, c16(Y.cols));
      j.x_valid = std::min(128, x_valid - m0);
      j.y_valid = y_valid;
      j.out = out; j.ld = ld; j.row0 = row0 + m0; j.col0 = col0; j.scale = scale;
      j.bias = bias;
      j.n_tiles = n_tiles; j.split = sp; j.n_split = n_split;
      j.x_f16 = X.f16; j.y_f16 = Y.f16;
      j.x_main = X.main; j.y_main = Y.main;
      jobs.push_back(j);
    }
  }
}
}  // namespace

int neat_weight_gradients(neat_ctx* c, const neat_grad_group* groups, int n_groups, float* flat, void* stream) {
  if (!c || !groups || n_groups <= 0 || !flat) return fail(NEAT_EINVAL, "bad argument");
  const Plan& P = c->plan;
  const neat_net_config& g = P.cfg;
  const int L = g.sdf_layers, HL = g.head_layers, H = g.sdf_hidden, E = P.E, F = g.feat, S = g.sdf_skip;
  const SdfSaveLayout fl = sdf_save_layout(L, true);
  const SdfBwdSaveLayout bl = sdf_bwd_layout(L);
  const HeadSaveLayout hfl = head_save_layout(HL);
  const HeadBwdSaveLayout hbl = head_bwd_layout(HL);
  const float rs2 = 0.70710678118654752f;
  std::vector<WJob> jobs;
  for (int gi = 0; gi < n_groups; ++gi) {
    const neat_grad_group& G = groups[gi];
    if (G.M <= 0) continue;
    const int nt = (G.M + TILE_M - 1) / TILE_M;
    // tile-range splits per GEMM: ~48 tiles per CTA, at most 32 (measured: 784 tiles -> 16 splits 1.29 -> 1.22 ms vs 8;
    // 6272 tiles -> 32 splits 10.9 -> 9.6..10.1 ms; more splits only add flush traffic).  NEAT_WGRAD_SPLIT overrides the cap.
    const int ns = std::max(1, std::min(c->wgrad_max_split, nt / std::max(1, c->wgrad_tiles_per_split)));
    if (G.sdf_fwd_save && G.sdf_bwd_save) {
      const Planes pe = aux_planes(G.sdf_fwd_save, fl.total, fl.pe, c16(E), 0);
      const Planes p0 = aux_planes(G.sdf_bwd_save, bl.total, bl.p_aux, c16(E), 0);
      const Planes ones = aux_planes(c->ones_tile, 0, 0, 16, 0);
      for (int l = 0; l < L; ++l) {
        const LinearDims d = P.sdf[l];
        float* Wg = flat + d.w_off;
        float* bg = flat + d.b_off;
        // ---- reverse-sweep term: z_bar_l^T u_l
        std::vector<std::pair<Planes, std::pair<int, int>>> xs;  // (planes, (rows valid, row0))
        if (l < L - 1) {
          xs.push_back({main_planes(G.sdf_bwd_save, bl.total, bl.zb + l * TILE_MAIN_BYTES, P.sdf_f[l].npad, 0), {d.out, 0}});
        } else {
          xs.push_back({aux_planes(G.sdf_bwd_save, bl.total, bl.zb_aux, 16, 0), {1, 0}});
          xs.push_back({main_planes(G.sdf_bwd_save, bl.total, bl.zb + l * TILE_MAIN_BYTES, F, 0), {F, 1}});
        }
        for (auto& xe : xs) {
          const Planes& X = xe.first;
          const int xv = xe.second.first, r0 = xe.second.second;
          if (l == 0) {
            add_gemm(jobs, X, xv, pe, E, Wg, d.in, r0, 0, 1.f, bg, nt, ns);
          } else if (l == S) {
            const Planes um = main_planes(G.sdf_fwd_save, fl.total, fl.u + (l - 1) * TILE_MAIN_BYTES, H - E, 0);
            add_gemm(jobs, X, xv, um, H - E, Wg, d.in, r0, 0, rs2, nullptr, nt, ns);
            add_gemm(jobs, X, xv, pe, E, Wg, d.in, r0, H - E, rs2, nullptr, nt, ns);
            // bias gradient is unscaled: a separate ones-only job through the aux planes (E columns, none written)
            add_gemm(jobs, X, xv, pe, 0, Wg, d.in, r0, 0, 1.f, bg, nt, ns);
          } else {
            const Planes um = main_planes(G.sdf_fwd_save, fl.total, fl.u + (l - 1) * TILE_MAIN_BYTES, d.in, 0);
            add_gemm(jobs, X, xv, um, d.in, Wg, d.in, r0, 0, 1.f, bg, nt, ns);
          }
        }
        // ---- tangent term: a_l^T p_in
        const Planes A = l < L - 1 ? main_planes(G.sdf_fwd_save, fl.total, fl.a + l * TILE_MAIN_BYTES, P.sdf_f[l].npad, 0) : ones;
        const int av = l < L - 1 ? d.out : 1;
        if (l == 0) {
          add_gemm(jobs, A, av, p0, E, Wg, d.in, 0, 0, 1.f, nullptr, nt, ns);
        } else if (l == S) {
          const Planes pm = main_planes(G.sdf_bwd_save, bl.total, bl.p + (l - 1) * TILE_MAIN_BYTES, H - E, 0);
          add_gemm(jobs, A, av, pm, H - E, Wg, d.in, 0, 0, rs2, nullptr, nt, ns);
          add_gemm(jobs, A, av, p0, E, Wg, d.in, 0, H - E, rs2, nullptr, nt, ns);
        } else {
          const Planes pm = main_planes(G.sdf_bwd_save, bl.total, bl.p + (l - 1) * TILE_MAIN_BYTES, d.in, 0);
          add_gemm(jobs, A, av, pm, d.in, Wg, d.in, 0, 0, 1.f, nullptr, nt, ns);
        }
      }
    }
    for (int h = 0; h < 2; ++h) {
      if (!G.head_fwd_save[h] || !G.head_bwd_save[h] || !G.feat_tiles) continue;
      const std::vector<LinearDims>& net = h == 0 ? P.rend : P.att;
      const int aux_in = net[0].in - F;
      for (int l = 0; l < HL; ++l) {
        const LinearDims d = net[l];
        float* Wg = flat + d.w_off;
        float* bg = flat + d.b_off;
        const Planes X = l < HL - 1 ? main_planes(G.head_bwd_save[h], hbl.total, hbl.zb + l * TILE_MAIN_BYTES, d.out, 0)
                                    : aux_planes(G.head_bwd_save[h], hbl.total, hbl.zb_aux, 16, 0);
        if (l == 0) {
          const Planes ft = main_planes(G.sdf_fwd_save, fl.total, fl.feat, F, 0);  // the bf16 copy in the save record
          const Planes ax = aux_planes(G.head_fwd_save[h], hfl.total, hfl.aux, c16(aux_in), 0);
          add_gemm(jobs, X, d.out, ft, F, Wg, d.in, 0, aux_in, 1.f, bg, nt, ns);
          add_gemm(jobs, X, d.out, ax, aux_in, Wg, d.in, 0, 0, 1.f, nullptr, nt, ns);
        } else {
          const Planes um = main_planes(G.head_fwd_save[h], hfl.total, hfl.u + (l - 1) * TILE_MAIN_BYTES, d.in, 0);
          add_gemm(jobs, X, d.out, um, d.in, Wg, d.in, 0, 0, 1.f, bg, nt, ns);
        }
      }
    }
  }
  if (jobs.empty()) return NEAT_OK;
  {
    // longest-processing-time-first: CTAs are handed to SMs in block order, so the expensive (job pairs) go first and
    // the cheap ones fill the tail.  A pair = the 128-row halves of one (GEMM, split); it stays adjacent (shared Y tiles).
    auto cost = [](const WJob& j) {
      const long tiles = j.n_tiles > j.split ? (j.n_tiles - j.split + j.n_split - 1) / j.n_split : 0;
      return tiles * (j.x_cols + j.n_cols);
    };
    auto same_group = [](const WJob& a, const WJob& b) {
      return a.x_base == b.x_base && a.x_hi == b.x_hi && a.y_base == b.y_base && a.y_hi == b.y_hi && a.split == b.split &&
             a.out == b.out && a.col0 == b.col0 && a.bias == b.bias;
    };
    std::vector<int> gid(jobs.size());
    std::vector<long> gcost;
    for (size_t i = 0; i < jobs.size(); ++i) {
      if (i > 0 && same_group(jobs[i - 1], jobs[i])) {
        gid[i] = gid[i - 1];
        gcost[gid[i]] += cost(jobs[i]);
      } else {
        gid[i] = static_cast<int>(gcost.size());
        gcost.push_back(cost(jobs[i]));
      }
    }
    std::vector<int> order(jobs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (gid[a] == gid[b]) return a < b;
      if (gcost[gid[a]] != gcost[gid[b]]) return gcost[gid[a]] > gcost[gid[b]];
      return gid[a] < gid[b];
    });
    std::vector<WJob> sorted(jobs.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = jobs[order[i]];
    jobs.swap(sorted);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (static_cast<int>(jobs.size()) > c->jobs_cap) {
    CK(cudaStreamSynchronize(st));
    cudaFree(c->jobs_dev);
    c->jobs_cap = static_cast<int>(jobs.size()) * 2;
    CK(cudaMalloc(&c->jobs_dev, sizeof(WJob) * c->jobs_cap));
  }
  // the job table only depends on buffer addresses and sizes, which are the same every step (persistent workspaces):
  // upload it when it changed, not every step (a > 64 KB copy from pageable memory blocks the host on the stream)
  if (c->jobs_last.size() != jobs.size() ||
      std::memcmp(c->jobs_last.data(), jobs.data(), sizeof(WJob) * jobs.size()) != 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone)
      return fail(NEAT_EINVAL, "weight_gradients: a new job table cannot be uploaded during stream capture (run one eager step first)");
    CK(cudaStreamSynchronize(st));  // a launch still reading the previous table must have finished
    c->jobs_last = jobs;  // (the copy below reads the context's own vector: it outlives the call)
    CK(cudaMemcpyAsync(c->jobs_dev, c->jobs_last.data(), sizeof(WJob) * jobs.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  wgrad_kernel<<<static_cast<int>(jobs.size()), WG_THREADS, sizeof(WgradSmem) + 1024, st>>>(c->jobs_dev);
  ++g_launches;
  CK(cudaGetLastError());
  return NEAT_OK;
}

long long neat_launch_count(void) { return g_launches.load(); }

int neat_set_precision(neat_ctx* c, int fast) {
  if (!c) return fail(NEAT_EINVAL, "null context");
  for (Program* p : {&c->prog_query, &c->prog_render, &c->prog_head[0], &c->prog_head[1], &c->prog_head_bwd[0],
                     &c->prog_head_bwd[1], &c->prog_sdf_bwd}) {
    p->fast = fast ? 1 : 0;
    for (int i = 0; i < p->n; ++i)
      p->s[i].comp = 1.0f + 0.35f * (fast ? 1.0f : 3.0f) * static_cast<float>(p->s[i].w.nk_main + p->s[i].w.nk_aux) * 5.9604645e-08f;
  }
  return NEAT_OK;
}

// ---------------------------------------------------------------- debug / bring-up

// Tests only: cap the persistent grid (several tiles per CTA at small point counts) and force the wgrad tile-range
// splitting (max_split pieces of >= tiles_per_split tiles); 0 restores the defaults.  Scratch sizes depend on the grid:
// set the cap before the first launch that sizes a workspace.
int neat_debug_set_grid_cap(neat_ctx* c, int max_ctas) {
  if (!c || max_ctas < 0) return fail(NEAT_EINVAL, "bad argument");
  c->grid_cap = max_ctas;
  return NEAT_OK;
}
int neat_debug_set_wgrad_split(neat_ctx* c, int max_split, int tiles_per_split) {
  if (!c || max_split < 0 || tiles_per_split < 0) return fail(NEAT_EINVAL, "bad argument");
  c->wgrad_max_split = max_split > 0 ? max_split : 32;
  c->wgrad_tiles_per_split = tiles_per_split > 0 ? tiles_per_split : 48;
  return NEAT_OK;
}

int neat_debug_set_flags(unsigned flags) {
  CK(cudaMemcpyToSymbol(g_dbg, &flags, sizeof(unsigned)));
  return NEAT_OK;
}
int neat_debug_set_l2_prefetch(int on) {
  CK(cudaMemcpyToSymbol(g_l2_prefetch, &on, sizeof(int)));
  return NEAT_OK;
}

}  // extern "C"
