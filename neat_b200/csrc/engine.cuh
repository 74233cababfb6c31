// Tile-MLP engine shared by all tensor-core kernels (v2: column-group pipelining).
//
// One CTA owns a tile of 128 points.  TMEM lane i == tile row i.  16 epilogue warps: warp w serves row quadrant
// q = w % 4 (the TMEM lanes a warp may touch) and, inside every 64-column group, the 16-column slice j = w / 4.
// The activation tile ("A") stays resident in shared memory as two bf16 planes (hi / lo split of the fp32
// value) in the UMMA no-swizzle canonical layout, chunk-major:
//     element (row r, column k)  ->  plane + (k / 8) * 2048 + r * 16 + (k % 8) * 2      [bytes]
// i.e. an 8-column chunk of all 128 rows is one contiguous 2 KB block.  The same bytes are a valid K-major
// operand (layer GEMMs) and a valid MN-major operand (weight-gradient GEMMs).  Columns [0,256) are the "main"
// segment, columns [256,304) the "aux" segment (positional encoding, view dirs, normals, ...).
//
// Pipelining: the accumulator of layer l lives in one 256-column half of TMEM while layer l+1 accumulates into
// the other.  All 16 warps work on column group 0 first, then 1, 2, 3; after each group they arrive on a_ready[g]
// and the MMA issuer starts the k-steps that read those 64 columns: the MMAs of layer l+1 run underneath the
// epilogue of layer l from its first quarter on.
//
// Weights are pre-packed (api.cu) into per-k-step slabs in exactly the B-operand layout, so a slab is one
// contiguous cp.async.bulk (1-D TMA) into a ring of shared-memory stages:
//     slab(k-step) = hi plane [2 chunks][npad rows][8] bf16, then lo plane (same shape)
// Warp 16 lane 0 is the weight producer, warp 17 lane 0 issues tcgen05.mma (3 per k-step: hi*hi+hi*lo+lo*hi).
#pragma once
#include "layout.h"
#include "umma.cuh"

namespace neat {

// k_stride: bytes between core matrices adjacent along K; mn_stride: along M/N
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr, uint32_t k_stride, uint32_t mn_stride) {
  return make_desc(saddr, k_stride, mn_stride);
}
// a descriptor whose start address is `bytes` further (the 14-bit address field cannot carry: smem < 256 KB)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

template <int STAGES>
struct alignas(1024) EngineSmem {
  uint8_t a_hi[A_PLANE_BYTES];
  uint8_t a_lo[A_PLANE_BYTES];
  uint8_t w[STAGES][W_STAGE_BYTES];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t a_ready[N_GROUPS];  // 16 arrivals each: lane 0 of every epilogue warp, once per column group and stage
  uint64_t aux_ready;          // 4 arrivals: lane 0 of the slice-0 warps
  uint64_t d_ready;
  uint64_t in_ready;           // bulk loads of input operand tiles (heads / backward kernels)
  uint64_t wr_done;            // 16 arrivals: every epilogue warp wrote (and fenced) its part of the A planes
  uint64_t rd_done;            // 1 arrival: the lead thread's bulk store finished reading the A planes
  uint32_t tmem_base;
};

template <int STAGES>
__device__ __forceinline__ void engine_init(EngineSmem<STAGES>& sm) {
  const int warp = warp_idx_uniform();
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    for (int g = 0; g < N_GROUPS; ++g) mbar_init(&sm.a_ready[g], EPI_WARPS);
    mbar_init(&sm.aux_ready, 4);
    mbar_init(&sm.d_ready, 1);
    mbar_init(&sm.in_ready, 1);
    mbar_init(&sm.wr_done, EPI_WARPS);
    mbar_init(&sm.rd_done, 1);
    fence_mbar_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// Warpgroup register reallocation: the producer / MMA warpgroup (whose loops live in uniform registers) shrinks to
// AUX_REGS per thread, the four epilogue warpgroups grow from the launch bound's 96 to EPI_REGS.  Measured motivation
// (round 2): the backward epilogues keep one 8-column unit of saved tensors in flight per thread because a second one
// spills at 96 registers (sdf_bwd 1.11 -> 1.30 ms with the spills), and every unit is an HBM round trip.
// The first statement of every role branch after engine_init (all four warps of a warpgroup execute the same one).
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int STAGES>
__device__ __forceinline__ void engine_fini(EngineSmem<STAGES>& sm) {
  tc_fence_before();
  __syncthreads();
  if (warp_idx_uniform() == EPI_WARPS) tmem_dealloc(sm.tmem_base, TMEM_COLS);
}

// producer warp (all lanes, converged; one elected lane issues): stream every slab of every step, for every tile.
// While the slabs of step i are streamed it also asks the L2 for what the epilogue of step i+1 will read from the
// tile's record (Step::pf_*), one slice per k-step: those epilogue loads are dependent-latency bound (a thread can keep
// one 8-column unit in flight), so turning their HBM misses (~1.2 us) into L2 hits (~0.35 us) is worth more than any
// instruction tuning -- was the hypothesis.  MEASURED: slower either way (issued by an epilogue thread: sdf_bwd 1.37 ->
// 1.57 ms; issued here: 1.42 -> 1.52 ms), so Program::pf_base stays nullptr unless NEAT_EPILOGUE_PREFETCH=1 asks for the
// experiment.  A second rejected idea from the same session: dropping the fp32 sigma' and zhat tensors and rebuilding
// them from the saved bf16 hi/lo activations (-14 % DRAM bytes for sdf_bwd) nearly doubled the executed instructions
// (294 M -> 515 M) and cost 0.1 ms: these epilogues are instruction/latency bound, not byte bound.
template <int STAGES>
__device__ __forceinline__ void producer_loop(EngineSmem<STAGES>& sm, const Program& prog, const uint8_t* packed,
                                              int n_tiles) {
  uint32_t stage = 0, phase = 0;
  const bool pf_on = prog.pf_base != nullptr && g_l2_prefetch != 0;
  for (int t = 0; t < n_tiles; ++t) {
    const uint64_t tile = blockIdx.x + static_cast<uint64_t>(t) * gridDim.x;
    for (int i = 0; i < prog.n; ++i) {
      const PLayer w = prog.s[i].w;
      const uint32_t slab = static_cast<uint32_t>(w.npad) * 64u;
      const uint32_t bytes = prog.fast ? slab / 2 : slab;  // the hi plane is the first half of a slab
      const int nk = w.nk_main + w.nk_aux;
      const uint8_t* src = packed + w.off;
      // prefetch target: the next step of this tile, or the first step of this CTA's next tile
      const bool wrap = i + 1 == prog.n;
      const Step& nx = prog.s[wrap ? 0 : i + 1];
      const bool pf = pf_on && (!wrap || t + 1 < n_tiles);
      const uint8_t* rec = prog.pf_base + (tile + (wrap ? gridDim.x : 0)) * prog.pf_stride;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(&sm.empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sm.full[stage], bytes);
          bulk_g2s(sm.w[stage], src + static_cast<size_t>(ks) * slab, bytes, &sm.full[stage]);
          if (pf) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const uint32_t total = nx.pf_bytes[r];
              const uint32_t piece = ((total + nk - 1) / nk + 127u) & ~127u;
              const uint32_t lo = piece * ks;
              if (lo < total) bulk_prefetch_l2(rec + nx.pf_off[r] + lo, min(piece, total - lo));
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
}

// MMA warp (all lanes, converged; one elected lane issues the tcgen05 instructions).  The loop is the critical path of a
// tile (the epilogue warps wait for it 35-45 % of their time, and 57 % of this warp's own samples are instruction issue),
// so it is kept to running descriptors (one add per k-step), a single iteration counter for stage / parity, and one asm
// block per k-step.
template <int STAGES>
__device__ __forceinline__ void mma_loop(EngineSmem<STAGES>& sm, const Program& prog, int n_tiles) {
  static_assert((STAGES & (STAGES - 1)) == 0, "stage count must be a power of two");
  uint32_t it = 0;  // k-steps issued so far: stage = it % STAGES, parity = (it / STAGES) & 1
  uint32_t a_phase = 0, aux_phase = 0;
  const uint32_t tmem = __shfl_sync(0xffffffffu, sm.tmem_base, 0);
  const uint64_t da_hi0 = make_desc_k(smem_u32(sm.a_hi), A_CHUNK_BYTES, 128);
  const uint64_t da_lo0 = make_desc_k(smem_u32(sm.a_lo), A_CHUNK_BYTES, 128);
  const uint32_t w0 = smem_u32(sm.w[0]);
  const uint32_t fast = prog.fast != 0 ? 1u : 0u;
  constexpr uint32_t A_KSTEP = 16 * (A_CHUNK_BYTES / 8) >> 4;  // descriptor-address units per 16 columns of A
  for (int t = 0; t < n_tiles; ++t) {
    for (int i = 0; i < prog.n; ++i) {
      const Step st = prog.s[i];
      const uint32_t npad = st.w.npad;
      const uint32_t idesc = make_idesc(TILE_M, npad, 0, 0, prog.a_f16, prog.b_f16);
      const uint32_t d = tmem + st.d_col;
      const uint64_t db0 = make_desc_k(w0, npad * 16, 128);  // stage 0, hi plane; lo plane = + npad * 32 bytes
      const uint32_t b_lo_off = (npad * 32) >> 4;
      uint32_t acc = 0;
      auto kstep = [&](uint64_t da_hi, uint64_t da_lo) {
        const uint32_t stage = it & (STAGES - 1);
        mbar_wait(&sm.full[stage], (it / STAGES) & 1);
        tc_fence_after();
        const uint64_t db_hi = db0 + stage * (W_STAGE_BYTES >> 4);
        if (elect_one()) umma_kstep_bf16x3(d, da_hi, da_lo, db_hi, db_hi + b_lo_off, idesc, acc, fast, &sm.empty[stage]);
        __syncwarp();
        acc = 1u;
        ++it;
      };
      if (st.wait_aux) {
        mbar_wait(&sm.aux_ready, aux_phase);
        aux_phase ^= 1;
        tc_fence_after();
      }
      const int nk_main = st.w.nk_main;
      uint64_t da_hi = da_hi0, da_lo = da_lo0;
      for (int g = 0; g < N_GROUPS; ++g) {
        if (st.wait_a) {
          mbar_wait(&sm.a_ready[g], a_phase);
          tc_fence_after();
        }
        const int k_end = min(nk_main, (g + 1) * (GROUP_COLS / 16));
        for (int ks = g * (GROUP_COLS / 16); ks < k_end; ++ks) {
          kstep(da_hi, da_lo);
          da_hi += A_KSTEP;
          da_lo += A_KSTEP;
        }
      }
      if (st.wait_a) a_phase ^= 1;
      da_hi = da_hi0 + (A_MAIN_COLS / 16) * A_KSTEP;
      da_lo = da_lo0 + (A_MAIN_COLS / 16) * A_KSTEP;
      for (int ks = 0; ks < st.w.nk_aux; ++ks) {
        kstep(da_hi, da_lo);
        da_hi += A_KSTEP;
        da_lo += A_KSTEP;
      }
      if (st.commit_d) {
        if (elect_one()) umma_commit(&sm.d_ready);
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------- epilogue-side helpers (warps 0..15)
struct Epi {
  int q, j, lane, row;  // row quadrant, 16-column slice inside a group, lane, tile row
  uint32_t tm;          // TMEM address of this warp's lanes, column 0
  uint32_t d_phase, wr_phase, rd_phase;
  bool lead;            // thread 0: issues bulk stores
};
template <int STAGES>
__device__ __forceinline__ Epi epi_make(const EngineSmem<STAGES>& sm) {
  Epi e;
  const int warp = warp_idx_uniform();
  e.lane = threadIdx.x & 31;
  e.q = warp & 3;
  e.j = warp >> 2;
  e.row = e.q * 32 + e.lane;
  e.tm = __shfl_sync(0xffffffffu, sm.tmem_base, 0) + (static_cast<uint32_t>(e.q * 32) << 16);
  e.d_phase = 0;
  e.wr_phase = 0;
  e.rd_phase = 0;
  e.lead = threadIdx.x == 0;
  return e;
}
template <int STAGES>
__device__ __forceinline__ void epi_wait_d(EngineSmem<STAGES>& sm, Epi& e) {
  mbar_wait(&sm.d_ready, e.d_phase);
  e.d_phase ^= 1;
  tc_fence_after();
}
// column offset of this warp's 16-column slice in group g
__device__ __forceinline__ int epi_col(const Epi& e, int g) { return g * GROUP_COLS + 16 * e.j; }
// The kernels whose epilogue also reads saved tensors from global memory walk their 64 columns per layer as 8
// "units" of 8 columns (unit u = group u/2, half u%2) and keep the loads of unit u+1 in flight while unit u is
// computed; the loads of unit 0 are issued BEFORE waiting for the accumulator, under the tail of the MMAs.
constexpr int N_UNITS = 2 * N_GROUPS;
__device__ __forceinline__ int epi_unit_col(const Epi& e, int u) { return (u >> 1) * GROUP_COLS + 16 * e.j + 8 * (u & 1); }
__device__ __forceinline__ uint4 ldg128(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// fp32 values of 8 columns from the raw hi / lo vectors of an operand tile row (F16: the tile holds fp16 pairs)
template <bool F16>
__device__ __forceinline__ void unpack_hilo8(const uint4& h, const uint4& l, float* v) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float h0, h1, l0, l1;
    unpack2<F16>(hw[i], h0, h1);
    unpack2<F16>(lw[i], l0, l1);
    v[2 * i] = h0 + l0;
    v[2 * i + 1] = h1 + l1;
  }
}
// fp32 per-tile tensors that only travel between epilogues (sigma', zhat, feat_bar) are stored [col / 4][row][4]: a thread's
// 4 consecutive columns are ONE 16-byte access and a warp's access is 512 contiguous bytes (the first layout, [col][row],
// cost one 4-byte access per column: 4x the load/store instructions of these LSU-bound epilogues)
__device__ __forceinline__ float4* f4_at(float* base, int col, int row) { return reinterpret_cast<float4*>(base) + (col >> 2) * TILE_M + row; }
__device__ __forceinline__ const float4* f4_at(const float* base, int col, int row) {
  return reinterpret_cast<const float4*>(base) + (col >> 2) * TILE_M + row;
}
__device__ __forceinline__ void f4_unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

// sigma'(z) = sigmoid(100 z) in [0, 1] travels between epilogues (sdf_render forward -> its normal pass -> both sweeps of
// sdf_bwd) as 16-bit FIXED POINT, n = rn(65535 sigma'): 0 and 1 (the linear branch of Softplus) are exact, the absolute
// error is <= 7.6e-6 everywhere.  Measured before adopting it (profiles/r02_whatif_parity.log, fp32 values rounded to
// this grid): no change of any parameter gradient beyond run-to-run noise, normals 6.5e-6 -> 8.2e-6 of their range.  As
// fp32 it was 8 KB of every 25 KB / point the forward writes and was read three times; the sweeps that read it are bound
// by exactly that traffic (profiles/r02_whatif_timing.log: sdf_bwd 1.22 ms -> 0.80 ms with its loads switched off).
// Layout of one layer's block (64 KB per tile): the operand tiles' own chunk-major order, [column / 8][row][8 columns]
// -- the 8 columns of a unit are ONE 16-byte access per thread and 512 contiguous bytes per warp, and the byte offset of
// (thread, unit) is the same expression for sigma', the hi / lo planes of a saved operand tile and the two planes of the
// zhat scratch (unit_off below): one offset register serves every stream of the backward epilogues.
constexpr int S1_LAYER_BYTES = 256 * TILE_M * 2;
// byte offset of this thread's 16 bytes of unit u (8 columns) in a chunk-major plane
__device__ __forceinline__ uint32_t unit_off(const Epi& e, int u) {
  return static_cast<uint32_t>(2 * e.j * A_CHUNK_BYTES + e.row * 16) + static_cast<uint32_t>((u >> 1) * 8 + (u & 1)) * A_CHUNK_BYTES;
}
__device__ __forceinline__ uint4* s1_at(uint8_t* layer, const Epi& e, int u) {
  return reinterpret_cast<uint4*>(layer + unit_off(e, u));
}
__device__ __forceinline__ const uint4* s1_at(const uint8_t* layer, const Epi& e, int u) {
  return reinterpret_cast<const uint4*>(layer + unit_off(e, u));
}
// GLOBAL layout of a saved MAIN operand tile (the tiles wgrad reads; bf16 pairs, hi plane then lo plane, 64 KB each):
//     plane + (row / 64) * 32 KB + (col / 8) * 1 KB + (row % 64) * 16 + (col % 8) * 2
// i.e. the shared-memory chunk-major order within each HALF of the points.  A half tile of any column range is then one
// contiguous piece per plane, so wgrad streams half tiles through a two-stage ring with single bulk copies and multiplies
// one half while the next one lands (as [chunk][128 rows] planes a half tile was 1 KB pieces: the copy engine's per-copy
// cost made that pipeline slower than none, round 1).  The records are written from registers, so their layout is free;
// aux tiles (12 KB planes, some of them bulk-stored out of shared memory) keep the shared-memory order.
constexpr int G_HALF_ROWS = TILE_M / 2;
constexpr int G_CHUNK_BYTES = G_HALF_ROWS * 16;                      // 1024
constexpr int G_HALF_PLANE_BYTES = (A_MAIN_COLS / 8) * G_CHUNK_BYTES;  // 32768
__device__ __forceinline__ uint32_t gtile_off(int chunk, int row) {
  return static_cast<uint32_t>((row >> 6) * G_HALF_PLANE_BYTES + chunk * G_CHUNK_BYTES + (row & 63) * 16);
}
// ... of this thread's unit u (see unit_off)
__device__ __forceinline__ uint32_t gunit_off(const Epi& e, int u) {
  return gtile_off(2 * e.j + (u >> 1) * 8 + (u & 1), e.row);
}

// zhat scratch of one layer (sdf_bwd): two such planes of float4, columns 0-3 / 4-7 of every chunk
constexpr int ZH_PLANE_BYTES = (256 / 8) * TILE_M * 16;
__device__ __forceinline__ float4* zh_at(uint8_t* layer, int half, const Epi& e, int u) {
  return reinterpret_cast<float4*>(layer + half * ZH_PLANE_BYTES + unit_off(e, u));
}
__device__ __forceinline__ uint32_t s1_pack2(float a, float b) {
  // 2^23 + x rounds x to the nearest integer (ulp = 1): the low 16 bits of the sum's mantissa are n
  return __byte_perm(__float_as_uint(fmaf(a, 65535.0f, 8388608.0f)), __float_as_uint(fmaf(b, 65535.0f, 8388608.0f)), 0x5410);
}
__device__ __forceinline__ uint4 s1_pack8(const float* d) {
  return make_uint4(s1_pack2(d[0], d[1]), s1_pack2(d[2], d[3]), s1_pack2(d[4], d[5]), s1_pack2(d[6], d[7]));
}
__device__ __forceinline__ void s1_unpack2(uint32_t w, float& a, float& b) {
  // bits 0x4B000000 | n are the float 2^23 + n; fma((2^23 + n), s, -2^23 s) = rn(n s) with a single rounding
  constexpr float s = 1.0f / 65535.0f, c = -8388608.0f * s;
  a = fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), s, c);
  b = fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)), s, c);
}
__device__ __forceinline__ void s1_unpack8(const uint4& w, float* o) {
  s1_unpack2(w.x, o[0], o[1]); s1_unpack2(w.y, o[2], o[3]); s1_unpack2(w.z, o[4], o[5]); s1_unpack2(w.w, o[6], o[7]);
}

// after this thread finished writing its slice of group g of the A tile (and reading that part of the accumulator).
// Every lane fences its own writes, the warp converges, ONE lane arrives: 16 arrivals per barrier phase instead of
// 512 serialized shared-memory atomics on one word.  All publish / store helpers must be called warp-uniformly.
__device__ __forceinline__ bool epi_elect() {
  __syncwarp();
  return (threadIdx.x & 31) == 0;
}
template <int STAGES>
__device__ __forceinline__ void epi_publish_group(EngineSmem<STAGES>& sm, int g) {
  fence_proxy_async();
  tc_fence_before();
  if (epi_elect()) mbar_arrive(&sm.a_ready[g]);
}
// a stage that wrote no main columns still has to arrive on every group (the MMA issuer waits on all of them)
template <int STAGES>
__device__ __forceinline__ void epi_publish_all(EngineSmem<STAGES>& sm) {
  fence_proxy_async();
  tc_fence_before();
  if (epi_elect()) {
#pragma unroll
    for (int g = 0; g < N_GROUPS; ++g) mbar_arrive(&sm.a_ready[g]);
  }
}
template <int STAGES>
__device__ __forceinline__ void epi_publish_aux(EngineSmem<STAGES>& sm) {  // slice-0 warps (j == 0) only
  fence_proxy_async();
  tc_fence_before();
  if (epi_elect()) mbar_arrive(&sm.aux_ready);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// The A planes may be overwritten only after the previous bulk store out of them finished READING shared memory.
// Only the lead thread can know (bulk_group completion is per thread); it tells the others through an mbarrier,
// so nobody sits in a CTA-wide barrier.  Must alternate with epi_store_* (one arrival per call).
template <int STAGES>
__device__ __forceinline__ void epi_planes_free(EngineSmem<STAGES>& sm, Epi& e) {
  if (e.lead) {
    bulk_wait_read0();
    mbar_arrive(&sm.rd_done);
  }
  mbar_wait(&sm.rd_done, e.rd_phase);
  e.rd_phase ^= 1;
}
// every epilogue thread calls this after writing + fencing its part; the lead thread waits for all 16 warps and then
// stores `plane_bytes` of both planes (starting at a_hi / a_lo) to dst / dst + plane_bytes.  Nobody else waits.
template <int STAGES>
__device__ __forceinline__ void epi_wrote(EngineSmem<STAGES>& sm) {
  if (epi_elect()) mbar_arrive(&sm.wr_done);
}
template <int STAGES>
__device__ __forceinline__ void epi_store_main(EngineSmem<STAGES>& sm, Epi& e, const uint8_t* a_hi, const uint8_t* a_lo,
                                               uint8_t* dst, int plane_bytes) {
  epi_wrote(sm);
  if (e.lead) {
    mbar_wait(&sm.wr_done, e.wr_phase);
    bulk_s2g(dst, a_hi, plane_bytes);
    bulk_s2g(dst + plane_bytes, a_lo, plane_bytes);
    bulk_commit();
  }
  e.wr_phase ^= 1;
}

// 16-byte store through the shared window (STS.128; a generic ST.E.128 pays the address translation)
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// write 16 / 8 consecutive columns [c0, ...) of row `row` (c0 % 8 == 0) into the A tile
template <bool F16>
__device__ __forceinline__ void store_a16(uint8_t* a_hi, uint8_t* a_lo, int row, int c0, const float* v) {
  const uint32_t off = (c0 >> 3) * A_CHUNK_BYTES + row * 16;
  const uint32_t s_hi = smem_u32(a_hi) + off, s_lo = smem_u32(a_lo) + off;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 hi, lo;
    split8<F16>(v + 8 * j, hi, lo);
    sts128(s_hi + j * A_CHUNK_BYTES, hi);
    sts128(s_lo + j * A_CHUNK_BYTES, lo);
  }
}
// The same columns to the A tile AND to a saved operand tile in global memory (same chunk-major layout, hi plane then
// lo plane, TILE_MAIN = 2 x 64 KB).  The save records used to leave the SM as 128 KB bulk (TMA) stores out of the A
// planes; the next epilogue then had to wait until the copy engine had finished READING the planes, i.e. every layer's
// store had to fit into the ~3 us of the next layer's MMAs -- 40 GB/s per SM, the whole HBM write bandwidth in bursts.
// Measured with the stores disabled: 20-27 % of every training kernel.  Written from registers (each warp stores 512
// contiguous bytes per chunk, streaming hint) nothing waits for them.  gtile == nullptr: shared memory only.
template <bool F16>
__device__ __forceinline__ void store_a16_save(uint8_t* a_hi, uint8_t* a_lo, uint8_t* gtile, int row, int c0, const float* v) {
  const uint32_t off = (c0 >> 3) * A_CHUNK_BYTES + row * 16;
  const uint32_t s_hi = smem_u32(a_hi) + off, s_lo = smem_u32(a_lo) + off;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 hi, lo;
    split8<F16>(v + 8 * j, hi, lo);
    sts128(s_hi + j * A_CHUNK_BYTES, hi);
    sts128(s_lo + j * A_CHUNK_BYTES, lo);
    if (gtile) {
      const uint32_t goff = gtile_off((c0 >> 3) + j, row);
      __stcs(reinterpret_cast<uint4*>(gtile + goff), hi);
      __stcs(reinterpret_cast<uint4*>(gtile + (A_MAIN_COLS / 8) * A_CHUNK_BYTES + goff), lo);
    }
  }
}
// The three steps of store_a16_save as separate calls, so that a kernel can put the group's publish (fence.proxy.async =
// MEMBAR + FENCE, then the mbarrier arrive the MMA warp is waiting for) BETWEEN the shared-memory stores and the global
// ones: the fence then does not have to wait for the global stores' round trip.
template <bool F16>
__device__ __forceinline__ void split16(const float* v, uint4 hi[2], uint4 lo[2]) {
  split8<F16>(v, hi[0], lo[0]);
  split8<F16>(v + 8, hi[1], lo[1]);
}
__device__ __forceinline__ void sts16(uint8_t* a_hi, uint8_t* a_lo, int row, int c0, const uint4 hi[2], const uint4 lo[2]) {
  const uint32_t off = (c0 >> 3) * A_CHUNK_BYTES + row * 16;
  const uint32_t s_hi = smem_u32(a_hi) + off, s_lo = smem_u32(a_lo) + off;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    sts128(s_hi + j * A_CHUNK_BYTES, hi[j]);
    sts128(s_lo + j * A_CHUNK_BYTES, lo[j]);
  }
}
// The saved operand tiles (read by the backward kernels and by wgrad) hold bf16 pairs whatever the shared-memory tile
// holds: tcgen05.mma kind::f16 faults ("illegal instruction", measured on B200) when A and B carry different 16-bit
// formats, and the backward tensors need bf16's exponent range.  16 / 8 fp32 values -> bf16 hi / lo -> global tile.
// lo_off: byte offset of the lo plane from the hi plane.  MAIN tile (lo_off = PLANE_MAIN): the half-tile-contiguous global
// layout (gtile_off); aux tile (lo_off = PLANE_AUX): the shared-memory order.
__device__ __forceinline__ void stg_bf16_pairs8(uint8_t* gtile, uint32_t lo_off, int row, int chunk, const float* v) {
  uint4 hi, lo;
  split8<false>(v, hi, lo);
  const uint32_t off = lo_off == (A_MAIN_COLS / 8) * A_CHUNK_BYTES ? gtile_off(chunk, row)
                                                                   : static_cast<uint32_t>(chunk * A_CHUNK_BYTES + row * 16);
  __stcs(reinterpret_cast<uint4*>(gtile + off), hi);
  __stcs(reinterpret_cast<uint4*>(gtile + lo_off + off), lo);
}
__device__ __forceinline__ void stg16(uint8_t* gtile, int row, int c0, const uint4 hi[2], const uint4 lo[2]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint32_t goff = gtile_off((c0 >> 3) + j, row);
    __stcs(reinterpret_cast<uint4*>(gtile + goff), hi[j]);
    __stcs(reinterpret_cast<uint4*>(gtile + (A_MAIN_COLS / 8) * A_CHUNK_BYTES + goff), lo[j]);
  }
}
template <bool F16>
__device__ __forceinline__ void store_a8_save(uint8_t* a_hi, uint8_t* a_lo, uint8_t* gtile, int row, int c0, const float* v) {
  uint4 hi, lo;
  split8<F16>(v, hi, lo);
  const uint32_t off = (c0 >> 3) * A_CHUNK_BYTES + row * 16;
  sts128(smem_u32(a_hi) + off, hi);
  sts128(smem_u32(a_lo) + off, lo);
  if (gtile) {
    const uint32_t goff = gtile_off(c0 >> 3, row);
    __stcs(reinterpret_cast<uint4*>(gtile + goff), hi);
    __stcs(reinterpret_cast<uint4*>(gtile + (A_MAIN_COLS / 8) * A_CHUNK_BYTES + goff), lo);
  }
}
template <bool F16>
__device__ __forceinline__ void store_a8(uint8_t* a_hi, uint8_t* a_lo, int row, int c0, const float* v) {
  uint4 hi, lo;
  split8<F16>(v, hi, lo);
  const uint32_t off = (c0 >> 3) * A_CHUNK_BYTES + row * 16;
  sts128(smem_u32(a_hi) + off, hi);
  sts128(smem_u32(a_lo) + off, lo);
}

// ---------------------------------------------------------------- activations
constexpr float SP_BETA = 100.0f;
constexpr float SP_THRESH = 20.0f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// bare MUFU.EX2 / MUFU.LG2: exp2f() wraps the MUFU in a denormal-range rescue (FSETP + 2 predicated FMUL per
// element); here an argument below -126 flushes to 0, i.e. softplus(z) = 0 for 100 z < -87 (true value < 1e-40)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// nn.Softplus(beta=100, threshold=20):  z if 100 z > 20 else log1p(exp(100 z)) / 100
__device__ __forceinline__ float softplus100(float z) {
  const float t = z * (SP_BETA * LOG2E);  // log2(e^(100 z))
  const float e = ex2_approx(fminf(t, SP_THRESH * LOG2E));
  const float sp = lg2_approx(1.0f + e) * (LN2 / SP_BETA);
  return t > SP_THRESH * LOG2E ? z : sp;
}
// softplus and its derivative sigmoid(100 z) (1 above the threshold, as autograd computes it)
__device__ __forceinline__ void softplus100_d1(float z, float& h, float& d1) {
  const float t = z * (SP_BETA * LOG2E);
  const float e = ex2_approx(fminf(t, SP_THRESH * LOG2E));
  const float ope = 1.0f + e;
  const float sp = lg2_approx(ope) * (LN2 / SP_BETA);
  const float sg = __fdividef(e, ope);
  const bool lin = t > SP_THRESH * LOG2E;
  h = lin ? z : sp;
  d1 = lin ? 1.0f : sg;
}

}  // namespace neat
