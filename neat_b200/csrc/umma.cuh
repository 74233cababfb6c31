// sm_100a primitives used by every MLP kernel: mbarrier, 1-D bulk TMA (cp.async.bulk),
// tcgen05.mma / tcgen05.ld / TMEM allocation, and the shared-memory matrix descriptor.
// Everything is inline PTX; there is no CUTLASS/CuTe dependency.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace neat {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Warp index the compiler can prove to be warp-uniform (a shuffle from lane 0): branches on it are uniform branches,
// so the role loops below run converged and their addresses / descriptors live in uniform registers.  With a plain
// `threadIdx.x >> 5` (or an `if (lane == 0)` around the loop) every tcgen05 / TMA instruction was wrapped in an
// ELECT + R2UR + BRA.U.ANY serialisation loop: ~95 SASS instructions per k-step on the MMA issuer's critical path.
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---------------------------------------------------------------- bulk async copy (TMA, 1-D)
// global -> shared, completion signalled on an mbarrier (complete_tx).  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk store (bulk_group completion)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
// ask the L2 to fetch [gmem, gmem + bytes) (16-byte aligned, size % 16 == 0); no completion tracking.
// Used (a) by wgrad for the tile ahead of the one being multiplied and (b) by the
// weight-producer warp of the backward kernels for the saved tensors the NEXT step's epilogue reads (engine.cuh).  The
// same hint issued by an epilogue thread made sdf_bwd slower (1.37 -> 1.57 ms): the issuing warp stalls, and the tile
// waits for its slowest warp.  g_l2_prefetch (neat_debug_set_l2_prefetch): 0 = all hints off.
__constant__ int g_l2_prefetch = 2;
// what-if switches for measurements only (neat_debug_set_flags; 0 in every product / test path):
//   1: wgrad skips X_hi * Y_lo      2: wgrad skips X_lo * Y_hi        4: sigma' rounded to 16-bit fixed point when saved
//   8: sdf_bwd builds zhat from the hi plane of a_l only
//  16: sdf_bwd epilogues skip their global loads   32: ... skip their global stores (p, z_bar; WRONG RESULTS, timing only)
//  64: sdf_render skips the u / a / feat save stores   128: sdf_render's normal pass skips the sigma' loads (timing only)
__constant__ unsigned g_dbg = 0;
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, uint32_t bytes) {
  if (g_l2_prefetch)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- L2 eviction-priority hints (per access)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void stg128_hint(void* ptr, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ float4 ldg128_hint(const void* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol)
               : "memory");
  return v;
}

// generic-proxy writes to smem -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; writes the TMEM base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layout:
// core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
//   lbo = byte distance between core matrices adjacent in the K direction   (K-major operand)
//   sbo = byte distance between core matrices adjacent in the M/N direction
// For an MN-major operand the two roles are exchanged by the hardware; see make_desc_mn.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// Instruction descriptor, kind::f16: {F16 | BF16} x {F16 | BF16} -> FP32, dense.  The two operand formats are separate
// fields (A: bits 7-9, B: bits 10-12; 0 = F16, 1 = BF16) and may differ.
//   a_mn / b_mn : 0 = K-major operand, 1 = MN-major operand;  a_f16 / b_f16 : 1 = the operand holds fp16, 0 = bf16
__device__ __forceinline__ uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn, uint32_t a_f16 = 0,
                                               uint32_t b_f16 = 0) {
  uint32_t d = 0;
  d |= 1u << 4;                     // D format  : F32
  d |= (a_f16 ? 0u : 1u) << 7;      // A format
  d |= (b_f16 ? 0u : 1u) << 10;     // B format
  d |= (a_mn & 1u) << 15;
  d |= (b_mn & 1u) << 16;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One k-step of the bf16x3 product as ONE instruction sequence: D (+)= Ahi*Bhi [+ Ahi*Blo + Alo*Bhi unless `fast`], then
// tcgen05.commit -> `bar`.  Issued by the elected lane; keeping the three MMAs, their predicates and the commit in one asm
// block spares the per-MMA predicate set-up and lets the caller keep everything in uniform registers.
__device__ __forceinline__ void umma_kstep_bf16x3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                  uint32_t idesc, uint32_t accumulate, uint32_t fast, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pacc, pfull;\n\t"
      "setp.ne.b32 pacc, %6, 0;\n\t"
      "setp.eq.b32 pfull, %7, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, pacc;\n\t"
      "@pfull tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %5, 1;\n\t"
      "@pfull tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, 1;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
      ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate), "r"(fast), "r"(smem_u32(bar))
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// warp w reads TMEM lanes [32*(w%4), +32): thread i gets lane 32*(w%4)+i, 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- hi/lo split operands ("x3" arithmetic)
// x ~= hi + lo with hi = rn16(x), lo = rn16(x - hi); A*B is evaluated as Ahi*Bhi + Ahi*Blo + Alo*Bhi with fp32 accumulation
// in TMEM.  Two 16-bit formats are used:
//   bf16 pairs (8 + 8 mantissa bits, fp32's exponent range): the BACKWARD tensors (z_bar, p: gradients of any magnitude);
//              ~2^-17 relative per operand
//   fp16 pairs (11 + 11 bits): the weights and every FORWARD activation (values of O(1e-4 .. 1e2): softplus / ReLU outputs,
//              positional encodings, normals, features); ~2^-22 relative per operand, i.e. fp32-class products.  Measured
//              reason (round 2): with bf16 pairs the SDF carried ~3e-5 absolute error, which the Laplace density divides
//              by beta -- at beta = 0.01 the composited end points missed the 1e-4 bound at 1024 rays, and the error of the
//              compositing weights dominated the gradient error of the rendering head.  Conversions saturate (no inf).
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}
// two fp32 -> packed bf16x2 (a in the low half) with ONE F2FP instruction; the scalar __float2bfloat16_rn compiles to
// F2F.BF16.F32, which issues on the quarter-rate XU pipe next to ex2/lg2 and was the epilogue's bottleneck
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// the two fp32 values of a packed pair
template <bool F16>
__device__ __forceinline__ void unpack2(uint32_t v, float& a, float& b) {
  if (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    a = f.x; b = f.y;
  } else {
    a = __uint_as_float(v << 16);
    b = __uint_as_float(v & 0xFFFF0000u);
  }
}
// hi/lo split of a pair: hi = rn16(x), lo = rn16(x - hi)
template <bool F16>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  float ha, hb;
  if (F16) hi = cvt_f16x2(a, b); else hi = cvt_bf16x2(a, b);
  unpack2<F16>(hi, ha, hb);
  if (F16) lo = cvt_f16x2(a - ha, b - hb); else lo = cvt_bf16x2(a - ha, b - hb);
}
// splits 8 consecutive fp32 values into two 16-byte vectors (hi, lo)
template <bool F16>
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  split2<F16>(v[0], v[1], hi.x, lo.x);
  split2<F16>(v[2], v[3], hi.y, lo.y);
  split2<F16>(v[4], v[5], hi.z, lo.z);
  split2<F16>(v[6], v[7], hi.w, lo.w);
}

}  // namespace neat
