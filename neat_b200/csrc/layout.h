// Plain (host + device) layout constants and program descriptors of the tile-MLP engine (engine.cuh).
#pragma once
#include <stdint.h>

namespace neat {

constexpr int TILE_M = 128;
constexpr int A_MAIN_COLS = 256;
constexpr int A_AUX_COLS = 48;
constexpr int A_CHUNK_BYTES = TILE_M * 16;                                     // 2048
constexpr int A_PLANE_BYTES = (A_MAIN_COLS + A_AUX_COLS) / 8 * A_CHUNK_BYTES;  // 77824
constexpr int W_STAGE_BYTES = 256 * 64;                                        // 16384 (npad = 256)
constexpr int MAX_STEPS = 28;
constexpr int EPI_WARPS = 16;                        // 4 row quadrants (TMEM lane groups) x 4 column groups
constexpr int N_GROUPS = EPI_WARPS / 4;              // column groups
constexpr int GROUP_COLS = A_MAIN_COLS / N_GROUPS;   // 64 columns per group
constexpr int EPI_THREADS = EPI_WARPS * 32;          // 512
// + one more warpgroup: weight producer warp, MMA issuer warp and two idle warps.  Registers are handed out per 4 warps,
// so 18 warps cost as much as 20 and cap every thread at 96; with a full fifth warpgroup that gives most of its registers
// back (setmaxnreg, engine.cuh: role_registers) the 16 epilogue warps run with 112.
constexpr int NUM_THREADS = EPI_THREADS + 128;
constexpr int EPI_REGS = 112, AUX_REGS = 32;         // (96 - 32) * 128 freed = (112 - 96) * 512 taken
constexpr uint32_t TMEM_COLS = 512;                  // two 256-column fp32 accumulators

struct PLayer {        // a packed weight matrix: (nk_main + nk_aux) slabs of npad*64 bytes, then padded fp32 bias
  uint32_t off;        // byte offset of slab 0 in the packed buffer
  uint32_t bias_off;   // byte offset of the zero-padded fp32 bias [npad]
  uint16_t npad;       // output width (multiple of 32, <= 256)
  uint8_t nk_main;     // k-steps (16 columns each) over A main columns [0, 16*nk_main)
  uint8_t nk_aux;      // k-steps over A aux columns
};

struct Step {          // one GEMM of a kernel's program
  PLayer w;
  uint16_t d_col;      // TMEM column of the accumulator
  uint8_t wait_a;      // 1: wait (group by group) until the epilogue published the main columns of the A tile
  uint8_t wait_aux;    // 1: wait until the aux columns were published
  uint8_t commit_d;    // 1: signal the epilogue when this GEMM (and all before it) completed
  // Accumulator bias compensation.  tcgen05.mma adds every K=16 block product into the fp32 accumulator with round-toward-
  // zero (measured: making the operands 32x more precise left the SDF error at 1.6e-5; a numpy model of the MLP with RZ
  // accumulation reproduces 1.69e-5, with round-to-nearest 1.0e-6 -- see DESIGN.md section 2.1).  n truncations shrink the
  // sum by ~0.35 n 2^-24 relative (0.5 ulp each, ulp / |acc| = 0.72 x 2^-23 on average, partial sums smaller than the
  // final one), so the forward epilogues read the accumulator as acc * comp, comp = 1 + 0.35 n 2^-24, folded into the
  // bias FMA: SDF error 1.7e-5 -> 2.8e-6 in the model, for free.
  float comp;
  // what the EPILOGUE of this step reads from the tile's read-only record (Program::pf_base + tile * pf_stride + off):
  // the weight-producer warp asks the L2 for it one step ahead, spread over the previous step's k-steps
  uint32_t pf_off[2];
  uint32_t pf_bytes[2];
};

struct Program {
  int n;
  int fast;  // 0: "x3" (hi*hi + hi*lo + lo*hi, the parity mode); 1: hi*hi only (~1e-2 .. 1e-3 accuracy)
  int a_f16, b_f16;  // operand formats of the tcgen05 instruction descriptor: 1 = fp16 pairs, 0 = bf16 pairs (umma.cuh)
  const uint8_t* pf_base;   // per-tile records the epilogues read (nullptr: no prefetching)
  uint64_t pf_stride;
  Step s[MAX_STEPS];
};

}  // namespace neat
