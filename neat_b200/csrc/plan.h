// Host-side plan: network dimensions, the flat fp32 parameter layout shared with Python, and the
// gather tables that turn the flat parameters into tcgen05-ready weight slabs (see engine.cuh).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/neat_b200.h"
#include "layout.h"

namespace neat {

struct LinearDims {
  int in, out;
  size_t w_off, b_off;  // float offsets into the flat parameter buffer (W is [out, in] row-major)
};

struct GatherTables {
  // bf16 hi/lo pairs: packed[dst_hi] , packed[dst_lo] (bf16 element indices) <- scale * flat[src]  (src < 0: zero)
  std::vector<int32_t> src;
  std::vector<uint32_t> dst_hi, dst_lo;
  std::vector<float> scale;
  // fp32 copies: packed_f32[fdst] <- flat[fsrc] (fsrc < 0: zero)
  std::vector<int32_t> fsrc;
  std::vector<uint32_t> fdst;
};

struct Plan {
  neat_net_config cfg;
  int E;   // SDF positional-encoding width 3 + 6*multires
  int Ev;  // view-dir encoding width (rendering head)
  std::vector<LinearDims> sdf, rend, att;
  size_t n_params = 0;

  // packed layers (device layout); see build_plan() for what each one multiplies
  std::vector<PLayer> sdf_f;    // forward layer l (last layer: sdf row only, npad 32)
  PLayer sdf_f_feat;            // last layer, feature rows 1..F
  std::vector<PLayer> sdf_t;    // transposed layer l : v = a * W_l
  std::vector<PLayer> rend_f, rend_t, att_f, att_t;
  PLayer rend_t0_aux, att_t0_aux;  // transposed first head layer restricted to the normal inputs
  uint32_t w_last_row_off = 0;     // fp32 copy of W_{L-1}[0, :] (byte offset)
  size_t packed_bytes = 0;
  GatherTables g;
};

void build_plan(const neat_net_config& cfg, Plan& plan);

}  // namespace neat
