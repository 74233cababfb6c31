// Host-side plan builder: flat parameter layout + gather tables for the packed weight slabs.
#include "plan.h"

#include <cmath>
#include <functional>
#include <stdexcept>

namespace neat {
namespace {

inline int ceil_to(int x, int m) { return (x + m - 1) / m * m; }

struct Src {
  long idx;  // float index into the flat parameter buffer, -1 = zero
  float scale;
};
using SrcFn = std::function<Src(int n, bool aux, int kk)>;  // weight feeding output n from A column kk of a segment
using BiasFn = std::function<long(int n)>;

struct Builder {
  Plan& p;
  size_t cur = 0;  // bytes
  explicit Builder(Plan& plan) : p(plan) {}

  PLayer add(int npad, int nk_main, int nk_aux, const SrcFn& f, const BiasFn& bf) {
    if (npad % 32 || npad > 256 || npad <= 0) throw std::runtime_error("npad must be a multiple of 32 in (0,256]");
    if (nk_main * 16 > A_MAIN_COLS || nk_aux * 16 > A_AUX_COLS) throw std::runtime_error("segment too wide");
    cur = (cur + 127) / 128 * 128;
    PLayer L{};
    L.off = static_cast<uint32_t>(cur);
    L.npad = static_cast<uint16_t>(npad);
    L.nk_main = static_cast<uint8_t>(nk_main);
    L.nk_aux = static_cast<uint8_t>(nk_aux);
    const size_t slab = static_cast<size_t>(npad) * 64;
    for (int ks = 0; ks < nk_main + nk_aux; ++ks) {
      const bool aux = ks >= nk_main;
      const int k0 = 16 * (aux ? ks - nk_main : ks);
      const size_t base_el = (cur + ks * slab) / 2;  // bf16 element index of the hi plane
      for (int c = 0; c < 2; ++c)
        for (int n = 0; n < npad; ++n)
          for (int e = 0; e < 8; ++e) {
            const Src s = f(n, aux, k0 + c * 8 + e);
            const size_t el = base_el + (static_cast<size_t>(c) * npad + n) * 8 + e;
            p.g.src.push_back(static_cast<int32_t>(s.idx));
            p.g.scale.push_back(s.scale);
            p.g.dst_hi.push_back(static_cast<uint32_t>(el));
            p.g.dst_lo.push_back(static_cast<uint32_t>(el + static_cast<size_t>(npad) * 16));
          }
    }
    cur += (nk_main + nk_aux) * slab;
    L.bias_off = static_cast<uint32_t>(cur);
    for (int n = 0; n < npad; ++n) {
      p.g.fsrc.push_back(static_cast<int32_t>(bf(n)));
      p.g.fdst.push_back(static_cast<uint32_t>(cur / 4 + n));
    }
    cur += static_cast<size_t>(npad) * 4;
    return L;
  }
};

const float RSQRT2 = static_cast<float>(1.0 / std::sqrt(2.0));

}  // namespace

void build_plan(const neat_net_config& cfg, Plan& P) {
  P = Plan{};
  P.cfg = cfg;
  const int L = cfg.sdf_layers, H = cfg.sdf_hidden, F = cfg.feat, S = cfg.sdf_skip;
  const int E = cfg.multires > 0 ? 3 + 6 * cfg.multires : 3;
  const int Ev = cfg.multires_view > 0 ? 3 + 6 * cfg.multires_view : 3;
  P.E = E;
  P.Ev = Ev;
  if (L < 3 || H % 32 || H > 256 || H <= E || F % 32 || F > 256 || E > A_AUX_COLS || Ev + 6 > A_AUX_COLS ||
      cfg.head_layers < 2 || cfg.head_hidden % 32 || cfg.head_hidden > 256 || (S >= 0 && (S < 2 || S >= L - 1)))
    throw std::runtime_error("unsupported network shape");

  // ---- flat parameter layout --------------------------------------------------------------
  size_t off = 0;
  auto push = [&](std::vector<LinearDims>& v, int in, int out) {
    LinearDims d{in, out, off, off + static_cast<size_t>(in) * out};
    off += static_cast<size_t>(in) * out + out;
    v.push_back(d);
  };
  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? E : H;
    int out = l == L - 1 ? 1 + F : H;
    if (l + 1 == S) out = H - E;
    push(P.sdf, in, out);
  }
  const int HH = cfg.head_hidden, HL = cfg.head_layers;
  const int rin = 3 + Ev + 3, ain = 9;
  for (int l = 0; l < HL; ++l) push(P.rend, l == 0 ? rin + F : HH, l == HL - 1 ? 3 : HH);
  for (int l = 0; l < HL; ++l) push(P.att, l == 0 ? ain + F : HH, l == HL - 1 ? 6 : HH);
  P.n_params = off;

  Builder B(P);
  auto W = [](const LinearDims& d, int n, int k) { return static_cast<long>(d.w_off + static_cast<size_t>(n) * d.in + k); };
  const BiasFn nobias = [](int) { return -1L; };

  // ---- ImplicitNetwork, forward --------------------------------------------------------------
  for (int l = 0; l < L; ++l) {
    const LinearDims d = P.sdf[l];
    if (l == L - 1) {
      P.sdf_f.push_back(B.add(32, ceil_to(d.in, 16) / 16, 0,
                              [&](int n, bool, int kk) { return Src{n == 0 && kk < d.in ? W(d, 0, kk) : -1, 1.f}; },
                              [&](int n) { return n == 0 ? static_cast<long>(d.b_off) : -1L; }));
      P.sdf_f_feat = B.add(ceil_to(F, 32), ceil_to(d.in, 16) / 16, 0,
                           [&](int n, bool, int kk) { return Src{n < F && kk < d.in ? W(d, 1 + n, kk) : -1, 1.f}; },
                           [&](int n) { return n < F ? static_cast<long>(d.b_off + 1 + n) : -1L; });
      continue;
    }
    const BiasFn bias = [&](int n) { return n < d.out ? static_cast<long>(d.b_off + n) : -1L; };
    if (l == 0) {
      P.sdf_f.push_back(B.add(ceil_to(d.out, 32), 0, ceil_to(E, 16) / 16,
                              [&](int n, bool, int kk) { return Src{n < d.out && kk < E ? W(d, n, kk) : -1, 1.f}; }, bias));
    } else if (l == S) {
      const int im = d.in - E;
      P.sdf_f.push_back(B.add(ceil_to(d.out, 32), ceil_to(im, 16) / 16, ceil_to(E, 16) / 16,
                              [&](int n, bool aux, int kk) {
                                if (n >= d.out) return Src{-1, 1.f};
                                if (!aux) return Src{kk < im ? W(d, n, kk) : -1, RSQRT2};
                                return Src{kk < E ? W(d, n, im + kk) : -1, RSQRT2};
                              },
                              bias));
    } else {
      P.sdf_f.push_back(B.add(ceil_to(d.out, 32), ceil_to(d.in, 16) / 16, 0,
                              [&](int n, bool, int kk) { return Src{n < d.out && kk < d.in ? W(d, n, kk) : -1, 1.f}; }, bias));
    }
  }
  // ---- ImplicitNetwork, transposed: out[j] = sum_n a[n] W_l[n][j] ---------------------------------
  for (int l = 0; l < L; ++l) {
    const LinearDims d = P.sdf[l];
    if (l == L - 1) {  // reduction over [feat (main) ; sdf (aux column 0)]
      P.sdf_t.push_back(B.add(ceil_to(d.in, 32), ceil_to(F, 16) / 16, 1,
                              [&](int j, bool aux, int kk) {
                                if (j >= d.in) return Src{-1, 1.f};
                                if (!aux) return Src{kk < F ? W(d, 1 + kk, j) : -1, 1.f};
                                return Src{kk == 0 ? W(d, 0, j) : -1, 1.f};
                              },
                              nobias));
    } else {
      const float sc = l == S ? RSQRT2 : 1.f;
      P.sdf_t.push_back(B.add(ceil_to(d.in, 32), ceil_to(d.out, 16) / 16, 0,
                              [&](int j, bool, int kk) { return Src{j < d.in && kk < d.out ? W(d, kk, j) : -1, sc}; },
                              nobias));
    }
  }
  // ---- heads ------------------------------------------------------------------------------------
  auto heads = [&](const std::vector<LinearDims>& net, int aux_in, std::vector<PLayer>& fw, std::vector<PLayer>& tr,
                   PLayer& t0aux) {
    const int n_l = static_cast<int>(net.size());
    for (int l = 0; l < n_l; ++l) {
      const LinearDims d = net[l];
      const BiasFn bias = [&](int n) { return n < d.out ? static_cast<long>(d.b_off + n) : -1L; };
      if (l == 0)
        fw.push_back(B.add(ceil_to(d.out, 32), ceil_to(F, 16) / 16, ceil_to(aux_in, 16) / 16,
                           [&](int n, bool aux, int kk) {
                             if (n >= d.out) return Src{-1, 1.f};
                             if (!aux) return Src{kk < F ? W(d, n, aux_in + kk) : -1, 1.f};
                             return Src{kk < aux_in ? W(d, n, kk) : -1, 1.f};
                           },
                           bias));
      else
        fw.push_back(B.add(ceil_to(d.out, 32), ceil_to(d.in, 16) / 16, 0,
                           [&](int n, bool, int kk) { return Src{n < d.out && kk < d.in ? W(d, n, kk) : -1, 1.f}; }, bias));
    }
    for (int l = 0; l < n_l; ++l) {
      const LinearDims d = net[l];
      if (l == n_l - 1) {  // reduction over the 3 / 6 outputs, which the backward kernel puts in aux columns
        tr.push_back(B.add(ceil_to(d.in, 32), 0, 1,
                           [&](int j, bool, int kk) { return Src{j < d.in && kk < d.out ? W(d, kk, j) : -1, 1.f}; }, nobias));
      } else if (l == 0) {  // feature part of the input gradient
        tr.push_back(B.add(ceil_to(F, 32), ceil_to(d.out, 16) / 16, 0,
                           [&](int j, bool, int kk) { return Src{j < F && kk < d.out ? W(d, kk, aux_in + j) : -1, 1.f}; },
                           nobias));
        t0aux = B.add(32, ceil_to(d.out, 16) / 16, 0,  // normal part: inputs aux_in-3 .. aux_in-1
                      [&](int j, bool, int kk) { return Src{j < 3 && kk < d.out ? W(d, kk, aux_in - 3 + j) : -1, 1.f}; },
                      nobias);
      } else {
        tr.push_back(B.add(ceil_to(d.in, 32), ceil_to(d.out, 16) / 16, 0,
                           [&](int j, bool, int kk) { return Src{j < d.in && kk < d.out ? W(d, kk, j) : -1, 1.f}; }, nobias));
      }
    }
  };
  heads(P.rend, rin, P.rend_f, P.rend_t, P.rend_t0_aux);
  heads(P.att, ain, P.att_f, P.att_t, P.att_t0_aux);

  // ---- fp32 copy of the sdf row of the last ImplicitNetwork layer (normal pass seed) ----------------
  B.cur = (B.cur + 127) / 128 * 128;
  P.w_last_row_off = static_cast<uint32_t>(B.cur);
  {
    const LinearDims d = P.sdf[L - 1];
    for (int k = 0; k < 256; ++k) {
      P.g.fsrc.push_back(k < d.in ? static_cast<int32_t>(W(d, 0, k)) : -1);
      P.g.fdst.push_back(static_cast<uint32_t>(B.cur / 4 + k));
    }
    B.cur += 256 * 4;
  }
  P.packed_bytes = (B.cur + 127) / 128 * 128;
}

}  // namespace neat
