// Host side of the junction block of VolSDFNetwork.forward (code/model/networks/neat_wfr_rend_a.py:457-496) and of the
// junction terms of VolSDFLoss (code/model/networks/loss_wfr.py:95-108): the two assignment problems the reference
// solves with scipy.optimize.linear_sum_assignment after .cpu().  The clustering itself runs on the GPU (dbscan.cuh);
// what is left is a 200 x 60 and a n_local x 1024 assignment plus a handful of projections, which take ~0.5 ms as
// numpy/scipy calls and ~30 us here -- and the GPU is idle while the host does them (one device->host hand-over per step).
//
// neat_linear_sum_assignment: rectangular min-cost assignment by shortest augmenting paths with dual updates
// (Jonker-Volgenant as restated for rectangular matrices by D. F. Crouse, "On implementing 2D rectangular assignment
// algorithms", IEEE TAES 2016 -- the algorithm scipy documents for linear_sum_assignment), float64 like scipy.
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "../../include/neat_b200.h"

namespace {

// cost: nr x nc row-major (ld = nc), nr <= nc.  col4row[i] = column assigned to row i.  Returns false if infeasible.
bool lsap_rows_le_cols(int nr, int nc, const double* cost, std::vector<int>& col4row) {
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(nr, 0.0), v(nc, 0.0), shortest(nc);
  std::vector<int> row4col(nc, -1), path(nc), remaining(nc);
  std::vector<char> SR(nr), SC(nc);
  col4row.assign(nr, -1);
  for (int cur = 0; cur < nr; ++cur) {
    std::fill(shortest.begin(), shortest.end(), INF);
    std::fill(SR.begin(), SR.end(), 0);
    std::fill(SC.begin(), SC.end(), 0);
    std::fill(path.begin(), path.end(), -1);
    // columns are scanned in reverse storage order so that ties resolve towards the lowest column index
    int n_rem = nc;
    for (int j = 0; j < nc; ++j) remaining[j] = nc - 1 - j;
    double min_val = 0.0;
    int i = cur, sink = -1;
    while (sink < 0) {
      SR[i] = 1;
      double lowest = INF;
      int index = -1;
      const double* row = cost + static_cast<size_t>(i) * nc;
      for (int it = 0; it < n_rem; ++it) {
        const int j = remaining[it];
        const double r = min_val + row[j] - u[i] - v[j];
        if (r < shortest[j]) {
          path[j] = i;
          shortest[j] = r;
        }
        if (shortest[j] < lowest || (shortest[j] == lowest && row4col[j] < 0)) {
          lowest = shortest[j];
          index = it;
        }
      }
      min_val = lowest;
      if (!(min_val < INF)) return false;
      const int j = remaining[index];
      if (row4col[j] < 0) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--n_rem];
    }
    u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (SR[r] && r != cur) u[r] += min_val - shortest[col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= min_val - shortest[j];
    int j = sink;
    for (;;) {
      const int r = path[j];
      row4col[j] = r;
      std::swap(col4row[r], j);
      if (r == cur) break;
    }
  }
  return true;
}

struct Cam {   // world -> camera rows of inverse(pose), and the 3x3 intrinsics
  float R[3][3], T[3], K[3][3];
};

// closed-form inverse of a rigid-or-not 4x4 (cofactors, double accumulation), rows 0..2 only
bool inverse_rows3(const float* m, float R[3][3], float T[3]) {
  double a[4][4], inv[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a[i][j] = m[4 * i + j];
  // Gauss-Jordan with partial pivoting
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) inv[i][j] = i == j;
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (a[piv][c] == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 4; ++j) {
        std::swap(a[piv][j], a[c][j]);
        std::swap(inv[piv][j], inv[c][j]);
      }
    const double d = 1.0 / a[c][c];
    for (int j = 0; j < 4; ++j) {
      a[c][j] *= d;
      inv[c][j] *= d;
    }
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      const double f = a[r][c];
      if (f == 0.0) continue;
      for (int j = 0; j < 4; ++j) {
        a[r][j] -= f * a[c][j];
        inv[r][j] -= f * inv[c][j];
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = static_cast<float>(inv[i][j]);
    T[i] = static_cast<float>(inv[i][3]);
  }
  return true;
}

// VolSDFNetwork.project2D (neat_wfr_rend_a.py:317-331) for one point, fp32, same epsilon / sign guard
inline void project2d(const Cam& c, bool calibrated, const float* X, float* uv) {
  float p[3];
  for (int i = 0; i < 3; ++i) p[i] = c.R[i][0] * X[0] + c.R[i][1] * X[1] + c.R[i][2] * X[2] + c.T[i];
  float x[3];
  if (calibrated) {
    x[0] = p[0]; x[1] = p[1]; x[2] = p[2];
  } else {
    for (int i = 0; i < 3; ++i) x[i] = c.K[i][0] * p[0] + c.K[i][1] * p[1] + c.K[i][2] * p[2];
  }
  const float den = x[2];
  const float sign = den >= 0.f ? 1.f : -1.f;
  const float eps = std::fabs(den) < 1e-8f ? 1e-8f : 0.f;
  const float d = den + eps * sign;
  uv[0] = x[0] / d;
  uv[1] = x[1] / d;
}

}  // namespace

extern "C" {

int neat_linear_sum_assignment(const double* cost, int n_rows, int n_cols, int* row_ind, int* col_ind) {
  if (n_rows < 0 || n_cols < 0 || (n_rows && n_cols && (!cost || !row_ind || !col_ind))) return NEAT_EINVAL;
  if (n_rows == 0 || n_cols == 0) return 0;
  for (size_t i = 0, n = static_cast<size_t>(n_rows) * n_cols; i < n; ++i)
    if (std::isnan(cost[i]) || cost[i] == -std::numeric_limits<double>::infinity()) return NEAT_EINVAL;
  std::vector<int> a;
  if (n_rows <= n_cols) {
    if (!lsap_rows_le_cols(n_rows, n_cols, cost, a)) return NEAT_EINVAL;
    for (int i = 0; i < n_rows; ++i) {
      row_ind[i] = i;
      col_ind[i] = a[i];
    }
    return n_rows;
  }
  std::vector<double> t(static_cast<size_t>(n_rows) * n_cols);
  for (int i = 0; i < n_rows; ++i)
    for (int j = 0; j < n_cols; ++j) t[static_cast<size_t>(j) * n_rows + i] = cost[static_cast<size_t>(i) * n_cols + j];
  if (!lsap_rows_le_cols(n_cols, n_rows, t.data(), a)) return NEAT_EINVAL;
  // pairs (row a[j], column j), returned sorted by row like scipy
  std::vector<std::pair<int, int>> pr(n_cols);
  for (int j = 0; j < n_cols; ++j) pr[j] = {a[j], j};
  std::sort(pr.begin(), pr.end());
  for (int j = 0; j < n_cols; ++j) {
    row_ind[j] = pr[j].first;
    col_ind[j] = pr[j].second;
  }
  return n_cols;
}

int neat_junction_match(const float* centroids, int n_centroids, const float* gt_vertices, int n_gt,
                        const float* pose, const float* intrinsics, const float* global_junctions, int n_global,
                        int use_median, float* local_out, int* n_local_out, int* global_rows, int* global_cols,
                        int* n_close_out, float* median_out) {
  if (!pose || !intrinsics || !local_out || !n_local_out || n_centroids < 0 || n_gt < 0 || n_global < 0)
    return NEAT_EINVAL;
  Cam cam;
  if (!inverse_rows3(pose, cam.R, cam.T)) return NEAT_EINVAL;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cam.K[i][j] = intrinsics[4 * i + j];
  *n_local_out = 0;
  if (n_close_out) *n_close_out = 0;
  if (median_out) *median_out = 10.f;
  const int C = n_centroids, J = n_gt;
  if (C == 0 || J == 0) return NEAT_OK;
  // 2D projections of the cluster centroids, pixel and calibrated (neat_wfr_rend_a.py:466-470)
  std::vector<float> j2d(2 * C), j2c(2 * C);
  for (int c = 0; c < C; ++c) {
    project2d(cam, false, centroids + 3 * c, &j2d[2 * c]);
    project2d(cam, true, centroids + 3 * c, &j2c[2 * c]);
  }
  // jcost[gt, cluster] = || j2d - gt ||_2 (fp32 as torch computes it), then the assignment in float64 (:471-473)
  std::vector<double> cost(static_cast<size_t>(J) * C);
  for (int g = 0; g < J; ++g)
    for (int c = 0; c < C; ++c) {
      const float dx = j2d[2 * c] - gt_vertices[2 * g], dy = j2d[2 * c + 1] - gt_vertices[2 * g + 1];
      cost[static_cast<size_t>(g) * C + c] = std::sqrt(dx * dx + dy * dy);
    }
  const int n_pairs = std::min(J, C);
  std::vector<int> a0(n_pairs), a1(n_pairs);
  if (neat_linear_sum_assignment(cost.data(), J, C, a0.data(), a1.data()) != n_pairs) return NEAT_EINVAL;
  float thresh = 10.f;
  if (use_median) {  // torch.median of the matched costs: the LOWER of the two middle values; NaN (no pairs) -> 10
    std::vector<float> sel(n_pairs);
    for (int k = 0; k < n_pairs; ++k) sel[k] = static_cast<float>(cost[static_cast<size_t>(a0[k]) * C + a1[k]]);
    std::sort(sel.begin(), sel.end());
    thresh = sel[(n_pairs - 1) / 2];
    if (median_out) *median_out = thresh;
  }
  int n = 0;
  for (int k = 0; k < n_pairs; ++k) {
    if (!(static_cast<float>(cost[static_cast<size_t>(a0[k]) * C + a1[k]]) < thresh)) continue;
    const int c = a1[k];
    float* o = local_out + 7 * n++;
    o[0] = centroids[3 * c]; o[1] = centroids[3 * c + 1]; o[2] = centroids[3 * c + 2];
    o[3] = j2d[2 * c]; o[4] = j2d[2 * c + 1];
    o[5] = j2c[2 * c]; o[6] = j2c[2 * c + 1];
  }
  *n_local_out = n;
  if (n == 0 || n_global == 0 || !global_junctions || !global_rows || !global_cols) return NEAT_OK;
  // the loss' assignment (loss_wfr.py:104-108): L1 distance in 3D + 0.1 x L1 distance of the calibrated projections
  std::vector<float> gcal(2 * static_cast<size_t>(n_global));
  for (int g = 0; g < n_global; ++g) project2d(cam, true, global_junctions + 3 * g, &gcal[2 * g]);
  std::vector<double> cost2(static_cast<size_t>(n) * n_global);
  for (int i = 0; i < n; ++i) {
    const float* o = local_out + 7 * i;
    for (int g = 0; g < n_global; ++g) {
      const float* G = global_junctions + 3 * g;
      const float d3 = std::fabs(o[0] - G[0]) + std::fabs(o[1] - G[1]) + std::fabs(o[2] - G[2]);
      const float d2 = std::fabs(o[5] - gcal[2 * g]) + std::fabs(o[6] - gcal[2 * g + 1]);
      cost2[static_cast<size_t>(i) * n_global + g] = d3 + 0.1f * d2;
    }
  }
  const int n2 = std::min(n, n_global);
  if (neat_linear_sum_assignment(cost2.data(), n, n_global, global_rows, global_cols) != n2) return NEAT_EINVAL;
  int close = 0;
  for (int k = 0; k < n2; ++k)
    if (cost2[static_cast<size_t>(global_rows[k]) * n_global + global_cols[k]] < 10.0) ++close;
  if (n_close_out) *n_close_out = close;
  return NEAT_OK;
}

}  // extern "C"
