// Host-side logic of the hot path: quadrature rules with deal.II semantics, FE_Q shape tables, locality
// ordering / row partition, column-block and singular-patch tables.  Integer / O(N) work only; all
// floating-point assembly work happens in bs_assembly.cu on the device.
#include "bs_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>

namespace bs {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }
const std::string &get_last_error() { return g_last_error; }

// ---------------------------------------------------------------------------------------------------------
// 1-D Gauss-Legendre on [0,1] (deal.II QGauss<1>): Newton on Legendre polynomials
// ---------------------------------------------------------------------------------------------------------
void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w) {
  BS_REQUIRE(n >= 1 && n <= 256, "gauss order out of range");
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    long double z = cosl(pi * (i + 0.75L) / (n + 0.5L));
    long double pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1.0L, p2 = 0.0L;
      for (int j = 1; j <= n; ++j) {
        long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0L);
      long double z1 = z;
      z = z1 - p1 / pp;
      if (fabsl(z - z1) < 1e-19L) break;
    }
    long double wi = 2.0L / ((1.0L - z * z) * pp * pp);
    x[i] = (double)((1.0L - z) / 2.0L);
    x[n - 1 - i] = (double)((1.0L + z) / 2.0L);
    w[i] = w[n - 1 - i] = (double)(wi / 2.0L);
  }
}

Rule2D tensor_rule2(const std::vector<double> &x1, const std::vector<double> &w1, const std::vector<double> &x2,
                    const std::vector<double> &w2) {
  Rule2D r;
  for (size_t j = 0; j < x2.size(); ++j)
    for (size_t i = 0; i < x1.size(); ++i) {  // first coordinate fastest
      r.xi.push_back(x1[i]);
      r.xi.push_back(x2[j]);
      r.w.push_back(w1[i] * w2[j]);
    }
  return r;
}
Rule2D tensor_rule(const std::vector<double> &x1, const std::vector<double> &w1) { return tensor_rule2(x1, w1, x1, w1); }

// QGaussOneOverR<2>(n, vertex_index, factor_out_singular_weight=true): Lachat-Watson
static Rule2D lw_vertex(int n, int v) {
  std::vector<double> x, w;
  gauss_legendre_01(n, x, w);
  Rule2D g = tensor_rule(x, w);
  const double pi4 = M_PI / 4.0;
  const int m = g.size();
  Rule2D r;
  r.xi.resize(4 * m);
  r.w.resize(2 * m);
  for (int q = 0; q < m; ++q) {
    double u = g.xi[2 * q], t = g.xi[2 * q + 1];
    double px = u, py = u * std::tan(pi4 * t);
    double ww = g.w[q] * pi4 / std::cos(pi4 * t);
    ww *= std::sqrt(px * px + py * py);
    r.xi[2 * q] = px;
    r.xi[2 * q + 1] = py;
    r.w[q] = ww;
    r.xi[2 * (m + q)] = py;
    r.xi[2 * (m + q) + 1] = px;
    r.w[m + q] = ww;
  }
  double theta = 0;
  if (v == 1) theta = M_PI / 2;
  if (v == 2) theta = -M_PI / 2;
  if (v == 3) theta = M_PI;
  if (v != 0) {
    double c = std::cos(theta), s = std::sin(theta);
    for (int q = 0; q < 2 * m; ++q) {
      double X = r.xi[2 * q] - .5, Y = r.xi[2 * q + 1] - .5;
      r.xi[2 * q] = c * X - s * Y + .5;
      r.xi[2 * q + 1] = s * X + c * Y + .5;
    }
  }
  return r;
}

// QGaussOneOverR<2>(n, Point<2> singularity, true)
static Rule2D lw_point(int n, double sx, double sy) {
  Rule2D quads[4] = {lw_vertex(n, 3), lw_vertex(n, 2), lw_vertex(n, 1), lw_vertex(n, 0)};
  const double ox[4] = {0, sx, 0, sx}, oy[4] = {0, 0, sy, sy};
  const double vx[4] = {0, 1, 0, 1}, vy[4] = {0, 0, 1, 1};
  Rule2D r;
  for (int b = 0; b < 4; ++b) {
    double dx = std::fabs(sx - vx[b]), dy = std::fabs(sy - vy[b]);
    double area = dx * dy;
    if (area > 1e-8)
      for (int q = 0; q < quads[b].size(); ++q) {
        r.xi.push_back(ox[b] + dx * quads[b].xi[2 * q]);
        r.xi.push_back(oy[b] + dy * quads[b].xi[2 * q + 1]);
        r.w.push_back(quads[b].w[q] * area);
      }
  }
  return r;
}

// QTelles<1>(n, s)
static void telles_1d(int n, double s, std::vector<double> &xo, std::vector<double> &wo) {
  std::vector<double> x, w;
  gauss_legendre_01(n, x, w);
  xo.clear();
  wo.clear();
  const double eb = 2 * s - 1, es = eb * eb - 1;
  const double gb = std::cbrt(eb * es + std::fabs(es)) + std::cbrt(eb * es - std::fabs(es)) + eb;
  for (int q = 0; q < n; ++q) {
    if (!(std::fabs(x[q] - s) > 1e-10)) continue;
    double g = 2 * x[q] - 1;
    double d = g - gb;
    double eta = (d * d * d + gb * (gb * gb + 3)) / (1 + 3 * gb * gb);
    double J = 3 * d * d / (1 + 3 * gb * gb);
    xo.push_back((eta + 1) / 2);
    wo.push_back(J * w[q]);
  }
}

// QSplit<2>(QDuffy(n, 1.), s)
static Rule2D qsplit_duffy(int n, double sx, double sy) {
  std::vector<double> x, w;
  gauss_legendre_01(n, x, w);
  Rule2D g = tensor_rule(x, w);
  const double vx[4] = {0, 1, 0, 1}, vy[4] = {0, 0, 1, 1};
  const int f[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
  Rule2D r;
  for (int k = 0; k < 4; ++k) {
    double b00 = vx[f[k][0]] - sx, b10 = vy[f[k][0]] - sy;  // first column
    double b01 = vx[f[k][1]] - sx, b11 = vy[f[k][1]] - sy;  // second column
    double J = std::fabs(b00 * b11 - b01 * b10);
    if (J < 1e-12) continue;
    for (int q = 0; q < g.size(); ++q) {
      double xh = g.xi[2 * q], yh = g.xi[2 * q + 1];
      double X = xh * (1 - yh), Y = xh * yh;  // beta = 1
      r.xi.push_back(sx + b00 * X + b01 * Y);
      r.xi.push_back(sy + b10 * X + b11 * Y);
      r.w.push_back(g.w[q] * xh * J);
    }
  }
  return r;
}

static Rule2D qiterated(int n, int k) {
  std::vector<double> x, w, xs, ws;
  gauss_legendre_01(n, x, w);
  for (int i = 0; i < k; ++i)
    for (int q = 0; q < n; ++q) {
      xs.push_back((x[q] + i) / k);
      ws.push_back(w[q] / k);
    }
  return tensor_rule(xs, ws);
}

int n_shape(int degree) { return degree == 1 ? 4 : 9; }

void unit_support_point(int degree, int a, double &sx, double &sy) {
  static const double q1[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
  static const double q2[9][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}, {0, .5}, {1, .5}, {.5, 0}, {.5, 1}, {.5, .5}};
  if (degree == 1) {
    sx = q1[a][0];
    sy = q1[a][1];
  } else {
    sx = q2[a][0];
    sy = q2[a][1];
  }
}

// ref: BEMProblem<3>::get_singular_quadrature, source/bem_stokes.cc:4912-4957
Rule2D make_singular_rule(int kind, int order, int fe_degree, int a) {
  BS_REQUIRE(fe_degree == 1 || fe_degree == 2, "fe_degree must be 1 or 2");
  BS_REQUIRE(a >= 0 && a < n_shape(fe_degree), "local index out of range");
  BS_REQUIRE(order >= 1 && order <= 64, "singular quadrature order out of range");
  double sx, sy;
  unit_support_point(fe_degree, a, sx, sy);
  if (kind == BS_SING_MIXED) return fe_degree > 1 ? qiterated(order, fe_degree) : lw_point(order, sx, sy);
  if (kind == BS_SING_DUFFY) return qsplit_duffy(order, sx, sy);
  if (kind == BS_SING_TELLES) {
    std::vector<double> x1, w1, x2, w2;
    telles_1d(order, sx, x1, w1);
    telles_1d(order, sy, x2, w2);
    return tensor_rule2(x1, w1, x2, w2);
  }
  throw Error(BS_ERR_INVALID, "unknown singular quadrature kind");
}

static void lagrange_1d(int degree, double x, double *l, double *d) {
  if (degree == 1) {
    l[0] = 1 - x;
    l[1] = x;
    d[0] = -1;
    d[1] = 1;
  } else {
    l[0] = 2 * (x - .5) * (x - 1);
    l[1] = -4 * x * (x - 1);
    l[2] = 2 * x * (x - .5);
    d[0] = 4 * x - 3;
    d[1] = -8 * x + 4;
    d[2] = 4 * x - 1;
  }
}

void shape_eval(int degree, double x, double y, double *phi, double *dphi) {
  double lx[3], dx[3], ly[3], dy[3];
  lagrange_1d(degree, x, lx, dx);
  lagrange_1d(degree, y, ly, dy);
  const int na = n_shape(degree);
  for (int a = 0; a < na; ++a) {
    double sx, sy;
    unit_support_point(degree, a, sx, sy);
    int ix = (int)std::lround(sx * degree), iy = (int)std::lround(sy * degree);
    phi[a] = lx[ix] * ly[iy];
    if (dphi) {
      dphi[2 * a] = dx[ix] * ly[iy];
      dphi[2 * a + 1] = lx[ix] * dy[iy];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// geometry: support points, locality order, partition
// ---------------------------------------------------------------------------------------------------------
static inline uint64_t spread3(uint64_t v) {
  v &= 0x1fffffULL;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

static void compute_support(Context &c);

void update_coordinates(Context &c) {
  compute_support(c);
  std::vector<double> sup_int((size_t)3 * c.N);
  for (int p = 0; p < c.N; ++p)
    for (int d = 0; d < 3; ++d) sup_int[(size_t)3 * p + d] = c.support[(size_t)3 * c.node_of_pos[p] + d];
  c.d_support.upload(sup_int, c.stream);
  c.d_map_nodes.upload(c.map_nodes, c.stream);
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

static void compute_support(Context &c) {
  const int N = c.N, na = c.na, nam = c.na_map;
  // support points = mapped unit support points (ref: DoFTools::map_dofs_to_support_points, bem_stokes.cc:2855)
  c.support.assign((size_t)3 * N, 0.0);
  std::vector<double> phi((size_t)na * nam);
  for (int a = 0; a < na; ++a) {
    double sx, sy;
    unit_support_point(c.fe_degree, a, sx, sy);
    shape_eval(c.map_degree, sx, sy, &phi[(size_t)a * nam], nullptr);
  }
  std::vector<char> seen(N, 0);
  for (int cell = 0; cell < c.ncell; ++cell)
    for (int a = 0; a < na; ++a) {
      int i = c.conn[(size_t)cell * na + a];
      BS_REQUIRE(i >= 0 && i < N, "conn_stokes entry out of range");
      if (seen[i]) continue;
      seen[i] = 1;
      double p[3] = {0, 0, 0};
      for (int b = 0; b < nam; ++b) {
        int m = c.conn_map[(size_t)cell * nam + b];
        BS_REQUIRE(m >= 0 && m < c.Nmap, "conn_map entry out of range");
        for (int d = 0; d < 3; ++d) p[d] += phi[(size_t)a * nam + b] * c.map_nodes[(size_t)3 * m + d];
      }
      for (int d = 0; d < 3; ++d) c.support[(size_t)3 * i + d] = p[d];
    }
  for (int i = 0; i < N; ++i) BS_REQUIRE(seen[i], "node without a cell");
}

// Support points, their Morton order and the row partition: host only (no device call)
void compute_node_order(Context &c) {
  const int N = c.N;
  compute_support(c);

  // Morton order of the support points -> spatially compact column blocks and row partitions
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < N; ++i)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], c.support[(size_t)3 * i + d]);
      hi[d] = std::max(hi[d], c.support[(size_t)3 * i + d]);
    }
  double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-300});
  std::vector<uint64_t> code(N);
  for (int i = 0; i < N; ++i) {
    uint64_t k = 0;
    for (int d = 0; d < 3; ++d) {
      double t = (c.support[(size_t)3 * i + d] - lo[d]) / ext;
      uint64_t q = (uint64_t)std::min(2097151.0, std::max(0.0, t * 2097151.0));
      k |= spread3(q) << d;
    }
    code[i] = k;
  }
  std::vector<int> owner(N, 0);
  const bool user_owner = !c.owner_in.empty();
  if (user_owner) {
    BS_REQUIRE((int)c.owner_in.size() == N, "owner_of_node length != n_nodes");
    owner = c.owner_in;
    for (int i = 0; i < N; ++i) BS_REQUIRE(owner[i] >= 0 && owner[i] < c.nranks, "owner_of_node out of range");
  }
  c.node_of_pos.resize(N);
  std::iota(c.node_of_pos.begin(), c.node_of_pos.end(), 0);
  std::stable_sort(c.node_of_pos.begin(), c.node_of_pos.end(), [&](int a, int b) {
    if (owner[a] != owner[b]) return owner[a] < owner[b];
    return code[a] < code[b];
  });
  c.pos_of_node.resize(N);
  for (int p = 0; p < N; ++p) c.pos_of_node[c.node_of_pos[p]] = p;
  c.part_start.assign(c.nranks + 1, 0);
  if (user_owner) {
    for (int i = 0; i < N; ++i) c.part_start[owner[i] + 1]++;
    for (int r = 0; r < c.nranks; ++r) c.part_start[r + 1] += c.part_start[r];
  } else {
    for (int r = 0; r <= c.nranks; ++r) c.part_start[r] = (int)(((long long)N * r) / c.nranks);
  }
  c.p0 = c.part_start[c.rank];
  c.p1 = c.part_start[c.rank + 1];
  c.rows_loc = (size_t)3 * (c.p1 - c.p0);
}

void build_geometry(Context &c) {
  const int N = c.N, na = c.na;
  compute_node_order(c);

  // upload
  std::vector<double> sup_int((size_t)3 * N);
  for (int p = 0; p < N; ++p)
    for (int d = 0; d < 3; ++d) sup_int[(size_t)3 * p + d] = c.support[(size_t)3 * c.node_of_pos[p] + d];
  c.d_support.upload(sup_int, c.stream);
  c.d_map_nodes.upload(c.map_nodes, c.stream);
  std::vector<int> conn_pos(c.conn.size());
  for (size_t k = 0; k < c.conn.size(); ++k) conn_pos[k] = c.pos_of_node[c.conn[k]];
  c.d_conn_pos.upload(conn_pos, c.stream);
  c.d_conn_map.upload(c.conn_map, c.stream);

  // patches (cells containing a node + FIRST local index, ref: bem_stokes.cc:2885-2895) by internal position
  std::vector<int> cnt(N + 1, 0);
  for (int cell = 0; cell < c.ncell; ++cell)
    for (int a = 0; a < na; ++a) {
      int p = conn_pos[(size_t)cell * na + a];
      bool first = true;
      for (int b = 0; b < a; ++b)
        if (conn_pos[(size_t)cell * na + b] == p) first = false;
      if (first) cnt[p + 1]++;
    }
  for (int p = 0; p < N; ++p) cnt[p + 1] += cnt[p];
  std::vector<int> pc(cnt[N]), pl(cnt[N]), fill(cnt.begin(), cnt.end() - 1);
  for (int cell = 0; cell < c.ncell; ++cell)
    for (int a = 0; a < na; ++a) {
      int p = conn_pos[(size_t)cell * na + a];
      bool first = true;
      for (int b = 0; b < a; ++b)
        if (conn_pos[(size_t)cell * na + b] == p) first = false;
      if (!first) continue;
      pc[fill[p]] = cell;
      pl[fill[p]] = a;
      fill[p]++;
    }
  c.d_patch_ptr.upload(cnt, c.stream);
  c.d_patch_cell.upload(pc, c.stream);
  c.d_patch_local.upload(pl, c.stream);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  c.have_geometry = true;
}

// ---------------------------------------------------------------------------------------------------------
// tables depending on quadrature: regular shape tables, column blocks, singular rule tables
// ---------------------------------------------------------------------------------------------------------
// Cell-split mode of K1: order the cells of a block as a sequence of pairs that share no node (the two thread sets of a CTA
// update their cells' tile entries concurrently), -1 where a cell has no partner.  Matching on the "shares no node" graph
// of the block: greedy start, then augmentation over pair swaps (blocks have <= 16 cells).  The thread sets may be one
// step apart, so a cell must not share a node with the other set's cell of the previous step either; the pairs are
// ordered (and oriented) to make that rare, and the steps where it cannot be avoided are flagged in `sync` (bit = step).
static std::vector<int> match_cells(const std::vector<int> &cells, const std::vector<int> &cpos, int na) {
  const int n = (int)cells.size();
  auto compatible = [&](int x, int y) {
    for (int a = 0; a < na; ++a)
      for (int b = 0; b < na; ++b)
        if (cpos[(size_t)cells[x] * na + a] == cpos[(size_t)cells[y] * na + b]) return false;
    return true;
  };
  std::vector<std::vector<char>> ok(n, std::vector<char>(n, 0));
  for (int x = 0; x < n; ++x)
    for (int y = x + 1; y < n; ++y) ok[x][y] = ok[y][x] = compatible(x, y) ? 1 : 0;
  std::vector<int> mate(n, -1);
  for (int x = 0; x < n; ++x) {  // greedy start: the compatible free cell with the fewest compatible alternatives
    if (mate[x] >= 0) continue;
    int best = -1, best_deg = 1 << 30;
    for (int y = 0; y < n; ++y) {
      if (y == x || mate[y] >= 0 || !ok[x][y]) continue;
      int deg = 0;
      for (int z = 0; z < n; ++z) deg += (z != x && mate[z] < 0 && ok[y][z]);
      if (deg < best_deg) best = y, best_deg = deg;
    }
    if (best >= 0) mate[x] = best, mate[best] = x;
  }
  // improve: two free cells x, y and a matched pair (u, v) with x-u and y-v compatible -> two pairs instead of one
  for (bool changed = true; changed;) {
    changed = false;
    for (int x = 0; x < n && !changed; ++x) {
      if (mate[x] >= 0) continue;
      for (int y = x + 1; y < n && !changed; ++y) {
        if (mate[y] >= 0) continue;
        for (int u = 0; u < n && !changed; ++u) {
          const int v = mate[u];
          if (v < 0) continue;
          if (ok[x][u] && ok[y][v]) {
            mate[x] = u, mate[u] = x, mate[y] = v, mate[v] = y;
            changed = true;
          }
        }
      }
    }
  }
  return mate;
}

static std::vector<int> pair_cells(const std::vector<int> &cells, const std::vector<int> &cpos, int na, int &unpaired,
                                   unsigned &sync) {
  const int n = (int)cells.size();
  auto compatible = [&](int x, int y) {
    if (x < 0 || y < 0) return true;
    for (int a = 0; a < na; ++a)
      for (int b = 0; b < na; ++b)
        if (cpos[(size_t)cells[x] * na + a] == cpos[(size_t)cells[y] * na + b]) return false;
    return true;
  };
  const std::vector<int> mate = match_cells(cells, cpos, na);
  std::vector<std::pair<int, int>> pairs;  // local indices, second = -1: no partner
  for (int x = 0; x < n; ++x) {
    if (mate[x] >= 0 && mate[x] < x) continue;
    pairs.emplace_back(x, mate[x]);
    if (mate[x] < 0) ++unpaired;
  }
  // order: next = a remaining pair (either orientation) without a cross conflict with the previous step, if there is one
  std::vector<int> out;
  std::vector<char> used(pairs.size(), 0);
  int pa = -1, pb = -1;
  sync = 0;
  for (size_t step = 0; step < pairs.size(); ++step) {
    int pick = -1;
    bool swap = false, clean = false;
    for (size_t k = 0; k < pairs.size() && !clean; ++k) {
      if (used[k]) continue;
      for (int o = 0; o < 2 && !clean; ++o) {
        const int a = o ? pairs[k].second : pairs[k].first, b = o ? pairs[k].first : pairs[k].second;
        if (a < 0) continue;  // the first set always has a cell
        const bool good = compatible(a, pb) && compatible(b, pa);
        if (pick < 0 || good) pick = (int)k, swap = (o == 1), clean = good;
      }
    }
    used[pick] = 1;
    const int a = swap ? pairs[pick].second : pairs[pick].first, b = swap ? pairs[pick].first : pairs[pick].second;
    if (!clean && step > 0) sync |= 1u << step;
    out.push_back(cells[a]);
    out.push_back(b >= 0 ? cells[b] : -1);
    pa = a;
    pb = b;
  }
  return out;
}

// Cell blocks of K1 (host only, no device call)
void build_cell_blocks(Context &c) {
  // ---- cell blocks --------------------------------------------------------------------------------------
  // Disjoint, spatially compact clusters of cells touching at most tj distinct nodes each: every (row node,
  // cell) pair is integrated exactly once.  Blocks that share a node get different colours; colours are
  // launched one after another, so the partial tiles of a shared node column are combined in a fixed order
  // (first colour stores, later colours read-modify-write) without atomics.
  ColumnBlocks &B = c.blocks;
  const int na = c.na;
  B.cs = cell_sets(c);
  B.tj = choose_tj(na, c.kp.type, c.nq_pad, B.cs);
  const int tj = B.tj;
  const int max_block_cells = (B.cs == 2) ? 16 : 32;  // padded to pairs, a block still has at most 32 cell slots
  BS_REQUIRE(tj >= na, "shared memory too small for one cell per block");
  std::vector<int> cpos((size_t)c.ncell * na);
  for (size_t k = 0; k < cpos.size(); ++k) cpos[k] = c.pos_of_node[c.conn[k]];
  // node -> cells (by position)
  std::vector<int> nptr(c.N + 1, 0);
  for (size_t k = 0; k < cpos.size(); ++k) nptr[cpos[k] + 1]++;
  for (int i = 0; i < c.N; ++i) nptr[i + 1] += nptr[i];
  std::vector<int> ncells(nptr[c.N]), nfill(nptr.begin(), nptr.end() - 1);
  for (int cell = 0; cell < c.ncell; ++cell)
    for (int a = 0; a < na; ++a) ncells[nfill[cpos[(size_t)cell * na + a]]++] = cell;
  // seed order: cells sorted by their smallest node position (the node order is a Morton curve)
  std::vector<int> corder(c.ncell);
  std::iota(corder.begin(), corder.end(), 0);
  std::vector<int> ckey(c.ncell);
  for (int cell = 0; cell < c.ncell; ++cell) {
    int m = cpos[(size_t)cell * na];
    for (int a = 1; a < na; ++a) m = std::min(m, cpos[(size_t)cell * na + a]);
    ckey[cell] = m;
  }
  std::stable_sort(corder.begin(), corder.end(), [&](int x, int y) { return ckey[x] < ckey[y]; });
  // cell centroids: the growth prefers, among the cheapest candidates, the one closest to the block's centre, which
  // turns 2x4 strips into 3x3 patches (more cells per tile, fewer shared node columns)
  std::vector<double> ccen((size_t)3 * c.ncell, 0.0);
  for (int cell = 0; cell < c.ncell; ++cell)
    for (int a = 0; a < na; ++a)
      for (int d = 0; d < 3; ++d) ccen[(size_t)3 * cell + d] += c.support[(size_t)3 * c.conn[(size_t)cell * na + a] + d] / na;
  std::vector<int> block_of(c.ncell, -1);
  std::vector<std::vector<int>> bcells, bnodes;
  std::vector<int> mark(c.N, -1);  // mark[node] == stamp of the growth in progress when the node is in it
  int stamp = 0;
  // Growth from a seed: repeatedly the unassigned neighbour cell adding the fewest new nodes.  Ties, strategy 0: closest
  // to the block's centre (3 x 3 patches), then lowest key.  Strategies 1-4 (cell-split mode, where a block should be a
  // 2 x 4 strip so that its cells pair up): closest to a line along one of the seed's two local axes through the
  // middle of one of its edges (the strip's two rows lie on either side of that line), then closest to the seed along it.
  auto grow = [&](int seed, int strategy, std::vector<int> &cells, std::vector<int> &nodes) {
    ++stamp;
    cells.clear();
    nodes.clear();
    double bsum[3] = {0, 0, 0}, org[3] = {0, 0, 0}, ax[3] = {0, 0, 0}, pe[3] = {0, 0, 0}, wu = 1.0, wv = 1.0;
    if (strategy > 0) {
      const int *sn = &c.conn[(size_t)seed * na];
      const int axis = (strategy - 1) >> 1, side = (strategy - 1) & 1;  // axis 0: local nodes 0->1, axis 1: 0->2
      const int e0 = 0, e1 = axis == 0 ? 1 : 2, f0 = axis == 0 ? 2 : 1, f1 = 3;  // edge (e0,e1) and the opposite edge (f0,f1)
      const int m0 = side ? f0 : e0, m1 = side ? f1 : e1;
      double an = 0, pd = 0;
      for (int d = 0; d < 3; ++d) {
        const double *X = c.support.data();
        org[d] = 0.5 * (X[(size_t)3 * sn[m0] + d] + X[(size_t)3 * sn[m1] + d]);
        ax[d] = X[(size_t)3 * sn[e1] + d] - X[(size_t)3 * sn[e0] + d];
        pe[d] = X[(size_t)3 * sn[f0] + d] - X[(size_t)3 * sn[e0] + d];
        an += ax[d] * ax[d];
      }
      an = std::sqrt(std::max(an, 1e-300));
      for (int d = 0; d < 3; ++d) ax[d] /= an, pd += pe[d] * ax[d];
      double pn = 0;
      for (int d = 0; d < 3; ++d) pe[d] -= pd * ax[d], pn += pe[d] * pe[d];
      pn = std::sqrt(std::max(pn, 1e-300));
      for (int d = 0; d < 3; ++d) pe[d] /= pn;
      wu = an;  // cell size along and across the strip: distances are compared in whole cells
      wv = pn;
    }
    auto add_cell = [&](int cell) {
      cells.push_back(cell);
      for (int d = 0; d < 3; ++d) bsum[d] += ccen[(size_t)3 * cell + d];
      for (int a = 0; a < na; ++a) {
        const int p = cpos[(size_t)cell * na + a];
        if (mark[p] != stamp) {
          mark[p] = stamp;
          nodes.push_back(p);
        }
      }
    };
    add_cell(seed);
    while (true) {
      int best = -1, best_new = 1 << 30, best_key = 1 << 30;
      double best_d1 = 1e300, best_d2 = 1e300;
      const double inv = 1.0 / (double)cells.size();
      for (size_t in = 0; in < nodes.size(); ++in) {
        const int p = nodes[in];
        for (int e = nptr[p]; e < nptr[p + 1]; ++e) {
          const int cell = ncells[e];
          if (block_of[cell] >= 0 || std::find(cells.begin(), cells.end(), cell) != cells.end()) continue;
          int nn = 0;
          for (int a = 0; a < na; ++a) nn += (mark[cpos[(size_t)cell * na + a]] != stamp);
          double d1 = 0, d2 = 0;  // primary and secondary tie-break distances
          if (strategy == 0) {
            for (int d = 0; d < 3; ++d) {
              const double t = ccen[(size_t)3 * cell + d] - bsum[d] * inv;
              d1 += t * t;
            }
          } else {
            double u = 0, v = 0;
            for (int d = 0; d < 3; ++d) {
              const double t = ccen[(size_t)3 * cell + d] - org[d];
              u += t * ax[d];
              v += t * pe[d];
            }
            d1 = std::floor(std::fabs(v) / wv);        // 0: one of the two rows next to the line
            d2 = std::floor(std::fabs(u) / wu + 0.5);  // column distance from the seed
          }
          const bool closer = d1 < best_d1 * (1.0 - 1e-6), same = !closer && d1 <= best_d1 * (1.0 + 1e-6);
          const bool closer2 = same && d2 < best_d2 * (1.0 - 1e-6), same2 = same && !closer2 && d2 <= best_d2 * (1.0 + 1e-6);
          if (nn < best_new || (nn == best_new && (closer || closer2 || (same2 && ckey[cell] < best_key)))) {
            best = cell;
            best_new = nn;
            best_key = ckey[cell];
            best_d1 = d1;
            best_d2 = d2;
          }
        }
      }
      if (best < 0 || (int)nodes.size() + best_new > tj || (int)cells.size() >= max_block_cells) break;
      add_cell(best);
    }
  };
  const int nstrategies = (B.cs == 2 && !std::getenv("BS_GROW_COMPACT")) ? 5 : 1;
  std::vector<int> tcells, tnodes, gcells, gnodes;
  // Seeds: cell-split mode continues next to the blocks already made (advancing front, first in first out), so that
  // the strips of a structured region line up end to end instead of leaving gaps shorter than a strip; otherwise, and
  // whenever the front is empty, the next unassigned cell of the Morton order.
  std::vector<int> front;
  size_t front_head = 0, next_morton = 0;
  std::vector<char> queued(c.ncell, 0);
  const bool use_front = nstrategies > 1 && !std::getenv("BS_NO_FRONT");
  while (true) {
    int seed = -1;
    while (use_front && front_head < front.size()) {
      const int cand = front[front_head++];
      if (block_of[cand] < 0) {
        seed = cand;
        break;
      }
    }
    if (seed < 0) {
      while (next_morton < corder.size() && block_of[corder[next_morton]] >= 0) ++next_morton;
      if (next_morton == corder.size()) break;
      seed = corder[next_morton];
    }
    // cell-split mode: the growth whose cells pair up best (fewest steps per cell), then the larger one
    double best_score = 1e300;
    for (int st = 0; st < nstrategies; ++st) {
      grow(seed, st, tcells, tnodes);
      double score = 0.0;
      if (nstrategies > 1) {
        const std::vector<int> mate = match_cells(tcells, cpos, na);
        int single = 0;
        for (int m : mate) single += (m < 0);
        const double steps = 0.5 * (double)(tcells.size() + single);
        // cost model per cell: a step (two cells' rule rows), a touched node column (write-out, fused product) and
        // the fixed cost of a CTA (prologue, barriers), in the proportions of the measured profile
        score = (0.16 * steps + 0.0094 * (double)tnodes.size() + 0.2) / (double)tcells.size() + 1e-6 * st;
      }
      if (score < best_score) best_score = score, gcells = tcells, gnodes = tnodes;
    }
    const int b = (int)bcells.size();
    bcells.push_back(gcells);
    bnodes.push_back(gnodes);
    for (int cell : gcells) block_of[cell] = b;
    if (use_front)
      for (int p : gnodes)
        for (int e = nptr[p]; e < nptr[p + 1]; ++e) {
          const int cell = ncells[e];
          if (block_of[cell] < 0 && !queued[cell]) queued[cell] = 1, front.push_back(cell);
        }
  }
  B.nblocks = (int)bcells.size();
  // greedy colouring of the block conflict graph (blocks sharing a node)
  std::vector<std::vector<int>> blocks_of_node(c.N);
  for (int b = 0; b < B.nblocks; ++b)
    for (int p : bnodes[b]) blocks_of_node[p].push_back(b);
  std::vector<int> colour(B.nblocks, -1);
  int ncol = 0;
  {
    std::vector<int> used;
    for (int b = 0; b < B.nblocks; ++b) {
      used.assign(ncol + 1, 0);
      for (int p : bnodes[b])
        for (int ob : blocks_of_node[p])
          if (colour[ob] >= 0) used[colour[ob]] = 1;
      int k = 0;
      while (k < ncol && used[k]) ++k;
      colour[b] = k;
      ncol = std::max(ncol, k + 1);
    }
  }
  // emit blocks grouped by colour; first-touch flag = lowest colour among the blocks sharing the node
  std::vector<int> border(B.nblocks);
  std::iota(border.begin(), border.end(), 0);
  std::stable_sort(border.begin(), border.end(), [&](int x, int y) { return colour[x] < colour[y]; });
  B.colour_start.assign(ncol + 1, 0);
  for (int b = 0; b < B.nblocks; ++b) B.colour_start[colour[b] + 1]++;
  for (int k = 0; k < ncol; ++k) B.colour_start[k + 1] += B.colour_start[k];
  B.cell_ptr.assign(1, 0);
  B.cells.clear();
  B.slots.clear();
  B.nodes.assign((size_t)B.nblocks * tj, -1);
  B.first.assign((size_t)B.nblocks * tj, 0);
  B.max_cells = 0;
  B.unpaired = 0;
  B.sync.assign(B.nblocks, 0u);
  B.sync_steps = B.steps = 0;
  long long touched = 0;
  for (int nb = 0; nb < B.nblocks; ++nb) {
    const int b = border[nb];
    std::vector<int> nodes = bnodes[b];
    std::sort(nodes.begin(), nodes.end());  // ascending positions: runs of consecutive columns coalesce
    touched += (long long)nodes.size();
    for (size_t sidx = 0; sidx < nodes.size(); ++sidx) {
      const int p = nodes[sidx];
      B.nodes[(size_t)nb * tj + sidx] = p;
      int minc = colour[b];
      for (int ob : blocks_of_node[p]) minc = std::min(minc, colour[ob]);
      B.first[(size_t)nb * tj + sidx] = (minc == colour[b]) ? 1 : 0;
    }
    std::vector<int> order = bcells[b];
    if (B.cs == 2) {
      unsigned sync = 0;
      order = pair_cells(bcells[b], cpos, na, B.unpaired, sync);
      B.sync[nb] = sync;
      B.steps += (long long)order.size() / 2;
      for (unsigned m = sync; m; m &= m - 1) ++B.sync_steps;
    }
    if (B.cs != 2) B.steps += (long long)order.size();
    for (int cell : order) {
      B.cells.push_back(cell);
      for (int a = 0; a < na; ++a) {
        if (cell < 0) {  // no partner in this step
          B.slots.push_back(0);
          continue;
        }
        const int p = cpos[(size_t)cell * na + a];
        const int sidx = (int)(std::lower_bound(nodes.begin(), nodes.end(), p) - nodes.begin());
        B.slots.push_back((signed char)sidx);
      }
    }
    B.cell_ptr.push_back((int)B.cells.size());
    B.max_cells = std::max(B.max_cells, (int)order.size());
  }
  B.node_touch_ratio = (double)touched / std::max(1, c.N);
}

void build_tables(Context &c) {
  BS_REQUIRE(c.have_geometry && c.have_quadrature, "geometry and quadrature must be set first");
  const int na = c.na, nam = c.na_map, nq = c.reg.size();
  c.nq = nq;
  c.nq_pad = (nq + 1) & ~1;  // keep every 7*nq_pad*8-byte cell record a multiple of 16 B for bulk copies
  std::vector<double> phi((size_t)nq * na), tab((size_t)nq * nam * 3);
  std::vector<double> ph(MAX_NA), dph(2 * MAX_NA);
  for (int q = 0; q < nq; ++q) {
    shape_eval(c.fe_degree, c.reg.xi[2 * q], c.reg.xi[2 * q + 1], ph.data(), nullptr);
    for (int a = 0; a < na; ++a) phi[(size_t)q * na + a] = ph[a];
    shape_eval(c.map_degree, c.reg.xi[2 * q], c.reg.xi[2 * q + 1], ph.data(), dph.data());
    for (int a = 0; a < nam; ++a) {
      tab[((size_t)q * nam + a) * 3 + 0] = ph[a];
      tab[((size_t)q * nam + a) * 3 + 1] = dph[2 * a];
      tab[((size_t)q * nam + a) * 3 + 2] = dph[2 * a + 1];
    }
  }
  c.d_phi_reg.upload(phi, c.stream);
  c.d_map_tab_reg.upload(tab, c.stream);
  {  // 1-D factors of the tensor-product shape functions: phi_a(q) = l_ix(a)(x_qx) * l_iy(a)(x_qy)
    const int n1 = (int)c.x1d.size(), nb1 = c.fe_degree + 1;
    std::vector<double> l1((size_t)n1 * nb1);
    for (int i = 0; i < n1; ++i) {
      double l[3], d[3];
      lagrange_1d(c.fe_degree, c.x1d[i], l, d);
      for (int b = 0; b < nb1; ++b) l1[(size_t)i * nb1 + b] = l[b];
    }
    c.d_l1d.upload(l1, c.stream);
  }

  build_cell_blocks(c);
  ColumnBlocks &B = c.blocks;
  c.d_blk_cell_ptr.upload(B.cell_ptr, c.stream);
  c.d_blk_cells.upload(B.cells, c.stream);
  c.d_blk_slots.upload(B.slots, c.stream);
  c.d_blk_nodes.upload(B.nodes, c.stream);
  c.d_blk_first.upload(B.first, c.stream);
  c.d_blk_sync.upload(B.sync, c.stream);
  if (std::getenv("BS_TRACE"))
    fprintf(stderr, "[bs] cell blocks: %d blocks, tj %d, cell sets %d, %lld steps (%d cells without partner, %lld steps behind a barrier), touch ratio %.3f\n",
            B.nblocks, B.tj, B.cs, B.steps, B.unpaired, B.sync_steps, B.node_touch_ratio);

  // singular rule tables: per rule, per point: phi[na], then (phi_map, dphi_x, dphi_y)[na_map], then weight
  if (c.have_singular) {
    BS_REQUIRE((int)c.sing.size() == na, "one singular rule per scalar local index required");
    const int rec = na + 3 * nam + 1;
    std::vector<double> st;
    c.sing_off.assign(na, 0);
    c.sing_nq.assign(na, 0);
    for (int a = 0; a < na; ++a) {
      c.sing_off[a] = (int)(st.size() / rec);
      c.sing_nq[a] = c.sing[a].size();
      for (int q = 0; q < c.sing[a].size(); ++q) {
        double x = c.sing[a].xi[2 * q], y = c.sing[a].xi[2 * q + 1];
        shape_eval(c.fe_degree, x, y, ph.data(), nullptr);
        for (int b = 0; b < na; ++b) st.push_back(ph[b]);
        shape_eval(c.map_degree, x, y, ph.data(), dph.data());
        for (int b = 0; b < nam; ++b) {
          st.push_back(ph[b]);
          st.push_back(dph[2 * b]);
          st.push_back(dph[2 * b + 1]);
        }
        st.push_back(c.sing[a].w[q]);
      }
    }
    c.d_sing_tab.upload(st, c.stream);
    c.d_sing_off.upload(c.sing_off, c.stream);
    c.d_sing_nq.upload(c.sing_nq, c.stream);
  }
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace bs

// ---------------------------------------------------------------------------------------------------------
// Host pre-pass (inputs of the hot path, O(N) sparse work the reference also does on the host):
// scalar mass matrix M_ab = sum_q phi_a phi_b JxW (ref: bem_stokes.cc:2499-2517), L2-projected unit normals
// M n = int phi n (3945-3998), M n_hat, l2 = n_hat^T M n_hat (4002-4005), rigid modes about `pole` and their duals
// M N_r (2626-2641, 2773).  The mass solve uses Jacobi-preconditioned CG to 1e-15 (the reference: Trilinos CG + AMG).
// ---------------------------------------------------------------------------------------------------------
namespace bs {

struct CsrMass {
  int n = 0;
  std::vector<int> ptr, col;
  std::vector<double> val;
  void mult(const double *x, double *y) const {
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int e = ptr[i]; e < ptr[i + 1]; ++e) s += val[e] * x[col[e]];
      y[i] = s;
    }
  }
};

void host_prepass(int fe_degree, int map_degree, int n_map_nodes, const double *euler_vec, int ncell, const int *conn_map,
                  int n_nodes, const int *conn, int quad_order, const double *pole, double *nhat, double *Mnhat, double *l2,
                  double *N_rigid, double *N_rigid_dual, double *area_out, double *support_out) {
  BS_REQUIRE((fe_degree == 1 || fe_degree == 2) && (map_degree == 1 || map_degree == 2), "FE degrees must be 1 or 2");
  const int na = n_shape(fe_degree), nam = n_shape(map_degree), N = n_nodes;
  std::vector<double> x1, w1;
  gauss_legendre_01(quad_order, x1, w1);
  Rule2D rule = tensor_rule(x1, w1);
  const int nq = rule.size();
  std::vector<double> phi((size_t)nq * na), pm((size_t)nq * nam), dpm((size_t)nq * nam * 2);
  for (int q = 0; q < nq; ++q) {
    shape_eval(fe_degree, rule.xi[2 * q], rule.xi[2 * q + 1], &phi[(size_t)q * na], nullptr);
    shape_eval(map_degree, rule.xi[2 * q], rule.xi[2 * q + 1], &pm[(size_t)q * nam], &dpm[(size_t)q * nam * 2]);
  }
  // sparsity: node -> neighbour nodes through cells
  std::vector<std::vector<int>> adj(N);
  for (int c = 0; c < ncell; ++c)
    for (int a = 0; a < na; ++a)
      for (int b = 0; b < na; ++b) adj[conn[(size_t)c * na + a]].push_back(conn[(size_t)c * na + b]);
  CsrMass M;
  M.n = N;
  M.ptr.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) {
    std::sort(adj[i].begin(), adj[i].end());
    adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
    M.ptr[i + 1] = M.ptr[i] + (int)adj[i].size();
  }
  M.col.resize(M.ptr[N]);
  M.val.assign(M.ptr[N], 0.0);
  for (int i = 0; i < N; ++i) std::copy(adj[i].begin(), adj[i].end(), M.col.begin() + M.ptr[i]);
  auto entry = [&](int i, int j) -> double & {
    const int *b = &M.col[M.ptr[i]], *e = &M.col[M.ptr[i + 1]];
    return M.val[M.ptr[i] + (int)(std::lower_bound(b, e, j) - b)];
  };
  std::vector<double> rhs((size_t)3 * N, 0.0);  // [c][i]
  double area = 0;
  std::vector<double> X((size_t)nam * 3), mloc((size_t)na * na), bloc((size_t)na * 3);
  for (int c = 0; c < ncell; ++c) {
    for (int a = 0; a < nam; ++a) {
      const int m = conn_map[(size_t)c * nam + a];
      for (int d = 0; d < 3; ++d) X[(size_t)3 * a + d] = euler_vec[(size_t)m + (size_t)d * n_map_nodes];
    }
    std::fill(mloc.begin(), mloc.end(), 0.0);
    std::fill(bloc.begin(), bloc.end(), 0.0);
    for (int q = 0; q < nq; ++q) {  // local element matrix and normal moments first, one scatter per cell afterwards
      double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
      for (int a = 0; a < nam; ++a)
        for (int d = 0; d < 3; ++d) {
          t1[d] += dpm[((size_t)q * nam + a) * 2] * X[(size_t)3 * a + d];
          t2[d] += dpm[((size_t)q * nam + a) * 2 + 1] * X[(size_t)3 * a + d];
        }
      const double nn[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
      const double J = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
      const double jxw = rule.w[q] * J;
      area += jxw;
      for (int a = 0; a < na; ++a) {
        const double pa = phi[(size_t)q * na + a];
        for (int d = 0; d < 3; ++d) bloc[(size_t)3 * a + d] += pa * (nn[d] / J) * jxw;
        for (int b = 0; b < na; ++b) mloc[(size_t)a * na + b] += pa * phi[(size_t)q * na + b] * jxw;
      }
    }
    for (int a = 0; a < na; ++a) {
      const int i = conn[(size_t)c * na + a];
      for (int d = 0; d < 3; ++d) rhs[(size_t)d * N + i] += bloc[(size_t)3 * a + d];
      for (int b = 0; b < na; ++b) entry(i, conn[(size_t)c * na + b]) += mloc[(size_t)a * na + b];
    }
  }
  // CG with Jacobi preconditioning, one solve per component
  std::vector<double> sol((size_t)3 * N, 0.0), r(N), z(N), p(N), Ap(N), dinv(N);
  for (int i = 0; i < N; ++i) dinv[i] = 1.0 / entry(i, i);
  for (int d = 0; d < 3; ++d) {
    double *x = &sol[(size_t)d * N];
    const double *b = &rhs[(size_t)d * N];
    double bn = 0;
    for (int i = 0; i < N; ++i) {
      r[i] = b[i];
      z[i] = dinv[i] * r[i];
      p[i] = z[i];
      bn += b[i] * b[i];
    }
    double rz = 0;
    for (int i = 0; i < N; ++i) rz += r[i] * z[i];
    for (int it = 0; it < 10 * N + 100 && bn > 0; ++it) {
      M.mult(p.data(), Ap.data());
      double pAp = 0;
      for (int i = 0; i < N; ++i) pAp += p[i] * Ap[i];
      const double alpha = rz / pAp;
      double rn = 0;
      for (int i = 0; i < N; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * Ap[i];
        rn += r[i] * r[i];
      }
      if (rn <= 1e-30 * bn) break;
      double rz2 = 0;
      for (int i = 0; i < N; ++i) {
        z[i] = dinv[i] * r[i];
        rz2 += r[i] * z[i];
      }
      const double beta = rz2 / rz;
      rz = rz2;
      for (int i = 0; i < N; ++i) p[i] = z[i] + beta * p[i];
    }
  }
  for (int i = 0; i < N; ++i) {
    const double nrm = std::sqrt(sol[i] * sol[i] + sol[(size_t)N + i] * sol[(size_t)N + i] + sol[(size_t)2 * N + i] * sol[(size_t)2 * N + i]);
    for (int d = 0; d < 3; ++d) nhat[(size_t)d * N + i] = sol[(size_t)d * N + i] / nrm;
  }
  double l2v = 0;
  for (int d = 0; d < 3; ++d) {
    M.mult(&nhat[(size_t)d * N], &Mnhat[(size_t)d * N]);
    for (int i = 0; i < N; ++i) l2v += nhat[(size_t)d * N + i] * Mnhat[(size_t)d * N + i];
  }
  if (l2) *l2 = l2v;
  if (area_out) *area_out = area;
  // support points of the unknown space and rigid modes about the pole
  std::vector<double> sup((size_t)3 * N, 0.0), pu((size_t)na * nam);
  for (int a = 0; a < na; ++a) {
    double sx, sy;
    unit_support_point(fe_degree, a, sx, sy);
    shape_eval(map_degree, sx, sy, &pu[(size_t)a * nam], nullptr);
  }
  for (int c = 0; c < ncell; ++c)
    for (int a = 0; a < na; ++a) {
      const int i = conn[(size_t)c * na + a];
      for (int d = 0; d < 3; ++d) {
        double s = 0;
        for (int b = 0; b < nam; ++b) s += pu[(size_t)a * nam + b] * euler_vec[(size_t)conn_map[(size_t)c * nam + b] + (size_t)d * n_map_nodes];
        sup[(size_t)3 * i + d] = s;
      }
    }
  if (support_out) std::copy(sup.begin(), sup.end(), support_out);
  if (N_rigid && N_rigid_dual) {
    const size_t n3 = (size_t)3 * N;
    std::fill(N_rigid, N_rigid + 6 * n3, 0.0);
    for (int i = 0; i < N; ++i) {
      const double x = sup[(size_t)3 * i] - pole[0], y = sup[(size_t)3 * i + 1] - pole[1], zc = sup[(size_t)3 * i + 2] - pole[2];
      for (int d = 0; d < 3; ++d) N_rigid[(size_t)d * n3 + (size_t)d * N + i] = 1.0;
      N_rigid[3 * n3 + (size_t)1 * N + i] = -zc;
      N_rigid[3 * n3 + (size_t)2 * N + i] = y;
      N_rigid[4 * n3 + (size_t)0 * N + i] = zc;
      N_rigid[4 * n3 + (size_t)2 * N + i] = -x;
      N_rigid[5 * n3 + (size_t)0 * N + i] = -y;
      N_rigid[5 * n3 + (size_t)1 * N + i] = x;
    }
    for (int rr = 0; rr < 6; ++rr)
      for (int d = 0; d < 3; ++d) M.mult(&N_rigid[rr * n3 + (size_t)d * N], &N_rigid_dual[rr * n3 + (size_t)d * N]);
  }
}

}  // namespace bs
