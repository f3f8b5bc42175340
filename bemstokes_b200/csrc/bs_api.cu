// extern "C" entry points of libbemstokes_b200 (see include/bemstokes_b200.h for the contract and the
// reference file:line each one replaces).  Exceptions never cross the ABI.
#include "bs_internal.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <numeric>

namespace bs {
const std::string &get_last_error();

struct Extra {  // per-context extras kept out of the header
  DBuf<int> d_node_of_pos;
  DBuf<double> vin, vout, vin2;
  DBuf<int> ir, ic;
  DBuf<unsigned char> d_flag;
  DBuf<double> dev_ref;  // device scratch in reference ordering
};
static Extra &extra(Context &c) {
  if (!c.extra) c.extra = new Extra();
  return *static_cast<Extra *>(c.extra);
}
static void drop_extra(Context &c) {
  delete static_cast<Extra *>(c.extra);
  c.extra = nullptr;
}

struct Timer {
  Context &c;
  double &acc;
  const char *name;
  double t0;
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  Timer(Context &c_, double &a, const char *n = "") : c(c_), acc(a), name(n), t0(now()) { cudaEventRecord(c.ev0, c.stream); }
  ~Timer() {
    cudaEventRecord(c.ev1, c.stream);
    cudaEventSynchronize(c.ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    acc += ms;
    if (std::getenv("BS_TRACE")) fprintf(stderr, "[bs trace rank %d] %s: device-span %.2f ms, host wall %.2f ms\n", c.rank, name, ms, 1e3 * (now() - t0));
  }
};
struct Mark {  // BS_TRACE sub-step marks
  Context &c;
  double t;
  bool on;
  Mark(Context &c_) : c(c_), t(Timer::now()), on(std::getenv("BS_TRACE") != nullptr) {}
  void operator()(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(c.stream);
    const double n = Timer::now();
    fprintf(stderr, "[bs trace rank %d]    %-28s %.2f ms\n", c.rank, what, 1e3 * (n - t));
    t = n;
  }
};

// reference-ordered vector (host or device per pointer mode) -> full internal device vector
static void to_internal(Context &c, const double *ref, int nextra, double *d_int, bool force_host = false) {
  Extra &e = extra(c);
  const size_t n = c.n3() + nextra;
  const double *src = ref;
  if (force_host || c.pointer_mode == BS_PTR_HOST) {
    e.dev_ref.alloc(std::max(e.dev_ref.n, n));
    BS_CUDA(cudaMemcpyAsync(e.dev_ref.p, ref, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    src = e.dev_ref.p;
  }
  perm_in(c, src, d_int, e.d_node_of_pos.p, nextra);
}
// internal entries [lo,hi) of a full-indexed internal device vector -> reference-ordered vector
static void from_internal(Context &c, const double *d_int_full, int nextra, double *ref, size_t lo, size_t hi,
                          bool force_host = false) {
  Extra &e = extra(c);
  const size_t n = c.n3() + nextra;
  if (force_host || c.pointer_mode == BS_PTR_HOST) {
    if (lo == 0 && hi == n) {
      e.dev_ref.alloc(std::max(e.dev_ref.n, n));
      perm_out(c, d_int_full, e.dev_ref.p, e.d_node_of_pos.p, nextra, lo, hi);
      BS_CUDA(cudaMemcpyAsync(ref, e.dev_ref.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      BS_CUDA(cudaStreamSynchronize(c.stream));
    } else {
      std::vector<double> h(hi - lo);
      BS_CUDA(cudaMemcpyAsync(h.data(), d_int_full + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      BS_CUDA(cudaStreamSynchronize(c.stream));
      for (size_t i = lo; i < hi; ++i) {
        if (i < c.n3()) ref[(size_t)c.node_of_pos[i / 3] + (i % 3) * c.N] = h[i - lo];
        else ref[i] = h[i - lo];
      }
    }
  } else {
    perm_out(c, d_int_full, ref, e.d_node_of_pos.p, nextra, lo, hi);
  }
}

static void alloc_matrix(Context &c, DBuf<double> &store, DMat &M, size_t rows, size_t cols) {
  store.alloc((c.rows_loc + MAX_RIGID) * c.ld + 2);
  // the assembly writes every (row, column < 3N) entry exactly once before anything reads it; only the padding
  // columns [3N, ld) and the MAX_RIGID spare rows have to be cleared
  if (c.rows_loc)
    BS_CUDA(cudaMemset2DAsync(store.p + c.n3(), c.ld * sizeof(double), 0, (c.ld - c.n3()) * sizeof(double), c.rows_loc, c.stream));
  BS_CUDA(cudaMemsetAsync(store.p + c.rows_loc * c.ld, 0, (MAX_RIGID * c.ld + 2) * sizeof(double), c.stream));
  M = DMat();  // also drops an implicit rank-1 term of the previous assembly
  M.p = store.p;
  M.rows = rows;
  M.cols = cols;
  M.ld = c.ld;
  M.owned = true;
}

static int nextra_of(Context &c, int which) { return which == BS_MAT_A ? c.num_rigid : 0; }

// the one accessor of the stored operators for every entry point: after bs_build_monolithic(keep_VK = 0) the
// monolithic matrix lives in V's storage, so V (with -K columns and rigid columns written into it) is gone
static const DMat &matrix_of(Context &c, int which) {
  BS_REQUIRE(which == BS_MAT_V || which == BS_MAT_K || which == BS_MAT_A, "unknown matrix id");
  if (which == BS_MAT_V && c.A_aliases_V) throw Error(BS_ERR_INVALID, "V was consumed by bs_build_monolithic(keep_VK=0)");
  const DMat &M = which == BS_MAT_V ? c.V : (which == BS_MAT_K ? c.K : c.A);
  BS_REQUIRE(M.valid(), which == BS_MAT_A ? "monolithic matrix not built" : "matrix not available");
  return M;
}

}  // namespace bs

using namespace bs;

#define BS_API_BEGIN try {
#define BS_API_END                                  \
  }                                                 \
  catch (const bs::Error &e) {                      \
    bs::set_last_error(e.what());                   \
    return e.code;                                  \
  }                                                 \
  catch (const std::exception &e) {                 \
    bs::set_last_error(e.what());                   \
    return BS_ERR_INVALID;                          \
  }                                                 \
  return BS_OK;

static Context &ctx_of(bs_context *h) {
  if (!h) throw Error(BS_ERR_INVALID, "null context");
  Context &c = h->c;
  BS_CUDA(cudaSetDevice(c.device));
  return c;
}

extern "C" {

static void build_constraint_tables(Context &c);

const char *bs_last_error(void) { return bs::get_last_error().c_str(); }
int bs_version(void) { return 100; }

int bs_create(bs_context **out, int device, int fe_degree, int map_degree) {
  BS_API_BEGIN
  BS_REQUIRE(out != nullptr, "null output pointer");
  BS_REQUIRE((fe_degree == 1 || fe_degree == 2) && (map_degree == 1 || map_degree == 2), "FE degrees must be 1 or 2");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(BS_ERR_NO_DEVICE, "no CUDA device visible: libbemstokes_b200 has no CPU fallback");
  BS_REQUIRE(device >= 0 && device < ndev, "device ordinal out of range");
  cudaDeviceProp prop;
  BS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    throw Error(BS_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                      ", this library is built for sm_100a (B200) only");
  BS_CUDA(cudaSetDevice(device));
  bs_context *h = new bs_context();
  Context &c = h->c;
  c.device = device;
  c.fe_degree = fe_degree;
  c.map_degree = map_degree;
  c.na = n_shape(fe_degree);
  c.na_map = n_shape(map_degree);
  c.sm_count = prop.multiProcessorCount;
  BS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.own_stream = true;
  BS_CUDA(cudaEventCreate(&c.ev0));
  BS_CUDA(cudaEventCreate(&c.ev1));
  *out = h;
  BS_API_END
}

int bs_destroy(bs_context *h) {
  BS_API_BEGIN
  if (!h) return BS_OK;
  Context &c = h->c;
  cudaSetDevice(c.device);
  cudaStreamSynchronize(c.stream);
  for (int r = 0; r < (int)c.peer_xbuf.size(); ++r)
    if (r != c.rank) {
      if (c.peer_xbuf[r]) cudaIpcCloseMemHandle(c.peer_xbuf[r]);
      if (c.peer_flags[r]) cudaIpcCloseMemHandle(c.peer_flags[r]);
    }
  drop_extra(c);
  gm_host_release(c);
  if (c.ev0) cudaEventDestroy(c.ev0);
  if (c.ev1) cudaEventDestroy(c.ev1);
  if (c.own_stream && c.stream) cudaStreamDestroy(c.stream);
  delete h;
  BS_API_END
}

int bs_set_pointer_mode(bs_context *h, int mode) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(mode == BS_PTR_HOST || mode == BS_PTR_DEVICE, "bad pointer mode");
  c.pointer_mode = mode;
  BS_API_END
}

int bs_set_stream(bs_context *h, void *s) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  if (c.own_stream) cudaStreamDestroy(c.stream);
  if (s) {
    c.stream = (cudaStream_t)s;
    c.own_stream = false;
  } else {
    BS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.own_stream = true;
  }
  BS_API_END
}

int bs_set_partition(bs_context *h, int rank, int nranks, const int *owner_of_node, int n_nodes) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  BS_REQUIRE(!c.have_geometry, "bs_set_partition must precede bs_set_geometry");
  c.rank = rank;
  c.nranks = nranks;
  c.owner_in.clear();
  if (owner_of_node) c.owner_in.assign(owner_of_node, owner_of_node + n_nodes);
  BS_API_END
}

int bs_prepass(bs_context *h, int pole_kind, const double *pole, double *nhat, double *Mnhat, double *l2gamma, double *N_rigid,
               double *N_rigid_dual, double *area, double *support_points, double *center_of_mass, double *pole_used,
               int *cg_iterations) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(nhat && Mnhat, "nhat / Mnhat output arrays are required");
  BS_REQUIRE((N_rigid == nullptr) == (N_rigid_dual == nullptr), "N_rigid and N_rigid_dual go together");
  BS_REQUIRE(pole_kind == BS_POLE_ORIGIN || pole_kind == BS_POLE_POINT || pole_kind == BS_POLE_BARICENTER, "unknown pole kind");
  BS_REQUIRE(pole_kind != BS_POLE_POINT || pole, "BS_POLE_POINT needs the point");
  const size_t n3 = c.n3();
  double *d_nh = c.wsd("pre.nhat", n3), *d_mn = c.wsd("pre.Mnhat", n3);
  double *d_nr = N_rigid ? c.wsd("pre.Nr", 6 * n3) : nullptr, *d_nrd = N_rigid ? c.wsd("pre.Nrd", 6 * n3) : nullptr;
  device_prepass(c, pole_kind, pole, d_nh, d_mn, d_nr, d_nrd, l2gamma, area, center_of_mass, pole_used, cg_iterations);
  from_internal(c, d_nh, 0, nhat, 0, n3);
  from_internal(c, d_mn, 0, Mnhat, 0, n3);
  if (N_rigid)
    for (int r = 0; r < 6; ++r) {
      from_internal(c, d_nr + r * n3, 0, N_rigid + r * n3, 0, n3);
      from_internal(c, d_nrd + r * n3, 0, N_rigid_dual + r * n3, 0, n3);
    }
  if (support_points) std::copy(c.support.begin(), c.support.end(), support_points);  // host copy, original node order
  BS_CUDA(cudaStreamSynchronize(c.stream));
  BS_API_END
}

int bs_get_owned_nodes(bs_context *h, int *n_owned, int *owned) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(c.have_geometry, "geometry not set");
  if (n_owned) *n_owned = c.p1 - c.p0;
  if (owned)
    for (int p = c.p0; p < c.p1; ++p) owned[p - c.p0] = c.node_of_pos[p];
  BS_API_END
}

int bs_set_geometry(bs_context *h, int n_map_nodes, const double *euler_vec, int ncell, const int *conn_map, int n_nodes,
                    const int *conn_stokes, const int *material_id) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(n_map_nodes > 0 && ncell > 0 && n_nodes > 0 && euler_vec && conn_map && conn_stokes, "bad geometry arguments");
  // Moving geometry on an unchanged mesh (the reference's per-frame compute_euler_vector, bem_stokes.cc:5677-5687):
  // keep ordering, partition, cell blocks, tables and matrix storage; only the coordinates are new.
  if (c.have_geometry && c.Nmap == n_map_nodes && c.N == n_nodes && c.ncell == ncell &&
      std::equal(c.conn_map.begin(), c.conn_map.end(), conn_map) && std::equal(c.conn.begin(), c.conn.end(), conn_stokes)) {
    for (int i = 0; i < n_map_nodes; ++i)
      for (int d = 0; d < 3; ++d) c.map_nodes[(size_t)3 * i + d] = euler_vec[(size_t)i + (size_t)d * n_map_nodes];
    update_coordinates(c);
    c.V = DMat();
    c.K = DMat();
    c.A = DMat();
    return BS_OK;
  }
  c.Nmap = n_map_nodes;
  c.N = n_nodes;
  c.ncell = ncell;
  c.map_nodes.resize((size_t)3 * n_map_nodes);
  for (int i = 0; i < n_map_nodes; ++i)
    for (int d = 0; d < 3; ++d) c.map_nodes[(size_t)3 * i + d] = euler_vec[(size_t)i + (size_t)d * n_map_nodes];
  c.conn_map.assign(conn_map, conn_map + (size_t)ncell * c.na_map);
  c.conn.assign(conn_stokes, conn_stokes + (size_t)ncell * c.na);
  c.material.assign(ncell, 0);
  if (material_id) c.material.assign(material_id, material_id + ncell);
  build_geometry(c);
  build_constraint_tables(c);  // constraints given before the geometry
  extra(c).d_node_of_pos.upload(c.node_of_pos, c.stream);
  c.ld = ((c.n3() + MAX_RIGID + 15) / 16) * 16;
  // a new geometry invalidates matrices and the factorised preconditioner (sizes and ordering changed)
  c.prec_kind = BS_PREC_NONE;
  c.lu_blocks.clear();
  c.V = DMat();
  c.K = DMat();
  c.A = DMat();
  if (c.have_quadrature) build_tables(c);
  BS_API_END
}

int bs_make_gauss_1d(int n, double *x, double *w) {
  try {
    std::vector<double> xs, ws;
    gauss_legendre_01(n, xs, ws);
    for (int i = 0; i < n; ++i) {
      x[i] = xs[i];
      w[i] = ws[i];
    }
    return n;
  } catch (const std::exception &e) {
    bs::set_last_error(e.what());
    return BS_ERR_INVALID;
  }
}

int bs_make_singular_rule(int kind, int order, int fe_degree, int local_index, int capacity, double *xi, double *w) {
  try {
    Rule2D r = make_singular_rule(kind, order, fe_degree, local_index);
    if (xi && w) {
      if (capacity < r.size()) throw Error(BS_ERR_INVALID, "capacity too small for the rule");
      std::copy(r.xi.begin(), r.xi.end(), xi);
      std::copy(r.w.begin(), r.w.end(), w);
    }
    return r.size();
  } catch (const bs::Error &e) {
    bs::set_last_error(e.what());
    return e.code;
  } catch (const std::exception &e) {
    bs::set_last_error(e.what());
    return BS_ERR_INVALID;
  }
}

int bs_set_quadrature(bs_context *h, int n1d, const double *x1d, const double *w1d) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(n1d >= 1 && n1d <= 32, "regular quadrature order must be in [1,32]");
  if (x1d && w1d) {
    c.x1d.assign(x1d, x1d + n1d);
    c.w1d.assign(w1d, w1d + n1d);
  } else {
    gauss_legendre_01(n1d, c.x1d, c.w1d);
  }
  c.reg = tensor_rule(c.x1d, c.w1d);
  c.have_quadrature = true;
  if (c.have_geometry) build_tables(c);
  BS_API_END
}

int bs_set_singular_quadrature(bs_context *h, int kind, int order) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  c.sing.clear();
  for (int a = 0; a < c.na; ++a) c.sing.push_back(make_singular_rule(kind, order, c.fe_degree, a));
  c.have_singular = true;
  if (c.have_geometry && c.have_quadrature) build_tables(c);
  BS_API_END
}

int bs_set_singular_rule(bs_context *h, int a, int nq, const double *xi, const double *w) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(a >= 0 && a < c.na && nq > 0 && xi && w, "bad singular rule arguments");
  if ((int)c.sing.size() != c.na) c.sing.assign(c.na, Rule2D());
  c.sing[a].xi.assign(xi, xi + 2 * (size_t)nq);
  c.sing[a].w.assign(w, w + nq);
  c.have_singular = true;
  for (int b = 0; b < c.na; ++b) c.have_singular = c.have_singular && c.sing[b].size() > 0;
  if (c.have_singular && c.have_geometry && c.have_quadrature) build_tables(c);
  BS_API_END
}

int bs_set_kernel(bs_context *h, int type, double eps, int wall_orientation, const double *wall_position) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(type >= 0 && type <= 2, "unknown kernel type");
  BS_REQUIRE(wall_orientation >= 0 && wall_orientation < 3, "wall orientation must be 0,1,2");
  bool retile = (type == BS_KERNEL_FREE) != (c.kp.type == BS_KERNEL_FREE);
  const int cs_before = cell_sets(c);
  c.kp.type = type;
  c.kp.eps = eps;
  c.kp.o = wall_orientation;
  c.kp.wall_pos = wall_position ? wall_position[wall_orientation] : 0.0;
  retile = retile || cell_sets(c) != cs_before;  // the cell blocks depend on the K1 variant (tile size, cell pairs)
  if (retile && c.have_geometry && c.have_quadrature) build_tables(c);
  BS_API_END
}

// internal tables of the owned constrained rows (after geometry + constraints are known)
static void build_constraint_tables(Context &c) {
  c.n_cons_owned = 0;
  if (c.cons_dof.empty() || !c.have_geometry) return;
  const int N = c.N;
  auto to_int = [&](int ref) { return 3 * c.pos_of_node[ref % N] + ref / N; };
  std::vector<int> rows, ptr(1, 0), cols;
  std::vector<double> coefs;
  std::vector<unsigned char> node(std::max(1, c.p1 - c.p0), 0);
  for (size_t k = 0; k < c.cons_dof.size(); ++k) {
    const int ref = c.cons_dof[k];
    BS_REQUIRE(ref >= 0 && ref < 3 * N, "constrained dof out of range");
    const int gi = to_int(ref), li = gi - 3 * c.p0;
    if (li < 0 || li >= (int)c.rows_loc) continue;  // another rank's row
    rows.push_back(li);
    for (int q = c.cons_ptr[k]; q < c.cons_ptr[k + 1]; ++q) {
      BS_REQUIRE(c.cons_col[q] >= 0 && c.cons_col[q] < 3 * N, "constraining dof out of range");
      cols.push_back(to_int(c.cons_col[q]));
      coefs.push_back(c.cons_coef[q]);
    }
    ptr.push_back((int)cols.size());
    if (ref < N) node[li / 3] = 1;  // the reference tests the x-component dof of the node (bem_stokes.cc:3078)
  }
  c.n_cons_owned = (int)rows.size();
  if (!c.n_cons_owned) return;
  if (cols.empty()) {
    cols.push_back(0);
    coefs.push_back(0.0);
  }
  c.d_cons_row.upload(rows, c.stream);
  c.d_cons_ptr.upload(ptr, c.stream);
  c.d_cons_col.upload(cols, c.stream);
  c.d_cons_coef.upload(coefs, c.stream);
  c.d_cons_node.upload(node, c.stream);
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

extern "C" int bs_set_constraints(bs_context *h, int n_constrained, const int *dof, const int *ptr, const int *cols, const double *coefs) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(n_constrained >= 0, "negative count");
  c.cons_dof.clear();
  c.cons_ptr.clear();
  c.cons_col.clear();
  c.cons_coef.clear();
  if (n_constrained > 0) {
    BS_REQUIRE(dof && ptr, "constraint arrays missing");
    BS_REQUIRE(ptr[n_constrained] == 0 || (cols && coefs), "constraint entries missing");
    c.cons_dof.assign(dof, dof + n_constrained);
    c.cons_ptr.assign(ptr, ptr + n_constrained + 1);
    c.cons_col.assign(cols, cols + ptr[n_constrained]);
    c.cons_coef.assign(coefs, coefs + ptr[n_constrained]);
  }
  build_constraint_tables(c);
  BS_API_END
}

extern "C" int bs_set_torque_mode(bs_context *h, const double *N_torque, const double *N_torque_dual, double rhs_value) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  if (!N_torque) {
    c.torque_on = false;
    c.torque_mode.clear();
    c.torque_dual.clear();
    return BS_OK;
  }
  BS_REQUIRE(c.have_geometry && N_torque_dual, "geometry first; both vectors are needed");
  c.torque_mode.assign(N_torque, N_torque + c.n3());
  c.torque_dual.assign(N_torque_dual, N_torque_dual + c.n3());
  c.torque_rhs = rhs_value;
  c.torque_on = true;
  BS_API_END
}

// flags of the wall-velocity unknowns for the fused (no-K) assembly; see Context::col_flags
static void build_flag_tables(Context &c) {
  c.n_flagged = 0;
  if (c.col_flags.empty() || !c.have_geometry) return;
  BS_REQUIRE(c.col_flags.size() == c.n3(), "column flags: 3N entries expected");
  std::vector<int> kcol(c.n3(), -1);
  int nf = 0;
  for (size_t p = 0; p < (size_t)c.N; ++p)
    for (int k = 0; k < 3; ++k)
      if (c.col_flags[(size_t)c.node_of_pos[p] + (size_t)k * c.N]) kcol[3 * p + k] = nf++;
  c.n_flagged = nf;
  c.ldk = ((size_t)nf + 1) & ~(size_t)1;
  if (nf) c.d_kcol.upload(kcol, c.stream);
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

extern "C" int bs_set_column_flags(bs_context *h, const unsigned char *col_is_K) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  c.col_flags.clear();
  if (col_is_K) {
    BS_REQUIRE(c.have_geometry, "geometry first");
    c.col_flags.assign(col_is_K, col_is_K + c.n3());
  }
  build_flag_tables(c);
  BS_API_END
}

static void assemble_common(Context &c, bool fused) {
  BS_REQUIRE(c.have_geometry && c.have_quadrature && c.have_singular, "geometry, quadrature and singular quadrature must be set");
  c.fused = fused;
  alloc_matrix(c, c.storeV, c.V, c.rows_loc, c.n3());
  if (!fused) alloc_matrix(c, c.storeK, c.K, c.rows_loc, c.n3());
  else {
    c.K = DMat();
    c.storeK.release();  // the point of the fused mode: the double-layer matrix is never materialised
    c.d_KX.alloc(c.rows_loc * c.panel_p + 2);
    c.d_KX.zero(c.stream);
    if (c.n_flagged > 0) c.d_Kflag.alloc(c.rows_loc * c.ldk + 2);  // every entry is stored by its first colour before anything reads it
  }
  c.A = DMat();
  c.A_aliases_V = false;
  {
    Timer t(c, c.stats.geometry_ms);
    launch_cell_geometry(c);
  }
  {
    Timer t(c, c.stats.assemble_regular_ms);
    launch_assembly_regular(c);
  }
  {
    Timer t(c, c.stats.assemble_singular_ms);
    launch_assembly_singular(c);
  }
  if (c.n_cons_owned) {  // constrained rows are not integrated: they hold the constraint equation (ref: 2970-2995)
    BS_REQUIRE(!fused, "hanging-node constraints are not supported by the fused (no-K) assembly");
    apply_constraint_rows(c, c.V, c.n3());
    apply_constraint_rows(c, c.K, c.n3());
  }
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

int bs_assemble_VK(bs_context *h) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  assemble_common(c, false);
  BS_API_END
}

// internal vector (3N) from a reference-ordered host vector after the tangential projection P v
static void project_to_internal(Context &c, const double *v, const double *nhat, const double *Mnhat, double l2, double *dst,
                                size_t stride) {
  const size_t n = c.n3();
  double d = 0;
  for (size_t i = 0; i < n; ++i) d += Mnhat[i] * v[i];
  d /= l2;
  for (size_t p = 0; p < (size_t)c.N; ++p)
    for (int k = 0; k < 3; ++k) {
      const size_t ref = (size_t)c.node_of_pos[p] + (size_t)k * c.N;
      dst[(3 * p + k) * stride] = v[ref] - d * nhat[ref];
    }
}

int bs_assemble_fused(bs_context *h, int num_rigid, const double *N_rigid, const double *nhat, const double *Mnhat,
                      double l2gamma, const double *shape_vel) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(num_rigid >= 0 && num_rigid <= MAX_RIGID && (num_rigid == 0 || N_rigid), "bad rigid modes");
  BS_REQUIRE(nhat && Mnhat && l2gamma > 0, "projector data missing");
  const size_t n = c.n3();
  const int pp = 3 + num_rigid + 1;
  BS_REQUIRE(pp <= MAX_PANEL, "panel too wide");
  c.panel_p = pp;
  c.panel_nr = num_rigid;
  c.h_panel.assign(n * pp, 0.0);
  for (size_t p = 0; p < (size_t)c.N; ++p)
    for (int k = 0; k < 3; ++k) c.h_panel[(3 * p + k) * pp + k] = 1.0;  // versors e_k (K correction, bem_stokes.cc:3044-3072)
  for (int r = 0; r < num_rigid; ++r) project_to_internal(c, N_rigid + (size_t)r * n, nhat, Mnhat, l2gamma, &c.h_panel[3 + r], pp);
  if (shape_vel) project_to_internal(c, shape_vel, nhat, Mnhat, l2gamma, &c.h_panel[3 + num_rigid], pp);
  c.d_panel.upload(c.h_panel, c.stream);
  assemble_common(c, true);
  BS_API_END
}

static void set_projector(Context &c, const double *nhat, const double *Mnhat, double l2) {
  BS_REQUIRE(nhat && Mnhat && l2 > 0, "projector data missing");
  c.d_nhat.alloc(c.n3() + 2);
  c.d_Mnhat.alloc(c.n3() + 2);
  to_internal(c, nhat, 0, c.d_nhat.p, true);
  to_internal(c, Mnhat, 0, c.d_Mnhat.p, true);
  c.l2gamma = l2;
  c.have_projector = true;
}

int bs_correct_V(bs_context *h, const double *nhat, const double *Mnhat, double l2gamma, double *Vn_out) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(c.V.valid(), "V not assembled");
  Timer t(c, c.stats.correct_ms, "bs_correct_V");
  Mark mark(c);
  set_projector(c, nhat, Mnhat, l2gamma);
  mark("set_projector");
  Extra &e = extra(c);
  e.vout.alloc(std::max(e.vout.n, c.n3() + MAX_RIGID + 2));
  double *vn_loc = e.vout.p + 3 * (size_t)c.p0;
  c.V.r1_u = c.V.r1_w = nullptr;  // V nhat of the raw operator (the correction is idempotent)
  c.V.r1_ncols = 0;
  gemv(c, c.V, c.d_nhat.p, vn_loc);
  mark("gemv V*nhat");
  if (Vn_out) from_internal(c, e.vout.p, 0, Vn_out, 3 * (size_t)c.p0, 3 * (size_t)c.p1, true);
  mark("Vn to host");
  // u = (nhat - V nhat) / l2 on the owned rows, w = M nhat:  V + u w^T.  Kept as the two vectors (DMat::r1_*): every reader of
  // the matrix applies the term, the read-modify-write pass over the whole matrix (86 GB at the benchmark size) is gone
  e.vin.alloc(std::max(e.vin.n, c.n3() + MAX_RIGID + 2));
  sub(c, c.d_nhat.p + 3 * (size_t)c.p0, vn_loc, e.vin.p, c.rows_loc);
  zero_constrained_entries(c, e.vin.p);  // "We correct only if we don't have constraints" (ref: 3024-3025)
  if (std::getenv("BS_EXPLICIT_RANK1")) {
    rank1_update(c, c.V, e.vin.p, c.d_Mnhat.p, 1.0 / l2gamma);
  } else {
    c.d_r1u.alloc(c.rows_loc + MAX_RIGID + 2);
    c.d_r1u.zero(c.stream);
    c.d_r1w.alloc(c.n3() + 2);
    copy(c, e.vin.p, c.d_r1u.p, c.rows_loc);
    scal(c, 1.0 / l2gamma, c.d_r1u.p, c.rows_loc);
    copy(c, c.d_Mnhat.p, c.d_r1w.p, c.n3());
    c.V.r1_u = c.d_r1u.p;
    c.V.r1_w = c.d_r1w.p;
    c.V.r1_ncols = c.n3();
  }
  BS_CUDA(cudaStreamSynchronize(c.stream));
  mark("rank-1 update");
  BS_API_END
}

int bs_correct_K(bs_context *h, int use_internal_alpha) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  if (c.fused) {
    // C_k = K e_k are the first three panel products; they enter the monolithic build (no K to correct in place)
    BS_REQUIRE(c.d_KX.p != nullptr, "fused assembly not run");
    std::vector<double> kx(c.rows_loc * c.panel_p);
    BS_CUDA(cudaMemcpyAsync(kx.data(), c.d_KX.p, kx.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    c.h_C.assign(3 * c.rows_loc, 0.0);
    for (size_t r = 0; r < c.rows_loc; ++r)
      for (int k = 0; k < 3; ++k) c.h_C[(size_t)k * c.rows_loc + r] = kx[r * c.panel_p + k];
    c.fused_alpha = use_internal_alpha;
    return BS_OK;
  }
  BS_REQUIRE(c.K.valid(), "K not assembled");
  Timer t(c, c.stats.correct_ms, "bs_correct_K");
  Mark mark(c);
  const size_t n = c.n3();
  const size_t ldx = (n + 4) & ~(size_t)3;  // multiple of 4: 32-byte vector loads in the tensor-path sweep
  std::vector<double> E(3 * ldx, 0.0);
  for (size_t p = 0; p < (size_t)c.N; ++p)
    for (int k = 0; k < 3; ++k) E[k * ldx + 3 * p + k] = 1.0;
  Extra &e = extra(c);
  e.vin2.upload(E, c.stream);
  e.vout.alloc(std::max(e.vout.n, 3 * c.rows_loc + 2));
  mark("versor upload");
  gemv_multi(c, c.K, 3, e.vin2.p, ldx, e.vout.p, c.rows_loc);
  mark("K * versors");
  k_correct_diag(c, c.K, e.vout.p, use_internal_alpha);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  mark("diag correction");
  BS_API_END
}

int bs_build_monolithic(bs_context *h, const unsigned char *col_is_K, int num_rigid, const double *N_rigid,
                        const double *N_rigid_dual, const double *nhat, const double *Mnhat, double l2gamma, int grid_type,
                        int imposed_component, double scaling, const double *shape_vel, int keep_VK, double *rhs_out) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(c.V.valid() && (c.K.valid() || c.fused), "V and K must be assembled (and corrected) first");
  BS_REQUIRE(num_rigid >= 0 && num_rigid <= MAX_RIGID, "num_rigid out of range");
  if (c.fused) {
    if (col_is_K) {
      BS_REQUIRE(c.n_flagged > 0 && c.col_flags.size() == c.n3() && std::equal(c.col_flags.begin(), c.col_flags.end(), col_is_K),
                 "fused (no-K) assembly with mixed boundary conditions: hand the same flags to bs_set_column_flags before bs_assemble_fused");
    } else {
      BS_REQUIRE(c.n_flagged == 0, "bs_set_column_flags was given flags, bs_build_monolithic none");
    }
    BS_REQUIRE(num_rigid == c.panel_nr && c.h_C.size() == 3 * c.rows_loc, "fused mode: call bs_assemble_fused and bs_correct_K first");
    keep_VK = 0;
  }
  BS_REQUIRE(num_rigid == 0 || (N_rigid && N_rigid_dual), "rigid modes missing");
  BS_REQUIRE(!(c.fused && (c.torque_on || c.n_cons_owned)), "torque unknown / constraints are not supported by the fused (no-K) assembly");
  BS_REQUIRE(!c.torque_on || num_rigid + 1 <= MAX_RIGID, "no room for the torque unknown");
  Timer t(c, c.stats.monolithic_ms, "bs_build_monolithic");
  set_projector(c, nhat, Mnhat, l2gamma);
  Extra &e = extra(c);
  const size_t n = c.n3();
  // with solve_with_torque the flagellum's angular velocity is one more unknown after the rigid ones: same column
  // (-scaling P K P N) and row (scaling N_dual) as a rigid mode of a Real grid (ref: 3143-3147, 3252-3256, 3340-3352)
  const int nr = num_rigid + (c.torque_on ? 1 : 0);
  auto mode = [&](int r) { return r < num_rigid ? N_rigid + (size_t)r * n : c.torque_mode.data(); };
  auto dual = [&](int r) { return r < num_rigid ? N_rigid_dual + (size_t)r * n : c.torque_dual.data(); };
  const int nvec = nr + 1;  // rigid modes (+ torque mode) + shape velocity
  c.num_rigid = nr;
  c.mono_size = n + nr;
  const bool last = (c.rank == c.nranks - 1);
  // ---- P N_r and P u_shape on the host (O(n) work, inputs of the boundary), then K * panel on the device
  const size_t ldx = (n + 4) & ~(size_t)3;  // multiple of 4: 32-byte vector loads in the tensor-path sweep
  std::vector<double> X((size_t)nvec * ldx, 0.0);
  auto project_into = [&](const double *v, double *dst_int) {
    double d = 0;
    for (size_t i = 0; i < n; ++i) d += Mnhat[i] * v[i];
    d /= l2gamma;
    for (size_t p = 0; p < (size_t)c.N; ++p)
      for (int k = 0; k < 3; ++k) {
        const size_t ref = (size_t)c.node_of_pos[p] + (size_t)k * c.N;
        dst_int[3 * p + k] = v[ref] - d * nhat[ref];
      }
  };
  for (int r = 0; r < nr; ++r) project_into(mode(r), &X[(size_t)r * ldx]);
  const bool use_shape = (grid_type == BS_GRID_REAL && shape_vel != nullptr);
  if (use_shape) project_into(shape_vel, &X[(size_t)nr * ldx]);
  e.vout.alloc(std::max(e.vout.n, (size_t)nvec * c.rows_loc + 2));
  if (!c.fused) {
    e.vin2.upload(X, c.stream);
    gemv_multi(c, c.K, nvec, e.vin2.p, ldx, e.vout.p, c.rows_loc);
  } else {
    // K_corr x = K x - D x + x with the uncorrected products of the assembly epilogue, (D x)_(i,j) = sum_k C_k[i,j] x_(i,k)
    std::vector<double> kx(c.rows_loc * c.panel_p), out((size_t)nvec * c.rows_loc, 0.0);
    BS_CUDA(cudaMemcpyAsync(kx.data(), c.d_KX.p, kx.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    const int pp = c.panel_p;
    for (int v = 0; v < nvec; ++v) {
      if (v == nr && !use_shape) continue;
      for (size_t r = 0; r < c.rows_loc; ++r) {
        const size_t node_pos = (size_t)c.p0 + r / 3;
        double val = kx[r * pp + 3 + v];
        for (int k = 0; k < 3; ++k) val -= c.h_C[(size_t)k * c.rows_loc + r] * c.h_panel[(3 * node_pos + k) * pp + 3 + v];
        if (!c.fused_alpha) val += c.h_panel[((size_t)3 * c.p0 + r) * pp + 3 + v];
        out[(size_t)v * c.rows_loc + r] = val;
      }
    }
    BS_CUDA(cudaMemcpyAsync(e.vout.p, out.data(), out.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
  }
  // second projection needs the full vectors: gather slices
  std::vector<double> Y((size_t)nvec * n, 0.0);
  {
    struct P_ { double *p; } full;
    full.p = c.wsd("mono.full", n + 2);
    std::vector<double> tmp(n);
    for (int v = 0; v < nvec; ++v) {
      exchange(c, BS_MAT_K, e.vout.p + (size_t)v * c.rows_loc, full.p + (c.nranks == 1 ? 0 : 0));
      if (c.nranks == 1) {
        BS_CUDA(cudaMemcpyAsync(tmp.data(), full.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      } else {
        BS_CUDA(cudaMemcpyAsync(tmp.data(), full.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      }
      BS_CUDA(cudaStreamSynchronize(c.stream));
      // internal -> apply P in internal ordering
      double d = 0;
      std::vector<double> &mn = tmp;  // alias for clarity
      (void)mn;
      for (size_t p = 0; p < (size_t)c.N; ++p)
        for (int k = 0; k < 3; ++k) d += Mnhat[(size_t)c.node_of_pos[p] + (size_t)k * c.N] * tmp[3 * p + k];
      d /= l2gamma;
      for (size_t p = 0; p < (size_t)c.N; ++p)
        for (int k = 0; k < 3; ++k)
          Y[(size_t)v * n + 3 * p + k] = tmp[3 * p + k] - d * nhat[(size_t)c.node_of_pos[p] + (size_t)k * c.N];
    }
  }
  // ---- matrix
  if (keep_VK) {
    alloc_matrix(c, c.storeA, c.A, c.rows_loc + (last ? nr : 0), n + nr);
    c.A_aliases_V = false;
  } else {
    c.A = c.V;
    c.A.rows = c.rows_loc + (last ? nr : 0);
    c.A.cols = n + nr;
    c.A.owned = false;
    c.A_aliases_V = true;
  }
  const unsigned char *d_flag = nullptr;
  if (col_is_K) {
    std::vector<unsigned char> f(n);
    for (size_t p = 0; p < (size_t)c.N; ++p)
      for (int k = 0; k < 3; ++k) f[3 * p + k] = col_is_K[(size_t)c.node_of_pos[p] + (size_t)k * c.N];
    e.d_flag.upload(f, c.stream);
    d_flag = e.d_flag.p;
  }
  if (c.fused && col_is_K) {
    // the flagged columns of A = the compact -K columns kept by the assembly, with the K correction of their 3 x 3
    // diagonal blocks: -(K - C_k + delta (1 - alpha))  (ref: 3076-3092 then 3194-3245)
    double *dC = c.wsd("mono.C", 3 * c.rows_loc + 2);
    BS_CUDA(cudaMemcpyAsync(dC, c.h_C.data(), 3 * c.rows_loc * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    scatter_flagged_columns(c, c.A, c.d_Kflag.p, c.ldk, c.d_kcol.p, dC, c.fused_alpha);
    BS_CUDA(cudaStreamSynchronize(c.stream));
  } else {
    DMat Vv = c.V;
    Vv.rows = c.rows_loc;
    Vv.cols = n;
    select_columns(c, c.A, Vv, c.K, d_flag, c.A_aliases_V);
  }
  {
    // the implicit V correction carries over to the V columns of A
    c.A.r1_u = c.V.r1_u;
    c.A.r1_w = c.V.r1_w;
    c.A.r1_ncols = c.V.r1_ncols;
    if (c.V.r1_u && col_is_K) {
      std::vector<double> wm(n);
      BS_CUDA(cudaMemcpyAsync(wm.data(), c.V.r1_w, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      BS_CUDA(cudaStreamSynchronize(c.stream));
      for (size_t p = 0; p < (size_t)c.N; ++p)
        for (int k = 0; k < 3; ++k)
          if (col_is_K[(size_t)c.node_of_pos[p] + (size_t)k * c.N]) wm[3 * p + k] = 0.0;
      c.d_r1wA.upload(wm, c.stream);
      BS_CUDA(cudaStreamSynchronize(c.stream));
      c.A.r1_w = c.d_r1wA.p;
    }
  }
  // rigid columns A(i, 3N+r) = -scaling * tmpN[r][i]  (ref: bem_stokes.cc:3247-3251)
  struct P2_ { double *p; } colbuf;
  colbuf.p = c.wsd("mono.col", c.rows_loc + 2);
  DMat Arows = c.A;
  Arows.rows = c.rows_loc;
  for (int r = 0; r < nr; ++r) {
    BS_CUDA(cudaMemcpyAsync(colbuf.p, &Y[(size_t)r * n + 3 * (size_t)c.p0], c.rows_loc * sizeof(double), cudaMemcpyHostToDevice,
                            c.stream));
    set_column(c, Arows, n + r, colbuf.p, -scaling);
    BS_CUDA(cudaStreamSynchronize(c.stream));
  }
  // rigid rows (ref: 3297-3339), stored after the last rank's node rows
  std::vector<double> rhs_int(n + nr, 0.0);
  if (use_shape && !c.torque_on)  // solve_with_torque: zero right-hand side on every node row (ref: 3190-3192)
    for (size_t i = 0; i < n; ++i) rhs_int[i] = Y[(size_t)nr * n + i];
  if (nr > 0) {
    std::vector<double> rows((size_t)nr * c.ld, 0.0);
    for (int r = 0; r < nr; ++r) {
      if (r >= num_rigid) {  // torque row
        rhs_int[n + r] = c.torque_rhs;
        for (size_t p = 0; p < (size_t)c.N; ++p)
          for (int k = 0; k < 3; ++k) rows[(size_t)r * c.ld + 3 * p + k] = scaling * dual(r)[c.node_of_pos[p] + (size_t)k * c.N];
      } else if (grid_type != BS_GRID_REAL) {
        rhs_int[n + r] = (r == imposed_component) ? 1.0 : 0.0;
        if (grid_type == BS_GRID_IMPOSED_VELOCITY) {
          rows[(size_t)r * c.ld + n + r] = scaling;
        } else {
          for (size_t p = 0; p < (size_t)c.N; ++p)
            for (int k = 0; k < 3; ++k)
              rows[(size_t)r * c.ld + 3 * p + k] = N_rigid_dual[(size_t)r * n + c.node_of_pos[p] + (size_t)k * c.N];
        }
      } else {
        for (size_t p = 0; p < (size_t)c.N; ++p)
          for (int k = 0; k < 3; ++k)
            rows[(size_t)r * c.ld + 3 * p + k] = scaling * N_rigid_dual[(size_t)r * n + c.node_of_pos[p] + (size_t)k * c.N];
      }
    }
    if (last)
      BS_CUDA(cudaMemcpyAsync(c.A.p + c.rows_loc * c.ld, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice,
                              c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
  }
  if (c.n_cons_owned) {  // constrained rows: 1 on the diagonal, -coefficients, nothing else; rhs 0 (ref: 3156-3183)
    DMat Arow = c.A;
    Arow.rows = c.rows_loc;
    apply_constraint_rows(c, Arow, n + nr);
    BS_CUDA(cudaStreamSynchronize(c.stream));
  }
  {
    const int N = c.N;
    for (size_t k = 0; k < c.cons_dof.size(); ++k) rhs_int[3 * (size_t)c.pos_of_node[c.cons_dof[k] % N] + c.cons_dof[k] / N] = 0.0;
  }
  if (rhs_out) {
    for (size_t p = 0; p < (size_t)c.N; ++p)
      for (int k = 0; k < 3; ++k) rhs_out[(size_t)c.node_of_pos[p] + (size_t)k * c.N] = rhs_int[3 * p + k];
    for (int r = 0; r < nr; ++r) rhs_out[n + r] = rhs_int[n + r];
  }
  c.has_rigid_rows = last && nr > 0;
  BS_API_END
}

int bs_matrix_size(bs_context *h, int which, int *rows, int *cols) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  const size_t m = c.full_vec_len(which);
  if (rows) *rows = (int)m;
  if (cols) *cols = (int)m;
  BS_API_END
}

int bs_vmult(bs_context *h, int which, const double *x, double *y) { return bs_vmult_multi(h, which, 1, x, y); }

int bs_vmult_multi(bs_context *h, int which, int nrhs, const double *X, double *Y) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(which >= 0 && which <= 2 && nrhs >= 1 && X && Y, "bad vmult arguments");
  const DMat &M = matrix_of(c, which);
  Extra &e = extra(c);
  const int nx = nextra_of(c, which);
  const size_t m = c.full_vec_len(which), mloc = c.local_vec_len(which);
  const size_t ldx = (m + 4) & ~(size_t)3;
  e.vin.alloc(std::max(e.vin.n, (size_t)nrhs * ldx));
  e.vout.alloc(std::max(e.vout.n, (size_t)nrhs * ldx));
  for (int k = 0; k < nrhs; ++k) to_internal(c, X + (size_t)k * m, nx, e.vin.p + (size_t)k * ldx);
  const size_t off = c.slice_offset(which);
  {
    cudaEventRecord(c.ev0, c.stream);
    if (nrhs == 1) gemv(c, M, e.vin.p, e.vout.p + off);
    else gemv_multi(c, M, nrhs, e.vin.p, ldx, e.vout.p + off, ldx);
    cudaEventRecord(c.ev1, c.stream);
  }
  for (int k = 0; k < nrhs; ++k) from_internal(c, e.vout.p + (size_t)k * ldx, nx, Y + (size_t)k * m, off, off + mloc);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  c.stats.vmult_ms_last = ms;
  BS_API_END
}

int bs_get_entries(bs_context *h, int which, int n, const int *rows, const int *cols, double *out) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  const DMat &M = matrix_of(c, which);
  BS_REQUIRE(n >= 0 && rows && cols && out, "bad arguments");
  const int N = c.N, n3 = 3 * N;
  std::vector<int> ir(n), ic(n);
  auto to_int = [&](int ref) {
    if (ref >= n3) return ref;
    return 3 * c.pos_of_node[ref % N] + ref / N;
  };
  for (int k = 0; k < n; ++k) {
    BS_REQUIRE(rows[k] >= 0 && rows[k] < (int)c.full_vec_len(which) && cols[k] >= 0 && cols[k] < (int)c.full_vec_len(which),
               "entry index out of range");
    const int r = to_int(rows[k]) - 3 * c.p0;
    BS_REQUIRE(r >= 0 && r < (int)M.rows, "row not owned by this rank");
    ir[k] = r;
    ic[k] = to_int(cols[k]);
  }
  Extra &e = extra(c);
  e.ir.upload(ir, c.stream);
  e.ic.upload(ic, c.stream);
  e.vout.alloc(std::max(e.vout.n, (size_t)n + 2));
  gather_entries(c, M, n, e.ir.p, e.ic.p, e.vout.p);
  BS_CUDA(cudaMemcpyAsync(out, e.vout.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  BS_CUDA(cudaStreamSynchronize(c.stream));
  BS_API_END
}

int bs_tangential_projector(bs_context *h, const double *in, double *out) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(c.have_projector, "projector data not set (bs_correct_V / bs_build_monolithic)");
  Extra &e = extra(c);
  const size_t n = c.n3();
  e.vin.alloc(std::max(e.vin.n, n + 2));
  to_internal(c, in, 0, e.vin.p);
  const double d = dot(c, c.d_Mnhat.p, e.vin.p, n);
  axpy(c, -d / c.l2gamma, c.d_nhat.p, e.vin.p, n);
  from_internal(c, e.vin.p, 0, out, 0, n);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  BS_API_END
}

int bs_precond_setup(bs_context *h, int which, int kind, int param) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  if (kind == BS_PREC_NONE) {
    c.prec_kind = kind;
    c.prec_which = which;
    return BS_OK;
  }
  const DMat &M = matrix_of(c, which);
  Timer t(c, c.stats.precond_setup_ms);
  c.prec_which = which;
  const size_t mloc = c.local_vec_len(which);
  const size_t off = c.slice_offset(which);
  if (kind == BS_PREC_JACOBI) {
    c.d_prec_diag.alloc(mloc + 2);
    extract_diag(c, M, off, c.d_prec_diag.p);
    std::vector<double> d(mloc);
    BS_CUDA(cudaMemcpyAsync(d.data(), c.d_prec_diag.p, mloc * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    for (auto &v : d) v = (v != 0.0) ? 1.0 / v : 1.0;
    BS_CUDA(cudaMemcpyAsync(c.d_prec_diag.p, d.data(), mloc * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
  } else if (kind == BS_PREC_DIRECT || kind == BS_PREC_BLOCK_DIRECT || kind == BS_PREC_BAND) {
    // BLOCK_DIRECT: the node rows of this rank (rigid unknowns pass through), optionally cut into diagonal blocks of at
    // most `param` rows; DIRECT / BAND: the whole matrix (one GPU).  ref: DirectPreconditioner::initialize
    // (source/direct_preconditioner.cc:10-23), band copy assemble_monolithic_preconditioner (bem_stokes.cc:3437-3475)
    size_t nb, max_block = 0;
    int band = 0;
    if (kind == BS_PREC_BLOCK_DIRECT) {
      nb = c.rows_loc;
      max_block = param > 0 ? (size_t)param : 0;
    } else {
      BS_REQUIRE(c.nranks == 1, "full direct / band preconditioner needs the whole matrix on one GPU; use BS_PREC_BLOCK_DIRECT");
      nb = mloc;
      if (kind == BS_PREC_BAND) {
        BS_REQUIRE(param > 0, "band preconditioner needs a positive bandwidth");
        band = param;
      }
    }
    precond_factor_blocks(c, M, off, nb, max_block, band);
    BS_CUDA(cudaStreamSynchronize(c.stream));
  } else {
    BS_REQUIRE(kind == BS_PREC_NONE, "unknown preconditioner kind");
  }
  c.prec_kind = kind;
  BS_API_END
}

int bs_precond_vmult(bs_context *h, const double *x, double *y) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  Extra &e = extra(c);
  const int which = c.prec_which;
  const int nx = nextra_of(c, which);
  const size_t m = c.full_vec_len(which), mloc = c.local_vec_len(which), off = c.slice_offset(which);
  e.vin.alloc(std::max(e.vin.n, m + 2));
  e.vout.alloc(std::max(e.vout.n, m + 2));
  to_internal(c, x, nx, e.vin.p);
  apply_precond(c, e.vin.p + off, e.vout.p + off);
  from_internal(c, e.vout.p, nx, y, off, off + mloc);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  BS_API_END
}

int bs_gmres(bs_context *h, int which, const double *b, double *x, double tol_abs, int max_steps, int max_n_tmp_vectors,
             int *iterations, double *final_residual) {
  int rc = BS_OK;
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(b && x, "null vectors");
  matrix_of(c, which);
  BS_REQUIRE(c.prec_kind == BS_PREC_NONE || c.prec_which == which, "preconditioner was set up for another matrix");
  Timer t(c, c.stats.solve_ms, "bs_gmres");
  Extra &e = extra(c);
  const int nx = nextra_of(c, which);
  const size_t m = c.full_vec_len(which), mloc = c.local_vec_len(which), off = c.slice_offset(which);
  struct P_ { double *p; } db, dx;
  db.p = c.wsd("api.b", m + 2);
  dx.p = c.wsd("api.x", m + 2);
  to_internal(c, b, nx, db.p);
  to_internal(c, x, nx, dx.p);
  (void)e;
  rc = gmres(c, which, db.p + off, dx.p + off, tol_abs, max_steps, max_n_tmp_vectors, iterations, final_residual);
  from_internal(c, dx.p, nx, x, off, off + mloc);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  if (rc == BS_ERR_NOT_CONVERGED) bs::set_last_error("GMRES did not converge within max_steps");
  }
  catch (const bs::Error &e) {
    bs::set_last_error(e.what());
    return e.code;
  }
  catch (const std::exception &e) {
    bs::set_last_error(e.what());
    return BS_ERR_INVALID;
  }
  return rc;
}

int bs_set_gmres_orthogonalization(bs_context *h, int kind) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(kind == BS_ORTHO_CGS2 || kind == BS_ORTHO_MGS, "unknown orthogonalisation");
  c.gmres_ortho = kind;
  BS_API_END
}

int bs_gmres_multi(bs_context *h, int which, int nrhs, const double *B, double *X, double tol_abs, int max_steps,
                   int max_n_tmp_vectors, int *iterations, double *final_residuals) {
  int rc = BS_OK;
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(B && X && nrhs >= 1, "bad arguments");
  matrix_of(c, which);
  BS_REQUIRE(c.prec_kind == BS_PREC_NONE || c.prec_which == which, "preconditioner was set up for another matrix");
  Timer t(c, c.stats.solve_ms);
  const int nx = nextra_of(c, which);
  const size_t m = c.full_vec_len(which), mloc = c.local_vec_len(which), off = c.slice_offset(which);
  const size_t ldv = (m + 3) & ~(size_t)1;
  struct P_ { double *p; } db, dx;
  db.p = c.wsd("api.b", (size_t)nrhs * ldv);
  dx.p = c.wsd("api.x", (size_t)nrhs * ldv);
  for (int k = 0; k < nrhs; ++k) {
    to_internal(c, B + (size_t)k * m, nx, db.p + (size_t)k * ldv);
    to_internal(c, X + (size_t)k * m, nx, dx.p + (size_t)k * ldv);
  }
  rc = gmres_batched(c, which, nrhs, db.p + off, dx.p + off, ldv, tol_abs, max_steps, max_n_tmp_vectors, iterations,
                     final_residuals);
  for (int k = 0; k < nrhs; ++k) from_internal(c, dx.p + (size_t)k * ldv, nx, X + (size_t)k * m, off, off + mloc);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  if (rc == BS_ERR_NOT_CONVERGED) bs::set_last_error("GMRES did not converge within max_steps for at least one right-hand side");
  }
  catch (const bs::Error &e) {
    bs::set_last_error(e.what());
    return e.code;
  }
  catch (const std::exception &e) {
    bs::set_last_error(e.what());
    return BS_ERR_INVALID;
  }
  return rc;
}

int bs_direct_solve(bs_context *h, int which, const double *b, double *x) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(c.nranks == 1, "bs_direct_solve needs the whole matrix on one GPU");
  const DMat &M = matrix_of(c, which);
  Timer t(c, c.stats.solve_ms);
  const int nx = nextra_of(c, which);
  const size_t m = c.full_vec_len(which);
  struct P_ { double *p; } lu, dx;
  struct PI_ { int *p; } piv;
  const size_t ldl = (m + 15) & ~(size_t)15;  // 128-byte aligned rows: the trailing update runs on the tensor path
  lu.p = c.wsd("direct.lu", m * ldl + 2);
  piv.p = c.wsi("direct.piv", m + 2);
  dx.p = c.wsd("api.x", m + 2);
  BS_CUDA(cudaMemsetAsync(lu.p, 0, m * ldl * sizeof(double), c.stream));
  BS_CUDA(cudaMemcpy2DAsync(lu.p, ldl * sizeof(double), M.p, M.ld * sizeof(double), m * sizeof(double), m,
                            cudaMemcpyDeviceToDevice, c.stream));
  add_rank1_block(c, M, 0, 0, m, lu.p, ldl);
  lu_factor(c, lu.p, m, ldl, piv.p);
  to_internal(c, b, nx, dx.p);
  lu_solve(c, lu.p, m, ldl, piv.p, dx.p);
  from_internal(c, dx.p, nx, x, 0, m);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  BS_API_END
}

int bs_kernel_eval(int device, int type, double eps, int o, int npts, const double *p, const double *pim, double *G,
                   double *W) {
  BS_API_BEGIN
  BS_REQUIRE(npts > 0 && p, "bad arguments");
  BS_REQUIRE(type >= 0 && type <= 2 && o >= 0 && o < 3, "bad kernel type / orientation");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw Error(BS_ERR_NO_DEVICE, "no CUDA device visible: libbemstokes_b200 has no CPU fallback");
  BS_CUDA(cudaSetDevice(device));
  DBuf<double> dp, dq, dG, dW;
  dp.upload(p, (size_t)3 * npts, 0);
  if (pim) dq.upload(pim, (size_t)3 * npts, 0);
  if (G) dG.alloc((size_t)9 * npts);
  if (W) dW.alloc((size_t)27 * npts);
  kernel_eval_device(type, eps, o, npts, dp.p, pim ? dq.p : nullptr, dG.p, dW.p, 0);
  if (G) BS_CUDA(cudaMemcpy(G, dG.p, sizeof(double) * 9 * npts, cudaMemcpyDeviceToHost));
  if (W) BS_CUDA(cudaMemcpy(W, dW.p, sizeof(double) * 27 * npts, cudaMemcpyDeviceToHost));
  BS_CUDA(cudaDeviceSynchronize());
  BS_API_END
}

int bs_evaluate_bie(bs_context *h, int npts, const double *points, const double *vel, const double *forces, double *out,
                    int on_boundary) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(npts > 0 && points && vel && forces && out, "bad arguments");
  struct P_ { double *p; } dp, dv, df, dout;
  dp.p = c.wsd("eval.pts", (size_t)3 * npts);
  BS_CUDA(cudaMemcpyAsync(dp.p, points, sizeof(double) * 3 * npts, cudaMemcpyHostToDevice, c.stream));
  dv.p = c.wsd("eval.vel", c.n3() + 2);
  df.p = c.wsd("eval.frc", c.n3() + 2);
  dout.p = c.wsd("eval.out", (size_t)3 * npts);
  to_internal(c, vel, 0, dv.p, true);
  to_internal(c, forces, 0, df.p, true);
  if (on_boundary)  // the reference's on-boundary routine accumulates into val_velocities (bem_stokes.cc:5543-5550)
    BS_CUDA(cudaMemcpyAsync(dout.p, out, sizeof(double) * 3 * npts, cudaMemcpyHostToDevice, c.stream));
  evaluate_bie(c, npts, dp.p, dv.p, df.p, dout.p, on_boundary != 0, on_boundary != 0);
  BS_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * 3 * npts, cudaMemcpyDeviceToHost, c.stream));
  BS_CUDA(cudaStreamSynchronize(c.stream));
  BS_API_END
}

int bs_host_prepass(int fe_degree, int map_degree, int n_map_nodes, const double *euler_vec, int ncell, const int *conn_map,
                    int n_nodes, const int *conn_stokes, int quad_order, const double *pole, double *nhat, double *Mnhat,
                    double *l2gamma, double *N_rigid, double *N_rigid_dual, double *area, double *support_points) {
  BS_API_BEGIN
  BS_REQUIRE(euler_vec && conn_map && conn_stokes && nhat && Mnhat, "bad arguments");
  const double origin[3] = {0, 0, 0};
  host_prepass(fe_degree, map_degree, n_map_nodes, euler_vec, ncell, conn_map, n_nodes, conn_stokes, quad_order,
               pole ? pole : origin, nhat, Mnhat, l2gamma, N_rigid, N_rigid_dual, area, support_points);
  BS_API_END
}

int bs_host_cell_blocks(int n_nodes, const double *nodes, int ncell, const int *conn, int kernel_type, int n1d,
                        int *sizes_out, int *cell_ptr, int *cells, unsigned *sync_mask, int *block_nodes,
                        unsigned char *first_touch, int *colour_start, int *pos_of_node) {
  BS_API_BEGIN
  BS_REQUIRE(n_nodes > 0 && ncell > 0 && nodes && conn && sizes_out && n1d >= 1, "bad arguments");
  BS_REQUIRE(kernel_type >= 0 && kernel_type <= 2, "unknown kernel type");
  Context c;  // host members only: nothing below touches the device
  c.N = c.Nmap = n_nodes;
  c.ncell = ncell;
  c.na = c.na_map = 4;
  c.map_nodes.assign(nodes, nodes + (size_t)3 * n_nodes);
  c.conn.assign(conn, conn + (size_t)4 * ncell);
  c.conn_map = c.conn;
  c.kp.type = kernel_type;
  c.kp.eps = 0.0;
  c.x1d.assign(n1d, 0.0);
  c.nq = n1d * n1d;
  c.nq_pad = (c.nq + 1) & ~1;
  compute_node_order(c);
  build_cell_blocks(c);
  const ColumnBlocks &B = c.blocks;
  sizes_out[0] = B.nblocks;
  sizes_out[1] = B.tj;
  sizes_out[2] = B.cs;
  sizes_out[3] = (int)B.colour_start.size() - 1;
  sizes_out[4] = (int)B.cells.size();
  sizes_out[5] = B.unpaired;
  BS_REQUIRE(B.tj <= 32 && B.colour_start.size() <= 65, "tiling exceeds the documented output sizes");
  if (cell_ptr) std::copy(B.cell_ptr.begin(), B.cell_ptr.end(), cell_ptr);
  if (cells) std::copy(B.cells.begin(), B.cells.end(), cells);
  if (sync_mask) std::copy(B.sync.begin(), B.sync.end(), sync_mask);
  if (block_nodes) std::copy(B.nodes.begin(), B.nodes.end(), block_nodes);
  if (first_touch) std::copy(B.first.begin(), B.first.end(), first_touch);
  if (colour_start) std::copy(B.colour_start.begin(), B.colour_start.end(), colour_start);
  if (pos_of_node) std::copy(c.pos_of_node.begin(), c.pos_of_node.end(), pos_of_node);
  BS_API_END
}

int bs_set_comm(bs_context *h, bs_allgatherv_fn ag, bs_allreduce_sum_fn ar, void *user) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  c.cb_allgatherv = ag;
  c.cb_allreduce = ar;
  c.cb_user = user;
  BS_API_END
}

int bs_exchange_export(bs_context *h, size_t max_vec_len, unsigned char *handles_out) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(handles_out != nullptr && max_vec_len > 0, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  BS_REQUIRE(c.nranks <= BS_MAX_RANKS, "the peer exchange supports at most 32 ranks");
  c.xchg_ld = (max_vec_len + 17) & ~(size_t)15;
  c.red_off = Context::XCHG_SLOTS * c.xchg_ld;  // then [2 parities][nranks][BS_RED_CAP] partial sums of the peers
  c.d_xchg.alloc(c.red_off + 2 * (size_t)c.nranks * BS_RED_CAP);
  c.d_xchg.zero(c.stream);
  c.d_flags.alloc(BS_FLAG_WORDS);
  c.d_flags.zero(c.stream);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  cudaIpcMemHandle_t hx, hf;
  BS_CUDA(cudaIpcGetMemHandle(&hx, c.d_xchg.p));
  BS_CUDA(cudaIpcGetMemHandle(&hf, c.d_flags.p));
  std::memcpy(handles_out, &hx, 64);
  std::memcpy(handles_out + 64, &hf, 64);
  BS_API_END
}

int bs_exchange_import(bs_context *h, int nranks, const unsigned char *all) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(nranks == c.nranks && all != nullptr, "nranks mismatch");
  BS_REQUIRE(nranks <= BS_MAX_RANKS, "the peer exchange supports at most 32 ranks");
  BS_REQUIRE(c.d_xchg.p != nullptr, "call bs_exchange_export first");
  c.peer_xbuf.assign(nranks, nullptr);
  c.peer_flags.assign(nranks, nullptr);
  for (int r = 0; r < nranks; ++r) {
    if (r == c.rank) {
      c.peer_xbuf[r] = c.d_xchg.p;
      c.peer_flags[r] = c.d_flags.p;
      continue;
    }
    cudaIpcMemHandle_t hx, hf;
    std::memcpy(&hx, all + (size_t)r * BS_IPC_EXPORT_BYTES, 64);
    std::memcpy(&hf, all + (size_t)r * BS_IPC_EXPORT_BYTES + 64, 64);
    BS_CUDA(cudaIpcOpenMemHandle(&c.peer_xbuf[r], hx, cudaIpcMemLazyEnablePeerAccess));
    BS_CUDA(cudaIpcOpenMemHandle(&c.peer_flags[r], hf, cudaIpcMemLazyEnablePeerAccess));
  }
  std::vector<double *> px(nranks);
  std::vector<unsigned long long *> pf(nranks);
  for (int r = 0; r < nranks; ++r) {
    px[r] = (double *)c.peer_xbuf[r];
    pf[r] = (unsigned long long *)c.peer_flags[r];
  }
  c.d_peer_xbuf.upload(px, c.stream);
  c.d_peer_flags.upload(pf, c.stream);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  c.epoch = 0;
  c.p2p = true;
  BS_API_END
}

int bs_get_stats(bs_context *h, bs_stats *out) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(out != nullptr, "null output");
  c.stats.n_cell_blocks = c.blocks.nblocks;
  c.stats.n_colours = c.blocks.colour_start.empty() ? 0 : (long long)c.blocks.colour_start.size() - 1;
  c.stats.node_touch_ratio = c.blocks.node_touch_ratio;
  c.stats.cell_sets = c.blocks.cs;
  c.stats.cell_steps = c.blocks.steps;
  c.stats.unpaired_cells = c.blocks.unpaired;
  c.stats.sync_steps = c.blocks.sync_steps;
  *out = c.stats;
  BS_API_END
}
int bs_reset_stats(bs_context *h) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  c.stats = bs_stats{};
  BS_API_END
}

int bs_bench_vmult_multi(bs_context *h, int which, int nrhs, int repeats, double *ms_per_call) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  const DMat &M = which == BS_MAT_V ? c.V : (which == BS_MAT_K ? c.K : c.A);
  BS_REQUIRE(M.valid() && repeats > 0 && nrhs >= 1, "matrix not available");
  Extra &e = extra(c);
  const size_t m = c.full_vec_len(which);
  const size_t ldx = (m + 4) & ~(size_t)3;
  e.vin.alloc(std::max(e.vin.n, (size_t)nrhs * ldx));
  e.vout.alloc(std::max(e.vout.n, (size_t)nrhs * ldx));
  fill(c, e.vin.p, 1.0, (size_t)nrhs * ldx);
  gemv_multi(c, M, nrhs, e.vin.p, ldx, e.vout.p, ldx);  // warm-up
  cudaEventRecord(c.ev0, c.stream);
  for (int i = 0; i < repeats; ++i) gemv_multi(c, M, nrhs, e.vin.p, ldx, e.vout.p, ldx);
  cudaEventRecord(c.ev1, c.stream);
  BS_CUDA(cudaEventSynchronize(c.ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  if (ms_per_call) *ms_per_call = ms / repeats;
  BS_API_END
}

int bs_bench_vmult(bs_context *h, int which, int repeats, double *ms_per_call) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  const DMat &M = which == BS_MAT_V ? c.V : (which == BS_MAT_K ? c.K : c.A);
  BS_REQUIRE(M.valid() && repeats > 0, "matrix not available");
  Extra &e = extra(c);
  const size_t m = c.full_vec_len(which);
  e.vin.alloc(std::max(e.vin.n, m + 2));
  e.vout.alloc(std::max(e.vout.n, m + 2));
  fill(c, e.vin.p, 1.0, m);
  gemv(c, M, e.vin.p, e.vout.p);  // warm-up
  cudaEventRecord(c.ev0, c.stream);
  for (int i = 0; i < repeats; ++i) gemv(c, M, e.vin.p, e.vout.p);
  cudaEventRecord(c.ev1, c.stream);
  BS_CUDA(cudaEventSynchronize(c.ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  if (ms_per_call) *ms_per_call = ms / repeats;
  BS_API_END
}

// LU micro-benchmark on a synthetic matrix (uniform entries in [-1,1), n on the diagonal): factorisation and the
// cooperative application timed with CUDA events on the context stream
__global__ void k_fill_test_matrix(double *A, size_t ld, int n) {
  const size_t i = blockIdx.y, j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n || j >= (size_t)n) return;
  unsigned long long h = (i * 0x9E3779B97F4A7C15ull) ^ (j * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull);
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  A[i * ld + j] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0 + (i == j ? (double)n : 0.0);
}
int bs_bench_lu(bs_context *h, int n, int apply_repeats, double *factor_ms, double *apply_ms, double *residual) {
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(n > 0 && apply_repeats > 0, "bad arguments");
  const size_t ld = ((size_t)n + 15) & ~(size_t)15;
  double *A = c.wsd("benchlu.A", (size_t)n * ld + 2);
  BS_CUDA(cudaMemsetAsync(A, 0, (size_t)n * ld * sizeof(double), c.stream));
  k_fill_test_matrix<<<dim3((n + 255) / 256, n), 256, 0, c.stream>>>(A, ld, n);
  DMat M;
  M.p = A, M.rows = n, M.cols = n, M.ld = ld;
  std::vector<Context::LuBlock> keep;
  keep.swap(c.lu_blocks);  // the benchmark must not disturb a preconditioner that is set up (buffers are re-used: set up again afterwards)
  cudaEventRecord(c.ev0, c.stream);
  precond_factor_blocks(c, M, 0, (size_t)n, 0, 0);
  cudaEventRecord(c.ev1, c.stream);
  BS_CUDA(cudaEventSynchronize(c.ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  if (factor_ms) *factor_ms = ms;
  double *x = c.wsd("benchlu.x", (size_t)n + 2), *y = c.wsd("benchlu.y", (size_t)n + 2), *r = c.wsd("benchlu.r", (size_t)n + 2);
  fill(c, x, 1.0, n);
  lu_apply_fast(c, c.lu_blocks[0], x, y, nullptr);
  cudaEventRecord(c.ev0, c.stream);
  for (int i = 0; i < apply_repeats; ++i) lu_apply_fast(c, c.lu_blocks[0], x, y, nullptr);
  cudaEventRecord(c.ev1, c.stream);
  BS_CUDA(cudaEventSynchronize(c.ev1));
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  if (apply_ms) *apply_ms = ms / apply_repeats;
  if (residual) {  // |A y - x|_inf / |x|_inf with the matrix regenerated (the factorisation ran in place on a copy)
    k_fill_test_matrix<<<dim3((n + 255) / 256, n), 256, 0, c.stream>>>(A, ld, n);
    gemv(c, M, y, r);
    std::vector<double> hr(n);
    BS_CUDA(cudaMemcpyAsync(hr.data(), r, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    double e = 0;
    for (double v : hr) e = std::max(e, std::fabs(v - 1.0));
    *residual = e;
  }
  c.lu_blocks.swap(keep);
  if (!c.lu_blocks.empty()) c.prec_kind = BS_PREC_NONE;  // its buffers were overwritten
  c.lu_blocks.clear();
  BS_API_END
}

}  // extern "C"

// ---- DN-operator route (ref: dirichlet_to_neumann_operator bem_stokes.cc:4072-4129, solve_system(false) 4163-4258) ----
// F_k = P V^-1 P K P u_k for nvec input velocities at once: one projection kernel, one multi-right-hand-side sweep over K
// (FP64 tensor path from 4 vectors), nvec V-solves advanced in lockstep (one sweep over V per iteration for all of them,
// or one LU for all with solve_directly), one projection.
__global__ void k_project_apply(double *X, size_t ldx, int nvec, const double *dots, const double *nhat, double inv_l2, size_t n) {
  const int k = blockIdx.y;
  const double d = dots[k] * inv_l2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    X[(size_t)k * ldx + i] -= d * nhat[i];
}
namespace bs {
// X_k <- P X_k = X_k - (Mnhat . X_k) / l2 * nhat for nvec full internal vectors
static void project_multi(Context &c, double *X, size_t ldx, int nvec) {
  const size_t n = c.n3();
  double *dots = c.wsd("dn.dots", 16);
  multi_dot(c, X, ldx, nvec, c.d_Mnhat.p, n, dots);
  k_project_apply<<<dim3((unsigned)std::min<size_t>((n + 255) / 256, 592), nvec), 256, 0, c.stream>>>(X, ldx, nvec, dots, c.d_nhat.p,
                                                                                                  1.0 / c.l2gamma, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
}  // namespace bs

extern "C" int bs_dn_operator_multi(bs_context *h, int nvec, const double *U, double *F, int solve_directly, double tol_abs,
                                    int max_steps, int max_n_tmp_vectors, int *iterations) {
  int rc = BS_OK;
  BS_API_BEGIN
  Context &c = ctx_of(h);
  BS_REQUIRE(nvec >= 1 && nvec <= Context::XCHG_SLOTS && U && F, "1 <= nvec <= 8 input velocities");
  BS_REQUIRE(c.have_projector, "projector data not set (bs_correct_V / bs_build_monolithic)");
  const DMat &K = matrix_of(c, BS_MAT_K);
  const DMat &V = matrix_of(c, BS_MAT_V);
  BS_REQUIRE(!solve_directly || c.nranks == 1, "solve_directly needs the whole matrix on one GPU");
  BS_REQUIRE(c.prec_kind == BS_PREC_NONE || c.prec_which == BS_MAT_V, "preconditioner was set up for another matrix");
  Timer t(c, c.stats.solve_ms, "bs_dn_operator_multi");
  const size_t n = c.n3(), off = c.slice_offset(BS_MAT_V), mloc = c.rows_loc;
  const size_t ldx = (n + 4) & ~(size_t)3;
  double *X = c.wsd("dn.x", (size_t)nvec * ldx), *Y = c.wsd("dn.y", (size_t)nvec * ldx), *Fv = c.wsd("dn.f", (size_t)nvec * ldx);
  BS_CUDA(cudaMemsetAsync(Y, 0, (size_t)nvec * ldx * sizeof(double), c.stream));
  BS_CUDA(cudaMemsetAsync(Fv, 0, (size_t)nvec * ldx * sizeof(double), c.stream));
  for (int k = 0; k < nvec; ++k) to_internal(c, U + (size_t)k * n, 0, X + (size_t)k * ldx);
  project_multi(c, X, ldx, nvec);                                       // P u
  gemv_multi(c, K, nvec, X, ldx, Y + off, ldx);                         // K P u on the owned rows
  if (c.nranks > 1)
    for (int k = 0; k < nvec; ++k) exchange(c, BS_MAT_K, Y + (size_t)k * ldx + off, Y + (size_t)k * ldx);
  project_multi(c, Y, ldx, nvec);                                       // P K P u
  if (solve_directly) {
    const size_t ldl = (n + 15) & ~(size_t)15;
    double *lu = c.wsd("direct.lu", n * ldl + 2);
    int *piv = c.wsi("direct.piv", n + 2);
    BS_CUDA(cudaMemsetAsync(lu, 0, n * ldl * sizeof(double), c.stream));
    BS_CUDA(cudaMemcpy2DAsync(lu, ldl * sizeof(double), V.p, V.ld * sizeof(double), n * sizeof(double), n, cudaMemcpyDeviceToDevice, c.stream));
    add_rank1_block(c, V, 0, 0, n, lu, ldl);
    lu_factor(c, lu, n, ldl, piv);
    for (int k = 0; k < nvec; ++k) {
      copy(c, Y + (size_t)k * ldx, Fv + (size_t)k * ldx, n);
      lu_solve(c, lu, n, ldl, piv, Fv + (size_t)k * ldx);
    }
    if (iterations)
      for (int k = 0; k < nvec; ++k) iterations[k] = 1;
  } else {
    std::vector<int> its(nvec, 0);
    rc = gmres_batched(c, BS_MAT_V, nvec, Y + off, Fv + off, ldx, tol_abs, max_steps, max_n_tmp_vectors, its.data(), nullptr);
    if (iterations) std::copy(its.begin(), its.end(), iterations);
    if (c.nranks > 1)
      for (int k = 0; k < nvec; ++k) exchange(c, BS_MAT_V, Fv + (size_t)k * ldx + off, Fv + (size_t)k * ldx);
  }
  (void)mloc;
  project_multi(c, Fv, ldx, nvec);                                      // P V^-1 P K P u
  for (int k = 0; k < nvec; ++k) from_internal(c, Fv + (size_t)k * ldx, 0, F + (size_t)k * n, 0, n);
  BS_CUDA(cudaStreamSynchronize(c.stream));
  if (rc == BS_ERR_NOT_CONVERGED) bs::set_last_error("GMRES did not converge within max_steps for at least one DN right-hand side");
  }
  catch (const bs::Error &e) {
    bs::set_last_error(e.what());
    return e.code;
  }
  catch (const std::exception &e) {
    bs::set_last_error(e.what());
    return BS_ERR_INVALID;
  }
  return rc;
}

// ---- FP64 FMA-chain microbenchmark: the denominator of the assembly roofline (not in MEASURED_PEAKS.json) ----
__global__ void k_fp64_peak(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, b = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
    a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int bs_bench_fp64_sustained(int device, double seconds, double *tflops) {
  BS_API_BEGIN
  BS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BS_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 15;
  DBuf<double> out;
  out.alloc((size_t)blocks * threads);
  cudaEvent_t e0, e1;
  BS_CUDA(cudaEventCreate(&e0));
  BS_CUDA(cudaEventCreate(&e1));
  // back-to-back launches for `seconds` (power-capped clocks), rate of the second half
  k_fp64_peak<<<blocks, threads>>>(out.p, iters);
  BS_CUDA(cudaDeviceSynchronize());
  const double fl = 2.0 * 8 * (double)iters * blocks * threads;
  cudaEventRecord(e0);
  k_fp64_peak<<<blocks, threads>>>(out.p, iters);
  cudaEventRecord(e1);
  BS_CUDA(cudaEventSynchronize(e1));
  float ms1 = 0;
  cudaEventElapsedTime(&ms1, e0, e1);
  const int n = std::max(4, (int)(seconds * 1e3 / std::max(ms1, 0.01f)));
  for (int i = 0; i < n / 2; ++i) k_fp64_peak<<<blocks, threads>>>(out.p, iters);
  cudaEventRecord(e0);
  for (int i = 0; i < n / 2; ++i) k_fp64_peak<<<blocks, threads>>>(out.p, iters);
  cudaEventRecord(e1);
  BS_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (tflops) *tflops = fl * (n / 2) / (ms * 1e-3) / 1e12;
  BS_API_END
}

extern "C" int bs_bench_fp64_peak(int device, double *tflops) {
  BS_API_BEGIN
  BS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BS_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 15;
  DBuf<double> out;
  out.alloc((size_t)blocks * threads);
  cudaEvent_t e0, e1;
  BS_CUDA(cudaEventCreate(&e0));
  BS_CUDA(cudaEventCreate(&e1));
  k_fp64_peak<<<blocks, threads>>>(out.p, 1 << 10);
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_fp64_peak<<<blocks, threads>>>(out.p, iters);
    cudaEventRecord(e1);
    BS_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * (double)iters * blocks * threads;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (tflops) *tflops = best;
  BS_API_END
}
