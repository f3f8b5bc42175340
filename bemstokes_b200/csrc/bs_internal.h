// Internal declarations of libbemstokes_b200 (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include "../../include/bemstokes_b200.h"

namespace bs {

// ---------------------------------------------------------------------------------------------------------
// errors: C++ exceptions inside, status codes at the ABI
// ---------------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string &m);

#define BS_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      throw bs::Error(BS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " at " + __FILE__ + \
                                       ":" + std::to_string(__LINE__));                                 \
  } while (0)
#define BS_REQUIRE(cond, msg)                                          \
  do {                                                                 \
    if (!(cond)) throw bs::Error(BS_ERR_INVALID, std::string(msg));    \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// device buffer
// ---------------------------------------------------------------------------------------------------------
// Grow-only: cudaMalloc / cudaFree are expensive (hundreds of ms once NCCL has mapped peers), so a buffer is
// reallocated only when it has to grow and nothing on a hot path allocates.
template <class T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;    // logical size
  size_t cap = 0;  // allocated size
  DBuf() = default;
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = cap = 0;
  }
  void alloc(size_t count) {
    if (count <= cap && p) {
      n = count;
      return;
    }
    release();
    if (count) BS_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    n = cap = count;
  }
  void upload(const std::vector<T> &h, cudaStream_t s) {
    alloc(h.size());
    if (h.size()) BS_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const T *h, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) BS_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) {
    if (n) BS_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// Row-block of a dense matrix, row-major, leading dimension ld (doubles), internal (node-major) ordering.
// Optional implicit rank-1 term: the matrix stands for  M + r1_u r1_w^T  on the columns < r1_ncols (the V correction of
// ref bem_stokes.cc:3017-3032 kept as two vectors instead of a read-modify-write pass over the whole matrix; every
// reader below - matvec, entry reads, diagonal, LU copies - applies it).
struct DMat {
  double *p = nullptr;
  size_t rows = 0, cols = 0, ld = 0;
  bool owned = false;
  const double *r1_u = nullptr;  // [rows] (zero on the rigid rows)
  const double *r1_w = nullptr;  // [r1_ncols]
  size_t r1_ncols = 0;
  bool valid() const { return p != nullptr; }
};

// ---------------------------------------------------------------------------------------------------------
// host-side rule / shape tables
// ---------------------------------------------------------------------------------------------------------
struct Rule2D {
  std::vector<double> xi;  // [nq][2]
  std::vector<double> w;   // [nq]
  int size() const { return (int)w.size(); }
};
void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w);
Rule2D tensor_rule(const std::vector<double> &x1, const std::vector<double> &w1);
Rule2D make_singular_rule(int kind, int order, int fe_degree, int local_index);
int n_shape(int degree);
void unit_support_point(int degree, int a, double &sx, double &sy);
// phi[a], dphi[a][2] of scalar FE_Q(degree) at (x,y), deal.II dof order
void shape_eval(int degree, double x, double y, double *phi, double *dphi);

// ---------------------------------------------------------------------------------------------------------
// Green-kernel parameters (device-visible POD)
// ---------------------------------------------------------------------------------------------------------
struct KernelParams {
  int type = BS_KERNEL_FREE;
  double eps = 0.0;
  int o = 1;             // wall orientation
  double wall_pos = 0.0; // wall_positions[0][o]
};

// ---------------------------------------------------------------------------------------------------------
// assembly tiling constants
// ---------------------------------------------------------------------------------------------------------
#ifndef BS_TI
#define BS_TI 64
#endif
#ifndef BS_CTAS_PER_SM
#define BS_CTAS_PER_SM 2
#endif
constexpr int TI = BS_TI;                    // collocation (row) nodes per CTA of the regular assembly pass
constexpr int CTAS_PER_SM = BS_CTAS_PER_SM;  // resident CTAs per SM the shared-memory tile is sized for
constexpr int MAX_NA = 9;     // Q2
constexpr int MAX_RIGID = 7;

// flag words of the peer exchange (own copy in d_flags, written by the peers through the IPC mapping)
constexpr int BS_MAX_RANKS = 32;
constexpr int BS_FLAG_XCHG = 0;    // [32] arrival epoch of every source rank's Krylov slice
constexpr int BS_FLAG_RED = 32;    // [32] arrival sequence number of every source rank's partial sums
constexpr int BS_FLAG_ERR = 64;    // non-zero: a wait timed out (a rank stopped answering)
constexpr int BS_FLAG_EPOCH = 65;  // this rank's exchange epoch (device-resident counter)
constexpr int BS_FLAG_SEQ = 66;    // this rank's reduction sequence number
constexpr int BS_FLAG_WORDS = 128;
constexpr size_t BS_RED_CAP = 8192;  // doubles per (parity, source rank) of the cross-rank reduction area

struct ColumnBlocks {
  int tj = 0;                       // column nodes per block
  int nblocks = 0;
  std::vector<int> cell_ptr;        // [nblocks+1]
  std::vector<int> cells;           // concatenated cell ids
  std::vector<signed char> slots;   // [ncells_total][na] slot of the cell's node in its block
  std::vector<int> nodes;           // [nblocks][tj] node position of every slot (-1 = unused)
  std::vector<unsigned char> first; // [nblocks][tj] 1 = this block's colour is the first to touch the node column
  std::vector<int> colour_start;    // [ncolours+1] block ranges per colour (blocks are sorted by colour)
  int max_cells = 0;
  int cs = 1;                       // cell sets: with cs == 2 the cell list of a block is a sequence of pairs of cells that
                                    // share no node (-1 = no partner), integrated concurrently by the two thread sets of K1
  int unpaired = 0;                 // cells without a partner (cs == 2)
  std::vector<unsigned> sync;       // [nblocks] cs == 2: bit s set = a cell of step s shares a node with the other set's cell
                                    // of step s-1 (the thread sets may be one step apart): CTA barrier before step s
  long long sync_steps = 0, steps = 0;
  double node_touch_ratio = 0;      // sum over blocks of touched nodes / N  (tile traffic amplification)
};

struct Context {
  int device = 0;
  int fe_degree = 1, map_degree = 1;
  int na = 4, na_map = 4;
  int pointer_mode = BS_PTR_HOST;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;

  // partition
  int rank = 0, nranks = 1;
  std::vector<int> owner_in;        // user supplied owner_of_node (may be empty)
  // geometry (host copies)
  int N = 0, Nmap = 0, ncell = 0;
  std::vector<double> map_nodes;    // [Nmap][3]
  std::vector<int> conn_map, conn;  // original ids
  std::vector<int> material;
  std::vector<double> support;      // [N][3] original order
  std::vector<int> pos_of_node;     // original node id -> internal position
  std::vector<int> node_of_pos;     // internal position -> original node id
  std::vector<int> part_start;      // [nranks+1] internal position ranges
  int p0 = 0, p1 = 0;               // my node-position range
  bool have_geometry = false;

  // quadrature
  std::vector<double> x1d, w1d;
  Rule2D reg;
  std::vector<Rule2D> sing;         // per scalar local index
  bool have_quadrature = false, have_singular = false;
  KernelParams kp;

  // device geometry
  DBuf<double> d_support;           // [N][3] internal order
  DBuf<double> d_map_nodes;         // [Nmap][3]
  DBuf<int> d_conn_pos;             // [ncell][na] internal positions
  DBuf<int> d_conn_map;             // [ncell][na_map]
  DBuf<double> d_cellq;             // [ncell][7][nq_pad]
  DBuf<double> d_cellq8;            // [ncell][nq_pad][8] point-major prescaled copy (free-space fast path of K1)
  DBuf<double> d_phi_reg;           // [nq][na]
  DBuf<double> d_l1d;               // [n1d][degree+1] 1-D Lagrange values at the 1-D rule points
  DBuf<double> d_map_tab_reg;       // [nq][na_map][3]  (phi, dphi_x, dphi_y)
  int nq = 0, nq_pad = 0;
  ColumnBlocks blocks;
  DBuf<int> d_blk_cell_ptr, d_blk_cells;
  DBuf<signed char> d_blk_slots;
  DBuf<int> d_blk_nodes;
  DBuf<unsigned char> d_blk_first;
  DBuf<unsigned> d_blk_sync;
  // singular pass tables
  DBuf<int> d_patch_ptr, d_patch_cell, d_patch_local;   // per internal position (CSR)
  DBuf<double> d_sing_tab;          // concatenated per rule: [nqs][ (na + 3*na_map + 1) ]
  std::vector<int> sing_off, sing_nq;
  DBuf<int> d_sing_off, d_sing_nq;

  // matrices (row block of this rank, internal ordering)
  size_t ld = 0;                    // leading dimension (doubles)
  size_t rows_loc = 0;              // 3*(p1-p0)
  DBuf<double> storeV, storeK, storeA;
  DMat V, K, A;
  int num_rigid = 0;
  bool A_aliases_V = false;
  bool has_rigid_rows = false;      // this rank stores the rigid rows (last rank)
  size_t mono_size = 0;             // 3N + num_rigid

  // fused "no-K" assembly: K is never stored; K * panel products are accumulated in the tile epilogue.
  // panel X[3N][panel_p] (internal rows, p fastest): columns 0..2 = versors e_k, 3..3+nr-1 = P N_r, last = P u_shape
  bool fused = false;
  int panel_p = 0, panel_nr = 0;
  DBuf<double> d_panel;             // [3N][panel_p]
  DBuf<double> d_KX;                // [rows_loc][panel_p] = K_loc * X (uncorrected K)
  std::vector<double> h_panel;      // host copy of X
  std::vector<double> h_C;          // [3][rows_loc] K e_k on the owned rows (after bs_correct_K in fused mode)
  int fused_alpha = 0;
  // mixed boundary conditions in the fused mode: the unknowns flagged here are wall velocities, whose columns of the
  // monolithic matrix are -K columns (ref: bem_stokes.cc:3194-3245).  K is not stored in this mode, so the assembly keeps
  // -K for exactly these columns in a compact side matrix.
  std::vector<unsigned char> col_flags;   // [3N] reference ordering (empty = none)
  DBuf<int> d_kcol;                       // [3N] internal column -> compact column of Kflag, -1 = not flagged
  DBuf<double> d_Kflag;                   // [rows_loc][ldk] = -K(:, flagged) (uncorrected)
  size_t ldk = 0;
  int n_flagged = 0;
  // projector data (internal ordering, full length 3N)
  DBuf<double> d_nhat, d_Mnhat;
  DBuf<double> d_r1u, d_r1w, d_r1wA;  // implicit V correction: u = (nhat - V nhat) / l2 on the owned rows, w = M nhat (masked copy for A)
  double l2gamma = 0.0;
  bool have_projector = false;

  // hanging-node constraints (ref: bem_stokes.cc:2970-2995, 3156-3183), reference ordering as given by the host
  std::vector<int> cons_dof, cons_ptr, cons_col;
  std::vector<double> cons_coef;
  DBuf<int> d_cons_row, d_cons_ptr, d_cons_col;   // owned constrained rows (local row index), CSR of (internal column, coefficient)
  DBuf<double> d_cons_coef;
  DBuf<unsigned char> d_cons_node;                // [p1-p0] 1 = the node's x-component dof is constrained (K correction skips it)
  int n_cons_owned = 0;
  // flagellum torque unknown (ref: solve_with_torque, bem_stokes.cc:3191, 3252-3256, 3340-3352)
  std::vector<double> torque_mode, torque_dual;
  double torque_rhs = 0.0;
  bool torque_on = false;

  int gmres_ortho = BS_ORTHO_CGS2;
  // preconditioner
  int prec_kind = BS_PREC_NONE;
  int prec_which = BS_MAT_A;
  DBuf<double> d_prec_diag;         // Jacobi
  // LU factors of the diagonal blocks the preconditioner solves with (one block = the whole matrix for
  // BS_PREC_DIRECT / BAND, the rank's own row block or pieces of it for BS_PREC_BLOCK_DIRECT)
  struct LuBlock {
    size_t off = 0, n = 0, ld = 0;  // local row offset, size, leading dimension of LU
    double *LU = nullptr;           // points into d_lu
    int *piv = nullptr, *perm = nullptr;   // into d_piv: LAPACK-style pivot rows and the equivalent gather permutation
    double *LinvT = nullptr, *UinvT = nullptr;  // into d_luinv: row-major inverses of the 512 x 512 diagonal blocks of L and U
  };
  std::vector<LuBlock> lu_blocks;
  DBuf<double> d_lu, d_luinv;
  DBuf<int> d_piv;

  // comm
  bs_allgatherv_fn cb_allgatherv = nullptr;
  bs_allreduce_sum_fn cb_allreduce = nullptr;
  void *cb_user = nullptr;
  // peer-memory exchange (NVLink P2P through CUDA IPC)
  static constexpr int XCHG_SLOTS = 8;   // replicated vectors in flight (multi-RHS lockstep solves)
  bool p2p = false;
  size_t xchg_ld = 0;               // doubles per slot
  size_t red_off = 0;               // offset of the reduction area [2][nranks][BS_RED_CAP] inside d_xchg
  DBuf<double> d_xchg;              // [XCHG_SLOTS][xchg_ld] replicated vector buffer written by all ranks, then the reduction area
  DBuf<unsigned long long> d_flags; // BS_FLAG_WORDS words, see the BS_FLAG_* layout below
  void *gm_host = nullptr;          // bs_gmres.cu: pinned status mirror of the device-resident GMRES
  std::vector<void *> peer_xbuf, peer_flags;   // mapped pointers, rank order (own entry = local pointer)
  DBuf<double *> d_peer_xbuf;
  DBuf<unsigned long long *> d_peer_flags;
  unsigned long long epoch = 0;

  // scratch: named, context-owned, grow-only workspaces for what used to be function-local buffers
  std::map<std::string, DBuf<double>> ws_d;
  std::map<std::string, DBuf<int>> ws_i;
  double *wsd(const char *key, size_t count) {
    DBuf<double> &b = ws_d[key];
    if (b.cap < count) b.alloc(count);
    return b.p;
  }
  int *wsi(const char *key, size_t count) {
    DBuf<int> &b = ws_i[key];
    if (b.cap < count) b.alloc(count);
    return b.p;
  }
  DBuf<double> d_tmp0, d_tmp1, d_tmp2, d_tmp3;
  DBuf<double> d_small;             // small reductions
  double *h_pinned = nullptr;       // pinned host scratch
  size_t h_pinned_n = 0;

  void *extra = nullptr;            // bs_api.cu private scratch
  bs_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  size_t n3() const { return (size_t)3 * N; }
  size_t local_vec_len(int which) const;     // length of this rank's slice of a `which` vector
  size_t full_vec_len(int which) const;
  size_t slice_offset(int which) const;      // offset of my slice in the full internal vector
};

// ---- geometry / tables (bs_host.cu) ---------------------------------------------------------------------
void host_prepass(int fe_degree, int map_degree, int n_map_nodes, const double *euler_vec, int ncell, const int *conn_map,
                  int n_nodes, const int *conn, int quad_order, const double *pole, double *nhat, double *Mnhat, double *l2,
                  double *N_rigid, double *N_rigid_dual, double *area_out, double *support_out);
// same quantities on the device from the context's geometry (bs_prepass.cu); results in internal ordering
void device_prepass(Context &c, int pole_kind, const double pole_in[3], double *d_nhat, double *d_Mnhat, double *d_Nr,
                    double *d_Nrd, double *h_l2, double *h_area, double *h_center_of_mass, double *h_pole_used, int *cg_iterations);
void compute_node_order(Context &c);    // host part of build_geometry: support points, Morton order, row partition
void build_cell_blocks(Context &c);     // host part of build_tables: cell blocks of K1 (colours, pairs)
void build_geometry(Context &c);
void update_coordinates(Context &c);    // same mesh, new map_nodes
void build_tables(Context &c);          // after geometry + quadrature known
// ---- assembly (bs_assembly.cu) ---------------------------------------------------------------------------
void launch_cell_geometry(Context &c);
void launch_assembly_regular(Context &c);
void launch_assembly_singular(Context &c);
constexpr int MAX_PANEL = 12;
int tile_planes(int na, int kernel_type);
size_t assembly_smem_bytes(int na, int planes, int tj, int nq_pad, int cs = 1);
int choose_tj(int na, int kernel_type, int nq_pad, int cs = 1);
int cell_sets(const Context &c);        // 2: K1 runs two thread sets on different cells of a block (see ColumnBlocks::cs)
void kernel_eval_device(int type, double eps, int o, int npts, const double *d_p, const double *d_pim, double *d_G,
                        double *d_W, cudaStream_t s);
// ---- linear algebra (bs_linalg.cu) ------------------------------------------------------------------------
// y[rows] = M x; skip != nullptr: device flag, the sweep returns at once when it is set (device-resident GMRES)
void gemv(Context &c, const DMat &M, const double *x, double *y, const int *skip = nullptr);
void gemv_multi(Context &c, const DMat &M, int nrhs, const double *X, size_t ldx, double *Y, size_t ldy, const int *skip = nullptr);
void rank1_update(Context &c, DMat &M, const double *u, const double *w, double scale);  // M += scale u w^T
void perm_in(Context &c, const double *src_dev, double *dst_int, const int *d_node_of_pos, int nextra);
void perm_out(Context &c, const double *src_int, double *dst_dev, const int *d_node_of_pos, int nextra, size_t lo, size_t hi);
double dot(Context &c, const double *a, const double *b, size_t n);                // device sync + host result
void multi_dot(Context &c, const double *basis, size_t ldb, int k, const double *w, size_t n, double *d_out);
void multi_axpy(Context &c, const double *basis, size_t ldb, int k, const double *d_coef, double sign, double *w, size_t n);
void axpy(Context &c, double a, const double *x, double *y, size_t n);
void scal(Context &c, double a, double *x, size_t n);
void scal_dev_inv(Context &c, const double *d_s, double *x, const double *src, size_t n);  // x = src / *d_s
void copy(Context &c, const double *src, double *dst, size_t n);
void sub(Context &c, const double *a, const double *b, double *out, size_t n);      // out = a - b
void mul_elem(Context &c, const double *a, const double *d, double *out, size_t n); // out = a*d
void fill(Context &c, double *x, double v, size_t n);
void k_correct_diag(Context &c, DMat &K, const double *Ck /*[3][rows] internal*/, int use_internal_alpha);
void apply_constraint_rows(Context &c, DMat &M, size_t ncols);   // constrained owned rows: 0 ... 1 (diagonal) ... -coefficient ...
void zero_constrained_entries(Context &c, double *v_loc);         // v_loc[row] = 0 on the constrained owned rows
void extract_diag(Context &c, const DMat &M, size_t row_offset, double *d_out);
void select_columns(Context &c, DMat &A, const DMat &V, const DMat &K, const unsigned char *d_flag, bool alias);
void set_column(Context &c, DMat &A, size_t col, const double *v, double scale);
// A(:, c) = Kflag(:, kcol[c]) for the flagged columns (Kflag = -K uncorrected), K correction applied on the node's own 3 x 3 block
void scatter_flagged_columns(Context &c, DMat &A, const double *Kflag, size_t ldk, const int *kcol, const double *Ck, int alpha);
void gather_entries(Context &c, const DMat &M, int n, const int *d_r, const int *d_c, double *d_out);
// dst[i][j] (n x n, leading dimension ldd) += r1_u[row_off + i] * r1_w[col_off + j] of M's implicit term (after a plain copy of a block of M)
void add_rank1_block(Context &c, const DMat &M, size_t row_off, size_t col_off, size_t n, double *dst, size_t ldd);
// ---- solvers (bs_solve.cu) --------------------------------------------------------------------------------
void lu_factor(Context &c, double *A, size_t n, size_t ld, int *piv);
// factorise local diagonal blocks of `M` (copied; band > 0 drops entries outside the reference-ordered band) for the preconditioner
void precond_factor_blocks(Context &c, const DMat &M, size_t col_off, size_t n_total, size_t max_block, int band);
void lu_apply_fast(Context &c, const Context::LuBlock &b, const double *in, double *out, const int *skip);
void lu_solve(Context &c, const double *LU, size_t n, size_t ld, const int *piv, double *x /* in/out */);
void apply_operator(Context &c, int which, const double *x_full, double *y_loc);
void exchange(Context &c, int which, const double *y_loc, double *x_full);
void p2p_scatter(Context &c, int which, const double *src_loc, int slot, const double *inv_norm2, double *basis_dst);
void p2p_wait(Context &c);
void p2p_signal(Context &c);                                         // bs_gmres.cu: device-resident epoch
void p2p_wait_only(Context &c, const int *skip, void *gm_status);
void apply_precond(Context &c, const double *in_loc, double *out_loc, const int *skip = nullptr);
bool gmres_device_eligible(Context &c, int nrhs, int max_tmp);
int gmres_device(Context &c, int which, int nrhs, const double *d_B, double *d_X, size_t ldv, double tol, int max_steps,
                 int max_tmp, int *iters, double *final_res);
void gm_host_release(Context &c);
int gmres(Context &c, int which, const double *d_b_loc, double *d_x_loc, double tol, int max_steps, int max_tmp,
          int *iters, double *final_res);
int gmres_batched(Context &c, int which, int nrhs, const double *d_B, double *d_X, size_t ldv, double tol, int max_steps,
                  int max_tmp, int *iters, double *final_res);
void evaluate_bie(Context &c, int npts, const double *d_pts, const double *d_vel, const double *d_forces, double *d_out,
                  bool on_boundary, bool accumulate);
void count_launch(Context &c, int n = 1);

}  // namespace bs

struct bs_context {
  bs::Context c;
};
