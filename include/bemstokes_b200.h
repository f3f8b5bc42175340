/* bemstokes_b200.h — C-ABI of the B200-native BEMStokes hot path.
 *
 * Drop-in boundary for ONE path of mathLab/BEMStokes: collocation assembly of the dense single-layer (V) and
 * double-layer (K) Stokes matrices and the GMRES / direct solve of the monolithic rigid-body system.
 * Everything behind these entry points runs on the GPU (sm_100a); there is no CPU fallback.
 *
 * Conventions
 *  - every function returns 0 on success, a negative bs_status otherwise; bs_last_error() gives the text
 *    (the reference throws deal.II exceptions, caught in source/main.cc:48-71).
 *  - all pointers are caller-owned HOST arrays borrowed for the duration of the call, unless the context was
 *    switched to BS_PTR_DEVICE with bs_set_pointer_mode (then the x/y/b vectors of vmult/gmres/precond are
 *    device pointers in the same reference ordering).
 *  - vectors use the reference's component-major ordering: dof (node i, component c) = i + c*N
 *    (DoFRenumbering::component_wise, source/bem_stokes.cc:1593); the monolithic vector appends the
 *    num_rigid (6, +1 with torque) rigid unknowns at 3N.. (bem_stokes.cc:3247-3251).  The library permutes
 *    to its own node-major, locality-sorted device layout internally.
 *  - one context per rank/GPU; a context is not re-entrant.
 *
 * "ref:" = file:line under the reference tree that the entry point replaces.
 */
#ifndef BEMSTOKES_B200_H
#define BEMSTOKES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bs_context bs_context;

typedef enum {
  BS_OK = 0,
  BS_ERR_INVALID = -1,     /* bad argument / call order */
  BS_ERR_CUDA = -2,        /* CUDA runtime failure (message has the CUDA error string) */
  BS_ERR_NO_DEVICE = -3,   /* no sm_100 device: the product path has no CPU fallback */
  BS_ERR_NOT_CONVERGED = -4, /* SolverControl::NoConvergence equivalent */
  BS_ERR_UNSUPPORTED = -5,
  BS_ERR_COMM = -6
} bs_status;

typedef enum { BS_KERNEL_FREE = 0, BS_KERNEL_FREE_SURFACE = 1, BS_KERNEL_NO_SLIP = 2 } bs_kernel_type;
typedef enum { BS_SING_MIXED = 0, BS_SING_DUFFY = 1, BS_SING_TELLES = 2 } bs_singular_kind;
typedef enum { BS_MAT_V = 0, BS_MAT_K = 1, BS_MAT_A = 2 } bs_matrix_id;
typedef enum { BS_PREC_NONE = 0, BS_PREC_JACOBI = 1, BS_PREC_DIRECT = 2, BS_PREC_BLOCK_DIRECT = 3, BS_PREC_BAND = 4 } bs_precond_kind;
typedef enum { BS_GRID_REAL = 0, BS_GRID_IMPOSED_FORCE = 1, BS_GRID_IMPOSED_VELOCITY = 2 } bs_grid_type;
typedef enum { BS_PTR_HOST = 0, BS_PTR_DEVICE = 1 } bs_pointer_mode;
/* Arnoldi orthogonalisation: BS_ORTHO_MGS = deal.II's modified Gram-Schmidt verbatim (sequential projections,
 * conditional second pass every 5th iteration; one reduction per basis vector); BS_ORTHO_CGS2 (default) =
 * classical Gram-Schmidt twice with two fused reductions per iteration.  Identical in exact arithmetic; CGS2 keeps
 * the basis orthogonal to machine precision, so its iteration count can be one lower when the reference's residual
 * estimate sits just above the tolerance. */
typedef enum { BS_ORTHO_CGS2 = 0, BS_ORTHO_MGS = 1 } bs_ortho_kind;

const char *bs_last_error(void);
int bs_version(void);

/* ---- life cycle ------------------------------------------------------------------------------------ */
/* ref: BEMProblem<3>::BEMProblem (source/bem_stokes.cc:138-205), fe_stokes / fe_map (bem_stokes.h:414-419).
 * fe_degree, map_degree in {1,2}.  device = CUDA ordinal. */
int bs_create(bs_context **ctx, int device, int fe_degree, int map_degree);
int bs_destroy(bs_context *ctx);
int bs_set_pointer_mode(bs_context *ctx, int mode);
/* cudaStream_t (as void*) all work of this context is ordered on; NULL = a private non-blocking stream. */
int bs_set_stream(bs_context *ctx, void *cuda_stream);

/* ---- row partition (ref: this_cpu_set, bem_stokes.cc:1599-1634; assembly filter 2877) ---------------- */
/* owner_of_node[N] gives the owning rank of every node (same array on every rank) or NULL for the
 * library's balanced split of its locality order.  Must be called before bs_set_geometry when nranks>1. */
int bs_set_partition(bs_context *ctx, int rank, int nranks, const int *owner_of_node, int n_nodes);
int bs_get_owned_nodes(bs_context *ctx, int *n_owned, int *owned_nodes /* capacity N, may be NULL */);

/* ---- inputs ----------------------------------------------------------------------------------------- */
/* ref: euler_vec + map_dh / dh_stokes connectivity (bem_stokes.cc:1819, 1851, 2874 get_dof_indices).
 * euler_vec: 3*n_map_nodes, component-major.  conn_*: per cell the scalar node ids in deal.II local order
 * (Q1: 4 lexicographic vertices; Q2: 4 vertices, 4 edge mid-points (x=0,x=1,y=0,y=1), centre).
 * material_id[ncell] may be NULL (all body, material 0; bem_stokes.cc:515-521). */
int bs_set_geometry(bs_context *ctx, int n_map_nodes, const double *euler_vec, int ncell, const int *conn_map,
                    int n_nodes, const int *conn_stokes, const int *material_id);
/* ref: "Internal Quadrature" ParsedQuadrature (bem_stokes.cc:151): tensor product of a 1-D rule on [0,1],
 * first coordinate fastest.  x1d == NULL -> Gauss-Legendre of order n1d built by the library. */
int bs_set_quadrature(bs_context *ctx, int n1d, const double *x1d, const double *w1d);
/* ref: get_singular_quadrature (bem_stokes.cc:4912-4957): rule family + order, built by the library with
 * deal.II semantics (QGaussOneOverR / QIterated / QSplit(QDuffy) / QTelles). */
int bs_set_singular_quadrature(bs_context *ctx, int kind, int order);
/* Alternative: hand over the rule of local scalar index a as deal.II built it (points on [0,1]^2). */
int bs_set_singular_rule(bs_context *ctx, int local_index, int nq, const double *xi, const double *w);
/* ref: stokes_kernel / fs_stokes_kernel / ns_stokes_kernel + reflect_kernel / no_slip_kernel dispatch
 * (bem_stokes.cc:5027-5069), wall orientation/position (2861-2870, 2918-2919). */
int bs_set_kernel(bs_context *ctx, int type, double epsilon, int wall_orientation, const double *wall_position);

/* Host helpers with deal.II semantics (so a stand-alone host needs no deal.II). Return the number of
 * points written; xi is [n][2]. capacity in points. */
int bs_make_gauss_1d(int n, double *x, double *w);
int bs_make_singular_rule(int kind, int order, int fe_degree, int local_index, int capacity, double *xi, double *w);

/* Host helper, no device needed: the tiling of the regular assembly pass (K1) for a Q1 mesh - cell blocks, colours and,
 * in the cell-split mode (Gauss 8, no regularisation), the pairs of cells the two thread sets of a CTA integrate
 * concurrently with the steps that start with a barrier - exactly as bs_set_geometry / bs_set_quadrature build it
 * (the loop nest it tiles: ref bem_stokes.cc:2871-2998).  For tests and tooling.  nodes [N][3], conn [ncell][4].
 * sizes_out[6] = n_blocks, nodes per block (stride of block_nodes / first_touch), cell sets (1 or 2), n_colours,
 * length of `cells`, cells without partner.  Arrays sized by the caller (NULL = skip): cell_ptr [ncell + 1],
 * cells [2 ncell] (-1 = no partner in this step), sync_mask [ncell] (bit s: step s of the block starts with a barrier),
 * block_nodes / first_touch [32 ncell] (node POSITION per slot, -1 unused), colour_start [65], pos_of_node [N]. */
int bs_host_cell_blocks(int n_nodes, const double *nodes, int ncell, const int *conn, int kernel_type, int n1d,
                        int *sizes_out, int *cell_ptr, int *cells, unsigned *sync_mask, int *block_nodes,
                        unsigned char *first_touch, int *colour_start, int *pos_of_node);

/* Host pre-pass with the reference's semantics (compute_center_of_mass_and_rigid_modes bem_stokes.cc:2440-2788,
 * compute_normal_vector 3922-4011) for hosts without deal.II: scalar mass matrix, L2-projected unit normals nhat
 * (= normal_vector_pure for a body-only mesh), Mnhat = M nhat, l2gamma = nhat^T M nhat, the six rigid modes about
 * `pole` (NULL = origin) and their duals M N_r (each 6 x 3N row-major, may be NULL), surface area, support points
 * [N][3] (may be NULL).  Pure host code, O(N); every vector component-major. */
int bs_host_prepass(int fe_degree, int map_degree, int n_map_nodes, const double *euler_vec, int ncell, const int *conn_map,
                    int n_nodes, const int *conn_stokes, int quad_order, const double *pole, double *nhat, double *Mnhat,
                    double *l2gamma, double *N_rigid, double *N_rigid_dual, double *area, double *support_points);

/* The same pre-pass on the device, from the geometry and regular quadrature the context holds (bs_set_geometry,
 * bs_set_quadrature): per-cell local mass matrices, matrix-free M x as a gather over node patches, the three normal
 * components solved together by Jacobi-preconditioned CG to 1e-15 (ref: compute_normal_vector bem_stokes.cc:3922-4011,
 * compute_center_of_mass_and_rigid_modes 2440-2788).  Outputs as bs_host_prepass, reference ordering (i + c*N),
 * host or device arrays per pointer mode (support_points, center_of_mass[3], pole_used[3] are always host arrays).
 * The rigid modes are taken about the origin, about `pole` (BS_POLE_POINT) or about the surface centroid
 * int y dS / int dS (BS_POLE_BARICENTER, ref: 2487-2493, 2540-2550), which is also returned in center_of_mass.
 * N_rigid / N_rigid_dual ([6][3N]), l2gamma, area, support_points, center_of_mass, pole_used, cg_iterations may be NULL.  Every rank computes
 * the full vectors (O(N) work, geometry is replicated).  Cells with repeated nodes are not supported here. */
typedef enum { BS_POLE_ORIGIN = 0, BS_POLE_POINT = 1, BS_POLE_BARICENTER = 2 } bs_pole_kind; /* "Force Pole to be used" */
int bs_prepass(bs_context *ctx, int pole_kind, const double *pole, double *nhat, double *Mnhat, double *l2gamma,
               double *N_rigid, double *N_rigid_dual, double *area, double *support_points, double *center_of_mass,
               double *pole_used, int *cg_iterations);

/* Hanging-node constraints (ref: the AffineConstraints of dh_stokes, bem_stokes.cc:2970-2995 in the assembly loop,
 * 3024-3025 / 3078 in the corrections, 3156-3183 in the monolithic build).  dof[k] (reference ordering, i + c*N) is
 * constrained to sum_q coefs[q] * x[cols[q]], q in [ptr[k], ptr[k+1]).  The row of a constrained dof is not integrated: in
 * V, K and A it holds the constraint equation (1 on the diagonal, -coef at the constraining dofs, rhs 0), and the V / K
 * corrections skip it.  Call before bs_assemble_VK; n_constrained = 0 clears.  Not available with bs_assemble_fused. */
int bs_set_constraints(bs_context *ctx, int n_constrained, const int *dof, const int *ptr, const int *cols, const double *coefs);
/* solve_with_torque (ref: bem_stokes.cc:1612-1634, 3143-3147, 3191, 3252-3256, 3340-3352): one more unknown after the
 * rigid ones - the flagellum's angular velocity - with the column -scaling P K P N_torque, the row scaling N_torque_dual
 * and the right-hand side rhs_value (-2 in the reference); every node row then has a zero right-hand side.  Takes effect in
 * the next bs_build_monolithic (whose vectors get num_rigid + 1 trailing entries); N_torque = NULL switches it off. */
int bs_set_torque_mode(bs_context *ctx, const double *N_torque, const double *N_torque_dual, double rhs_value);

/* ---- assembly (ref: BEMProblem::assemble_stokes_system, bem_stokes.cc:2840-3435) ----------------------- */
/* K1 (regular Gauss pass) + K2 (singular pass) -> row-block of V and K on the device (2871-3000). */
int bs_assemble_VK(bs_context *ctx);
/* Fused "no-K" assembly for sizes where V and K do not fit together (BASELINE config 4: 294 918 DoF, 696 GB per
 * matrix): same K1/K2 passes, but the double-layer tile is multiplied in the tile epilogue with the panel
 * [e_0 e_1 e_2 | P N_r | P u_shape] instead of being stored, which is all the monolithic system of a body-only
 * problem needs from K (K correction 3044-3098, projected rigid columns 3120-3148, rhs 3127-3132).  Afterwards
 * call bs_correct_V, bs_correct_K and bs_build_monolithic as usual (A aliases V; col_is_K as given to
 * bs_set_column_flags); BS_MAT_K is not available. */
/* Mixed boundary conditions with the fused assembly: col_is_K[3N] flags the unknowns that are wall velocities, whose
 * columns of the monolithic matrix are -K columns (ref: the index-set logic of bem_stokes.cc:3194-3245).  K is not stored in
 * the fused mode, so bs_assemble_fused keeps -K for exactly these columns in a compact side matrix (rows x flagged columns)
 * and bs_build_monolithic, called with the same flags, moves them into A.  NULL clears.  Call before bs_assemble_fused. */
int bs_set_column_flags(bs_context *ctx, const unsigned char *col_is_K);
int bs_assemble_fused(bs_context *ctx, int num_rigid, const double *N_rigid, const double *nhat, const double *Mnhat,
                      double l2gamma, const double *shape_vel);
/* V <- V + (nhat - V nhat)(M nhat)^T / l2 on owned rows (3004-3036). nhat = normal_vector_pure,
 * Mnhat = M_normal_vector_pure, l2 = l2normGamma_pure.  Vn_out (3N, may be NULL) receives V*nhat
 * computed BEFORE the correction ("Check on the V operator Norm"). */
int bs_correct_V(bs_context *ctx, const double *nhat, const double *Mnhat, double l2gamma, double *Vn_out);
/* K(i+jN, i+kN) -= (K e_k)[i+jN]; += delta_jk unless use_internal_alpha (3044-3098). */
int bs_correct_K(bs_context *ctx, int use_internal_alpha);
/* Monolithic matrix + rhs (3120-3357); constrained rows as set by bs_set_constraints, torque unknown by bs_set_torque_mode.
 * col_is_K[3N] (NULL = all V): column j of A is -K(:,j) when set, V(:,j) otherwise (the reference derives it
 * from the body / wall index sets, 3194-3245).  N_rigid, N_rigid_dual: num_rigid x 3N row-major.
 * shape_vel (3N, may be NULL) is used for grid_type Real.  rhs_out: 3N+num_rigid.
 * keep_VK = 0 lets A alias V's storage (memory at scale), 1 keeps V intact. */
int bs_build_monolithic(bs_context *ctx, const unsigned char *col_is_K, int num_rigid, const double *N_rigid,
                        const double *N_rigid_dual, const double *nhat, const double *Mnhat, double l2gamma,
                        int grid_type, int imposed_component, double scaling, const double *shape_vel,
                        int keep_VK, double *rhs_out);

/* ---- operators (ref: TrilinosWrappers::SparseMatrix::vmult call sites) -------------------- */
int bs_matrix_size(bs_context *ctx, int which, int *rows, int *cols);
int bs_vmult(bs_context *ctx, int which, const double *x, double *y);
/* X, Y: nrhs vectors stored one after another (nrhs x size). */
int bs_vmult_multi(bs_context *ctx, int which, int nrhs, const double *X, double *Y);
/* ref: V_matrix(i,j) / K_matrix(i,j) / monolithic_system_matrix(i,i) element reads (3196-3243, 4355-4361,
 * 3412-3413).  rows/cols in reference ordering; rows must be owned by this rank. */
int bs_get_entries(bs_context *ctx, int which, int n, const int *rows, const int *cols, double *out);
/* tangential_projector_body (4142-4151) with the nhat/Mnhat/l2 given to bs_build_monolithic/correct_V. */
int bs_tangential_projector(bs_context *ctx, const double *in, double *out);

/* ---- preconditioner (ref: DirectPreconditioner, source/direct_preconditioner.cc:10-23; Jacobi 4296-4300;
 *      band copy assemble_monolithic_preconditioner 3437-3505) ---------------------------------------- */
/* kind BS_PREC_BAND: param = band width; BS_PREC_BLOCK_DIRECT: param = largest diagonal block in rows (0 = the whole
 * row block of the rank is one block); both ignored otherwise.  The LU factors persist until the next call: the host keeps
 * them across frames exactly as the reference keeps direct_trilinos_preconditioner (bem_stokes.cc:5768-5779). */
int bs_precond_setup(bs_context *ctx, int which, int kind, int bandwidth_or_block);
int bs_precond_vmult(bs_context *ctx, const double *x, double *y);

/* ---- solvers (ref: solve_system 4158-4508; deal.II SolverGMRES semantics) ----------------- */
/* Left-preconditioned restarted GMRES, x is the initial guess on entry.  max_n_tmp_vectors as
 * gmres_additional_data (restart length = max_n_tmp_vectors-2).  Returns BS_ERR_NOT_CONVERGED after
 * max_steps like SolverControl. */
int bs_gmres(bs_context *ctx, int which, const double *b, double *x, double tol_abs, int max_steps,
             int max_n_tmp_vectors, int *iterations, double *final_residual);
int bs_set_gmres_orthogonalization(bs_context *ctx, int kind);
/* nrhs independent systems (the 6 rigid-body resistance problems), B and X nrhs x size. */
int bs_gmres_multi(bs_context *ctx, int which, int nrhs, const double *B, double *X, double tol_abs, int max_steps,
                   int max_n_tmp_vectors, int *iterations, double *final_residuals);
/* DN-operator route, batched (ref: dirichlet_to_neumann_operator bem_stokes.cc:4072-4129, its 6 + 1 calls in
 * solve_system(false) 4163-4258): F_k = P V^-1 P K P u_k for nvec <= 8 velocities at once (U, F: nvec x 3N).  One
 * multi-right-hand-side sweep over K, then the nvec V-systems advanced in lockstep by one multi-right-hand-side sweep over
 * V per GMRES iteration (solve_directly = 0; the preconditioner set up for BS_MAT_V applies), or one LU of V for all of them
 * (solve_directly = 1).  Needs V and K both stored and corrected, and the projector data.  iterations: nvec counts (may be NULL). */
int bs_dn_operator_multi(bs_context *ctx, int nvec, const double *U, double *F, int solve_directly, double tol_abs,
                         int max_steps, int max_n_tmp_vectors, int *iterations);
/* ref: TrilinosWrappers::SolverDirect (4261-4267): dense LU with partial pivoting on the device. */
int bs_direct_solve(bs_context *ctx, int which, const double *b, double *x);

/* ---- field evaluation (ref: BEMProblem::evaluate_stokes_bie bem_stokes.cc:5366-5451 and
 *      evaluate_stokes_bie_on_boundary 5454-5560) ---------------------------------------------------------------
 * u_a(x_i) = sum_cells sum_q [G_ab f_b - (W n)_ab u_b] JxW with the kernel of bs_set_kernel.  points: npts x 3
 * (x,y,z per point); vel, forces: 3N reference ordering; out: 3*npts component-major (i + a*npts).  on_boundary=1:
 * free-space kernel, cells with a support point within 1e-3 of x_i use the singular rule, result ACCUMULATED into
 * out like the reference. */
int bs_evaluate_bie(bs_context *ctx, int npts, const double *points, const double *vel, const double *forces, double *out,
                    int on_boundary);

/* ---- kernel point evaluation (ref: StokesKernel::value_tens / value_tens2 kernel.cc:61-104,
 *      FreeSurfaceStokesKernel::value_tens_image(2), NoSlipWallStokesKernel::value_tens_image(2)) -------
 * Evaluated by the same device functions the assembly kernels use.  p, p_image: npts x 3.
 * G_out: npts x 9 (may be NULL), W_out: npts x 27 (may be NULL). */
int bs_kernel_eval(int device, int type, double epsilon, int wall_orientation, int npts, const double *p,
                   const double *p_image, double *G_out, double *W_out);

/* ---- multi-GPU plumbing -------------------------------------------------------------------------------
 * The library does not own a communicator; the host passes the exchange steps of the solve as callbacks
 * (the reference gets them from Epetra: Import in vmult, Allreduce in dots).  Buffers are
 * device pointers; the callback must order its work on `stream`.  allgatherv gathers counts[r] doubles of
 * every rank into recv at displs[r]. */
typedef int (*bs_allgatherv_fn)(void *user, const double *send, int sendcount, double *recv, const int *counts,
                                const int *displs, void *stream);
typedef int (*bs_allreduce_sum_fn)(void *user, double *buf, int count, void *stream);
int bs_set_comm(bs_context *ctx, bs_allgatherv_fn allgatherv, bs_allreduce_sum_fn allreduce, void *user);
/* Peer-memory exchange (NVLink P2P), replacing the allgatherv callback: every rank exports CUDA-IPC handles of
 * its replicated Krylov-vector buffer and of its arrival flags (bs_exchange_export, 2 x 64 bytes), the host
 * all-gathers the handles of all ranks (rank order) and hands them to bs_exchange_import.  From then on the kernel
 * that normalises a new Krylov vector stores its slice straight into every peer's buffer, a release store raises the
 * per-source flag on each peer, and the next matvec starts after an acquire-wait on its own flags: the exchange is
 * fused with the producing kernel and no collective call is made.  The dot-product reductions keep using the
 * allreduce callback.  max_vec_len = 3N + num_rigid of the largest system that will be solved. */
#define BS_IPC_EXPORT_BYTES 128
int bs_exchange_export(bs_context *ctx, size_t max_vec_len, unsigned char *handles_out /* BS_IPC_EXPORT_BYTES */);
int bs_exchange_import(bs_context *ctx, int nranks, const unsigned char *all_handles /* nranks x BS_IPC_EXPORT_BYTES */);

/* ---- timers / counters (ref: Teuchos timers bem_stokes.cc:19-23) ---------------------------------------- */
typedef struct {
  double assemble_regular_ms, assemble_singular_ms, geometry_ms, correct_ms, monolithic_ms;
  double precond_setup_ms, solve_ms, vmult_ms_last;
  long long kernel_launches; /* number of this library's kernels launched since bs_reset_stats */
  long long pairs_regular, pairs_singular;
  /* tiling of the regular pass: disjoint cell blocks, colours (launches), tile traffic amplification */
  long long n_cell_blocks, n_colours;
  double node_touch_ratio;
  /* last device-resident GMRES solve: time on the stream, of which in the sweeps over the matrix (CUDA events around each) */
  double gmres_stream_ms_last, gmres_matvec_ms_last;
  long long gmres_sweeps_last;
  /* regular pass, cell-split mode (Q1 unknowns, Gauss 8, no regularisation): thread sets working on different cells of a
   * block (2, else 1), steps of all blocks (pairs of cells), cells without a partner, steps that start with a barrier */
  long long cell_sets, cell_steps, unpaired_cells, sync_steps;
} bs_stats;
int bs_get_stats(bs_context *ctx, bs_stats *out);
int bs_reset_stats(bs_context *ctx);

/* ---- benchmarking helpers: device-resident repeat loops timed with CUDA events on the context stream ---- */
int bs_bench_vmult(bs_context *ctx, int which, int repeats, double *ms_per_call);
int bs_bench_vmult_multi(bs_context *ctx, int which, int nrhs, int repeats, double *ms_per_call);
/* dense LU (DirectPreconditioner's factorisation, (2/3) n^3 flops) and its application on a synthetic n x n matrix;
 * residual = |A y - b|_inf for b = 1.  Invalidates a preconditioner that was set up on this context. */
int bs_bench_lu(bs_context *ctx, int n, int apply_repeats, double *factor_ms, double *apply_ms, double *residual);
int bs_bench_fp64_peak(int device, double *tflops);                              /* burst, best of 3 */
int bs_bench_fp64_sustained(int device, double seconds, double *tflops);         /* back-to-back under the power cap */

#ifdef __cplusplus
}
#endif
#endif /* BEMSTOKES_B200_H */
