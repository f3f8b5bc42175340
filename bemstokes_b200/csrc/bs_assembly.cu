// Collocation assembly of V (single layer) and K (double layer) on sm_100a — replaces the (cell, node, q)
// loop nest of BEMProblem::assemble_stokes_system (ref: source/bem_stokes.cc:2871-3000).
//
//   K0  k_cell_geometry      FEValues::reinit for every cell with the regular rule: y_q, n_q*JxW_q, JxW_q (and a
//                            point-major copy with the constants folded in for the free-space fast path)
//   K1  k_assemble_regular   CTA tile = TI (64) collocation nodes x one cell block (<= tj nodes, disjoint cells,
//                            coloured); the block's cells are streamed through shared memory by bulk-async copies (TMA
//                            engine) in a full/empty mbarrier ring; per-CTA shared-memory tile over cells, one coalesced
//                            write-out per CTA: plain stores where this colour is the first to touch a node column,
//                            RED.ADD.F64 otherwise (colours are separate launches -> fixed summation order); fused
//                            variant multiplies the K tile with the panel instead of storing it.
//                            Cell-split mode (Q1 unknowns, Gauss 8, no regularisation; cell_sets()): one thread per
//                            (node, cell), the CTA's two thread sets on two cells of the block without a common node;
//                            free space / free surface: 2-D moment formulation (integrate_free_lin2d,
//                            integrate_free_surface_lin2d), no slip: coefficient x tensor sums (integrate_no_slip).
//                            Otherwise: QS threads per node split the rows of the tensor rule (or V warps / K warps for
//                            Q2), per-thread sum-factorised register accumulators over q, partner exchange by shuffles.
//   K2  k_assemble_singular  one warp per owned collocation node: the cells containing the node are
//                            integrated with the singular rule of that local index (geometry evaluated on
//                            the fly) and added to the node's three rows.
#include "bs_internal.h"
#include "bs_green.cuh"

#ifndef BS_MOM2D
#define BS_MOM2D 1   // cell-split free-space kernel: moments in both directions (0: per-row expansion)
#endif
#ifndef BS_NOSLIP_FAST
#define BS_NOSLIP_FAST 1   // cell-split no-slip kernel: coefficient x tensor sums (0: entry-by-entry green_eval)
#endif
#ifndef BS_ROWS2
#define BS_ROWS2 1   // cell-split free-space kernel: 0 one rule row per thread in flight, 1 two rows (measured +2-3 %), n > 1: n rows
#endif
#ifndef BS_QX_UNROLL
#define BS_QX_UNROLL 2
#endif

namespace bs {

constexpr int QX_UNROLL = BS_QX_UNROLL;  // quadrature points of one rule row in flight per thread

void count_launch(Context &c, int n) { c.stats.kernel_launches += n; }

// ---------------------------------------------------------------------------------------------------------
// small PTX wrappers: mbarrier + bulk async copy global -> shared
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Tile write-out primitive: what == 1 plain store (first colour to touch the column), what == 2 fire-and-forget L2
// reduction (later colours; no load latency on the critical path), what == 0 nothing.  Predicated, no branch: the
// lanes of a warp mix all three.  Blocks of one colour never share a node and colours are separate launches, so
// every address receives its addends in a fixed order.
__device__ __forceinline__ void store_or_reduce(double *p, double v, int what) {
  asm volatile(
      "{\n"
      ".reg .pred ps, pr;\n"
      "setp.eq.s32 ps, %2, 1;\n"
      "setp.eq.s32 pr, %2, 2;\n"
      "@ps st.global.f64 [%0], %1;\n"
      "@pr red.global.add.f64 [%0], %1;\n"
      "}\n" ::"l"(p),
      "d"(v), "r"(what)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// K0: per-cell quadrature data with the regular rule.  cellq[cell][7][nq_pad] = y(3), n*JxW(3), JxW
// ---------------------------------------------------------------------------------------------------------
__global__ void k_cell_geometry(int ncell, int nq, int nq_pad, int nam, const int *__restrict__ conn_map,
                                const double *__restrict__ map_nodes, const double *__restrict__ tab /*[nq][nam][3]*/,
                                const double *__restrict__ w, double *__restrict__ cellq, double *__restrict__ cellq8) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)ncell * nq) return;
  const int cell = (int)(gid / nq), q = (int)(gid % nq);
  double y[3] = {0, 0, 0}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
  for (int a = 0; a < nam; ++a) {
    const int m = conn_map[(size_t)cell * nam + a];
    const double ph = tab[((size_t)q * nam + a) * 3], dx = tab[((size_t)q * nam + a) * 3 + 1],
                 dy = tab[((size_t)q * nam + a) * 3 + 2];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double X = map_nodes[(size_t)3 * m + d];
      y[d] = fma(ph, X, y[d]);
      t1[d] = fma(dx, X, t1[d]);
      t2[d] = fma(dy, X, t2[d]);
    }
  }
  const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2], nz = t1[0] * t2[1] - t1[1] * t2[0];
  const double J = sqrt(nx * nx + ny * ny + nz * nz);
  double *o = cellq + (size_t)cell * 7 * nq_pad;
  o[0 * nq_pad + q] = y[0];
  o[1 * nq_pad + q] = y[1];
  o[2 * nq_pad + q] = y[2];
  o[3 * nq_pad + q] = w[q] * nx;  // n*JxW = (nn/|nn|) * w|nn|
  o[4 * nq_pad + q] = w[q] * ny;
  o[5 * nq_pad + q] = w[q] * nz;
  o[6 * nq_pad + q] = w[q] * J;
  // point-major copy for the free-space fast path of K1, constants folded in:
  //   y(3), JxW/(8 pi), 6*n(3), 0      (k = 3/(4 pi) (R.n JxW) r^-5 R(x)R = (R.6n) r^-2 * [JxW/(8 pi) r^-3] R(x)R)
  double *o8 = cellq8 + ((size_t)cell * nq_pad + q) * 8;
  const double s6 = 6.0 / J;
  o8[0] = y[0];
  o8[1] = y[1];
  o8[2] = y[2];
  o8[3] = (w[q] * J) * BS_INV_8PI;
  o8[4] = s6 * nx;
  o8[5] = s6 * ny;
  o8[6] = s6 * nz;
  // Pad slot of point q: constant number q of a bilinear (Q1-mapped) cell y = y00 + xi a + eta b + xi eta c, read by the
  // 2-D moment formulation of K1 (integrate_free_lin2d): 0-11 = y00, a, b, c; 12 + 6 v + t for the value v = (i,j) of the
  // symmetric 6-vector: t = 0 a_i b_j + b_i a_j, 1 a_i a_j, 2 b_i b_j, 3 a_i c_j + c_i a_j, 4 b_i c_j + c_i b_j, 5 c_i c_j.
  double cst = 0.0;
  if (nam == 4 && q < 48) {
    double X[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int d = 0; d < 3; ++d) X[a][d] = map_nodes[(size_t)3 * conn_map[(size_t)cell * 4 + a] + d];
    double A[3], B[3], C[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      A[d] = X[1][d] - X[0][d];
      B[d] = X[2][d] - X[0][d];
      C[d] = (X[3][d] - X[2][d]) - (X[1][d] - X[0][d]);
    }
    if (q < 12) {
      const int d = q % 3;
      cst = q < 3 ? X[0][d] : (q < 6 ? A[d] : (q < 9 ? B[d] : C[d]));
    } else {
      const int v = (q - 12) / 6, tt = (q - 12) % 6;
      const int i = v < 3 ? 0 : (v < 5 ? 1 : 2), j = v < 3 ? v : (v < 5 ? v - 2 : 2);
      cst = tt == 0 ? fma(A[i], B[j], B[i] * A[j])
          : tt == 1 ? A[i] * A[j]
          : tt == 2 ? B[i] * B[j]
          : tt == 3 ? fma(A[i], C[j], C[i] * A[j])
          : tt == 4 ? fma(B[i], C[j], C[i] * B[j]) : C[i] * C[j];
    }
  }
  o8[7] = cst;
}

void launch_cell_geometry(Context &c) {
  c.d_cellq.alloc((size_t)c.ncell * 7 * c.nq_pad);
  c.d_cellq.zero(c.stream);
  c.d_cellq8.alloc((size_t)c.ncell * 8 * c.nq_pad);
  c.d_cellq8.zero(c.stream);
  struct P_ { double *p; } dw;
  dw.p = c.wsd("geom.w", c.reg.w.size());
  BS_CUDA(cudaMemcpyAsync(dw.p, c.reg.w.data(), sizeof(double) * c.reg.w.size(), cudaMemcpyHostToDevice, c.stream));
  const long long total = (long long)c.ncell * c.nq;
  const int bs_ = 256;
  k_cell_geometry<<<(unsigned)((total + bs_ - 1) / bs_), bs_, 0, c.stream>>>(c.ncell, c.nq, c.nq_pad, c.na_map, c.d_conn_map.p,
                                                                            c.d_map_nodes.p, c.d_map_tab_reg.p, dw.p,
                                                                            c.d_cellq.p, c.d_cellq8.p);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

// ---------------------------------------------------------------------------------------------------------
// K1: regular pass
// ---------------------------------------------------------------------------------------------------------
struct RegParams {
  int p0, p1, N;            // owned row positions [p0,p1), number of nodes
  int nq, nq_pad, tj;
  const double *support;    // [N][3]
  const int *conn_pos;      // [ncell][NA]
  const double *cellq;      // [ncell][7][nq_pad]
  const double *cellq8;     // [ncell][nq_pad][8] point-major, prescaled (free-space fast path)
  const double *l1d;        // [n1d][NB1] 1-D Lagrange values at the 1-D rule points (NB1 = degree+1)
  int n1d;
  const int *blk_cell_ptr, *blk_cells;
  const signed char *blk_slots;
  const int *blk_nodes;            // [nblocks][tj] node position per slot, -1 unused
  const unsigned char *blk_first;  // [nblocks][tj] first-touch flag
  const unsigned *blk_sync;        // [nblocks] cell-split mode: steps that start with a CTA barrier (bit = step)
  int blk_begin;                   // first block of the colour being launched
  double *V, *K;
  size_t ld;
  KernelParams kp;
  // fused mode: K is not stored; KX[row][pp] += K_tile * X  (X = panel [3N][pp])
  const double *panel;
  double *KX;
  int pp;
  // fused mode with mixed boundary conditions: -K of the flagged columns goes to the compact matrix Kflag
  const int *kcol;   // [3N] internal column -> compact column, -1 = not flagged; nullptr = no flags
  double *Kflag;
  size_t ldk;
};

// nodes per cell block in the cell-split mode of K1 (see cell_sets)
__host__ __device__ constexpr int tj_cell_split(int kernel_type) { return kernel_type == BS_KERNEL_FREE ? 15 : 20; }
constexpr int ACC_LD = TI + 1;  // padded row-node stride of the shared accumulators (bank-conflict free both ways)
// stride between consecutive values of the shared tile [value][slot][row]; the +5 spreads the three values a
// write-out lane group reads for one node over different banks
__host__ __device__ inline int acc_vstride(int tj) { return tj * ACC_LD + 5; }

// cell records in flight per CTA (bulk-copy ring): 3 when a record is small, else 2
#ifndef BS_CELL_STAGES
#define BS_CELL_STAGES 3
#endif
__host__ __device__ inline int cell_stages(int nq_pad) { return nq_pad <= 64 ? BS_CELL_STAGES : 2; }
// doubles reserved for the 1-D shape table and its x-flipped copy (2 * (degree+1) * n1d <= max(nq_pad, 96))
__host__ __device__ inline int l1d_doubles(int nq_pad) { return nq_pad > 96 ? nq_pad : 96; }

// Value planes of the shared tile: V and K together, except for the Q1 image kernels, which integrate the two layers
// in two launches (LAYER 1, 2) so that a 9-value tile still covers a 16-node block.
int tile_planes(int na, int kernel_type) {
  const int nv = (kernel_type == BS_KERNEL_FREE) ? 6 : 9;
  return (na == 4 && kernel_type != BS_KERNEL_FREE) ? nv : 2 * nv;
}

// Cell-split mode (cs == 2; Q1 unknowns, Gauss 8, no regularisation): one thread per collocation node integrates
// a whole cell (all rows of the rule, all four shape functions) and the CTA's two thread sets work on two cells of the
// block that share no node, so nothing is exchanged between threads and the per-cell tile update is paid once per
// 8 rule rows instead of once per 4.  A ring stage then holds a pair of cell records.
int cell_sets(const Context &c) {
  if (c.na != 4 || c.kp.eps != 0.0 || c.x1d.size() != 8 || std::getenv("BS_NO_CELLSPLIT")) return 1;
  if (c.kp.type == BS_KERNEL_NO_SLIP) return 2;                    // per-point Green evaluation, any mapping
  return (c.na_map == 4 && !std::getenv("BS_NO_LINROWS")) ? 2 : 1;  // free space / free surface: moment formulation
}
__host__ __device__ inline int ring_stages(int nq_pad, int cs) { return cs == 2 ? 2 : cell_stages(nq_pad); }

size_t assembly_smem_bytes(int na, int planes, int tj, int nq_pad, int cs) {
  (void)na;
  size_t ring = (size_t)ring_stages(nq_pad, cs) * cs * 8 * nq_pad * 8;
  ring = std::max(ring, (size_t)3 * tj * MAX_PANEL * 8);  // K-only launches stage the fused panel rows in the ring
  return (size_t)planes * acc_vstride(tj) * 8 + ring + (size_t)l1d_doubles(nq_pad) * 8 + 64;  // tile, ring, shape table, mbarriers
}

int choose_tj(int na, int kernel_type, int nq_pad, int cs) {
  const int planes = tile_planes(na, kernel_type);
  const size_t budget = (227 * 1024) / CTAS_PER_SM - 1024 - 1024;  // minus static arrays (768 B) / per-CTA reserve
  if (cs == 2) {  // cell-split blocks: 2 x 4 cells (15 nodes) with the 12-plane tile of the free-space kernel, 3 x 4 cells
                  // (20 nodes) with the 9-plane tiles of the image kernels; compile-time constants of those kernels
    const int t = tj_cell_split(kernel_type);
    BS_REQUIRE(assembly_smem_bytes(na, planes, t, nq_pad, cs) <= budget, "cell-split tile does not fit (BS_TI / BS_CTAS_PER_SM changed?)");
    return t;
  }
  int tj = 2;
  for (int t = 2; t <= 32; t += 2)
    if (assembly_smem_bytes(na, planes, t, nq_pad, cs) <= budget) tj = t;
  return tj;
}

// index of (i,j) in the stored value vector
template <int NV>
__device__ __forceinline__ int vidx(int i, int j) {
  if (NV == 9) return 3 * i + j;
  const int a = i < j ? i : j, b = i < j ? j : i;
  return a == 0 ? b : (a == 1 ? 2 + b : 5);  // (0,0)=0 (0,1)=1 (0,2)=2 (1,1)=3 (1,2)=4 (2,2)=5
}

// 1-D node index (ix, iy) of scalar shape function a: phi_a(q) = l_ix(x_q) * l_iy(y_q)  (deal.II FE_Q order)
template <int NA>
__device__ __forceinline__ constexpr int shape_ix(int a) {
  return NA == 4 ? (a & 1) : (a == 0 ? 0 : a == 1 ? 2 : a == 2 ? 0 : a == 3 ? 2 : a == 4 ? 0 : a == 5 ? 2 : 1);
}
template <int NA>
__device__ __forceinline__ constexpr int shape_iy(int a) {
  return NA == 4 ? (a >> 1) : (a == 0 ? 0 : a == 1 ? 0 : a == 2 ? 2 : a == 3 ? 2 : a == 4 ? 1 : a == 5 ? 1 : a == 6 ? 0 : a == 7 ? 2 : 1);
}

// Free-space fast path: one quadrature point of the point-major prescaled cell record (see K0).
//   G JxW = c3 R(x)R + c1 I,  -S.n JxW = ck R(x)R   with c1 = J' / r, c3 = c1 / r^2, ck = (R.6n) c3 / r^2.
// The x-direction sums are taken over the six products P = R(x)R with the weights l_b*c3 and l_b*ck (and the
// isotropic part l_b*c1 as a 13th scalar): 55 FP64 instructions per point for Q1 instead of 64.
template <int NB1, int MODE>
struct FreePoint {
  double P[6], sg[NB1], sk[NB1], c1;
};

template <int NB1, int MODE>
__device__ __forceinline__ void free_point(const double *__restrict__ c8, const double *__restrict__ lrow,
                                           const double (&x)[3], FreePoint<NB1, MODE> &pt) {
  const double2 a = *reinterpret_cast<const double2 *>(c8);
  const double2 b = *reinterpret_cast<const double2 *>(c8 + 2);
  const double Rx = a.x - x[0], Ry = a.y - x[1], Rz = b.x - x[2];
  const double r2 = fma(Rx, Rx, fma(Ry, Ry, Rz * Rz));
  const double ri = rsqrt_normal(r2);
  const double ri2 = ri * ri;
  const double c1 = b.y * ri;
  const double c3 = c1 * ri2;
  pt.c1 = c1;
  pt.P[0] = Rx * Rx;
  pt.P[1] = Rx * Ry;
  pt.P[2] = Rx * Rz;
  pt.P[3] = Ry * Ry;
  pt.P[4] = Ry * Rz;
  pt.P[5] = Rz * Rz;
  if (MODE != 1) {
#pragma unroll
    for (int bb = 0; bb < NB1; ++bb) pt.sg[bb] = lrow[bb] * c3;
  }
  if (MODE != 0) {
    const double2 n01 = *reinterpret_cast<const double2 *>(c8 + 4);
    const double n2 = c8[6];
    const double Rn = fma(Rx, n01.x, fma(Ry, n01.y, Rz * n2));
    const double ck = (Rn * ri2) * c3;
#pragma unroll
    for (int bb = 0; bb < NB1; ++bb) pt.sk[bb] = lrow[bb] * ck;
  }
}

template <int NB1, int MODE, int NACC>
__device__ __forceinline__ void free_accumulate(const FreePoint<NB1, MODE> &pt, const double *__restrict__ lrow,
                                                double (&tmp)[NB1][NACC], double (&tmpI)[NB1]) {
#pragma unroll
  for (int bb = 0; bb < NB1; ++bb) {
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      if (MODE == 2) {
        tmp[bb][v] = fma(pt.sg[bb], pt.P[v], tmp[bb][v]);
        tmp[bb][6 + v] = fma(pt.sk[bb], pt.P[v], tmp[bb][6 + v]);
      } else {
        tmp[bb][v] = fma(MODE == 0 ? pt.sg[bb] : pt.sk[bb], pt.P[v], tmp[bb][v]);
      }
    }
    if (MODE != 1) tmpI[bb] = fma(lrow[bb], pt.c1, tmpI[bb]);
  }
}

// Three-stage split of the same arithmetic for the fully unrolled rows: A = load, R, r^2, 1/r (and R.6n),
// B = coefficients c1, c3, ck, C = products R(x)R, weights and the accumulation FMAs.  With A of point q+2 and B of
// point q+1 in flight next to C of point q every dependent chain has two points' worth of independent work to
// hide behind (one warp per scheduler cannot rely on its neighbour for that).
struct FreeA {
  double R[3], ri, Jp, Rn;
};
struct FreeB {
  double R[3], c1, c3, ck;
};
template <int MODE>
__device__ __forceinline__ void free_stage_a(const double *__restrict__ c8, const double (&x)[3], FreeA &a) {
  const double2 u = *reinterpret_cast<const double2 *>(c8);
  const double2 v = *reinterpret_cast<const double2 *>(c8 + 2);
  a.R[0] = u.x - x[0];
  a.R[1] = u.y - x[1];
  a.R[2] = v.x - x[2];
  a.Jp = v.y;
  const double r2 = fma(a.R[0], a.R[0], fma(a.R[1], a.R[1], a.R[2] * a.R[2]));
  a.ri = rsqrt_normal(r2);
  a.Rn = 0.0;
  if (MODE != 0) {
    const double2 n01 = *reinterpret_cast<const double2 *>(c8 + 4);
    a.Rn = fma(a.R[0], n01.x, fma(a.R[1], n01.y, a.R[2] * c8[6]));
  }
}
template <int MODE>
__device__ __forceinline__ void free_stage_b(const FreeA &a, FreeB &b) {
  const double ri2 = a.ri * a.ri;
  b.R[0] = a.R[0];
  b.R[1] = a.R[1];
  b.R[2] = a.R[2];
  b.c1 = a.Jp * a.ri;
  b.c3 = b.c1 * ri2;
  b.ck = (MODE != 0) ? (a.Rn * ri2) * b.c3 : 0.0;
}
template <int NB1, int MODE, int NACC>
__device__ __forceinline__ void free_stage_c(const FreeB &b, const double *__restrict__ lrow, double (&tmp)[NB1][NACC],
                                             double (&tmpI)[NB1]) {
  const double P[6] = {b.R[0] * b.R[0], b.R[0] * b.R[1], b.R[0] * b.R[2], b.R[1] * b.R[1], b.R[1] * b.R[2], b.R[2] * b.R[2]};
#pragma unroll
  for (int bb = 0; bb < NB1; ++bb) {
    const double l = lrow[bb];
    const double sg = l * b.c3, sk = l * b.ck;
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      if (MODE == 2) {
        tmp[bb][v] = fma(sg, P[v], tmp[bb][v]);
        tmp[bb][6 + v] = fma(sk, P[v], tmp[bb][6 + v]);
      } else {
        tmp[bb][v] = fma(MODE == 0 ? sg : sk, P[v], tmp[bb][v]);
      }
    }
    if (MODE != 1) tmpI[bb] = fma(l, b.c1, tmpI[bb]);
  }
}

// Software-pipelined integration of one (row, cell) pair with the free-space kernel: the dependent chain of
// point q+1 (load, r^2, rsqrt, coefficients) is issued together with the independent accumulation FMAs of point
// q, across the rows of the rule as well.  N1C > 0: compile-time 1-D rule size (fully unrolled rows).
//
// `lx_s` is the 1-D shape table used in the x direction.  For Q1 with two threads per row the odd thread gets the
// table with its two columns swapped: its accumulator slot s then holds shape function s^1, so that the pair
// exchanges only the slots the partner finalises (half the shuffles, see cell_pass).
template <int NA, int MODE, int QS, int N1C, int NACC>
__device__ __forceinline__ void integrate_free(const double *__restrict__ c8, const double *__restrict__ lx_s,
                                               const double *__restrict__ ly_s, int n1rt, const double (&x)[3], int part,
                                               double (&acc)[NA][NACC]) {
  constexpr int NB1 = (NA == 4) ? 2 : 3;
  const int n1 = N1C > 0 ? N1C : n1rt;
  const double *l1d_s = lx_s;
  double accI[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) accI[a] = 0.0;
  FreePoint<NB1, MODE> cur;  // two-stage pipeline state (run-time rule size)
  FreeA sa;                  // three-stage pipeline state (compile-time rule size)
  FreeB sb;
  if (part < n1) {
    if (N1C > 0) {
      FreeA a0;
      free_stage_a<MODE>(c8 + (size_t)8 * part * n1, x, a0);
      free_stage_b<MODE>(a0, sb);
      free_stage_a<MODE>(c8 + (size_t)8 * (part * n1 + 1), x, sa);
    } else {
      free_point<NB1, MODE>(c8 + (size_t)8 * part * n1, l1d_s, x, cur);
    }
  }
  for (int qy = part; qy < n1; qy += QS) {
    double tmp[NB1][NACC], tmpI[NB1];
#pragma unroll
    for (int bb = 0; bb < NB1; ++bb) {
      tmpI[bb] = 0.0;
#pragma unroll
      for (int v = 0; v < NACC; ++v) tmp[bb][v] = 0.0;
    }
    const double *crow = c8 + (size_t)8 * qy * n1;
    const int qyn = (qy + QS < n1) ? qy + QS : qy;  // this thread's next row (or a harmless re-read at the end)
    if (N1C > 0) {
      const double *nrow = c8 + (size_t)8 * qyn * n1;
#pragma unroll
      for (int qx = 0; qx < N1C; ++qx) {
        FreeB nb;
        free_stage_b<MODE>(sa, nb);                                                         // point qx+1
        FreeA na;
        free_stage_a<MODE>(qx + 2 < N1C ? crow + 8 * (qx + 2) : nrow + 8 * (qx + 2 - N1C), x, na);  // point qx+2
        free_stage_c<NB1, MODE, NACC>(sb, l1d_s + qx * NB1, tmp, tmpI);                     // point qx
        sb = nb;
        sa = na;
      }
    } else {
#pragma unroll 2
      for (int qx = 0; qx + 1 < n1; ++qx) {
        FreePoint<NB1, MODE> nxt;
        free_point<NB1, MODE>(crow + 8 * (qx + 1), l1d_s + (qx + 1) * NB1, x, nxt);
        free_accumulate<NB1, MODE, NACC>(cur, l1d_s + qx * NB1, tmp, tmpI);
        cur = nxt;
      }
      // last point of the row; the chain of the first point of this thread's next row rides along
      FreePoint<NB1, MODE> nxt;
      free_point<NB1, MODE>(c8 + (size_t)8 * qyn * n1, l1d_s, x, nxt);
      free_accumulate<NB1, MODE, NACC>(cur, l1d_s + (n1 - 1) * NB1, tmp, tmpI);
      cur = nxt;
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const double l = ly_s[qy * NB1 + shape_iy<NA>(a)];
#pragma unroll
      for (int v = 0; v < NACC; ++v) acc[a][v] = fma(tmp[shape_ix<NA>(a)][v], l, acc[a][v]);
      if (MODE != 1) accI[a] = fma(tmpI[shape_ix<NA>(a)], l, accI[a]);
    }
  }
  if (MODE != 1) {
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      acc[a][0] += accI[a];
      acc[a][3] += accI[a];
      acc[a][5] += accI[a];
    }
  }
}

// Free-space kernel, Q1 unknowns on a bilinear (Q1) mapping, Gauss 8 (N1C == -8): rows of the tensor rule are straight
// lines, y(xi, eta_q) = y(0, eta_q) + xi b(eta_q), hence R = R0 + xi B and every product R_i R_j is the quadratic
// A_v + xi C_v + xi^2 D_v with A = R0 (x) R0, C = R0 (x) B + B (x) R0, D = B (x) B constant along the row.  With
// l_0 = 1 - xi, l_1 = xi the x-direction sums sum_q c_q l_b(xi_q) R_i R_j collapse to the four scalar moments
// m_k = sum_q c_q xi_q^k (k = 0..3) per layer: 10 FMAs per point (4 + 4 + 2 for the isotropic part) instead of 6
// products and 26 FMAs, and one expansion per row (3 FMAs per value and shape function).  r^2, 1/r and R.n keep
// the per-point R = y_q - x, so the coefficients c_q are the same numbers as before; per point 29 FP64 instructions
// instead of 55, per row of 8 points about 380 instead of 490.
//   xi_s: [8][4] = xi_q, xi_q^2, xi_q^3, - ; xi_s[32] = xi_0, xi_s[33] = 1 / (xi_7 - xi_0)
template <int MODE>
__device__ __forceinline__ void free_stage_c_lin(const FreeB &b, const double *__restrict__ xq, double (&mg)[4], double (&mk)[4],
                                                 double (&mi)[2]) {
  const double2 x12 = *reinterpret_cast<const double2 *>(xq);
  const double x3 = xq[2];
  if (MODE != 1) {
    mg[0] += b.c3;
    mg[1] = fma(b.c3, x12.x, mg[1]);
    mg[2] = fma(b.c3, x12.y, mg[2]);
    mg[3] = fma(b.c3, x3, mg[3]);
    mi[0] += b.c1;
    mi[1] = fma(b.c1, x12.x, mi[1]);
  }
  if (MODE != 0) {
    mk[0] += b.ck;
    mk[1] = fma(b.ck, x12.x, mk[1]);
    mk[2] = fma(b.ck, x12.y, mk[2]);
    mk[3] = fma(b.ck, x3, mk[3]);
  }
}

template <int MODE, int QS, int NACC>
__device__ __forceinline__ void integrate_free_lin(const double *__restrict__ c8, const double *__restrict__ xi_s,
                                                   const double *__restrict__ ly_s, const double (&x)[3], int part,
                                                   double (&acc)[4][NACC]) {
  constexpr int N1 = 8;
  constexpr int KO = (MODE == 2) ? 6 : 0;  // offset of the double-layer values in acc
  const bool flip = (QS == 2) && (part == 1);  // odd partner: accumulator slot s holds shape function s^1 (see cell_pass)
  const double xi0 = xi_s[32], inv_dxi = xi_s[33];
  double accI[4] = {0.0, 0.0, 0.0, 0.0};
  FreeA sa;
  FreeB sb;
  if (part < N1) {
    FreeA a0;
    free_stage_a<MODE>(c8 + (size_t)8 * part * N1, x, a0);
    free_stage_b<MODE>(a0, sb);
    free_stage_a<MODE>(c8 + (size_t)8 * (part * N1 + 1), x, sa);
  }
  for (int qy = part; qy < N1; qy += QS) {
    double mg[4] = {0.0, 0.0, 0.0, 0.0}, mk[4] = {0.0, 0.0, 0.0, 0.0}, mi[2] = {0.0, 0.0};
    const double *crow = c8 + (size_t)8 * qy * N1;
    const int qyn = (qy + QS < N1) ? qy + QS : qy;  // this thread's next row (or a harmless re-read at the end)
    const double *nrow = c8 + (size_t)8 * qyn * N1;
#pragma unroll
    for (int qx = 0; qx < N1; ++qx) {
      FreeB nb;
      free_stage_b<MODE>(sa, nb);                                                                  // point qx+1
      FreeA na;
      free_stage_a<MODE>(qx + 2 < N1 ? crow + 8 * (qx + 2) : nrow + 8 * (qx + 2 - N1), x, na);      // point qx+2
      free_stage_c_lin<MODE>(sb, xi_s + 4 * qx, mg, mk, mi);                                       // point qx
      sb = nb;
      sa = na;
    }
    // ---- the row's straight line and the expansion of the moments
    double R0[3], B[3];
    {
      const double2 u0 = *reinterpret_cast<const double2 *>(crow), u7 = *reinterpret_cast<const double2 *>(crow + 8 * (N1 - 1));
      const double z0 = crow[2], z7 = crow[8 * (N1 - 1) + 2];
      B[0] = (u7.x - u0.x) * inv_dxi;
      B[1] = (u7.y - u0.y) * inv_dxi;
      B[2] = (z7 - z0) * inv_dxi;
      R0[0] = fma(-xi0, B[0], u0.x - x[0]);
      R0[1] = fma(-xi0, B[1], u0.y - x[1]);
      R0[2] = fma(-xi0, B[2], z0 - x[2]);
    }
    // moments of l_b xi^k: l_0 = 1 - xi, l_1 = xi; slot order swapped for the odd partner
    double Mg[3][2], Mk[3][2], Mi[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (MODE != 1) {
        const double d = mg[k] - mg[k + 1], e = mg[k + 1];
        Mg[k][0] = flip ? e : d;
        Mg[k][1] = flip ? d : e;
      }
      if (MODE != 0) {
        const double d = mk[k] - mk[k + 1], e = mk[k + 1];
        Mk[k][0] = flip ? e : d;
        Mk[k][1] = flip ? d : e;
      }
    }
    if (MODE != 1) {
      const double d = mi[0] - mi[1], e = mi[1];
      Mi[0] = flip ? e : d;
      Mi[1] = flip ? d : e;
    }
    const double ly0 = ly_s[qy * 2], ly1 = ly_s[qy * 2 + 1];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) {
        const int v = (i == 0) ? j : (i == 1 ? 2 + j : 5);
        const double Av = R0[i] * R0[j];
        const double Cv = (i == j) ? 2.0 * (R0[i] * B[i]) : fma(R0[i], B[j], B[i] * R0[j]);
        const double Dv = B[i] * B[j];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (MODE != 1) {
            const double tg = fma(Dv, Mg[2][b], fma(Cv, Mg[1][b], Av * Mg[0][b]));
            acc[b][v] = fma(tg, ly0, acc[b][v]);
            acc[b + 2][v] = fma(tg, ly1, acc[b + 2][v]);
          }
          if (MODE != 0) {
            const double tk = fma(Dv, Mk[2][b], fma(Cv, Mk[1][b], Av * Mk[0][b]));
            acc[b][KO + v] = fma(tk, ly0, acc[b][KO + v]);
            acc[b + 2][KO + v] = fma(tk, ly1, acc[b + 2][KO + v]);
          }
        }
      }
    if (MODE != 1) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        accI[b] = fma(Mi[b], ly0, accI[b]);
        accI[b + 2] = fma(Mi[b], ly1, accI[b + 2]);
      }
    }
  }
  if (MODE != 1) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      acc[a][0] += accI[a];
      acc[a][3] += accI[a];
      acc[a][5] += accI[a];
    }
  }
}

// Expansion of the 2-D moments of one layer (see integrate_free_lin2d): out[a][v] = sum_km P_v[k][m] N_a[k][m] with
// N_a[k][m] = sum_q c_q phi_a xi^k eta^m (shape function a = ix + 2 iy, phi_a = l_ix(xi) l_iy(eta), l_0 = 1 - t, l_1 = t)
// and P_v the 9 coefficients of R_i R_j on the bilinear cell, R00 = y00 - x; cell constants from the record's pad slots.
// iso != nullptr: the 2 x 2 moments of the isotropic part c_1, added to the diagonal values.  The sums are formed in
// registers (24 independent chains of 9 FMAs; the cell constants are plain loads with no store in between).
__device__ __forceinline__ void expand_moments2d(const double (&M)[4][4], const double (*iso)[2], const double (&R00)[3],
                                                 const double *__restrict__ c8, double (&out)[4][6]) {
  auto cst = [&](int k) { return c8[8 * k + 7]; };
  double A[3], B[3], C[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    A[d] = cst(3 + d);
    B[d] = cst(6 + d);
    C[d] = cst(9 + d);
  }
  double Ni[4] = {0.0, 0.0, 0.0, 0.0};
  if (iso) {
    Ni[3] = iso[1][1];
    Ni[2] = iso[0][1] - iso[1][1];
    Ni[1] = iso[1][0] - iso[1][1];
    Ni[0] = (iso[0][0] - iso[1][0]) - Ni[2];
  }
  double N[4][3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double X0[4], X1[4];  // l_0(xi) xi^k, l_1(xi) xi^k against eta^m, m <= 3
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      X1[m] = M[k + 1][m];
      X0[m] = M[k][m] - M[k + 1][m];
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      N[2][k][m] = X0[m + 1];
      N[0][k][m] = X0[m] - X0[m + 1];
      N[3][k][m] = X1[m + 1];
      N[1][k][m] = X1[m] - X1[m + 1];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) {
      const int v = (i == 0) ? j : (i == 1 ? 2 + j : 5);
      double P[3][3];
      P[0][0] = R00[i] * R00[j];
      P[1][0] = (i == j) ? 2.0 * (R00[i] * A[i]) : fma(R00[i], A[j], A[i] * R00[j]);
      P[0][1] = (i == j) ? 2.0 * (R00[i] * B[i]) : fma(R00[i], B[j], B[i] * R00[j]);
      P[1][1] = fma(R00[i], C[j], fma(C[i], R00[j], cst(12 + 6 * v)));
      P[2][0] = cst(12 + 6 * v + 1);
      P[0][2] = cst(12 + 6 * v + 2);
      P[2][1] = cst(12 + 6 * v + 3);
      P[1][2] = cst(12 + 6 * v + 4);
      P[2][2] = cst(12 + 6 * v + 5);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double o = P[0][0] * N[a][0][0];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int m = 0; m < 3; ++m)
            if (k + m > 0) o = fma(P[k][m], N[a][k][m], o);
        if (iso && i == j) o += Ni[a];
        out[a][v] = o;
      }
    }
}

// Cell-split mode (one thread integrates the whole cell): the moment formulation in both directions.  On a bilinear cell
// R = R00 + xi a + eta b + xi eta c, so every product R_i R_j is a polynomial with 9 coefficients P_km (k, m <= 2) in
// (xi, eta), and with the bilinear shape functions the cell integrals are contractions of P with the 16 scalar moments
// M_km = sum_q c_q xi_q^k eta_q^m (k, m <= 3) per layer.  Per rule row: the four xi-moments of integrate_free_lin, then
// 16 FMAs per layer into M (instead of the 60-FMA expansion of the row); per cell: one expansion, 9 FMAs per value and
// shape function, added to the thread's tile entries.  Cell constants come from the pad slots of the record (K0).
template <int MODE>
__device__ __forceinline__ void integrate_free_lin2d(const double *__restrict__ c8, const double *__restrict__ xi_s,
                                                     const double (&x)[3], double *const (&dst)[4], int vs) {
  constexpr int N1 = 8;
  constexpr int KO = (MODE == 2) ? 6 : 0;  // first tile plane (relative to dst) of the double layer
  double Mg[4][4], Mk[4][4], Mi[2][2];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int m = 0; m < 4; ++m) Mg[k][m] = Mk[k][m] = 0.0;
  Mi[0][0] = Mi[0][1] = Mi[1][0] = Mi[1][1] = 0.0;
  auto add_row = [&](const double (&mg)[4], const double (&mk)[4], const double (&mi)[2], int qy) {
    const double2 e12 = *reinterpret_cast<const double2 *>(xi_s + 4 * qy);  // eta, eta^2 (same 1-D rule)
    const double e3 = xi_s[4 * qy + 2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (MODE != 1) {
        Mg[k][0] += mg[k];
        Mg[k][1] = fma(mg[k], e12.x, Mg[k][1]);
        Mg[k][2] = fma(mg[k], e12.y, Mg[k][2]);
        Mg[k][3] = fma(mg[k], e3, Mg[k][3]);
      }
      if (MODE != 0) {
        Mk[k][0] += mk[k];
        Mk[k][1] = fma(mk[k], e12.x, Mk[k][1]);
        Mk[k][2] = fma(mk[k], e12.y, Mk[k][2]);
        Mk[k][3] = fma(mk[k], e3, Mk[k][3]);
      }
    }
    if (MODE != 1) {
      Mi[0][0] += mi[0];
      Mi[0][1] = fma(mi[0], e12.x, Mi[0][1]);
      Mi[1][0] += mi[1];
      Mi[1][1] = fma(mi[1], e12.x, Mi[1][1]);
    }
  };
#if BS_ROWS2
  // NR rule rows at a time: independent point pipelines per thread (more instruction-level parallelism
  // for the dependent chains of stage A / B; one warp per scheduler and CTA cannot rely on its neighbour for that)
  constexpr int NR = (BS_ROWS2 == 1) ? 2 : BS_ROWS2;
  static_assert(N1 % NR == 0, "rows in flight must divide the rule size");
  FreeA sa[NR];
  FreeB sb[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    FreeA a0;
    free_stage_a<MODE>(c8 + 8 * N1 * u, x, a0);
    free_stage_b<MODE>(a0, sb[u]);
    free_stage_a<MODE>(c8 + 8 * N1 * u + 8, x, sa[u]);
  }
  for (int qy = 0; qy < N1; qy += NR) {
    double mg[NR][4], mk[NR][4], mi[NR][2];
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      mi[u][0] = mi[u][1] = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) mg[u][k] = mk[u][k] = 0.0;
    }
    const double *crow = c8 + (size_t)8 * qy * N1;
    const int qn = (qy + NR < N1) ? qy + NR : qy;  // the pipelines' next rows (harmless re-read at the end)
    const double *nrow = c8 + (size_t)8 * qn * N1;
#pragma unroll
    for (int qx = 0; qx < N1; ++qx) {
      FreeB nb[NR];
      FreeA na[NR];
#pragma unroll
      for (int u = 0; u < NR; ++u) free_stage_b<MODE>(sa[u], nb[u]);
#pragma unroll
      for (int u = 0; u < NR; ++u)
        free_stage_a<MODE>((qx + 2 < N1 ? crow + 8 * (qx + 2) : nrow + 8 * (qx + 2 - N1)) + 8 * N1 * u, x, na[u]);
#pragma unroll
      for (int u = 0; u < NR; ++u) free_stage_c_lin<MODE>(sb[u], xi_s + 4 * qx, mg[u], mk[u], mi[u]);
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        sb[u] = nb[u];
        sa[u] = na[u];
      }
    }
#pragma unroll
    for (int u = 0; u < NR; ++u) add_row(mg[u], mk[u], mi[u], qy + u);
  }
#else
  FreeA sa;
  FreeB sb;
  {
    FreeA a0;
    free_stage_a<MODE>(c8, x, a0);
    free_stage_b<MODE>(a0, sb);
    free_stage_a<MODE>(c8 + 8, x, sa);
  }
  for (int qy = 0; qy < N1; ++qy) {
    double mg[4] = {0.0, 0.0, 0.0, 0.0}, mk[4] = {0.0, 0.0, 0.0, 0.0}, mi[2] = {0.0, 0.0};
    const double *crow = c8 + (size_t)8 * qy * N1;
    const double *nrow = c8 + (size_t)8 * ((qy + 1 < N1) ? qy + 1 : qy) * N1;  // next row (harmless re-read at the end)
#pragma unroll
    for (int qx = 0; qx < N1; ++qx) {
      FreeB nb;
      free_stage_b<MODE>(sa, nb);                                                                  // point qx+1
      FreeA na;
      free_stage_a<MODE>(qx + 2 < N1 ? crow + 8 * (qx + 2) : nrow + 8 * (qx + 2 - N1), x, na);      // point qx+2
      free_stage_c_lin<MODE>(sb, xi_s + 4 * qx, mg, mk, mi);                                       // point qx
      sb = nb;
      sa = na;
    }
    add_row(mg, mk, mi, qy);
  }
#endif
  // ---- expansion, one layer at a time; the sums are added to the thread's tile entries in one pass per layer
  const double R00[3] = {c8[7] - x[0], c8[15] - x[1], c8[23] - x[2]};
  double out[4][6];
  if (MODE != 1) {
    expand_moments2d(Mg, Mi, R00, c8, out);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int v = 0; v < 6; ++v) dst[a][(size_t)v * vs] += out[a][v];
  }
  if (MODE != 0) {
    expand_moments2d(Mk, nullptr, R00, c8, out);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int v = 0; v < 6; ++v) dst[a][(size_t)(KO + v) * vs] += out[a][v];
  }
}

// Free-surface image system, cell-split mode: the direct part (R = y - x) and the image part (R = y - x_im) are two
// free-space kernels on the same cell, one layer per launch (MODE 0 or 1): two sets of 2-D moments, two expansions that
// differ in R00 only, combined with the sign of the image term into the 9 unsymmetric tile values
// (ref: source/free_surface_kernel.cc:19-72, 135-209).
template <int MODE>
__device__ __forceinline__ void integrate_free_surface_lin2d(const double *__restrict__ c8, const double *__restrict__ xi_s,
                                                             const double (&x)[3], const double (&xim)[3], int o,
                                                             double *const (&dst)[4], int vs) {
  static_assert(MODE == 0 || MODE == 1, "one layer per launch");
  constexpr int N1 = 8;
  double MA[4][4], MB[4][4], IA[2][2], IB[2][2];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int m = 0; m < 4; ++m) MA[k][m] = MB[k][m] = 0.0;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int m = 0; m < 2; ++m) IA[k][m] = IB[k][m] = 0.0;
  FreeA saA, saB;
  FreeB sbA, sbB;
  {
    FreeA a0;
    free_stage_a<MODE>(c8, x, a0);
    free_stage_b<MODE>(a0, sbA);
    free_stage_a<MODE>(c8, xim, a0);
    free_stage_b<MODE>(a0, sbB);
    free_stage_a<MODE>(c8 + 8, x, saA);
    free_stage_a<MODE>(c8 + 8, xim, saB);
  }
  for (int qy = 0; qy < N1; ++qy) {
    double mA[4] = {0.0, 0.0, 0.0, 0.0}, mB[4] = {0.0, 0.0, 0.0, 0.0}, iA[2] = {0.0, 0.0}, iB[2] = {0.0, 0.0};
    double dummy[4] = {0.0, 0.0, 0.0, 0.0};
    const double *crow = c8 + (size_t)8 * qy * N1;
    const double *nrow = c8 + (size_t)8 * ((qy + 1 < N1) ? qy + 1 : qy) * N1;
#pragma unroll
    for (int qx = 0; qx < N1; ++qx) {
      FreeB nbA, nbB;
      free_stage_b<MODE>(saA, nbA);
      free_stage_b<MODE>(saB, nbB);
      FreeA naA, naB;
      const double *rec = qx + 2 < N1 ? crow + 8 * (qx + 2) : nrow + 8 * (qx + 2 - N1);
      free_stage_a<MODE>(rec, x, naA);
      free_stage_a<MODE>(rec, xim, naB);
      if (MODE == 0) {
        free_stage_c_lin<0>(sbA, xi_s + 4 * qx, mA, dummy, iA);
        free_stage_c_lin<0>(sbB, xi_s + 4 * qx, mB, dummy, iB);
      } else {
        free_stage_c_lin<1>(sbA, xi_s + 4 * qx, dummy, mA, iA);
        free_stage_c_lin<1>(sbB, xi_s + 4 * qx, dummy, mB, iB);
      }
      sbA = nbA;
      sbB = nbB;
      saA = naA;
      saB = naB;
    }
    const double2 e12 = *reinterpret_cast<const double2 *>(xi_s + 4 * qy);
    const double e3 = xi_s[4 * qy + 2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      MA[k][0] += mA[k];
      MA[k][1] = fma(mA[k], e12.x, MA[k][1]);
      MA[k][2] = fma(mA[k], e12.y, MA[k][2]);
      MA[k][3] = fma(mA[k], e3, MA[k][3]);
      MB[k][0] += mB[k];
      MB[k][1] = fma(mB[k], e12.x, MB[k][1]);
      MB[k][2] = fma(mB[k], e12.y, MB[k][2]);
      MB[k][3] = fma(mB[k], e3, MB[k][3]);
    }
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        IA[k][0] += iA[k];
        IA[k][1] = fma(iA[k], e12.x, IA[k][1]);
        IB[k][0] += iB[k];
        IB[k][1] = fma(iB[k], e12.x, IB[k][1]);
      }
    }
  }
  const double R00a[3] = {c8[7] - x[0], c8[15] - x[1], c8[23] - x[2]};
  const double R00b[3] = {c8[7] - xim[0], c8[15] - xim[1], c8[23] - xim[2]};
  double oa[4][6], ob[4][6];
  expand_moments2d(MA, MODE == 0 ? IA : nullptr, R00a, c8, oa);
  expand_moments2d(MB, MODE == 0 ? IB : nullptr, R00b, c8, ob);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double sg = (i == o) ? -1.0 : 1.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) dst[a][(size_t)(3 * i + j) * vs] += fma(sg, ob[a][vidx<6>(i, j)], oa[a][vidx<6>(i, j)]);
    }
}


// Free-surface image system on the same fast path: G_fs = G(R) + s_i G(R_im), K likewise, with s_i = -1 on the row of
// the wall normal and +1 otherwise (ref: source/free_surface_kernel.cc:19-72, 135-209).  Both terms are free-space
// kernels, so one layer at a time (MODE 0 or 1, the layer-split launches) is accumulated as two symmetric 6-vectors
// (direct and image part) with the 55-instruction formulation; the sign is applied once per cell when the 2 x 6
// sums are expanded to the 9 unsymmetric tile values.  Two-stage software pipeline over the points of a rule row.
template <int NA, int MODE, int QS>
__device__ __forceinline__ void integrate_free_surface(const double *__restrict__ c8, const double *__restrict__ lx_s,
                                                       const double *__restrict__ ly_s, int n1, const double (&x)[3],
                                                       const double (&xim)[3], int o, int part, double (&out)[NA][9]) {
  constexpr int NB1 = (NA == 4) ? 2 : 3;
  double acc[NA][12], accI[NA][2];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    accI[a][0] = accI[a][1] = 0.0;
#pragma unroll
    for (int v = 0; v < 12; ++v) acc[a][v] = 0.0;
  }
  auto point = [&](const double *rec, FreeB &pa, FreeB &pb) {
    FreeA a;
    free_stage_a<MODE>(rec, x, a);
    free_stage_b<MODE>(a, pa);
    free_stage_a<MODE>(rec, xim, a);
    free_stage_b<MODE>(a, pb);
  };
  FreeB ca, cb;
  if (part < n1) point(c8 + (size_t)8 * part * n1, ca, cb);
  for (int qy = part; qy < n1; qy += QS) {
    double tA[NB1][6], tB[NB1][6], tIA[NB1], tIB[NB1];
#pragma unroll
    for (int bb = 0; bb < NB1; ++bb) {
      tIA[bb] = tIB[bb] = 0.0;
#pragma unroll
      for (int v = 0; v < 6; ++v) tA[bb][v] = tB[bb][v] = 0.0;
    }
    const double *crow = c8 + (size_t)8 * qy * n1;
    const int qyn = (qy + QS < n1) ? qy + QS : qy;
    const double *nrow = c8 + (size_t)8 * qyn * n1;
#pragma unroll 2
    for (int qx = 0; qx < n1; ++qx) {
      FreeB na_, nb_;
      point(qx + 1 < n1 ? crow + 8 * (qx + 1) : nrow, na_, nb_);  // next point (first of this thread's next row at the end)
      free_stage_c<NB1, MODE, 6>(ca, lx_s + qx * NB1, tA, tIA);
      free_stage_c<NB1, MODE, 6>(cb, lx_s + qx * NB1, tB, tIB);
      ca = na_;
      cb = nb_;
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const double l = ly_s[qy * NB1 + shape_iy<NA>(a)];
#pragma unroll
      for (int v = 0; v < 6; ++v) {
        acc[a][v] = fma(tA[shape_ix<NA>(a)][v], l, acc[a][v]);
        acc[a][6 + v] = fma(tB[shape_ix<NA>(a)][v], l, acc[a][6 + v]);
      }
      if (MODE == 0) {
        accI[a][0] = fma(tIA[shape_ix<NA>(a)], l, accI[a][0]);
        accI[a][1] = fma(tIB[shape_ix<NA>(a)], l, accI[a][1]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double sg = (i == o) ? -1.0 : 1.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double v = fma(sg, acc[a][6 + vidx<6>(i, j)], acc[a][vidx<6>(i, j)]);
        if (MODE == 0 && i == j) v += fma(sg, accI[a][1], accI[a][0]);
        out[a][3 * i + j] += v;  // the caller's accumulators (zero, or the tile values in the cell-split mode)
      }
    }
}

// Free-surface image system with the moment formulation of integrate_free_lin (Q1 on bilinear cells, Gauss 8): the direct
// part R = y - x and the image part R_im = y - x_im are both free-space kernels along the same straight rows (same B,
// different R0), one layer per launch (MODE 0 or 1).  Three-stage software pipeline over the points as in integrate_free.
template <int MODE, int QS>
__device__ __forceinline__ void integrate_free_surface_lin(const double *__restrict__ c8, const double *__restrict__ xi_s,
                                                           const double *__restrict__ ly_s, const double (&x)[3],
                                                           const double (&xim)[3], int o, int part, double (&out)[4][9]) {
  static_assert(MODE == 0 || MODE == 1, "one layer per launch");
  constexpr int N1 = 8;
  const bool flip = (QS == 2) && (part == 1);
  const double xi0 = xi_s[32], inv_dxi = xi_s[33];
  double acc[4][12], accI[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    accI[a][0] = accI[a][1] = 0.0;
#pragma unroll
    for (int v = 0; v < 12; ++v) acc[a][v] = 0.0;
  }
  FreeA saA, saB;
  FreeB sbA, sbB;
  if (part < N1) {
    FreeA a0;
    const double *r0 = c8 + (size_t)8 * part * N1;
    free_stage_a<MODE>(r0, x, a0);
    free_stage_b<MODE>(a0, sbA);
    free_stage_a<MODE>(r0, xim, a0);
    free_stage_b<MODE>(a0, sbB);
    free_stage_a<MODE>(r0 + 8, x, saA);
    free_stage_a<MODE>(r0 + 8, xim, saB);
  }
  for (int qy = part; qy < N1; qy += QS) {
    double mA[4] = {0.0, 0.0, 0.0, 0.0}, mB[4] = {0.0, 0.0, 0.0, 0.0}, iA[2] = {0.0, 0.0}, iB[2] = {0.0, 0.0};
    double dummy[4] = {0.0, 0.0, 0.0, 0.0};
    const double *crow = c8 + (size_t)8 * qy * N1;
    const int qyn = (qy + QS < N1) ? qy + QS : qy;
    const double *nrow = c8 + (size_t)8 * qyn * N1;
#pragma unroll
    for (int qx = 0; qx < N1; ++qx) {
      FreeB nbA, nbB;
      free_stage_b<MODE>(saA, nbA);
      free_stage_b<MODE>(saB, nbB);
      FreeA naA, naB;
      const double *rec = qx + 2 < N1 ? crow + 8 * (qx + 2) : nrow + 8 * (qx + 2 - N1);
      free_stage_a<MODE>(rec, x, naA);
      free_stage_a<MODE>(rec, xim, naB);
      if (MODE == 0) {
        free_stage_c_lin<0>(sbA, xi_s + 4 * qx, mA, dummy, iA);
        free_stage_c_lin<0>(sbB, xi_s + 4 * qx, mB, dummy, iB);
      } else {
        free_stage_c_lin<1>(sbA, xi_s + 4 * qx, dummy, mA, iA);
        free_stage_c_lin<1>(sbB, xi_s + 4 * qx, dummy, mB, iB);
      }
      sbA = nbA;
      sbB = nbB;
      saA = naA;
      saB = naB;
    }
    double R0a[3], R0b[3], B[3];
    {
      const double2 u0 = *reinterpret_cast<const double2 *>(crow), u7 = *reinterpret_cast<const double2 *>(crow + 8 * (N1 - 1));
      const double z0 = crow[2], z7 = crow[8 * (N1 - 1) + 2];
      B[0] = (u7.x - u0.x) * inv_dxi;
      B[1] = (u7.y - u0.y) * inv_dxi;
      B[2] = (z7 - z0) * inv_dxi;
      const double y0[3] = {u0.x, u0.y, z0};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        R0a[d] = fma(-xi0, B[d], y0[d] - x[d]);
        R0b[d] = fma(-xi0, B[d], y0[d] - xim[d]);
      }
    }
    double MA[3][2], MB[3][2], IA[2], IB[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dA = mA[k] - mA[k + 1], eA = mA[k + 1], dB = mB[k] - mB[k + 1], eB = mB[k + 1];
      MA[k][0] = flip ? eA : dA;
      MA[k][1] = flip ? dA : eA;
      MB[k][0] = flip ? eB : dB;
      MB[k][1] = flip ? dB : eB;
    }
    if (MODE == 0) {
      const double dA = iA[0] - iA[1], eA = iA[1], dB = iB[0] - iB[1], eB = iB[1];
      IA[0] = flip ? eA : dA;
      IA[1] = flip ? dA : eA;
      IB[0] = flip ? eB : dB;
      IB[1] = flip ? dB : eB;
    }
    const double ly0 = ly_s[qy * 2], ly1 = ly_s[qy * 2 + 1];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) {
        const int v = (i == 0) ? j : (i == 1 ? 2 + j : 5);
        const double Dv = B[i] * B[j];
        const double AvA = R0a[i] * R0a[j], AvB = R0b[i] * R0b[j];
        const double CvA = (i == j) ? 2.0 * (R0a[i] * B[i]) : fma(R0a[i], B[j], B[i] * R0a[j]);
        const double CvB = (i == j) ? 2.0 * (R0b[i] * B[i]) : fma(R0b[i], B[j], B[i] * R0b[j]);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const double tA = fma(Dv, MA[2][b], fma(CvA, MA[1][b], AvA * MA[0][b]));
          const double tB = fma(Dv, MB[2][b], fma(CvB, MB[1][b], AvB * MB[0][b]));
          acc[b][v] = fma(tA, ly0, acc[b][v]);
          acc[b + 2][v] = fma(tA, ly1, acc[b + 2][v]);
          acc[b][6 + v] = fma(tB, ly0, acc[b][6 + v]);
          acc[b + 2][6 + v] = fma(tB, ly1, acc[b + 2][6 + v]);
        }
      }
    if (MODE == 0) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        accI[b][0] = fma(IA[b], ly0, accI[b][0]);
        accI[b + 2][0] = fma(IA[b], ly1, accI[b + 2][0]);
        accI[b][1] = fma(IB[b], ly0, accI[b][1]);
        accI[b + 2][1] = fma(IB[b], ly1, accI[b + 2][1]);
      }
    }
  }
  // G_fs = G(R) + s_i G(R_im), s_i = -1 on the row of the wall normal (ref: source/free_surface_kernel.cc:19-72, 135-209)
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double sg = (i == o) ? -1.0 : 1.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double v = fma(sg, acc[a][6 + vidx<6>(i, j)], acc[a][vidx<6>(i, j)]);
        if (MODE == 0 && i == j) v += fma(sg, accI[a][1], accI[a][0]);
        out[a][3 * i + j] += v;  // the caller's accumulators (zero, or the tile values in the cell-split mode)
      }
    }
}

// No-slip wall (Blake-type image system exactly as coded in the reference, source/no_slip_wall_kernel.cc:23-116, 127-199;
// green_eval<BS_KERNEL_NO_SLIP> is the literal transcription), cell-split mode.  With R = y - x, Q = y - x_im, h0 the wall
// distance of x, s = 2 h0 (h0 - Q_o), e_i = -1 on the row of the wall normal and +1 otherwise, the entries are
//   G_ij 8 pi = r^-3 R_i R_j + (-q^-3 - 3 e_i s q^-5) Q_i Q_j + d_ij (1/r - 1/q + e_i s q^-3) - 2 h0 q^-3 e_i (d_io Q_j - d_jo Q_i)
//   S_ij 4 pi / 3 = -(R.n) r^-5 R_i R_j + (Q.n) q^-5 (1 + 5 e_i s q^-2) Q_i Q_j
//                   + e_i q^-5 [ -2 h0^2 n_i Q_j + 2 h0 Q_o (n_i Q_j - n_j Q_i) - s d_ij Q_i^2 n_i + 2 h0 (Q.n) d_io Q_j ]
// i.e. a handful of scalar coefficients per point times the tensors R(x)R, Q(x)Q, n(x)Q and the vector Q.  The thread sums
// coefficient x tensor per rule row (once unweighted, once weighted with xi: l_1 = xi, l_0 = 1 - xi) and forms the nine
// entries once per row: 100 (single layer) / 150 (double layer) FP64 instructions per point instead of 184 / 245 for the
// entry-by-entry evaluation.  Checked against the literal form to 5e-16 (tests/test_oracle_extras_cpu.py restates it).
// MODE 0: single layer (sums over the rows in registers), MODE 1: double layer (added to the tile row by row).
template <int MODE>
__device__ __forceinline__ void integrate_no_slip(const double *__restrict__ cq, int nqp, const double *__restrict__ l1d_s,
                                                  const double (&x)[3], const double (&xim)[3], int o,
                                                  double *const (&dst)[4], int vs) {
  static_assert(MODE == 0 || MODE == 1, "one layer per launch");
  constexpr int N1 = 8;
  constexpr int NZ = (MODE == 0) ? 20 : 33;
  const double h0 = 0.5 * (x[o == 0 ? 0 : (o == 1 ? 1 : 2)] - xim[o == 0 ? 0 : (o == 1 ? 1 : 2)]);  // Q_o - R_o = x_o - x_im,o
  const double h2 = 2.0 * h0;
  double acc[4][9];
  if (MODE == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int v = 0; v < 9; ++v) acc[a][v] = 0.0;
  }
  // two-stage software pipeline over the points: loads, R, Q and the two 1/r of point q+1 are in flight while the
  // sums of point q are formed (one warp per scheduler cannot rely on its neighbour to hide the rsqrt chains)
  struct Pt {
    double R[3], Q[3], ri, qi, e[3];  // e: JxW (single layer) or n JxW (double layer)
  };
  auto stage_a = [&](int q, Pt &p) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double yq = cq[d * nqp + q];
      p.R[d] = yq - x[d];
      p.Q[d] = yq - xim[d];
    }
    if (MODE == 0) {
      p.e[0] = cq[6 * nqp + q];
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d) p.e[d] = cq[(3 + d) * nqp + q];
    }
    p.ri = rsqrt_normal(fma(p.R[0], p.R[0], fma(p.R[1], p.R[1], p.R[2] * p.R[2])));
    p.qi = rsqrt_normal(fma(p.Q[0], p.Q[0], fma(p.Q[1], p.Q[1], p.Q[2] * p.Q[2])));
  };
  Pt cur;
  stage_a(0, cur);
  for (int qy = 0; qy < N1; ++qy) {
    double Z0[NZ], Z1[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) Z0[k] = Z1[k] = 0.0;
#pragma unroll 2
    for (int qx = 0; qx < N1; ++qx) {
      const int q = qy * N1 + qx;
      const double xi = l1d_s[qx * 2 + 1];
      Pt nxt;
      stage_a(q + 1 < N1 * N1 ? q + 1 : q, nxt);  // harmless re-read at the end
      const double (&R)[3] = cur.R;
      const double (&Q)[3] = cur.Q;
      const double ri = cur.ri, qi = cur.qi;
      const double ri2 = ri * ri, qi2 = qi * qi, ri3 = ri2 * ri, qi3 = qi2 * qi;
      const double Qo = (o == 0) ? Q[0] : (o == 1 ? Q[1] : Q[2]);
      const double s = h2 * (h0 - Qo);
      const double PR[6] = {R[0] * R[0], R[0] * R[1], R[0] * R[2], R[1] * R[1], R[1] * R[2], R[2] * R[2]};
      const double PQ[6] = {Q[0] * Q[0], Q[0] * Q[1], Q[0] * Q[2], Q[1] * Q[1], Q[1] * Q[2], Q[2] * Q[2]};
      const double PQo[3] = {Qo * Q[0], Qo * Q[1], Qo * Q[2]};
      if (MODE == 0) {
        const double cj = cur.e[0] * BS_INV_8PI;
        const double a3 = -(cj * qi3), a1 = cj * ri3, sa3 = s * a3, u4 = 3.0 * (sa3 * qi2), d1 = cj * (ri - qi);
        // weights: R(x)R, Q(x)Q rows != o, Q(x)Q row o, diagonal rows != o, diagonal row o, Q (antisymmetric part)
        const double w[6] = {a1, a3 + u4, a3 - u4, d1 - sa3, d1 + sa3, -(h2 * a3)};
        double wx[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) wx[k] = w[k] * xi;
#pragma unroll
        for (int v = 0; v < 6; ++v) {
          Z0[v] = fma(w[0], PR[v], Z0[v]);
          Z1[v] = fma(wx[0], PR[v], Z1[v]);
          Z0[6 + v] = fma(w[1], PQ[v], Z0[6 + v]);
          Z1[6 + v] = fma(wx[1], PQ[v], Z1[6 + v]);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Z0[12 + j] = fma(w[2], PQo[j], Z0[12 + j]);
          Z1[12 + j] = fma(wx[2], PQo[j], Z1[12 + j]);
          Z0[17 + j] = fma(w[5], Q[j], Z0[17 + j]);
          Z1[17 + j] = fma(wx[5], Q[j], Z1[17 + j]);
        }
        Z0[15] += w[3];
        Z1[15] += wx[3];
        Z0[16] += w[4];
        Z1[16] += wx[4];
      } else {
        const double nJ[3] = {cur.e[0], cur.e[1], cur.e[2]};
        const double Rn = fma(R[0], nJ[0], fma(R[1], nJ[1], R[2] * nJ[2]));
        const double Qn = fma(Q[0], nJ[0], fma(Q[1], nJ[1], Q[2] * nJ[2]));
        const double qi5 = qi3 * qi2;
        const double b1 = Rn * (ri3 * ri2), b2 = Qn * qi5, c2 = 5.0 * (s * (b2 * qi2));
        const double hq = h2 * qi5;
        // weights: R(x)R, Q(x)Q rows != o, Q(x)Q row o, n(x)Q, its antisymmetric part, diagonal Q_i^2 n_i, Q (row o)
        const double w[7] = {-b1, b2 + c2, b2 - c2, -(h0 * hq), -(hq * Qo), -(s * qi5), h2 * b2};
        double wx[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) wx[k] = w[k] * xi;
#pragma unroll
        for (int v = 0; v < 6; ++v) {
          Z0[v] = fma(w[0], PR[v], Z0[v]);
          Z1[v] = fma(wx[0], PR[v], Z1[v]);
          Z0[6 + v] = fma(w[1], PQ[v], Z0[6 + v]);
          Z1[6 + v] = fma(wx[1], PQ[v], Z1[6 + v]);
        }
        double NQ[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) NQ[3 * i + j] = nJ[i] * Q[j];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          Z0[15 + k] = fma(w[3], NQ[k], Z0[15 + k]);
          Z1[15 + k] = fma(wx[3], NQ[k], Z1[15 + k]);
        }
        const double AS[3] = {NQ[1] - NQ[3], NQ[2] - NQ[6], NQ[5] - NQ[7]};  // (0,1), (0,2), (1,2)
        const double DG[3] = {NQ[0] * Q[0], NQ[4] * Q[1], NQ[8] * Q[2]};     // Q_i^2 n_i
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Z0[12 + j] = fma(w[2], PQo[j], Z0[12 + j]);
          Z1[12 + j] = fma(wx[2], PQo[j], Z1[12 + j]);
          Z0[24 + j] = fma(w[4], AS[j], Z0[24 + j]);
          Z1[24 + j] = fma(wx[4], AS[j], Z1[24 + j]);
          Z0[27 + j] = fma(w[5], DG[j], Z0[27 + j]);
          Z1[27 + j] = fma(wx[5], DG[j], Z1[27 + j]);
          Z0[30 + j] = fma(w[6], Q[j], Z0[30 + j]);
          Z1[30 + j] = fma(wx[6], Q[j], Z1[30 + j]);
        }
      }
      cur = nxt;
    }
    // ---- the nine entries of the row for the two x-direction shape functions, then the y direction
    const double ly0 = l1d_s[qy * 2], ly1 = l1d_s[qy * 2 + 1];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      double T[NZ];
#pragma unroll
      for (int k = 0; k < NZ; ++k) T[k] = (b == 1) ? Z1[k] : Z0[k] - Z1[k];
      double E[9];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const bool io = (i == o);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int v = vidx<6>(i, j);
          double val = T[v] + (io ? T[12 + j] : T[6 + v]);
          if (MODE == 0) {
            if (i == j) val += io ? T[16] : T[15];
            // -2 h0 q^-3 e_i (d_io Q_j - d_jo Q_i): row o (j != o) and column o (i != o) both get + 2 h0 q^-3 Q (T[17..19])
            if (i != j) val += io ? T[17 + j] : ((j == o) ? T[17 + i] : 0.0);
          } else {
            // antisymmetric part A_ij = c3 (n_i Q_j - n_j Q_i): stored (0,1), (0,2), (1,2)
            double e = T[15 + 3 * i + j];
            if (i < j) e -= T[24 + (i == 0 ? j - 1 : 2)];
            if (i > j) e += T[24 + (j == 0 ? i - 1 : 2)];
            if (i == j) e += T[27 + i];
            if (io) e += T[30 + j];
            val += io ? -e : e;
          }
          E[3 * i + j] = val;
        }
      }
      const double f0 = (MODE == 0) ? ly0 : -BS_3_4PI * ly0, f1 = (MODE == 0) ? ly1 : -BS_3_4PI * ly1;
#pragma unroll
      for (int v = 0; v < 9; ++v) {
        if (MODE == 0) {
          acc[b][v] = fma(E[v], f0, acc[b][v]);
          acc[b + 2][v] = fma(E[v], f1, acc[b + 2][v]);
        } else {
          dst[b][(size_t)v * vs] = fma(E[v], f0, dst[b][(size_t)v * vs]);
          dst[b + 2][(size_t)v * vs] = fma(E[v], f1, dst[b + 2][(size_t)v * vs]);
        }
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int v = 0; v < 9; ++v) dst[a][(size_t)v * vs] += acc[a][v];
  }
}

// One (row, cell) integration over this thread's share of the tensor rule.  MODE 0: single layer only, 1: double
// layer only, 2: both.  Sum-factorised: x-direction into NB1 temporaries per value, y-direction once per row of
// the rule; then the QS partial sums of a row (adjacent lanes) are combined by shuffles and lane `part` adds its
// share of the shape functions into the shared tile [value][slot][row].
// FAST: free-space kernel without regularisation, `cq` is the point-major prescaled record.
// TILE_ACC (cell-split mode, QS == 1): the thread owns its tile entries for the duration of the cell, so the accumulators
// start from the tile values and are stored back - no zeroing, no add pass.
template <int NA, int KT, int MODE, int QS, bool HAS_EPS, bool FAST, int N1C, bool KLOW = false, bool TILE_ACC = false>
__device__ __forceinline__ void cell_pass(const double *__restrict__ cq, const double *__restrict__ l1d_s, int n1, int nqp,
                                          const double (&x)[3], const double (&xim)[3], double eps, int o, bool ok,
                                          int part, const int (&slot)[NA], double *__restrict__ acc_s, int tj, int rl) {
  constexpr int NV = GreenTraits<KT>::NV;
  constexpr int NB1 = (NA == 4) ? 2 : 3;
  constexpr int NACC = (MODE == 2) ? 2 * NV : NV;
  constexpr int VOFF = (MODE == 1 && !KLOW) ? NV : 0;  // KLOW: double-layer-only launch, K lives in planes 0..NV-1
  double acc[NA][NACC];
  double *dst[NA];
  // the free-surface integrators keep their own partial sums and combine them into acc at the end: preloading acc would
  // only lengthen its live range there, so that kernel adds to the tile after the cell
  // 2-D moment formulation (free space, cell-split): the expansion at the end of the cell adds to the tile itself
  constexpr bool MOM2D = TILE_ACC && KT != BS_KERNEL_NO_SLIP && N1C == -8 && BS_MOM2D;
  if constexpr (MOM2D) {
    if (ok) {
      const int vs = acc_vstride(tj);
      double *d4[NA];
#pragma unroll
      for (int a = 0; a < NA; ++a) d4[a] = acc_s + (size_t)VOFF * vs + (size_t)slot[a] * ACC_LD + rl;
      if constexpr (KT == BS_KERNEL_FREE) integrate_free_lin2d<MODE>(cq, l1d_s + 32, x, d4, vs);
      else if constexpr (MODE != 2) integrate_free_surface_lin2d<MODE>(cq, l1d_s + 32, x, xim, o, d4, vs);  // layer-split launches
    }
    return;
  }
  if constexpr (TILE_ACC && KT == BS_KERNEL_NO_SLIP && !HAS_EPS && MODE != 2 && BS_NOSLIP_FAST) {
    if (ok) {
      const int vs = acc_vstride(tj);
      double *d4[NA];
#pragma unroll
      for (int a = 0; a < NA; ++a) d4[a] = acc_s + (size_t)VOFF * vs + (size_t)slot[a] * ACC_LD + rl;
      integrate_no_slip<MODE>(cq, nqp, l1d_s, x, xim, o, d4, vs);
    }
    return;
  }
  constexpr bool PRELOAD = TILE_ACC && KT != BS_KERNEL_FREE_SURFACE;
  if constexpr (PRELOAD) {
    static_assert(QS == 1, "tile accumulators: one thread per node and cell");
    if (ok) {
      const int vs = acc_vstride(tj);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        dst[a] = acc_s + (size_t)VOFF * vs + (size_t)slot[a] * ACC_LD + rl;
#pragma unroll
        for (int v = 0; v < NACC; ++v) acc[a][v] = dst[a][(size_t)v * vs];
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int v = 0; v < NACC; ++v) acc[a][v] = 0.0;
  }
  constexpr bool FLIP = FAST && QS == 2 && NA == 4;
  if (FAST) {
    const double *lx_s = FLIP ? l1d_s + part * (n1 * NB1) : l1d_s;  // odd partner: x-flipped copy of the table
    if constexpr (KT == BS_KERNEL_FREE) {
      if constexpr (N1C < 0) {
        static_assert(NA == 4 && N1C == -8, "linear-row fast path: Q1, Gauss 8");
        if (ok) integrate_free_lin<MODE, QS, NACC>(cq, l1d_s + 32, l1d_s, x, part, acc);
      } else {
        if (ok) integrate_free<NA, MODE, QS, N1C, NACC>(cq, lx_s, l1d_s, n1, x, part, acc);
      }
    } else {
      if constexpr (KT == BS_KERNEL_FREE_SURFACE && MODE != 2) {
        if constexpr (N1C < 0) {
          static_assert(NA == 4 && N1C == -8, "linear-row fast path: Q1, Gauss 8");
          if (ok) integrate_free_surface_lin<MODE, QS>(cq, l1d_s + 32, l1d_s, x, xim, o, part, acc);
        } else {
          if (ok) integrate_free_surface<NA, MODE, QS>(cq, lx_s, l1d_s, n1, x, xim, o, part, acc);
        }
      }
    }
  } else {
  for (int qy = part; ok && qy < n1; qy += QS) {
    double tmp[NB1][NACC];
#pragma unroll
    for (int b = 0; b < NB1; ++b)
#pragma unroll
      for (int v = 0; v < NACC; ++v) tmp[b][v] = 0.0;
    const int q0 = qy * n1;
#pragma unroll QX_UNROLL
    for (int qx = 0; qx < n1; ++qx) {
      const int q = q0 + qx;
      double R[3], Rim[3], nJ[3], g[NV], k[NV];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double yq = cq[d * nqp + q];
        R[d] = yq - x[d];
        Rim[d] = yq - xim[d];
        nJ[d] = cq[(3 + d) * nqp + q];
      }
      green_eval<KT>(R, Rim, nJ, cq[6 * nqp + q], HAS_EPS ? eps : 0.0, o, g, k);
#pragma unroll
      for (int b = 0; b < NB1; ++b) {
        const double l = l1d_s[qx * NB1 + b];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (MODE == 2) {
            tmp[b][v] = fma(g[v], l, tmp[b][v]);
            tmp[b][NV + v] = fma(k[v], l, tmp[b][NV + v]);
          } else {
            tmp[b][v] = fma(MODE == 0 ? g[v] : k[v], l, tmp[b][v]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const double l = l1d_s[qy * NB1 + shape_iy<NA>(a)];
#pragma unroll
      for (int v = 0; v < NACC; ++v) acc[a][v] = fma(tmp[shape_ix<NA>(a)][v], l, acc[a][v]);
    }
  }
  }
  if constexpr (TILE_ACC) {
    if (ok) {
      const int vs = acc_vstride(tj);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        if constexpr (!PRELOAD) dst[a] = acc_s + (size_t)VOFF * vs + (size_t)slot[a] * ACC_LD + rl;
#pragma unroll
        for (int v = 0; v < NACC; ++v) {
          if constexpr (PRELOAD) dst[a][(size_t)v * vs] = acc[a][v];
          else dst[a][(size_t)v * vs] += acc[a][v];
        }
      }
    }
    return;
  }
  if (FLIP) {
    // slots 1 and 3 hold the partner's shape functions (0^1, 2^1 in its numbering): send them, finalise 0 and 2
    const int vs = acc_vstride(tj);
#pragma unroll
    for (int a = 0; a < NA; a += 2) {
      double *dst = acc_s + (size_t)VOFF * vs + (size_t)slot[a] * ACC_LD + rl;  // slot[] arrives flipped for part 1
#pragma unroll
      for (int v = 0; v < NACC; ++v) {
        acc[a][v] += __shfl_xor_sync(0xffffffffu, acc[a + 1][v], 1);
        dst[(size_t)v * vs] += acc[a][v];
      }
    }
    return;
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
#pragma unroll
    for (int v = 0; v < NACC; ++v) {
#pragma unroll
      for (int m = 1; m < QS; m <<= 1) acc[a][v] += __shfl_xor_sync(0xffffffffu, acc[a][v], m);
    }
    if ((a % QS) == part) {
      const int vs = acc_vstride(tj);
      double *dst = acc_s + (size_t)VOFF * vs + (size_t)slot[a] * ACC_LD + rl;
#pragma unroll
      for (int v = 0; v < NACC; ++v) dst[(size_t)v * vs] += acc[a][v];
    }
  }
}

// Regular pass.  TI collocation nodes per CTA; QS threads per node share the rows of the tensor rule, and with
// VS == 2 a second set of warps integrates the double layer while the first integrates the single layer
// (half the accumulator registers per thread -> twice the resident warps).  A CTA has TI*QS*VS threads.
constexpr int MAXC = 32;  // cells per block (a block touches at most tj <= 32 nodes)

// LAYER 0: both layers in one tile; 1: single layer only; 2: double layer only (two launches, half the tile per node).
// CS == 2: cell-split mode (see cell_sets): QS == VS == 1, thread set t / TI integrates cell 2*step + set of the block.
template <int NA, int KT, int LAYER, int QS, int VS, bool HAS_EPS, bool FUSED, int N1C, int CS = 1>
__global__ void __launch_bounds__(TI *QS *VS *CS, CTAS_PER_SM) k_assemble_regular(const RegParams P) {
  // point-major prescaled cell records, pipelined points: free space, and the Q1 free-surface image system (its two
  // layers are integrated by separate launches; with Q2 the 9 x 12 partial sums would not fit the register file)
  constexpr bool FAST = !HAS_EPS && (KT == BS_KERNEL_FREE || (KT == BS_KERNEL_FREE_SURFACE && NA == 4 && LAYER != 0));
  constexpr int CQ = FAST ? 8 : 7;                           // doubles per quadrature point in a cell record
  constexpr int NV = GreenTraits<KT>::NV;
  constexpr int NV2 = (LAYER == 0) ? 2 * NV : NV;  // value planes of the tile
  constexpr int KPL = (LAYER == 0) ? NV : 0;       // first plane of the double layer
  static_assert(LAYER == 0 || VS == 1, "layer-split launches use one thread set");
  static_assert(CS == 1 || (CS == 2 && QS == 1 && VS == 1), "cell-split mode: one thread per node and cell");
  constexpr int NB1 = (NA == 4) ? 2 : 3;
  constexpr int NT = TI * QS * VS * CS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tj = (CS == 2) ? tj_cell_split(KT) : P.tj, nqp = (CS == 2) ? 64 : P.nq_pad, n1 = (CS == 2) ? 8 : P.n1d;
  const int ns = ring_stages(nqp, CS);
  double *cellbuf = reinterpret_cast<double *>(smem_raw);                   // [ns][CS][7][nqp] or [ns][CS][nqp][8]
  const size_t ring_doubles = max((size_t)ns * CS * 8 * nqp, (size_t)3 * tj * MAX_PANEL);
  double *l1d_s = cellbuf + ring_doubles;                                   // [n1][NB1] 1-D shape values (+ x-flipped copy)
  double *acc_s = l1d_s + l1d_doubles(nqp);                                 // [NV2][tj][ACC_LD]
  uint64_t *bars = reinterpret_cast<uint64_t *>(acc_s + (size_t)NV2 * acc_vstride(tj));  // full[3], empty[3]
  uint64_t *full = bars, *empty = bars + 3;
  __shared__ int s_cells[MAXC];          // block metadata staged once: no dependent global loads per cell
  __shared__ int s_conn[MAXC * NA];
  __shared__ signed char s_slots[MAXC * NA];
  constexpr bool FLIP = FAST && QS == 2 && NA == 4;  // see cell_pass

  const int t = threadIdx.x;
  const int vpart = (VS == 2) ? t / (TI * QS) : 0;   // warp-uniform role: 0 single layer, 1 double layer
  const int cset = (CS == 2) ? t / TI : 0;           // warp-uniform: which cell of a pair this thread integrates
  const int tt = t - (vpart + cset) * (TI * QS);
  const int rl = tt / QS, part = tt - rl * QS;
  // row tile fastest: co-resident CTAs integrate the same cell block for different rows, so its cell records and
  // metadata stay in L1/L2 instead of being re-streamed from HBM per row tile
  const unsigned bx = blockIdx.y, by = blockIdx.x;
  const int blk = P.blk_begin + bx;
  const int p = P.p0 + by * TI + rl;
  const bool row_ok = p < P.p1;
  const int cs = P.blk_cell_ptr[blk], ce = P.blk_cell_ptr[blk + 1];
  const uint32_t cell_bytes = (uint32_t)(CQ * nqp * sizeof(double));
  const double *cell_src = FAST ? P.cellq8 : P.cellq;

  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NT / 32);  // one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // first records of the ring right away: their HBM latency hides behind the tile clear and the metadata staging
    for (int i = 0; i < ns - 1 && cs + CS * i < ce; ++i) {
      int cid[CS], nvalid = 0;
#pragma unroll
      for (int u = 0; u < CS; ++u) {
        cid[u] = P.blk_cells[cs + CS * i + u];
        nvalid += cid[u] >= 0;
      }
      mbar_expect_tx(&full[i], nvalid * cell_bytes);
#pragma unroll
      for (int u = 0; u < CS; ++u)
        if (cid[u] >= 0)
          bulk_g2s(cellbuf + (size_t)(i * CS + u) * CQ * nqp, cell_src + (size_t)cid[u] * CQ * nqp, cell_bytes, &full[i]);
    }
  }
  for (int i = t; i < n1 * NB1; i += NT) {
    const double l = P.l1d[i];
    l1d_s[i] = l;
    if (FLIP) l1d_s[n1 * NB1 + (i ^ 1)] = l;  // columns swapped (NB1 == 2)
  }
  if (N1C < 0 && t < 8) {  // linear-row fast path: powers of the 1-D rule points xi_q = l_1(x_q), end points of a row
    const double xi = P.l1d[2 * t + 1];
    double *xt = l1d_s + 32 + 4 * t;
    xt[0] = xi;
    xt[1] = xi * xi;
    xt[2] = xi * xi * xi;
    xt[3] = 0.0;
    if (t == 0) {
      const double x0 = P.l1d[1], x7 = P.l1d[2 * 7 + 1];
      l1d_s[64] = x0;
      l1d_s[65] = 1.0 / (x7 - x0);
    }
  }
  if constexpr (NA == 4) {
    // a block has at most MAXC * 4 <= NT metadata entries: one per thread, the two dependent global loads are in flight
    // while the tile is cleared
    static_assert(MAXC * 4 <= NT, "one metadata entry per thread");
    const bool has = t < (ce - cs) * NA;
    const int cell = has ? P.blk_cells[cs + t / NA] : -1;  // -1: no partner cell in this step (cell-split mode)
    const signed char sl = has ? P.blk_slots[(size_t)cs * NA + t] : (signed char)0;
    const int ntile = NV2 * acc_vstride(tj), half = (ntile / 2) & ~1;
    for (int i = 2 * t; i < half; i += 2 * NT) *reinterpret_cast<double2 *>(acc_s + i) = make_double2(0.0, 0.0);
    const int cn = (has && cell >= 0) ? P.conn_pos[(size_t)cell * NA + t % NA] : -1;
    for (int i = half + 2 * t; i + 1 < ntile; i += 2 * NT) *reinterpret_cast<double2 *>(acc_s + i) = make_double2(0.0, 0.0);
    if ((ntile & 1) && t == 0) acc_s[ntile - 1] = 0.0;
    if (has) {
      if (t % NA == 0) s_cells[t / NA] = cell;
      s_conn[t] = cn;
      s_slots[t] = sl;
    }
  } else {
    for (int i = t; i < (ce - cs) * NA; i += NT) {
      const int cell = P.blk_cells[cs + i / NA];
      if (i % NA == 0) s_cells[i / NA] = cell;
      s_conn[i] = cell >= 0 ? P.conn_pos[(size_t)cell * NA + i % NA] : -1;
      s_slots[i] = P.blk_slots[(size_t)cs * NA + i];
    }
    for (int i = t; i < NV2 * acc_vstride(tj); i += NT) acc_s[i] = 0.0;
  }
  __syncthreads();
  double x[3] = {0, 0, 0};
  if (row_ok) {
    x[0] = P.support[(size_t)3 * p];
    x[1] = P.support[(size_t)3 * p + 1];
    x[2] = P.support[(size_t)3 * p + 2];
  }
  double xim[3] = {x[0], x[1], x[2]};
  if (KT != BS_KERNEL_FREE) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d == P.kp.o) xim[d] = x[d] - 2.0 * (x[d] - P.kp.wall_pos);  // ref: bem_stokes.cc:2918-2919
  }
  const double eps = HAS_EPS ? P.kp.eps : 0.0;
  const int o = P.kp.o;

  // Cell ring: thread 0 keeps ns-1 records in flight ahead of the one being integrated.  A stage is refilled once
  // every warp has released it (empty barrier) - warps are not held in lock step by a CTA barrier per cell.
  int st = 0, st_fill = ns - 1;             // stage of cell `it`, stage of cell `it + ns - 1`
  uint32_t ph_full = 0, ph_empty = 0;       // parity bits, one per stage
  const int nsteps = (ce - cs) / CS;  // cell-split mode: the host pads the cell list of a block to whole pairs
  // Cell-split mode: the two cells of a step share no node, but the thread sets may be one step apart (two ring stages);
  // where a cell shares a node with the other set's cell of the previous step (flagged by the host, about one step per
  // block), all warps finish that step first.
  const unsigned sync_mask = (CS == 2) ? P.blk_sync[blk] : 0u;
  for (int step = 0; step < nsteps; ++step) {
    // (a named barrier per pair of warps owning the same tile rows measured 8 % slower than the CTA barrier: the warps
    // of a CTA drift apart; a CTA barrier at every step measured the same as this)
    if (CS == 2 && ((sync_mask >> step) & 1u)) __syncthreads();
    if (t == 0 && step + ns - 1 < nsteps) {
      if (step >= 1) {  // the stage was last used by step-1
        mbar_wait(&empty[st_fill], (ph_empty >> st_fill) & 1u);
        ph_empty ^= 1u << st_fill;
      }
      int nvalid = 0;
#pragma unroll
      for (int u = 0; u < CS; ++u) nvalid += s_cells[(step + ns - 1) * CS + u] >= 0;
      mbar_expect_tx(&full[st_fill], nvalid * cell_bytes);
#pragma unroll
      for (int u = 0; u < CS; ++u) {
        const int cid = s_cells[(step + ns - 1) * CS + u];
        if (cid >= 0)
          bulk_g2s(cellbuf + (size_t)(st_fill * CS + u) * CQ * nqp, cell_src + (size_t)cid * CQ * nqp, cell_bytes,
                   &full[st_fill]);
      }
    }
    const int it = step * CS + cset;  // this thread's cell of the step
    int slot[NA];
    bool sing = (CS == 2) && s_cells[it] < 0;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      slot[a] = s_slots[it * NA + (FLIP ? (a ^ part) : a)];
      sing |= (s_conn[it * NA + a] == p);
    }
    mbar_wait(&full[st], (ph_full >> st) & 1u);
    ph_full ^= 1u << st;
    const double *cq = cellbuf + (size_t)(st * CS + cset) * CQ * nqp;
    const bool ok = row_ok && !sing;  // singular (node in cell) pairs are integrated by K2 (ref: 2885-2908)
    if (VS == 2) {
      if (vpart == 0) cell_pass<NA, KT, 0, QS, HAS_EPS, FAST, N1C>(cq, l1d_s, n1, nqp, x, xim, eps, o, ok, part, slot, acc_s, tj, rl);
      else cell_pass<NA, KT, 1, QS, HAS_EPS, FAST, N1C>(cq, l1d_s, n1, nqp, x, xim, eps, o, ok, part, slot, acc_s, tj, rl);
    } else if (LAYER == 1) {
      cell_pass<NA, KT, 0, QS, HAS_EPS, FAST, N1C, false, CS == 2>(cq, l1d_s, n1, nqp, x, xim, eps, o, ok, part, slot, acc_s, tj, rl);
    } else if (LAYER == 2) {
      cell_pass<NA, KT, 1, QS, HAS_EPS, FAST, N1C, true, CS == 2>(cq, l1d_s, n1, nqp, x, xim, eps, o, ok, part, slot, acc_s, tj, rl);
    } else {
      cell_pass<NA, KT, 2, QS, HAS_EPS, FAST, N1C, false, CS == 2>(cq, l1d_s, n1, nqp, x, xim, eps, o, ok, part, slot, acc_s, tj, rl);
    }
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(&empty[st]);  // this warp is done with the record
    st = (st + 1 == ns) ? 0 : st + 1;
    st_fill = (st_fill + 1 == ns) ? 0 : st_fill + 1;
  }
  __syncthreads();  // the tile is complete.  (Tried: every tile row belongs to one warp when VS == 1, so each warp could write out its
                    // own rows after a __syncwarp and skip this barrier, whose skew is 8 % of the samples - measured 782 ms
                    // instead of 470 ms per assembly: the warps of a CTA drift into different code regions of a 68 KB kernel.)

  // ---- combine the tile with global memory: rows 3*(p-p0)+i, columns 3*node(slot)+j.  The first colour that
  // touches a node column stores, later colours (launched after this one) add: fixed summation order.
  // Each warp takes whole matrix rows; its lanes walk the 3*tj tile columns of the row in (slot, component) order,
  // so runs of consecutive node positions become contiguous 8-byte stores / reductions (node positions ascend
  // within a block).  Per-lane column metadata lives in registers: no division or metadata load in the loop.
  const int rows_tile = min(TI, P.p1 - (P.p0 + (int)by * TI));
  const int *nodes = P.blk_nodes + (size_t)blk * tj;
  const unsigned char *first = P.blk_first + (size_t)blk * tj;
  const int lane = t & 31, warp = t >> 5;
  constexpr int NWARP = NT / 32;
  // A warp takes two consecutive row nodes at a time and its lanes walk the 2 * 3*tj tile columns of the pair, so
  // that the 48 columns of a 16-node block fill three warp steps completely instead of 2 x (32 + 16) lanes.
  constexpr int MAXSTEP = 6;  // 2 * 3*tj <= 192 columns
  const int vs = acc_vstride(tj);
  const int E = 3 * tj;
  // per-lane column metadata, loop invariant: shared-tile offset of the value for each matrix-row component i (+ row
  // of the pair), global column, and what to do with it (0 nothing, 1 store, 2 reduce) -- the row loop is branch free
  int soff[MAXSTEP][3], gcol[MAXSTEP], todo[MAXSTEP], kc[MAXSTEP];
  bool second[MAXSTEP];
  const bool mixed = FUSED && LAYER != 1 && P.kcol != nullptr;  // uniform: -K of the flagged columns is kept (mixed BC)
#pragma unroll
  for (int sidx = 0; sidx < MAXSTEP; ++sidx) {
    const int e2 = lane + 32 * sidx;
    second[sidx] = e2 >= E;
    const int e = e2 - (second[sidx] ? E : 0);
    const int sl = e / 3, j = e - 3 * sl;
    const int node = (e2 < 2 * E) ? nodes[sl] : -1;
    const bool valid = node >= 0;
    todo[sidx] = valid ? (first[sl] != 0 ? 1 : 2) : 0;
    gcol[sidx] = valid ? 3 * node + j : 0;
    kc[sidx] = (mixed && valid) ? P.kcol[3 * node + j] : -1;
#pragma unroll
    for (int i = 0; i < 3; ++i) soff[sidx][i] = valid ? vidx<NV>(i, j) * vs + sl * ACC_LD + (second[sidx] ? 1 : 0) : 0;
  }
  // cell-split mode, fused: this thread's share of the panel rows is fetched now and staged after the write-out
  constexpr int NPF = (CS == 2) ? (3 * tj_cell_split(KT) * MAX_PANEL + NT - 1) / NT : 1;
  double pf[NPF];
  constexpr bool prefetch_panel = CS == 2;  // measured + 2 % on the default workload
  if (prefetch_panel && FUSED && LAYER != 1) {
#pragma unroll
    for (int u = 0; u < NPF; ++u) {
      const int idx = t + u * NT;
      const int col = idx / MAX_PANEL, q = idx - col * MAX_PANEL;
      const int sl = col / 3, j = col - 3 * sl;
      const int node = (sl < tj) ? nodes[sl] : -1;
      pf[u] = (node >= 0 && q < P.pp) ? P.panel[((size_t)3 * node + j) * P.pp + q] : 0.0;
    }
  }
  for (int r_ = 2 * warp; r_ < rows_tile; r_ += 2 * NWARP) {
    const bool has2 = r_ + 1 < rows_tile;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const size_t rowoff = ((size_t)3 * (by * TI + r_) + i) * P.ld;
#pragma unroll
      for (int sidx = 0; sidx < MAXSTEP; ++sidx) {
        if (32 * sidx >= 2 * E) break;  // warp-uniform
        const size_t off = rowoff + (second[sidx] ? (size_t)3 * P.ld : 0) + gcol[sidx];
        const int what = (second[sidx] && !has2) ? 0 : todo[sidx];
        const double *as = acc_s + soff[sidx][i] + r_;
        if (LAYER != 2) store_or_reduce(P.V + off, as[0], what);
        if (!FUSED && LAYER != 1) store_or_reduce(P.K + off, as[(size_t)KPL * vs], what);
        if (mixed) {
          const size_t offk = ((size_t)3 * (by * TI + r_) + i + (second[sidx] ? 3 : 0)) * P.ldk + (kc[sidx] >= 0 ? kc[sidx] : 0);
          store_or_reduce(P.Kflag + offk, -as[(size_t)KPL * vs], kc[sidx] >= 0 ? what : 0);
        }
      }
    }
  }
  if (FUSED && LAYER != 1) {
    // K tile (rows x 3*tj) times the panel rows of this block's nodes; the single-layer part of the shared tile is
    // free after its write-out and stages the panel rows (double-layer-only launches use the idle cell ring).  Products go to KX with L2 reductions (every block adds to
    // the same rows: summation order of these few panel columns is not fixed; the stored matrix stays deterministic).
    __syncthreads();
    const int pp = P.pp;
    constexpr int PS = MAX_PANEL;  // padded panel stride in shared memory (zero filled): fixed-trip inner loops
    double *xs = (LAYER == 0) ? acc_s : cellbuf;  // [3*tj][PS]
    if (prefetch_panel) {
#pragma unroll
      for (int u = 0; u < NPF; ++u)
        if (t + u * NT < 3 * tj * PS) xs[t + u * NT] = pf[u];
    } else {
      for (int idx = t; idx < 3 * tj * PS; idx += NT) {
        const int col = idx / PS, q = idx - col * PS;
        const int sl = col / 3, j = col - 3 * sl;
        const int node = nodes[sl];
        xs[idx] = (node >= 0 && q < pp) ? P.panel[((size_t)3 * node + j) * pp + q] : 0.0;
      }
    }
    __syncthreads();
    // TPR threads per row node split the panel columns; each keeps the three matrix rows of its node in registers,
    // so a K value is read from the tile once and every thread is busy (TI*TPR == NT).
    constexpr int TPR = NT / TI, PSH = PS / TPR;
    static_assert(TPR * TI == NT && PSH * TPR == PS && PSH % 2 == 0, "fused epilogue thread layout");
    const double *kacc = acc_s + (size_t)KPL * vs;
    const int r_ = t / TPR, h_ = t - r_ * TPR;
    if (r_ < rows_tile) {
      double y[3][PSH];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < PSH; ++q) y[i][q] = 0.0;
      for (int sl = 0; sl < tj; ++sl) {
        double kv[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) kv[v] = kacc[(size_t)v * vs + (size_t)sl * ACC_LD + r_];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double2 *xr = reinterpret_cast<const double2 *>(xs + (3 * sl + j) * PS + h_ * PSH);  // broadcast reads
#pragma unroll
          for (int q2 = 0; q2 < PSH / 2; ++q2) {
            const double2 xv = xr[q2];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              y[i][2 * q2] = fma(kv[vidx<NV>(i, j)], xv.x, y[i][2 * q2]);
              y[i][2 * q2 + 1] = fma(kv[vidx<NV>(i, j)], xv.y, y[i][2 * q2 + 1]);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double *dst = P.KX + ((size_t)3 * (by * TI + r_) + i) * pp + h_ * PSH;
#pragma unroll
        for (int q = 0; q < PSH; ++q)
          if (h_ * PSH + q < pp) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + q), "d"(y[i][q]) : "memory");
      }
    }
  }
}

template <int NA, int KT, int LAYER, int QS, int VS, int CS = 1>
static void launch_reg_layer(Context &c, RegParams P, int nrow_tiles, size_t smem, int colour) {
  if constexpr (CS == 2) {  // cell-split mode (cell_sets): Gauss 8, no regularisation; moment formulation except for no-slip
    static_assert(NA == 4 && (LAYER == 0) == (KT == BS_KERNEL_FREE), "cell-split mode: Q1 unknowns");
    BS_REQUIRE(cell_sets(c) == 2 && c.blocks.cs == 2, "cell blocks were not built for the cell-split kernel");
    BS_REQUIRE(P.tj == tj_cell_split(KT) && P.nq_pad == 64 && P.n1d == 8, "cell-split kernel: compile-time block and rule sizes");
    constexpr int NC = (KT == BS_KERNEL_NO_SLIP) ? 0 : -8;
    auto kern = c.fused ? k_assemble_regular<NA, KT, LAYER, QS, VS, false, true, NC, 2>
                        : k_assemble_regular<NA, KT, LAYER, QS, VS, false, false, NC, 2>;
    BS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const std::vector<int> &cs = c.blocks.colour_start;
    const int nb = cs[colour + 1] - cs[colour];
    for (int b0 = 0; b0 < nb; b0 += 65535) {  // gridDim.y limit
      P.blk_begin = cs[colour] + b0;
      kern<<<dim3(nrow_tiles, std::min(nb - b0, 65535)), TI * QS * VS * CS, smem, c.stream>>>(P);
      BS_CUDA(cudaGetLastError());
      count_launch(c);
    }
    return;
  } else {
  // the free-space kernel has a variant with the 1-D rule size fixed at compile time (Gauss 8, the order of the
  // reference's parameter files): fully unrolled, software-pipelined rows
  const bool n8 = (KT == BS_KERNEL_FREE) && c.kp.eps == 0.0 && P.n1d == 8;
  constexpr int N8 = (KT == BS_KERNEL_FREE) ? 8 : 0;
  // Q1 unknowns on a Q1 (bilinear) mapping: the rows of the rule are straight lines -> moment formulation (N1C = -8)
  constexpr bool LINK = (KT == BS_KERNEL_FREE) || (KT == BS_KERNEL_FREE_SURFACE && LAYER != 0);  // kernels with the moment formulation
  constexpr int L8 = (LINK && NA == 4) ? -8 : N8;
  const bool lin8 = LINK && NA == 4 && c.na_map == 4 && c.kp.eps == 0.0 && P.n1d == 8 && !std::getenv("BS_NO_LINROWS");
  auto kern = c.fused ? ((c.kp.eps == 0.0) ? (lin8 ? k_assemble_regular<NA, KT, LAYER, QS, VS, false, true, L8>
                                                   : (n8 ? k_assemble_regular<NA, KT, LAYER, QS, VS, false, true, N8>
                                                         : k_assemble_regular<NA, KT, LAYER, QS, VS, false, true, 0>))
                                           : k_assemble_regular<NA, KT, LAYER, QS, VS, true, true, 0>)
                      : ((c.kp.eps == 0.0) ? (lin8 ? k_assemble_regular<NA, KT, LAYER, QS, VS, false, false, L8>
                                                   : (n8 ? k_assemble_regular<NA, KT, LAYER, QS, VS, false, false, N8>
                                                         : k_assemble_regular<NA, KT, LAYER, QS, VS, false, false, 0>))
                                           : k_assemble_regular<NA, KT, LAYER, QS, VS, true, false, 0>);
  BS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const std::vector<int> &cs = c.blocks.colour_start;
  const int nb = cs[colour + 1] - cs[colour];
  for (int b0 = 0; b0 < nb; b0 += 65535) {  // gridDim.y limit
    P.blk_begin = cs[colour] + b0;
    kern<<<dim3(nrow_tiles, std::min(nb - b0, 65535)), TI * QS * VS, smem, c.stream>>>(P);
    BS_CUDA(cudaGetLastError());
    count_launch(c);
  }
  }
}

// SPLIT: the two layers in two launches per colour (single layer, then double layer)
template <int NA, int KT, bool SPLIT, int QS, int VS, int CS = 1>
static void launch_reg(Context &c, RegParams P, int nrow_tiles, size_t smem) {
  BS_REQUIRE(c.blocks.max_cells <= MAXC, "cell block larger than MAXC");
  const std::vector<int> &cs = c.blocks.colour_start;
  for (size_t k = 0; k + 1 < cs.size(); ++k) {  // one launch (pair) per colour, stream order = summation order
    if (cs[k + 1] - cs[k] <= 0) continue;
    if constexpr (SPLIT) {
      launch_reg_layer<NA, KT, 1, QS, VS, CS>(c, P, nrow_tiles, smem, (int)k);
      launch_reg_layer<NA, KT, 2, QS, VS, CS>(c, P, nrow_tiles, smem, (int)k);
    } else {
      launch_reg_layer<NA, KT, 0, QS, VS, CS>(c, P, nrow_tiles, smem, (int)k);
    }
  }
}

void launch_assembly_regular(Context &c) {
  RegParams P;
  P.p0 = c.p0;
  P.p1 = c.p1;
  P.N = c.N;
  P.nq = c.nq;
  P.nq_pad = c.nq_pad;
  P.n1d = (int)c.x1d.size();
  P.tj = c.blocks.tj;
  P.support = c.d_support.p;
  P.conn_pos = c.d_conn_pos.p;
  P.cellq = c.d_cellq.p;
  P.cellq8 = c.d_cellq8.p;
  P.l1d = c.d_l1d.p;
  P.blk_cell_ptr = c.d_blk_cell_ptr.p;
  P.blk_cells = c.d_blk_cells.p;
  P.blk_slots = c.d_blk_slots.p;
  P.blk_nodes = c.d_blk_nodes.p;
  P.blk_first = c.d_blk_first.p;
  P.blk_sync = c.d_blk_sync.p;
  P.blk_begin = 0;
  P.V = c.V.p;
  P.K = c.K.p;
  P.ld = c.ld;
  P.kp = c.kp;
  P.panel = c.fused ? c.d_panel.p : nullptr;
  P.KX = c.fused ? c.d_KX.p : nullptr;
  P.pp = c.panel_p;
  P.kcol = (c.fused && c.n_flagged > 0) ? c.d_kcol.p : nullptr;
  P.Kflag = c.d_Kflag.p;
  P.ldk = c.ldk;
  const int nrow_tiles = (c.p1 - c.p0 + TI - 1) / TI;
  if (nrow_tiles == 0) return;
  
  const int grid = nrow_tiles;
  const size_t smem = assembly_smem_bytes(c.na, tile_planes(c.na, c.kp.type), c.blocks.tj, c.nq_pad, c.blocks.cs);
  BS_REQUIRE(c.blocks.cs == cell_sets(c), "cell blocks out of date (kernel or quadrature changed after the tables were built)");
  const bool q2 = (c.na == 9);
  switch (c.kp.type) {
    // <NA, kernel, two sequential passes?, threads per row over q, thread sets over {V,K}>
    case BS_KERNEL_FREE:
      if (q2) launch_reg<9, BS_KERNEL_FREE, false, 1, 2>(c, P, grid, smem);      // 256 threads
      else if (c.blocks.cs == 2) launch_reg<4, BS_KERNEL_FREE, false, 1, 1, 2>(c, P, grid, smem);  // cell-split pairs
      else launch_reg<4, BS_KERNEL_FREE, false, 2, 1>(c, P, grid, smem);         // 128 threads (V/K thread split measured slower)
      break;
    case BS_KERNEL_FREE_SURFACE:
      if (q2) launch_reg<9, BS_KERNEL_FREE_SURFACE, false, 1, 2>(c, P, grid, smem);
      else if (c.blocks.cs == 2) launch_reg<4, BS_KERNEL_FREE_SURFACE, true, 1, 1, 2>(c, P, grid, smem);
      else launch_reg<4, BS_KERNEL_FREE_SURFACE, true, 2, 1>(c, P, grid, smem);
      break;
    case BS_KERNEL_NO_SLIP:
      if (q2) launch_reg<9, BS_KERNEL_NO_SLIP, false, 1, 2>(c, P, grid, smem);
      else if (c.blocks.cs == 2) launch_reg<4, BS_KERNEL_NO_SLIP, true, 1, 1, 2>(c, P, grid, smem);
      else launch_reg<4, BS_KERNEL_NO_SLIP, true, 2, 1>(c, P, grid, smem);
      break;
    default:
      throw Error(BS_ERR_INVALID, "unknown kernel type");
  }
  c.stats.pairs_regular += (long long)(c.p1 - c.p0) * c.ncell * c.nq;
}

// ---------------------------------------------------------------------------------------------------------
// K2: singular pass — one warp per owned node, lanes over the points of the singular rule
// ---------------------------------------------------------------------------------------------------------
struct SingParams {
  int p0, p1, N, na, nam;
  const double *support, *map_nodes;
  const int *conn_pos, *conn_map;
  const int *patch_ptr, *patch_cell, *patch_local;
  const double *tab;          // records of (na + 3*nam + 1) doubles
  const int *sing_off, *sing_nq;
  double *V, *K;
  size_t ld;
  KernelParams kp;
  const double *panel;  // fused mode (K not stored)
  double *KX;
  int pp;
  const int *kcol;      // fused mode, mixed boundary conditions (see RegParams)
  double *Kflag;
  size_t ldk;
};

constexpr int SING_WARPS = 4;

template <int NA, int NAM, int KT, bool FUSED>
__global__ void __launch_bounds__(32 * SING_WARPS) k_assemble_singular(const SingParams P) {
  constexpr int NV = GreenTraits<KT>::NV;
  constexpr int NV2 = 2 * NV;
  constexpr int CH = (NA == 4) ? 4 : 3;          // shape functions per register chunk
  constexpr int REC = NA + 3 * NAM + 1;
  __shared__ double red[SING_WARPS][CH * NV2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = P.p0 + blockIdx.x * SING_WARPS + wid;
  if (p >= P.p1) return;
  const double x[3] = {P.support[(size_t)3 * p], P.support[(size_t)3 * p + 1], P.support[(size_t)3 * p + 2]};
  double xim[3] = {x[0], x[1], x[2]};
  if (KT != BS_KERNEL_FREE) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d == P.kp.o) xim[d] = x[d] - 2.0 * (x[d] - P.kp.wall_pos);
  }
  const size_t row0 = (size_t)3 * (p - P.p0);
  for (int e = P.patch_ptr[p]; e < P.patch_ptr[p + 1]; ++e) {
    const int cell = P.patch_cell[e], al = P.patch_local[e];
    const int off = P.sing_off[al], nqs = P.sing_nq[al];
    double X[NAM][3];
#pragma unroll
    for (int a = 0; a < NAM; ++a) {
      const int m = P.conn_map[(size_t)cell * NAM + a];
#pragma unroll
      for (int d = 0; d < 3; ++d) X[a][d] = P.map_nodes[(size_t)3 * m + d];
    }
    for (int a0 = 0; a0 < NA; a0 += CH) {
      double acc[CH][NV2];
#pragma unroll
      for (int a = 0; a < CH; ++a)
#pragma unroll
        for (int v = 0; v < NV2; ++v) acc[a][v] = 0.0;
      for (int q = lane; q < nqs; q += 32) {
        const double *rec = P.tab + (size_t)(off + q) * REC;
        double y[3] = {0, 0, 0}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < NAM; ++a) {
          const double ph = rec[NA + 3 * a], dx = rec[NA + 3 * a + 1], dy = rec[NA + 3 * a + 2];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            y[d] = fma(ph, X[a][d], y[d]);
            t1[d] = fma(dx, X[a][d], t1[d]);
            t2[d] = fma(dy, X[a][d], t2[d]);
          }
        }
        const double w = rec[NA + 3 * NAM];
        const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2],
                     nz = t1[0] * t2[1] - t1[1] * t2[0];
        const double JxW = w * sqrt(nx * nx + ny * ny + nz * nz);
        const double nJ[3] = {w * nx, w * ny, w * nz};
        double R[3], Rim[3], g[NV], k[NV];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          R[d] = y[d] - x[d];
          Rim[d] = y[d] - xim[d];
        }
        green_eval<KT>(R, Rim, nJ, JxW, P.kp.eps, P.kp.o, g, k);
#pragma unroll
        for (int a = 0; a < CH; ++a) {
          if (a0 + a < NA) {
            const double ph = rec[a0 + a];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              acc[a][v] = fma(g[v], ph, acc[a][v]);
              acc[a][NV + v] = fma(k[v], ph, acc[a][NV + v]);
            }
          }
        }
      }
#pragma unroll
      for (int a = 0; a < CH; ++a)
#pragma unroll
        for (int v = 0; v < NV2; ++v) {
          double s = acc[a][v];
#pragma unroll
          for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
          if (lane == 0) red[wid][a * NV2 + v] = s;
        }
      __syncwarp();
      // scatter-add: CH shape functions x 3 x 3 entries x 2 matrices
      for (int idx = lane; idx < CH * 18; idx += 32) {
        const int a = idx / 18, r = idx - a * 18;
        const int mat = r / 9, ij = r - mat * 9, i = ij / 3, j = ij - 3 * i;
        if (a0 + a < NA && !(FUSED && mat == 1)) {
          const int cpos = P.conn_pos[(size_t)cell * NA + a0 + a];
          const double val = red[wid][a * NV2 + mat * NV + vidx<NV>(i, j)];
          double *M = mat == 0 ? P.V : P.K;
          M[(row0 + i) * P.ld + (size_t)3 * cpos + j] += val;
        }
        if (FUSED && mat == 1 && a0 + a < NA && P.kcol) {  // mixed BC: -K of the flagged columns
          const int cpos = P.conn_pos[(size_t)cell * NA + a0 + a];
          const int kcc = P.kcol[3 * cpos + j];
          if (kcc >= 0) P.Kflag[(row0 + i) * P.ldk + kcc] -= red[wid][a * NV2 + NV + vidx<NV>(i, j)];
        }
      }
      if (FUSED) {
        // K block of this (node, cell) pair times the panel rows of the cell's nodes; this warp owns the rows
        for (int idx = lane; idx < 3 * P.pp; idx += 32) {
          const int i = idx / P.pp, q = idx - i * P.pp;
          double y = 0.0;
          for (int a = 0; a < CH; ++a) {
            if (a0 + a >= NA) break;
            const int cpos = P.conn_pos[(size_t)cell * NA + a0 + a];
            for (int j = 0; j < 3; ++j)
              y = fma(red[wid][a * NV2 + NV + vidx<NV>(i, j)], P.panel[((size_t)3 * cpos + j) * P.pp + q], y);
          }
          P.KX[(row0 + i) * P.pp + q] += y;
        }
      }
      __syncwarp();
    }
  }
}

template <int NA, int NAM, int KT>
static void launch_sing(Context &c, const SingParams &P) {
  const int n = c.p1 - c.p0;
  if (n <= 0) return;
  if (c.fused) k_assemble_singular<NA, NAM, KT, true><<<(n + SING_WARPS - 1) / SING_WARPS, 32 * SING_WARPS, 0, c.stream>>>(P);
  else k_assemble_singular<NA, NAM, KT, false><<<(n + SING_WARPS - 1) / SING_WARPS, 32 * SING_WARPS, 0, c.stream>>>(P);
  BS_CUDA(cudaGetLastError());
}

template <int KT>
static void launch_sing_kt(Context &c, const SingParams &P) {
  if (c.na == 4 && c.na_map == 4) launch_sing<4, 4, KT>(c, P);
  else if (c.na == 4 && c.na_map == 9) launch_sing<4, 9, KT>(c, P);
  else if (c.na == 9 && c.na_map == 4) launch_sing<9, 4, KT>(c, P);
  else launch_sing<9, 9, KT>(c, P);
}

void launch_assembly_singular(Context &c) {
  BS_REQUIRE(c.have_singular, "singular quadrature not set");
  SingParams P;
  P.p0 = c.p0;
  P.p1 = c.p1;
  P.N = c.N;
  P.na = c.na;
  P.nam = c.na_map;
  P.support = c.d_support.p;
  P.map_nodes = c.d_map_nodes.p;
  P.conn_pos = c.d_conn_pos.p;
  P.conn_map = c.d_conn_map.p;
  P.patch_ptr = c.d_patch_ptr.p;
  P.patch_cell = c.d_patch_cell.p;
  P.patch_local = c.d_patch_local.p;
  P.tab = c.d_sing_tab.p;
  P.sing_off = c.d_sing_off.p;
  P.sing_nq = c.d_sing_nq.p;
  P.V = c.V.p;
  P.K = c.K.p;
  P.ld = c.ld;
  P.kp = c.kp;
  P.panel = c.fused ? c.d_panel.p : nullptr;
  P.KX = c.fused ? c.d_KX.p : nullptr;
  P.pp = c.panel_p;
  P.kcol = (c.fused && c.n_flagged > 0) ? c.d_kcol.p : nullptr;
  P.Kflag = c.d_Kflag.p;
  P.ldk = c.ldk;
  switch (c.kp.type) {
    case BS_KERNEL_FREE: launch_sing_kt<BS_KERNEL_FREE>(c, P); break;
    case BS_KERNEL_FREE_SURFACE: launch_sing_kt<BS_KERNEL_FREE_SURFACE>(c, P); break;
    case BS_KERNEL_NO_SLIP: launch_sing_kt<BS_KERNEL_NO_SLIP>(c, P); break;
    default: throw Error(BS_ERR_INVALID, "unknown kernel type");
  }
  count_launch(c);
  long long pairs = 0;
  for (int a = 0; a < c.na; ++a) pairs += c.sing_nq[a];
  c.stats.pairs_singular += pairs * c.ncell;  // every cell has one singular node per local index
}

// ---------------------------------------------------------------------------------------------------------
// bs_kernel_eval: G (9) and W (27) through the same device functions (W_ijk = S_ij with n = e_k, sign restored)
// ---------------------------------------------------------------------------------------------------------
template <int KT>
__device__ void eval_one(const double *p, const double *pim, double eps, int o, double *G, double *W) {
  constexpr int NV = GreenTraits<KT>::NV;
  double R[3] = {p[0], p[1], p[2]}, Q[3] = {pim[0], pim[1], pim[2]};
  for (int kk = 0; kk < 3; ++kk) {
    double nJ[3] = {kk == 0 ? 1.0 : 0.0, kk == 1 ? 1.0 : 0.0, kk == 2 ? 1.0 : 0.0};
    double g[NV], k[NV];
    green_eval<KT>(R, Q, nJ, 1.0, eps, o, g, k);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        if (G && kk == 0) G[3 * i + j] = g[vidx<NV>(i, j)];
        if (W) W[9 * i + 3 * j + kk] = -k[vidx<NV>(i, j)];
      }
  }
}

__global__ void k_kernel_eval(int type, double eps, int o, int npts, const double *p, const double *pim, double *G,
                              double *W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  double *Gi = G ? G + 9 * (size_t)i : nullptr, *Wi = W ? W + 27 * (size_t)i : nullptr;
  const double *pi_ = pim ? pim + 3 * (size_t)i : p + 3 * (size_t)i;
  if (type == BS_KERNEL_FREE) eval_one<BS_KERNEL_FREE>(p + 3 * (size_t)i, pi_, eps, o, Gi, Wi);
  else if (type == BS_KERNEL_FREE_SURFACE) eval_one<BS_KERNEL_FREE_SURFACE>(p + 3 * (size_t)i, pi_, eps, o, Gi, Wi);
  else eval_one<BS_KERNEL_NO_SLIP>(p + 3 * (size_t)i, pi_, eps, o, Gi, Wi);
}

void kernel_eval_device(int type, double eps, int o, int npts, const double *d_p, const double *d_pim, double *d_G,
                        double *d_W, cudaStream_t s) {
  k_kernel_eval<<<(npts + 127) / 128, 128, 0, s>>>(type, eps, o, npts, d_p, d_pim, d_G, d_W);
  BS_CUDA(cudaGetLastError());
}

}  // namespace bs
