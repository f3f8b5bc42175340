// Dense FP64 linear-algebra kernels of the solve path: the HBM-streaming GEMV that replaces
// TrilinosWrappers::SparseMatrix::vmult on the dense-as-CRS matrices (ref call sites: SURVEY §8a), its
// multi-right-hand-side variant, fused multi-dot / multi-axpy for Gram-Schmidt, and the element-wise
// correction / monolithic-build kernels (ref: source/bem_stokes.cc:3004-3098, 3152-3357).
#include "bs_internal.h"

namespace bs {

// ---------------------------------------------------------------------------------------------------------
// GEMV  y[r] = sum_c A[r][c] x[c] ; A row-major with even ld, streamed once with 16-byte loads.
// One warp owns RPW consecutive rows (x is re-used from registers across them), lanes stride the columns.
// ---------------------------------------------------------------------------------------------------------
constexpr int GEMV_WARPS = 8;
constexpr int GEMV_RPW = 2;  // rows per warp: 2 streams 7.2 TB/s, 4 only 6.6 TB/s (measured, 87 GB matrix)

__device__ __forceinline__ double2 ld_stream2(const double *p) {
  double2 v;
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// launch bounds: 64 registers, 4 CTAs = 32 warps per SM (3 CTAs at 76 registers measured 0.4 % slower)
template <int RPW>
__global__ void __launch_bounds__(32 * GEMV_WARPS, 4) k_gemv(const double *__restrict__ A, size_t ld, size_t rows, size_t cols,
                                                          const double *__restrict__ x, double *__restrict__ y,
                                                          const int *skip, const double *__restrict__ r1_u,
                                                          const double *__restrict__ r1_dot) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * GEMV_WARPS + (threadIdx.x >> 5);
  const size_t r0 = warp * RPW;
  if (r0 >= rows) return;
  const double *a[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) a[r] = A + ((r0 + r < rows) ? (r0 + r) : r0) * ld;
  double acc[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) acc[r] = 0.0;
  const size_t cols2 = cols & ~(size_t)1;
  size_t c = (size_t)lane * 2;
  // main loop, 4 column steps in flight
  for (; c + 3 * 64 < cols2; c += 4 * 64) {
    double2 xv[4], av[4][RPW];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[u] = *reinterpret_cast<const double2 *>(x + c + u * 64);
#pragma unroll
      for (int r = 0; r < RPW; ++r) av[u][r] = ld_stream2(a[r] + c + u * 64);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[r] = fma(av[u][r].x, xv[u].x, fma(av[u][r].y, xv[u].y, acc[r]));
  }
  for (; c < cols2; c += 64) {
    const double2 xv = *reinterpret_cast<const double2 *>(x + c);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const double2 av = ld_stream2(a[r] + c);
      acc[r] = fma(av.x, xv.x, fma(av.y, xv.y, acc[r]));
    }
  }
  if ((cols & 1) && lane == 0) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) acc[r] = fma(a[r][cols - 1], x[cols - 1], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    double s = acc[r];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0 && r0 + r < rows) y[r0 + r] = r1_u ? fma(r1_u[r0 + r], *r1_dot, s) : s;  // + u (w . x): implicit rank-1 term
  }
}

// dots[k] = r1_w . X_k for the implicit rank-1 term of M (nullptr when M has none)
static const double *rank1_dots(Context &c, const DMat &M, int nrhs, const double *X, size_t ldx) {
  if (!M.r1_u) return nullptr;
  double *dots = c.wsd("r1.dots", 16);
  multi_dot(c, X, ldx, nrhs, M.r1_w, M.r1_ncols, dots);
  return dots;
}

void gemv(Context &c, const DMat &M, const double *x, double *y, const int *skip) {
  if (M.rows == 0) return;
  BS_REQUIRE((M.ld & 1) == 0, "matrix ld must be even");
  const double *r1_dot = rank1_dots(c, M, 1, x, 0);
  const size_t warps = (M.rows + GEMV_RPW - 1) / GEMV_RPW;
  const unsigned grid = (unsigned)((warps + GEMV_WARPS - 1) / GEMV_WARPS);
  k_gemv<GEMV_RPW><<<grid, 32 * GEMV_WARPS, 0, c.stream>>>(M.p, M.ld, M.rows, M.cols, x, y, skip, M.r1_u, r1_dot);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

// multi-RHS: Y[k][r] = sum_c A[r][c] X[k][c], NR right-hand sides per pass.  Every matrix element is read once
// from HBM and used NR times; RPW rows per warp re-use the (L1/L2 resident) x values from registers, U column
// steps are in flight per lane.
template <int NR, int RPW, int U>
__global__ void __launch_bounds__(32 * GEMV_WARPS) k_gemv_multi(const double *__restrict__ A, size_t ld, size_t rows,
                                                                size_t cols, const double *__restrict__ X, size_t ldx,
                                                                double *__restrict__ Y, size_t ldy, const int *skip,
                                                                const double *__restrict__ r1_u, const double *__restrict__ r1_dots) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * GEMV_WARPS + (threadIdx.x >> 5);
  const size_t r0 = warp * RPW;
  if (r0 >= rows) return;
  const double *a[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) a[r] = A + ((r0 + r < rows) ? (r0 + r) : r0) * ld;
  double acc[RPW][NR];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[r][k] = 0.0;
  const size_t cols2 = cols & ~(size_t)1;
  size_t c = (size_t)lane * 2;
  for (; c + (U - 1) * 64 < cols2; c += U * 64) {
    double2 av[U][RPW];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < RPW; ++r) av[u][r] = ld_stream2(a[r] + c + u * 64);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const double2 xv = *reinterpret_cast<const double2 *>(X + (size_t)k * ldx + c + u * 64);
#pragma unroll
        for (int r = 0; r < RPW; ++r) acc[r][k] = fma(av[u][r].x, xv.x, fma(av[u][r].y, xv.y, acc[r][k]));
      }
  }
  for (; c < cols2; c += 64) {
    double2 av[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) av[r] = ld_stream2(a[r] + c);
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const double2 xv = *reinterpret_cast<const double2 *>(X + (size_t)k * ldx + c);
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[r][k] = fma(av[r].x, xv.x, fma(av[r].y, xv.y, acc[r][k]));
    }
  }
  if ((cols & 1) && lane == 0) {
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int k = 0; k < NR; ++k) acc[r][k] = fma(a[r][cols - 1], X[(size_t)k * ldx + cols - 1], acc[r][k]);
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      double s = acc[r][k];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
      if (lane == 0 && r0 + r < rows) Y[(size_t)k * ldy + r0 + r] = r1_u ? fma(r1_u[r0 + r], r1_dots[k], s) : s;
    }
}

#ifndef BS_GEMVM_RPW
#define BS_GEMVM_RPW 2
#endif
#ifndef BS_GEMVM_U
#define BS_GEMVM_U 4
#endif

template <int NR>
static void launch_gemv_multi(Context &c, const DMat &M, const double *X, size_t ldx, double *Y, size_t ldy, const int *skip,
                              const double *r1_dots) {
  constexpr int RPW = BS_GEMVM_RPW, U = BS_GEMVM_U;
  const size_t warps = (M.rows + RPW - 1) / RPW;
  const unsigned grid = (unsigned)((warps + GEMV_WARPS - 1) / GEMV_WARPS);
  k_gemv_multi<NR, RPW, U><<<grid, 32 * GEMV_WARPS, 0, c.stream>>>(M.p, M.ld, M.rows, M.cols, X, ldx, Y, ldy, skip, M.r1_u, r1_dots);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

// ---------------------------------------------------------------------------------------------------------
// Multi-RHS sweep on the FP64 tensor path: Y[n][rows] = A[rows x cols] * X[n][cols]^T for 4 <= n <= 8 right-hand sides
// (batched rigid-body resistance solves, projected rigid columns).  DMMA m8n8k4: a warp owns MB blocks of 8 matrix
// rows (see k_gemm_dmma); the L1/LSU traffic for the vectors is 1/MB of the matrix stream instead of n times it
// (k_gemv_multi).  The columns are split over gridDim.y for enough warps in flight; k_sum_ksplit adds the partial
// results in fixed order.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int DMMA_WARPS = 4;

// Lane (g = lane/4, kk = lane%4) streams A[row g][k0 + 2kk, +1] with one 16-byte load per 8-row block and step and
// feeds the two elements to two MMAs whose k-slot kk stands for column k0+2kk resp. k0+2kk+1 (the sum over k does not
// care about the order); the B fragments X[g][k0 + 2kk, +1] are one 16-byte load per step shared by all MB blocks.
// (Measured: 32-byte LDG.256 loads - a full line per row and instruction - are slower here, 15.9 vs 14.5 ms.)
template <int MB, int U>
__global__ void __launch_bounds__(32 * DMMA_WARPS) k_gemm_dmma(const double *__restrict__ A, size_t ld, size_t rows, size_t cols,
                                                              const double *__restrict__ X, size_t ldx, int nrhs,
                                                              double *__restrict__ P, size_t ldp, size_t kchunk, const int *skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31, g = lane >> 2, kk = lane & 3;
  const size_t warp = (size_t)blockIdx.x * DMMA_WARPS + (threadIdx.x >> 5);
  const size_t r0 = warp * (8 * MB);
  if (r0 >= rows) return;
  const size_t kbeg = (size_t)blockIdx.y * kchunk, kend = min(cols, kbeg + kchunk);  // kchunk is a multiple of 8
  const double *a[MB];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) a[mb] = A + min(r0 + 8 * mb + g, rows - 1) * ld + 2 * kk;
  const bool xon = g < nrhs;
  const double *x = X + (size_t)(xon ? g : 0) * ldx + 2 * kk;
  double c[MB][2][2];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) c[mb][0][0] = c[mb][0][1] = c[mb][1][0] = c[mb][1][1] = 0.0;
  size_t k0 = kbeg;
  for (; k0 + 8 * U <= kend; k0 += 8 * U) {  // full steps: no masking
    double2 av[U][MB], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) av[u][mb] = ld_stream2(a[mb] + k0 + 8 * u);
      xv[u] = xon ? *reinterpret_cast<const double2 *>(x + k0 + 8 * u) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        dmma_m8n8k4(c[mb][0][0], c[mb][0][1], av[u][mb].x, xv[u].x);
        dmma_m8n8k4(c[mb][1][0], c[mb][1][1], av[u][mb].y, xv[u].y);
      }
  }
  for (; k0 < kend; k0 += 8) {  // tail: the matrix is zero padded up to ld, the vectors are masked
    const size_t col = k0 + 2 * kk;
    double2 xv = make_double2(0.0, 0.0);
    if (xon && col < kend) xv.x = x[k0];
    if (xon && col + 1 < kend) xv.y = x[k0 + 1];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const double2 av = ld_stream2(a[mb] + k0);
      dmma_m8n8k4(c[mb][0][0], c[mb][0][1], av.x, xv.x);
      dmma_m8n8k4(c[mb][1][0], c[mb][1][1], av.y, xv.y);
    }
  }
  // C fragment: lane holds rows g, right-hand sides 2kk and 2kk+1
  double *Pk = P + (size_t)blockIdx.y * 8 * ldp;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    const size_t row = r0 + 8 * mb + g;
    if (row < rows) {
      if (2 * kk < nrhs) Pk[(size_t)(2 * kk) * ldp + row] = c[mb][0][0] + c[mb][1][0];
      if (2 * kk + 1 < nrhs) Pk[(size_t)(2 * kk + 1) * ldp + row] = c[mb][0][1] + c[mb][1][1];
    }
  }
}

__global__ void k_sum_ksplit(size_t rows, int nrhs, int ksplit, const double *__restrict__ P, size_t ldp, double *__restrict__ Y,
                             size_t ldy, const int *skip, const double *__restrict__ r1_u, const double *__restrict__ r1_dots) {
  if (skip && *skip) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= rows || n >= nrhs) return;
  double s = 0.0;
  for (int k = 0; k < ksplit; ++k) s += P[((size_t)k * 8 + n) * ldp + i];
  Y[(size_t)n * ldy + i] = r1_u ? fma(r1_u[i], r1_dots[n], s) : s;
}

#ifndef BS_DMMA_MB
#define BS_DMMA_MB 2
#endif
#ifndef BS_DMMA_U
#define BS_DMMA_U 4
#endif

static void gemm_dmma(Context &c, const DMat &M, int nrhs, const double *X, size_t ldx, double *Y, size_t ldy, const int *skip,
                      const double *r1_dots) {
  constexpr int MB = BS_DMMA_MB, U = BS_DMMA_U;
  const size_t row_warps = (M.rows + 8 * MB - 1) / (8 * MB);
  const unsigned gx = (unsigned)((row_warps + DMMA_WARPS - 1) / DMMA_WARPS);
  // enough CTAs for ~8 waves of 16 per SM; at least 2048 columns per chunk
  int ksplit = (int)std::min<size_t>(16, std::max<size_t>(1, ((size_t)c.sm_count * 16 * 8 + gx - 1) / gx));
  ksplit = (int)std::min<size_t>(ksplit, std::max<size_t>(1, M.cols / 2048));
  size_t kchunk = (M.cols + ksplit - 1) / ksplit;
  kchunk = (kchunk + 7) & ~(size_t)7;
  ksplit = (int)((M.cols + kchunk - 1) / kchunk);
  const size_t ldp = (M.rows + 1) & ~(size_t)1;
  double *P = c.wsd("gemm.partial", (size_t)ksplit * 8 * ldp);
  k_gemm_dmma<MB, U><<<dim3(gx, ksplit), 32 * DMMA_WARPS, 0, c.stream>>>(M.p, M.ld, M.rows, M.cols, X, ldx, nrhs, P, ldp, kchunk, skip);
  BS_CUDA(cudaGetLastError());
  k_sum_ksplit<<<dim3((unsigned)((M.rows + 255) / 256), nrhs), 256, 0, c.stream>>>(M.rows, nrhs, ksplit, P, ldp, Y, ldy, skip, M.r1_u, r1_dots);
  BS_CUDA(cudaGetLastError());
  count_launch(c, 2);
}

void gemv_multi(Context &c, const DMat &M, int nrhs, const double *X, size_t ldx, double *Y, size_t ldy, const int *skip) {
  if (M.rows == 0) return;
  BS_REQUIRE((ldx & 1) == 0, "multi-vector ld must be even");
  int done = 0;
  while (done < nrhs) {
    const int nr = std::min(8, nrhs - done);
    const double *Xp = X + (size_t)done * ldx;
    double *Yp = Y + (size_t)done * ldy;
    const double *r1_dots = rank1_dots(c, M, nr, Xp, ldx);
    const bool aligned16 = (ldx % 2 == 0) && (reinterpret_cast<uintptr_t>(Xp) % 16 == 0) && (M.ld % 8 == 0) &&
                           (reinterpret_cast<uintptr_t>(M.p) % 16 == 0);
    if (nr >= 4 && aligned16 && !std::getenv("BS_NO_DMMA")) {  // FP64 tensor path; up to 3 right-hand sides the FMA kernel is HBM bound
      gemm_dmma(c, M, nr, Xp, ldx, Yp, ldy, skip, r1_dots);
      done += nr;
      continue;
    }
    switch (nr) {
      case 1: launch_gemv_multi<1>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      case 2: launch_gemv_multi<2>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      case 3: launch_gemv_multi<3>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      case 4: launch_gemv_multi<4>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      case 5: launch_gemv_multi<5>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      case 6: launch_gemv_multi<6>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      case 7: launch_gemv_multi<7>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
      default: launch_gemv_multi<8>(c, M, Xp, ldx, Yp, ldy, skip, r1_dots); break;
    }
    done += nr;
  }
}

// ---------------------------------------------------------------------------------------------------------
// element-wise matrix kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void k_rank1(double *M, size_t ld, size_t rows, size_t cols, const double *__restrict__ u,
                        const double *__restrict__ w, double scale) {
  const size_t r = blockIdx.y;
  if (r >= rows) return;
  const double ur = u[r] * scale;
  double *row = M + r * ld;
  for (size_t cc = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cc < cols; cc += (size_t)gridDim.x * blockDim.x)
    row[cc] = fma(ur, w[cc], row[cc]);
}

void rank1_update(Context &c, DMat &M, const double *u, const double *w, double scale) {
  if (M.rows == 0) return;
  const unsigned gx = (unsigned)std::min<size_t>((M.cols + 255) / 256, 64);
  size_t done = 0;
  while (done < M.rows) {  // gridDim.y limit 65535
    const size_t nr = std::min<size_t>(M.rows - done, 65535);
    k_rank1<<<dim3(gx, (unsigned)nr), 256, 0, c.stream>>>(M.p + done * M.ld, M.ld, nr, M.cols, u + done, w, scale);
    BS_CUDA(cudaGetLastError());
    count_launch(c);
    done += nr;
  }
}

// K(3k+j, 3(p0+k)+m) -= C[m][3k+j] ; += delta_jm unless use_internal_alpha   (ref: bem_stokes.cc:3076-3092)
__global__ void k_correct_diag(double *K, size_t ld, int nloc, int p0, const double *__restrict__ C, size_t ldc, int alpha,
                               const unsigned char *__restrict__ skip_node) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nloc * 9) return;
  const int k = idx / 9, jm = idx - 9 * k, j = jm / 3, m = jm - 3 * j;
  if (skip_node && skip_node[k]) return;  // constrained node (ref: bem_stokes.cc:3078)
  const size_t row = (size_t)3 * k + j;
  double v = K[row * ld + (size_t)3 * (p0 + k) + m] - C[(size_t)m * ldc + row];
  if (j == m && !alpha) v += 1.0;
  K[row * ld + (size_t)3 * (p0 + k) + m] = v;
}

void k_correct_diag(Context &c, DMat &K, const double *Ck, int use_internal_alpha) {
  const int nloc = c.p1 - c.p0;
  if (nloc == 0) return;
  k_correct_diag<<<(nloc * 9 + 255) / 256, 256, 0, c.stream>>>(K.p, K.ld, nloc, c.p0, Ck, c.rows_loc, use_internal_alpha,
                                                               c.n_cons_owned ? c.d_cons_node.p : nullptr);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

__global__ void k_extract_diag(const double *M, size_t ld, size_t rows, size_t row_offset, double *out, const double *r1_u,
                               const double *r1_w, size_t r1_ncols) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double v = M[r * ld + row_offset + r];
  if (r1_u && row_offset + r < r1_ncols) v = fma(r1_u[r], r1_w[row_offset + r], v);
  out[r] = v;
}
void extract_diag(Context &c, const DMat &M, size_t row_offset, double *d_out) {
  if (!M.rows) return;
  k_extract_diag<<<(unsigned)((M.rows + 255) / 256), 256, 0, c.stream>>>(M.p, M.ld, M.rows, row_offset, d_out, M.r1_u, M.r1_w, M.r1_ncols);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

// A(:, j) = flag[j] ? -K(:, j) : V(:, j)   for j < ncols (3N).  alias: A is V's storage (only -K columns written)
__global__ void k_select_columns(double *A, const double *__restrict__ V, const double *__restrict__ K, size_t ld,
                                 size_t rows, size_t ncols, const unsigned char *__restrict__ flag, int alias) {
  const size_t r = blockIdx.y;
  if (r >= rows) return;
  for (size_t cc = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cc < ncols; cc += (size_t)gridDim.x * blockDim.x) {
    const bool isK = flag && flag[cc];
    if (isK) A[r * ld + cc] = -K[r * ld + cc];
    else if (!alias) A[r * ld + cc] = V[r * ld + cc];
  }
}
void select_columns(Context &c, DMat &A, const DMat &V, const DMat &K, const unsigned char *d_flag, bool alias) {
  const size_t rows = V.rows;
  if (!rows) return;
  if (alias && !d_flag) return;
  const unsigned gx = (unsigned)std::min<size_t>((V.cols + 255) / 256, 64);
  size_t done = 0;
  while (done < rows) {
    const size_t nr = std::min<size_t>(rows - done, 65535);
    k_select_columns<<<dim3(gx, (unsigned)nr), 256, 0, c.stream>>>(A.p + done * A.ld, V.p + done * V.ld,
                                                                  K.valid() ? K.p + done * K.ld : nullptr, A.ld, nr, V.cols,
                                                                  d_flag, alias ? 1 : 0);
    BS_CUDA(cudaGetLastError());
    count_launch(c);
    done += nr;
  }
}

// rows of this rank x all columns; flagged columns only are written.  Ck[m][row] = (K e_m)[row] on the owned rows.
__global__ void k_scatter_flagged(double *A, size_t ld, size_t rows, size_t ncols, const double *__restrict__ Kflag, size_t ldk,
                                  const int *__restrict__ kcol, const double *__restrict__ Ck, size_t ldc, int p0, int alpha) {
  const size_t r = blockIdx.y;
  if (r >= rows) return;
  const size_t node_row = r / 3, jrow = r - 3 * node_row;  // row = (local node, component j)
  for (size_t cc = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cc < ncols; cc += (size_t)gridDim.x * blockDim.x) {
    const int kc = kcol[cc];
    if (kc < 0) continue;
    double v = Kflag[r * ldk + kc];  // = -K(r, cc)
    const size_t node_col = cc / 3, m = cc - 3 * node_col;
    if (node_col == (size_t)p0 + node_row) {  // own diagonal block: -(K - C_m[r] + delta_jm (1 - alpha))
      v += Ck[m * ldc + r];
      if (jrow == m && !alpha) v -= 1.0;
    }
    A[r * ld + cc] = v;
  }
}
void scatter_flagged_columns(Context &c, DMat &A, const double *Kflag, size_t ldk, const int *kcol, const double *Ck, int alpha) {
  const size_t rows = c.rows_loc, ncols = c.n3();
  if (!rows) return;
  const unsigned gx = (unsigned)std::min<size_t>((ncols + 255) / 256, 64);
  for (size_t done = 0; done < rows; done += 65535) {
    const size_t nr = std::min<size_t>(rows - done, 65535);
    k_scatter_flagged<<<dim3(gx, (unsigned)nr), 256, 0, c.stream>>>(A.p + done * A.ld, A.ld, nr, ncols, Kflag + done * ldk, ldk, kcol,
                                                                  Ck + done, rows, c.p0 + (int)(done / 3), alpha);
    BS_CUDA(cudaGetLastError());
    count_launch(c);
  }
}

__global__ void k_set_column(double *A, size_t ld, size_t rows, size_t col, const double *__restrict__ v, double scale) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) A[r * ld + col] = v ? scale * v[r] : scale;
}
void set_column(Context &c, DMat &A, size_t col, const double *v, double scale) {
  if (!A.rows) return;
  k_set_column<<<(unsigned)((A.rows + 255) / 256), 256, 0, c.stream>>>(A.p, A.ld, A.rows, col, v, scale);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

__global__ void k_gather_entries(const double *M, size_t ld, int n, const int *r, const int *cidx, double *out, const double *r1_u,
                                 const double *r1_w, size_t r1_ncols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = M[(size_t)r[i] * ld + cidx[i]];
  if (r1_u && (size_t)cidx[i] < r1_ncols) v = fma(r1_u[r[i]], r1_w[cidx[i]], v);
  out[i] = v;
}
void gather_entries(Context &c, const DMat &M, int n, const int *d_r, const int *d_c, double *d_out) {
  if (!n) return;
  k_gather_entries<<<(n + 255) / 256, 256, 0, c.stream>>>(M.p, M.ld, n, d_r, d_c, d_out, M.r1_u, M.r1_w, M.r1_ncols);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

// one CTA per constrained owned row: the row becomes the constraint equation x_ii - sum coef_k x_k = 0
__global__ void k_constraint_rows(double *M, size_t ld, size_t ncols, size_t diag_off, const int *rows, const int *ptr, const int *cols,
                                  const double *coefs) {
  const int r = rows[blockIdx.x];
  double *row = M + (size_t)r * ld;
  for (size_t j = threadIdx.x; j < ncols; j += blockDim.x) row[j] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    row[diag_off + r] = 1.0;
    for (int k = ptr[blockIdx.x]; k < ptr[blockIdx.x + 1]; ++k) row[cols[k]] = -coefs[k];
  }
}
void apply_constraint_rows(Context &c, DMat &M, size_t ncols) {
  if (!c.n_cons_owned) return;
  k_constraint_rows<<<c.n_cons_owned, 256, 0, c.stream>>>(M.p, M.ld, ncols, (size_t)3 * c.p0, c.d_cons_row.p, c.d_cons_ptr.p,
                                                          c.d_cons_col.p, c.d_cons_coef.p);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
__global__ void k_zero_rows(double *v, const int *rows, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[rows[i]] = 0.0;
}
void zero_constrained_entries(Context &c, double *v_loc) {
  if (!c.n_cons_owned) return;
  k_zero_rows<<<(c.n_cons_owned + 255) / 256, 256, 0, c.stream>>>(v_loc, c.d_cons_row.p, c.n_cons_owned);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

__global__ void k_add_rank1_block(double *dst, size_t ldd, size_t n, const double *__restrict__ u, const double *__restrict__ w,
                                  size_t wcols) {
  const size_t i = blockIdx.y;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n && j < wcols; j += (size_t)gridDim.x * blockDim.x)
    dst[i * ldd + j] = fma(u[i], w[j], dst[i * ldd + j]);
}
void add_rank1_block(Context &c, const DMat &M, size_t row_off, size_t col_off, size_t n, double *dst, size_t ldd) {
  if (!M.r1_u || n == 0 || col_off >= M.r1_ncols) return;
  const unsigned gx = (unsigned)std::min<size_t>((n + 255) / 256, 64);
  for (size_t done = 0; done < n; done += 65535) {
    const size_t nr = std::min<size_t>(n - done, 65535);
    k_add_rank1_block<<<dim3(gx, (unsigned)nr), 256, 0, c.stream>>>(dst + done * ldd, ldd, n, M.r1_u + row_off + done, M.r1_w + col_off,
                                                                  M.r1_ncols - col_off);
    BS_CUDA(cudaGetLastError());
    count_launch(c);
  }
}

// ---------------------------------------------------------------------------------------------------------
// vector kernels
// ---------------------------------------------------------------------------------------------------------
// reference ordering (i + c*N) <-> internal ordering (3*pos + c); `extra` trailing entries copied through
__global__ void k_perm_in(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ node_of_pos,
                          size_t N, int extra) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * N) {
    const size_t pos = i / 3, cc = i - 3 * pos;
    dst[i] = src[(size_t)node_of_pos[pos] + cc * N];
  } else if (i < 3 * N + extra) {
    dst[i] = src[i];
  }
}
__global__ void k_perm_out(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ node_of_pos,
                           size_t N, int extra, size_t lo, size_t hi) {
  // only internal entries [lo,hi) are written (this rank's slice); src is indexed by full internal index
  const size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  if (i < 3 * N) {
    const size_t pos = i / 3, cc = i - 3 * pos;
    dst[(size_t)node_of_pos[pos] + cc * N] = src[i];
  } else if (i < 3 * N + extra) {
    dst[i] = src[i];
  }
}

// ---- reductions -------------------------------------------------------------------------------------------
// out[k] = sum_i basis[k*ldb + i] * w[i], k < nk  (one CTA per (k, chunk); atomicAdd of partials into zeroed out)
__global__ void k_multi_dot(const double *__restrict__ basis, size_t ldb, int nk, const double *__restrict__ w, size_t n,
                            double *out) {
  const int k = blockIdx.y;
  const double *b = basis + (size_t)k * ldb;
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s = fma(b[i], w[i], s);
  __shared__ double red[8];
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int m = 4; m > 0; m >>= 1) s += __shfl_xor_sync(0xffu, s, m);
    if (threadIdx.x == 0) atomicAdd(out + k, s);
  }
}

// deterministic two-stage variant: partial[k][chunk] then a second tiny kernel sums chunks in fixed order
__global__ void k_multi_dot_partial(const double *__restrict__ basis, size_t ldb, int nk, const double *__restrict__ w,
                                    size_t n, double *partial, int nchunks) {
  const int k = blockIdx.y;
  const double *b = basis + (size_t)k * ldb;
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s = fma(b[i], w[i], s);
  __shared__ double red[8];
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int m = 4; m > 0; m >>= 1) s += __shfl_xor_sync(0xffu, s, m);
    if (threadIdx.x == 0) partial[(size_t)k * nchunks + blockIdx.x] = s;
  }
}
__global__ void k_sum_partials(const double *partial, int nchunks, int nk, double *out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nk) return;
  double s = 0.0;
  for (int j = 0; j < nchunks; ++j) s += partial[(size_t)k * nchunks + j];
  out[k] = s;
}

void multi_dot(Context &c, const double *basis, size_t ldb, int k, const double *w, size_t n, double *d_out) {
  if (k <= 0) return;
  const int nchunks = (int)std::min<size_t>(std::max<size_t>((n + 256 * 8 - 1) / (256 * 8), 1), 64);
  c.d_tmp2.alloc(std::max(c.d_tmp2.n, (size_t)k * nchunks));
  k_multi_dot_partial<<<dim3(nchunks, k), 256, 0, c.stream>>>(basis, ldb, k, w, n, c.d_tmp2.p, nchunks);
  BS_CUDA(cudaGetLastError());
  k_sum_partials<<<(k + 63) / 64, 64, 0, c.stream>>>(c.d_tmp2.p, nchunks, k, d_out);
  BS_CUDA(cudaGetLastError());
  count_launch(c, 2);
}

// w[i] += sign * sum_k coef[k] * basis[k][i]
__global__ void k_multi_axpy(const double *__restrict__ basis, size_t ldb, int nk, const double *__restrict__ coef, double sign,
                             double *w, size_t n) {
  extern __shared__ double cs[];
  for (int k = threadIdx.x; k < nk; k += blockDim.x) cs[k] = sign * coef[k];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double s = w[i];
    for (int k = 0; k < nk; ++k) s = fma(cs[k], basis[(size_t)k * ldb + i], s);
    w[i] = s;
  }
}
void multi_axpy(Context &c, const double *basis, size_t ldb, int k, const double *d_coef, double sign, double *w, size_t n) {
  if (k <= 0 || n == 0) return;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 1184);
  k_multi_axpy<<<grid, 256, k * sizeof(double), c.stream>>>(basis, ldb, k, d_coef, sign, w, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

__global__ void k_axpy(double a, const double *__restrict__ x, double *y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = fma(a, x[i], y[i]);
}
__global__ void k_scal(double a, double *x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= a;
}
__global__ void k_scal_dev_inv(const double *s, double *x, const double *src, size_t n) {
  const double a = 1.0 / *s;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = src[i] * a;
}
__global__ void k_sub(const double *a, const double *b, double *o, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = a[i] - b[i];
}
__global__ void k_mul(const double *a, const double *d, double *o, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = a[i] * d[i];
}
__global__ void k_fill(double *x, double v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}
static inline unsigned vgrid(size_t n) { return (unsigned)std::min<size_t>(std::max<size_t>((n + 255) / 256, 1), 1184); }

void axpy(Context &c, double a, const double *x, double *y, size_t n) {
  if (!n) return;
  k_axpy<<<vgrid(n), 256, 0, c.stream>>>(a, x, y, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void scal(Context &c, double a, double *x, size_t n) {
  if (!n) return;
  k_scal<<<vgrid(n), 256, 0, c.stream>>>(a, x, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void scal_dev_inv(Context &c, const double *d_s, double *x, const double *src, size_t n) {
  if (!n) return;
  k_scal_dev_inv<<<vgrid(n), 256, 0, c.stream>>>(d_s, x, src, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void copy(Context &c, const double *src, double *dst, size_t n) {
  if (n) BS_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
}
void sub(Context &c, const double *a, const double *b, double *out, size_t n) {
  if (!n) return;
  k_sub<<<vgrid(n), 256, 0, c.stream>>>(a, b, out, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void mul_elem(Context &c, const double *a, const double *d, double *out, size_t n) {
  if (!n) return;
  k_mul<<<vgrid(n), 256, 0, c.stream>>>(a, d, out, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void fill(Context &c, double *x, double v, size_t n) {
  if (!n) return;
  k_fill<<<vgrid(n), 256, 0, c.stream>>>(x, v, n);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

double dot(Context &c, const double *a, const double *b, size_t n) {
  c.d_small.alloc(std::max<size_t>(c.d_small.n, 512));
  multi_dot(c, a, 0, 1, b, n, c.d_small.p);
  double h = 0;
  BS_CUDA(cudaMemcpyAsync(&h, c.d_small.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  BS_CUDA(cudaStreamSynchronize(c.stream));
  return h;
}

// the permutation kernels need the node_of_pos table on the device; it is owned by the context
void perm_in(Context &c, const double *src_dev, double *dst_int, const int *d_node_of_pos, int nextra) {
  const size_t n = c.n3() + nextra;
  k_perm_in<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(src_dev, dst_int, d_node_of_pos, (size_t)c.N, nextra);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void perm_out(Context &c, const double *src_int, double *dst_dev, const int *d_node_of_pos, int nextra, size_t lo, size_t hi) {
  if (hi <= lo) return;
  k_perm_out<<<(unsigned)((hi - lo + 255) / 256), 256, 0, c.stream>>>(src_int, dst_dev, d_node_of_pos, (size_t)c.N, nextra, lo, hi);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

}  // namespace bs
