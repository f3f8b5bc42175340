// Device solvers: left-preconditioned restarted GMRES with deal.II SolverGMRES semantics (ref call sites
// source/bem_stokes.cc:4116, 4332; semantics SURVEY A.7), dense blocked LU with partial pivoting (replaces
// TrilinosWrappers::SolverDirect / Amesos KLU behind DirectPreconditioner, source/direct_preconditioner.cc).
#include "bs_internal.h"
#include <cmath>
#include <algorithm>
#include <chrono>
#include <cstdlib>

namespace bs {

size_t Context::full_vec_len(int which) const { return which == BS_MAT_A ? n3() + num_rigid : n3(); }
size_t Context::slice_offset(int) const { return (size_t)3 * p0; }
size_t Context::local_vec_len(int which) const {
  return rows_loc + ((which == BS_MAT_A && rank == nranks - 1) ? (size_t)num_rigid : 0);
}

static const DMat &mat_of(Context &c, int which) {
  const DMat &M = which == BS_MAT_V ? c.V : (which == BS_MAT_K ? c.K : c.A);
  BS_REQUIRE(M.valid(), which == BS_MAT_A ? "monolithic matrix not built" : "matrix not assembled");
  return M;
}

void apply_operator(Context &c, int which, const double *x_full, double *y_loc) { gemv(c, mat_of(c, which), x_full, y_loc); }

// slice of every rank -> replicated full vector (the Epetra Import of the reference's vmult)
void exchange(Context &c, int which, const double *y_loc, double *x_full) {
  const size_t mloc = c.local_vec_len(which);
  if (c.nranks == 1) {
    if (y_loc != x_full) copy(c, y_loc, x_full, mloc);
    return;
  }
  if (c.p2p && c.full_vec_len(which) <= c.xchg_ld) {
    // peer stores over NVLink instead of a collective call: every rank writes its slice into slot 0 of all replicated
    // buffers, publishes it and waits for the others
    p2p_scatter(c, which, y_loc, 0, nullptr, nullptr);
    p2p_wait(c);
    copy(c, c.d_xchg.p, x_full, c.full_vec_len(which));
    // the next scatter into slot 0 must not overtake a peer that is still copying: a second handshake closes the epoch
    p2p_wait(c);
    return;
  }
  BS_REQUIRE(c.cb_allgatherv != nullptr, "nranks > 1 but no communicator callbacks set (bs_set_comm)");
  std::vector<int> counts(c.nranks), displs(c.nranks);
  for (int r = 0; r < c.nranks; ++r) {
    counts[r] = 3 * (c.part_start[r + 1] - c.part_start[r]) + ((which == BS_MAT_A && r == c.nranks - 1) ? c.num_rigid : 0);
    displs[r] = 3 * c.part_start[r];
  }
  int rc = c.cb_allgatherv(c.cb_user, y_loc, (int)mloc, x_full, counts.data(), displs.data(), (void *)c.stream);
  if (rc != 0) throw Error(BS_ERR_COMM, "allgatherv callback failed");
}

static void allreduce(Context &c, double *d_buf, int count) {
  if (c.nranks == 1) return;
  BS_REQUIRE(c.cb_allreduce != nullptr, "nranks > 1 but no communicator callbacks set (bs_set_comm)");
  int rc = c.cb_allreduce(c.cb_user, d_buf, count, (void *)c.stream);
  if (rc != 0) throw Error(BS_ERR_COMM, "allreduce callback failed");
}

// ---------------------------------------------------------------------------------------------------------
// Peer-memory exchange over NVLink (CUDA-IPC mapped buffers): the kernel that produces a Krylov vector stores this
// rank's slice into the replicated buffer of EVERY rank; a release store per peer publishes it; the consumer
// acquire-waits on its own flag array before the matvec.  No collective call, no staging copy.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_p2p_scatter(const double *__restrict__ src, size_t n, const double *inv_norm2, double *basis_dst,
                              double *const *peer_bufs, int nranks, size_t dst_off) {
  const double a = inv_norm2 ? rsqrt(*inv_norm2) : 1.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = src[i] * a;
    if (basis_dst) basis_dst[i] = v;
    for (int r = 0; r < nranks; ++r) peer_bufs[r][dst_off + i] = v;  // own buffer included; peers over NVLink
  }
}
// store my slice (optionally normalised by 1/sqrt(*inv_norm2), optionally also into the local basis) into slot
// `slot` of every rank's replicated buffer
void p2p_scatter(Context &c, int which, const double *src_loc, int slot, const double *inv_norm2, double *basis_dst) {
  BS_REQUIRE(slot < Context::XCHG_SLOTS, "too many vectors in flight for the peer exchange buffer");
  BS_REQUIRE(c.full_vec_len(which) <= c.xchg_ld, "exchange buffer too small (bs_exchange_export max_vec_len)");
  const size_t mloc = c.local_vec_len(which);
  const size_t off = (size_t)slot * c.xchg_ld + c.slice_offset(which);
  const unsigned grid = (unsigned)std::min<size_t>(std::max<size_t>((mloc + 255) / 256, 1), 592);
  k_p2p_scatter<<<grid, 256, 0, c.stream>>>(src_loc, mloc, inv_norm2, basis_dst, c.d_peer_xbuf.p, c.nranks, off);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
// publish everything scattered so far and wait until every rank has published the same epoch
void p2p_wait(Context &c) {
  p2p_signal(c);  // the epoch is a device-resident counter (bs_gmres.cu), shared with the device-resident solver
  p2p_wait_only(c, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------------------
// LU with partial pivoting, blocked right-looking, everything on the device
// ---------------------------------------------------------------------------------------------------------
constexpr int LU_NB = 32;

// factorise the panel A[k0:n, k0:k0+nb] with one CTA; piv[k0+j] = absolute pivot row
__global__ void __launch_bounds__(1024) k_lu_panel(double *A, size_t ld, int n, int k0, int nb, int *piv, int *info) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = 0; j < nb; ++j) {
    const int col = k0 + j;
    // pivot search
    double best = -1.0;
    int bi = col;
    for (int r = col + tid; r < n; r += blockDim.x) {
      const double v = fabs(A[(size_t)r * ld + col]);
      if (v > best) {
        best = v;
        bi = r;
      }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, m);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_val[wid] = best;
      s_idx[wid] = bi;
    }
    __syncthreads();
    if (wid == 0) {
      best = (lane < (blockDim.x >> 5)) ? s_val[lane] : -1.0;
      bi = (lane < (blockDim.x >> 5)) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, m);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (ov > best || (ov == best && oi < bi)) {
          best = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_piv = bi;
        piv[col] = bi;
      }
    }
    __syncthreads();
    const int pr = s_piv;
    // swap rows col <-> pr inside the panel
    if (pr != col && tid < nb) {
      const double a = A[(size_t)col * ld + k0 + tid], b = A[(size_t)pr * ld + k0 + tid];
      A[(size_t)col * ld + k0 + tid] = b;
      A[(size_t)pr * ld + k0 + tid] = a;
    }
    __syncthreads();
    const double pv = A[(size_t)col * ld + col];
    const double dinv = pv != 0.0 ? 1.0 / pv : 0.0;  // singular column: leave it, reported through `info`
    if (pv == 0.0 && tid == 0) atomicCAS(info, 0, col + 1);
    // scale column and update the rest of the panel: thread <-> row
    for (int r = col + 1 + tid; r < n; r += blockDim.x) {
      double *row = A + (size_t)r * ld;
      const double l = row[col] * dinv;
      row[col] = l;
      for (int jj = j + 1; jj < nb; ++jj) row[k0 + jj] = fma(-l, A[(size_t)col * ld + k0 + jj], row[k0 + jj]);
    }
    __syncthreads();
  }
}

// Multi-CTA panel factorisation (cooperative launch: all CTAs resident, software grid barrier on a monotonic
// counter).  Rows col..n of the panel are split into contiguous chunks, one per CTA; per column: local pivot
// candidates -> barrier -> every CTA reduces the candidates (same winner everywhere), CTA 0 swaps the two rows inside
// the panel -> barrier -> all CTAs scale their part of the column and apply the rank-1 update to the rest of the panel.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*((volatile unsigned int *)counter) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// Register-resident variant: every thread keeps its (at most RPT) panel rows of 32 columns in registers for the
// whole panel; global memory is touched only for the pivot candidates, the two rows exchanged per column (published
// through a scratch buffer that also broadcasts the pivot row) and the final write-back.  Rows are assigned once
// per panel: row r -> CTA (r-k0)/chunk, thread ((r-k0)%chunk)%256.  Two grid barriers per column.
constexpr int LU_RPT = 2;
__global__ void __launch_bounds__(256) k_lu_panel_coop(double *A, size_t ld, int n, int k0, int nb, int *piv, double *cand_val,
                                                        int *cand_idx, double *xrow /*[2][32]*/, unsigned int *counter,
                                                        int chunk, int *info) {
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_piv;
  __shared__ double s_prow[LU_NB];
  const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int row[LU_RPT];
  double a[LU_RPT][LU_NB];
#pragma unroll
  for (int rr = 0; rr < LU_RPT; ++rr) {
    const int local = rr * 256 + tid;
    row[rr] = (local < chunk && k0 + g * chunk + local < n) ? k0 + g * chunk + local : -1;
#pragma unroll
    for (int j = 0; j < LU_NB; ++j) a[rr][j] = (row[rr] >= 0 && j < nb) ? A[(size_t)row[rr] * ld + k0 + j] : 0.0;
  }
  unsigned int nbar = 0;
#pragma unroll
  for (int j = 0; j < LU_NB; ++j) {
    if (j < nb) {  // nb is uniform over the grid
      const int col = k0 + j;
      // ---- pivot candidates from registers
      double best = -1.0;
      int bi = 0x7fffffff;
#pragma unroll
      for (int rr = 0; rr < LU_RPT; ++rr)
        if (row[rr] >= col) {
          const double v = fabs(a[rr][j]);
          if (v > best || (v == best && row[rr] < bi)) {
            best = v;
            bi = row[rr];
          }
        }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, m);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (ov > best || (ov == best && oi < bi)) {
          best = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_val[wid] = best;
        s_idx[wid] = bi;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < 8; ++w)
          if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) {
            best = s_val[w];
            bi = s_idx[w];
          }
        cand_val[g] = best;
        cand_idx[g] = bi;
      }
      grid_barrier(counter, (++nbar) * G);
      if (wid == 0) {  // every CTA reduces the G candidates: identical winner (largest value, lowest row on ties)
        best = -1.0;
        bi = 0x7fffffff;
        for (int q = lane; q < G; q += 32) {
          const double v = ((volatile double *)cand_val)[q];
          const int ix = ((volatile int *)cand_idx)[q];
          if (v > best || (v == best && ix < bi)) {
            best = v;
            bi = ix;
          }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, best, m);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
          if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
          }
        }
        if (lane == 0) s_piv = bi;
      }
      __syncthreads();
      const int pr = s_piv;
      if (g == 0 && tid == 0) piv[col] = pr;
      // ---- publish the two rows: xrow[0] = old row `col`, xrow[1] = old row `pr` (= the new pivot row)
#pragma unroll
      for (int rr = 0; rr < LU_RPT; ++rr) {
        if (row[rr] == col) {
#pragma unroll
          for (int jj = 0; jj < LU_NB; ++jj) xrow[jj] = a[rr][jj];
        }
        if (row[rr] == pr) {
#pragma unroll
          for (int jj = 0; jj < LU_NB; ++jj) xrow[LU_NB + jj] = a[rr][jj];
        }
      }
      grid_barrier(counter, (++nbar) * G);
      if (tid < LU_NB) s_prow[tid] = ((volatile double *)xrow)[LU_NB + tid];
      __syncthreads();
#pragma unroll
      for (int rr = 0; rr < LU_RPT; ++rr) {
        if (row[rr] == col) {  // receives the pivot row
#pragma unroll
          for (int jj = 0; jj < LU_NB; ++jj) a[rr][jj] = s_prow[jj];
        } else if (row[rr] == pr) {  // receives the old row `col` and is eliminated like every other row below
#pragma unroll
          for (int jj = 0; jj < LU_NB; ++jj) a[rr][jj] = ((volatile double *)xrow)[jj];
        }
      }
      const double dinv = s_prow[j] != 0.0 ? 1.0 / s_prow[j] : 0.0;
      if (s_prow[j] == 0.0 && g == 0 && tid == 0) atomicCAS(info, 0, col + 1);
#pragma unroll
      for (int rr = 0; rr < LU_RPT; ++rr)
        if (row[rr] > col) {
          const double l = a[rr][j] * dinv;
          a[rr][j] = l;
#pragma unroll
          for (int jj = j + 1; jj < LU_NB; ++jj) a[rr][jj] = fma(-l, s_prow[jj], a[rr][jj]);
        }
      __syncthreads();  // s_prow is rewritten in the next column
    }
  }
#pragma unroll
  for (int rr = 0; rr < LU_RPT; ++rr)
    if (row[rr] >= 0)
#pragma unroll
      for (int j = 0; j < LU_NB; ++j)
        if (j < nb) A[(size_t)row[rr] * ld + k0 + j] = a[rr][j];
}

// apply the panel's row swaps to columns outside the panel
__global__ void k_lu_swap(double *A, size_t ld, int n, int k0, int nb, const int *piv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n || (c >= k0 && c < k0 + nb)) return;
  for (int j = 0; j < nb; ++j) {
    const int r = k0 + j, pr = piv[r];
    if (pr != r) {
      const double a = A[(size_t)r * ld + c], b = A[(size_t)pr * ld + c];
      A[(size_t)r * ld + c] = b;
      A[(size_t)pr * ld + c] = a;
    }
  }
}

constexpr int LU_OUT = 256;  // outer block: trailing update of the columns right of it is one k=256 GEMM (halves the C traffic of k=128)

// Rows [j0, j0+nb) of U for every column c >= j0+nb:  u = L_jj^{-1} (A[j-rows][c] - L[j-rows][K0:j0) * U[K0:j0)[c]).
// Columns inside the current outer block [K0, K0+Bw) already carry the rank-32 updates of the earlier inner panels
// (left term empty); columns right of it are updated lazily here (left-looking inside the outer block).
// One thread per column; the L rows of this step are staged in shared memory.
__global__ void __launch_bounds__(128) k_lu_u12_step(double *A, size_t ld, int n, int K0, int Bw, int j0, int nb) {
  extern __shared__ double Ls[];  // [LU_NB][LU_OUT + 1]: L[j0+r][K0 + k], k < j0 + nb - K0
  const int kw = j0 + nb - K0;
  for (int i = threadIdx.x; i < nb * kw; i += blockDim.x) {
    const int r = i / kw, k = i - r * kw;
    Ls[r * (LU_OUT + 1) + k] = A[(size_t)(j0 + r) * ld + K0 + k];
  }
  __syncthreads();
  const int c = j0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double u[LU_NB];
#pragma unroll
  for (int r = 0; r < LU_NB; ++r) u[r] = (r < nb) ? A[(size_t)(j0 + r) * ld + c] : 0.0;
  if (c >= K0 + Bw) {  // lazy part: subtract L[j-rows][K0:j0) * U[K0:j0)[c]
    for (int k = 0; k < j0 - K0; ++k) {
      const double uk = A[(size_t)(K0 + k) * ld + c];
#pragma unroll
      for (int r = 0; r < LU_NB; ++r) u[r] = fma(-Ls[r * (LU_OUT + 1) + k], uk, u[r]);
    }
  }
  const int d0 = j0 - K0;  // unit-lower triangular solve with L_jj
#pragma unroll
  for (int r = 0; r < LU_NB; ++r) {
#pragma unroll
    for (int k = 0; k < r; ++k) u[r] = fma(-Ls[r * (LU_OUT + 1) + d0 + k], u[k], u[r]);
  }
#pragma unroll
  for (int r = 0; r < LU_NB; ++r)
    if (r < nb) A[(size_t)(j0 + r) * ld + c] = u[r];
}

// C[r0.., c0..] -= L[r.., k0:k0+kw) * U[k0:k0+kw)[c..]  for rows >= row_begin, columns in [col_begin, col_end).
// 128 x 128 tile per CTA, 256 threads, 8 x 8 register tile per thread, k in chunks of 16 through shared memory.
constexpr int GM_T = 128, GM_K = 16;
__global__ void __launch_bounds__(256) k_lu_gemm(double *A, size_t ld, int n, int k0, int kw, int row_begin, int col_begin,
                                                 int col_end) {
  __shared__ double Ls[GM_K][GM_T + 2];
  __shared__ double Us[GM_K][GM_T + 2];
  const int r0 = row_begin + blockIdx.y * GM_T, c0 = col_begin + blockIdx.x * GM_T;
  const int tid = threadIdx.x;
  // interleaved register tile: thread (ty, tx) owns rows ty + 16 i and columns tx + 16 j, so that the shared-memory
  // reads of a warp are either broadcasts (rows) or 16 consecutive doubles (columns): no bank conflicts
  const int ty = tid / 16, tx = tid % 16;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  for (int kk = 0; kk < kw; kk += GM_K) {
    __syncthreads();
    for (int i = tid; i < GM_T * GM_K; i += 256) {  // L tile: rows r0.., columns k0+kk.. (16 contiguous doubles per row)
      const int r = i / GM_K, k = i % GM_K;
      Ls[k][r] = (r0 + r < n && kk + k < kw) ? A[(size_t)(r0 + r) * ld + k0 + kk + k] : 0.0;
    }
    for (int i = tid; i < GM_K * GM_T; i += 256) {  // U tile: rows k0+kk.., columns c0.. (contiguous)
      const int k = i / GM_T, cc = i % GM_T;
      Us[k][cc] = (c0 + cc < col_end && kk + k < kw) ? A[(size_t)(k0 + kk + k) * ld + c0 + cc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GM_K; ++k) {
      double l[8], u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) l[i] = Ls[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = Us[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(l[i], u[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 16 * i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = c0 + tx + 16 * j;
      if (cc < col_end) A[(size_t)r * ld + cc] -= acc[i][j];
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Trailing update on the FP64 tensor path (DMMA m8n8k4):  C -= L * U  with the same index conventions as k_lu_gemm.
// The k range runs through shared memory in slices of 16 columns, three stages deep (cp.async, zero-filled at the
// matrix edges).  Fragment loads are bank-conflict free: A rows are padded to 20 doubles and B rows to 68 doubles,
// so that the 16 lanes of a half-warp (4 rows x 4 k for A, 4 k x 4 columns for B) hit 16 different 8-byte banks.
// Needs 16-byte aligned rows (ld, k0, col_begin even); k_lu_gemm stays as the fallback for odd edges.
// ---------------------------------------------------------------------------------------------------------
constexpr int DG_BM = 128, DG_BN = 64, DG_KC = 16, DG_STAGES = 3;
constexpr int DG_APAD = DG_KC + 4, DG_BPAD = DG_BN + 4;
constexpr size_t DG_STAGE_DOUBLES = (size_t)DG_BM * DG_APAD + (size_t)DG_KC * DG_BPAD;
constexpr size_t DG_SMEM = DG_STAGES * DG_STAGE_DOUBLES * sizeof(double);

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CTA tile 128 x 64, 8 warps as 4 x 2, warp tile 32 x 32 (4 x 4 MMA tiles, 32 accumulators = 64 registers per lane): two
// CTAs are resident per SM, so that the C read-modify-write of one tile overlaps the MMAs of the other
__global__ void __launch_bounds__(256, 2) k_lu_gemm_dmma(double *A, size_t ld, int n, int k0, int kw, int row_begin, int col_begin,
                                                         int col_end) {
  extern __shared__ __align__(16) double dg_smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = wid >> 1, wn = wid & 1;
  const int r0 = row_begin + blockIdx.y * DG_BM, c0 = col_begin + blockIdx.x * DG_BN;
  const int nk = (kw + DG_KC - 1) / DG_KC;
  auto load_stage = [&](int st, int kk) {
    double *As = dg_smem + (size_t)st * DG_STAGE_DOUBLES, *Bs = As + (size_t)DG_BM * DG_APAD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // L tile: 128 rows x 16 columns, 8 chunks of 16 bytes per row
      const int ch = tid + i * 256, row = ch >> 3, cq = ch & 7;
      const int gr = r0 + row, gk = kk + 2 * cq;
      const int valid = gr < n ? min(max(kw - gk, 0), 2) : 0;
      const double *src = valid ? A + (size_t)gr * ld + k0 + gk : A;
      cp_async16(As + (size_t)row * DG_APAD + 2 * cq, src, 8 * valid);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {  // U tile: 16 rows x 64 columns, 32 chunks per row
      const int ch = tid + i * 256, row = ch >> 5, cq = ch & 31;
      const int gk = kk + row, gc = c0 + 2 * cq;
      const int valid = gk < kw ? min(max(col_end - gc, 0), 2) : 0;
      const double *src = valid ? A + (size_t)(k0 + gk) * ld + gc : A;
      cp_async16(Bs + (size_t)row * DG_BPAD + 2 * cq, src, 8 * valid);
    }
  };
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int st = 0; st < DG_STAGES - 1; ++st) {
    if (st < nk) load_stage(st, st * DG_KC);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(DG_STAGES - 2) : "memory");
    __syncthreads();  // stage kt has landed for every thread; stage (kt-1) is free for the next prefetch
    if (kt + DG_STAGES - 1 < nk) load_stage((kt + DG_STAGES - 1) % DG_STAGES, (kt + DG_STAGES - 1) * DG_KC);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double *As = dg_smem + (size_t)(kt % DG_STAGES) * DG_STAGE_DOUBLES, *Bs = As + (size_t)DG_BM * DG_APAD;
    const double *ap = As + (size_t)(wm * 32 + g) * DG_APAD + t;
    const double *bp = Bs + (size_t)t * DG_BPAD + wn * 32 + g;
#pragma unroll
    for (int k4 = 0; k4 < DG_KC / 4; ++k4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = ap[(size_t)i * 8 * DG_APAD + 4 * k4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bp[(size_t)4 * k4 * DG_BPAD + 8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + wm * 32 + 8 * i + g;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = c0 + wn * 32 + 8 * j + 2 * t;
      double *p = A + (size_t)r * ld + cc;
      if (cc + 1 < col_end) {
        double2 v = *reinterpret_cast<double2 *>(p);
        v.x -= acc[i][j][0];
        v.y -= acc[i][j][1];
        *reinterpret_cast<double2 *>(p) = v;
      } else if (cc < col_end) {
        p[0] -= acc[i][j][0];
      }
    }
  }
}

// C -= L U on the tensor path when the rows are 16-byte aligned, on the FMA kernel otherwise
static void lu_trailing_update(Context &c, double *A, size_t ld, int n, int k0, int kw, int row_begin, int col_begin, int col_end) {
  const int ncols = col_end - col_begin, nrows = n - row_begin;
  if (ncols <= 0 || nrows <= 0 || kw <= 0) return;
  const bool aligned = (ld % 2 == 0) && (k0 % 2 == 0) && (col_begin % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  if (aligned && !std::getenv("BS_NO_DMMA")) {
    const dim3 grid((ncols + DG_BN - 1) / DG_BN, (nrows + DG_BM - 1) / DG_BM);
    static bool attr = false;
    if (!attr) {
      BS_CUDA(cudaFuncSetAttribute(k_lu_gemm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DG_SMEM));
      attr = true;
    }
    k_lu_gemm_dmma<<<grid, 256, DG_SMEM, c.stream>>>(A, ld, n, k0, kw, row_begin, col_begin, col_end);
  } else {
    const dim3 grid((ncols + GM_T - 1) / GM_T, (nrows + GM_T - 1) / GM_T);
    k_lu_gemm<<<grid, 256, 0, c.stream>>>(A, ld, n, k0, kw, row_begin, col_begin, col_end);
  }
  count_launch(c);
}

// Right-looking blocked LU with partial pivoting: inner panels of 32 columns (one CTA each), immediate rank-32
// updates only inside the current 128-column outer block, lazy U rows for the columns right of it, then one k=128
// GEMM for the trailing matrix.
void lu_factor(Context &c, double *A, size_t n_, size_t ld, int *piv) {
  const int n = (int)n_;
  const size_t sm_u12 = (size_t)LU_NB * (LU_OUT + 1) * sizeof(double);
  if (sm_u12 > 48 * 1024) BS_CUDA(cudaFuncSetAttribute(k_lu_u12_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_u12));
  const bool trace = std::getenv("BS_TRACE") != nullptr;
  double tacc[5] = {0, 0, 0, 0, 0};
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tlast = 0;
  if (trace) {
    cudaStreamSynchronize(c.stream);
    tlast = tnow();
  }
  int *info = c.wsi("lu.info", 4);
  BS_CUDA(cudaMemsetAsync(info, 0, sizeof(int), c.stream));
  auto tick = [&](int k) {
    if (!trace) return;
    cudaStreamSynchronize(c.stream);
    const double t = tnow();
    tacc[k] += t - tlast;
    tlast = t;
  };
  for (int K0 = 0; K0 < n; K0 += LU_OUT) {
    const int Bw = std::min(LU_OUT, n - K0);
    for (int j0 = K0; j0 < K0 + Bw; j0 += LU_NB) {
      const int nb = std::min(LU_NB, K0 + Bw - j0);
      const int rows_left = n - j0;
      // rows per CTA: at most 256*LU_RPT (register-resident rows), at least 64 so that tiny panels use few CTAs
      int chunk = std::max(64, (rows_left + c.sm_count - 1) / c.sm_count);
      if (chunk > 256 * LU_RPT || rows_left <= 1024) {
        k_lu_panel<<<1, 1024, 0, c.stream>>>(A, ld, n, j0, nb, piv, info);  // short panel, or taller than all SMs can hold
      } else {
        const int G = (rows_left + chunk - 1) / chunk;
        double *cand_val = c.wsd("lu.cand", 256 + 2 * LU_NB);
        double *xrow = cand_val + 256;
        int *cand_idx = c.wsi("lu.cand_idx", 256 + 4);
        unsigned int *counter = reinterpret_cast<unsigned int *>(cand_idx + 256);
        BS_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), c.stream));
        int n_ = n, j0_ = j0, nb_ = nb;
        size_t ld_ = ld;
        void *args[] = {&A, &ld_, &n_, &j0_, &nb_, &piv, &cand_val, &cand_idx, &xrow, &counter, &chunk, &info};
        BS_CUDA(cudaLaunchCooperativeKernel((void *)k_lu_panel_coop, dim3(G), dim3(256), args, 0, c.stream));
      }
      tick(0);
      k_lu_swap<<<(n + 255) / 256, 256, 0, c.stream>>>(A, ld, n, j0, nb, piv);
      count_launch(c, 2);
      tick(1);
      const int right = n - j0 - nb;
      if (right > 0) {
        k_lu_u12_step<<<(right + 127) / 128, 128, sm_u12, c.stream>>>(A, ld, n, K0, Bw, j0, nb);
        count_launch(c);
      }
      tick(2);
      const int inside = K0 + Bw - (j0 + nb), below = n - (j0 + nb);
      if (inside > 0 && below > 0) lu_trailing_update(c, A, ld, n, j0, nb, j0 + nb, j0 + nb, K0 + Bw);  // rank-32 update of the rest of the outer block
      tick(3);
    }
    const int rem = n - K0 - Bw;
    if (rem > 0) lu_trailing_update(c, A, ld, n, K0, Bw, K0 + Bw, K0 + Bw, n);  // trailing matrix: one GEMM with k = Bw
    tick(4);
  }
  if (trace)
    fprintf(stderr, "[bs trace] lu_factor n=%d: panel %.1f ms, swaps %.1f ms, U rows %.1f ms, inner GEMM %.1f ms, trailing GEMM %.1f ms\n", n,
            1e3 * tacc[0], 1e3 * tacc[1], 1e3 * tacc[2], 1e3 * tacc[3], 1e3 * tacc[4]);
  BS_CUDA(cudaGetLastError());
  int h_info = 0;
  BS_CUDA(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  BS_CUDA(cudaStreamSynchronize(c.stream));
  if (h_info != 0)
    throw Error(BS_ERR_INVALID, "LU factorisation: the matrix is singular (zero pivot in column " + std::to_string(h_info - 1) + ")");
}

// x <- P x (sequential swaps, tiny kernel), then blocked forward / backward substitution
__global__ void k_apply_piv(double *x, const int *piv, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    for (int i = 0; i < n; ++i) {
      const int p = piv[i];
      if (p != i) {
        const double t = x[i];
        x[i] = x[p];
        x[p] = t;
      }
    }
}
// solve the diagonal block (unit lower or upper) for rows [k0,k0+nb) with one warp
__global__ void k_tri_diag(const double *LU, size_t ld, int k0, int nb, double *x, int upper) {
  __shared__ double xs[LU_NB];
  const int lane = threadIdx.x;
  if (lane < nb) xs[lane] = x[k0 + lane];
  __syncwarp();
  if (!upper) {
    for (int j = 0; j < nb; ++j) {
      const double xj = xs[j];
      if (lane > j && lane < nb) xs[lane] = fma(-LU[(size_t)(k0 + lane) * ld + k0 + j], xj, xs[lane]);
      __syncwarp();
    }
  } else {
    for (int j = nb - 1; j >= 0; --j) {
      if (lane == j) xs[j] = xs[j] / LU[(size_t)(k0 + j) * ld + k0 + j];
      __syncwarp();
      const double xj = xs[j];
      if (lane < j) xs[lane] = fma(-LU[(size_t)(k0 + lane) * ld + k0 + j], xj, xs[lane]);
      __syncwarp();
    }
  }
  if (lane < nb) x[k0 + lane] = xs[lane];
}
// x[r] -= sum_{j<nb} LU[r][k0+j] * x[k0+j] for r in [r_lo, r_hi): one warp per row
__global__ void k_tri_update(const double *LU, size_t ld, int k0, int nb, double *x, int r_lo, int r_hi) {
  const int r = r_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= r_hi) return;
  double s = (lane < nb) ? LU[(size_t)r * ld + k0 + lane] * x[k0 + lane] : 0.0;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) x[r] -= s;
}

void lu_solve(Context &c, const double *LU, size_t n_, size_t ld, const int *piv, double *x) {
  const int n = (int)n_;
  k_apply_piv<<<1, 32, 0, c.stream>>>(x, piv, n);
  count_launch(c);
  for (int k0 = 0; k0 < n; k0 += LU_NB) {
    const int nb = std::min(LU_NB, n - k0);
    k_tri_diag<<<1, 32, 0, c.stream>>>(LU, ld, k0, nb, x, 0);
    const int rem = n - k0 - nb;
    if (rem > 0) k_tri_update<<<(rem + 7) / 8, 256, 0, c.stream>>>(LU, ld, k0, nb, x, k0 + nb, n);
    count_launch(c, 2);
  }
  for (int k0 = ((n - 1) / LU_NB) * LU_NB; k0 >= 0; k0 -= LU_NB) {
    const int nb = std::min(LU_NB, n - k0);
    k_tri_diag<<<1, 32, 0, c.stream>>>(LU, ld, k0, nb, x, 1);
    if (k0 > 0) k_tri_update<<<(k0 + 7) / 8, 256, 0, c.stream>>>(LU, ld, k0, nb, x, 0, k0);
    count_launch(c, 2);
  }
  BS_CUDA(cudaGetLastError());
}


// ---------------------------------------------------------------------------------------------------------
// Preconditioner application  x = U^-1 L^-1 P b  as ONE cooperative kernel per diagonal block (the substitution
// above takes four launches per 32 columns).  The triangular systems are solved in steps of TB = 512 unknowns whose
// diagonal blocks are inverted once after the factorisation, so that BOTH halves of a step are matrix-vector products
// spread over the whole grid and the sweep is bandwidth bound instead of latency bound:
//     x_k = inv(T_kk) r_k          512 rows, one warp each
//     r_i -= T_ik x_k              all remaining rows, one warp per row, TRB rows in flight per warp
// with a grid barrier after each half (2 n / 512 barriers per sweep; 128-unknown steps solved by one owner CTA took
// 13 ms at 36 864 unknowns, dominated by the 578 dependent owner steps).  Any fixed linear operator is admissible as a
// preconditioner, so the explicit inverses of the diagonal blocks do not affect what GMRES converges to.
// ---------------------------------------------------------------------------------------------------------
constexpr int TB = 512;

// perm = the gather form of LAPACK's sequential row interchanges: (P b)[i] = b[perm[i]]
__global__ void k_piv_to_perm(const int *piv, int *perm, int n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int i = 0; i < n; ++i) {
    const int p = piv[i];
    if (p != i) {
      const int t = perm[i];
      perm[i] = perm[p];
      perm[p] = t;
    }
  }
}

// inverses of the TB x TB diagonal blocks, row-major: Linv[k][i][j] = inv(L_kk)(i, j) (unit lower), Uinv likewise (upper).
// One CTA per diagonal block, thread j builds column j by substitution directly in the output (its own column only).
__global__ void __launch_bounds__(TB) k_tri_inverse(const double *LU, size_t ld, int n, double *Linv, double *Uinv) {
  const int k0 = blockIdx.x * TB, kb = min(TB, n - k0), j = threadIdx.x;
  const double *T = LU + (size_t)k0 * ld + k0;
  double *X = Linv + (size_t)blockIdx.x * TB * TB;
  for (int i = 0; i < TB; ++i) X[(size_t)i * TB + j] = (i == j) ? 1.0 : 0.0;
  if (j < kb)
    for (int i = j + 1; i < kb; ++i) {
      double sacc = 0.0;
      for (int q = j; q < i; ++q) sacc = fma(T[(size_t)i * ld + q], X[(size_t)q * TB + j], sacc);
      X[(size_t)i * TB + j] = -sacc;
    }
  X = Uinv + (size_t)blockIdx.x * TB * TB;
  for (int i = 0; i < TB; ++i) X[(size_t)i * TB + j] = (i == j && j >= kb) ? 1.0 : 0.0;
  if (j < kb) {
    const double dj = T[(size_t)j * ld + j];
    X[(size_t)j * TB + j] = dj != 0.0 ? 1.0 / dj : 0.0;
    for (int i = j - 1; i >= 0; --i) {
      double sacc = 0.0;
      for (int q = i + 1; q <= j; ++q) sacc = fma(T[(size_t)i * ld + q], X[(size_t)q * TB + j], sacc);
      const double di = T[(size_t)i * ld + i];
      X[(size_t)i * TB + j] = di != 0.0 ? -sacc / di : 0.0;
    }
  }
}

// dot products of TRB matrix rows (row r0 + u * stride, columns [0, ncols) of a TB-wide block starting at `base`) with
// the vector v (shared memory); all loads of a batch are issued before the first reduction.  Lane 0 of the warp hands
// the TRB results to `sink(row, value)`.
constexpr int TRB = 4;
template <class Sink>
__device__ __forceinline__ void tri_rows_dot(const double *base, size_t ld, int r0, int stride, int r_end, int ncols,
                                             const double *__restrict__ v, int lane, Sink sink) {
  for (int rb = r0; rb < r_end; rb += TRB * stride) {
    double s[TRB];
#pragma unroll
    for (int u = 0; u < TRB; ++u) s[u] = 0.0;
#pragma unroll
    for (int h = 0; h < TB / 256; ++h) {  // 256 columns per pass: 8 doubles per lane and row
      const int c0 = 256 * h + 8 * lane;
      double2 a[TRB][4];
#pragma unroll
      for (int u = 0; u < TRB; ++u) {
        const int r = rb + u * stride;
        const double *row = base + (size_t)r * ld + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[u][q] = make_double2(0.0, 0.0);
          if (r < r_end) {
            if (c0 + 2 * q + 1 < ncols) a[u][q] = *reinterpret_cast<const double2 *>(row + 2 * q);
            else if (c0 + 2 * q < ncols) a[u][q].x = row[2 * q];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 vv = *reinterpret_cast<const double2 *>(v + c0 + 2 * q);
#pragma unroll
        for (int u = 0; u < TRB; ++u) s[u] = fma(a[u][q].x, vv.x, fma(a[u][q].y, vv.y, s[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < TRB; ++u) {
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], m);
    }
    double mine = 0.0;
#pragma unroll
    for (int u = 0; u < TRB; ++u)
      if (lane == u) mine = s[u];
    const int r = rb + lane * stride;
    if (lane < TRB && r < r_end) sink(r, mine);
  }
}

__global__ void __launch_bounds__(256) k_lu_apply_coop(const double *LU, size_t ld, int n, const int *perm, const double *Linv,
                                                       const double *Uinv, const double *in, double *x, double *xk_g /*[TB]*/,
                                                       unsigned int *counter, const int *skip) {
  if (skip && *skip) return;
  __shared__ __align__(16) double vs[TB];
  const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gw = blockIdx.x * 8 + wid, total_warps = G * 8;
  const int nblk = (n + TB - 1) / TB;
  unsigned int nbar = 0;
  for (int i = blockIdx.x * 256 + tid; i < n; i += G * 256) __stcg(x + i, in[perm[i]]);
  grid_barrier(counter, (++nbar) * G);
  for (int sweep = 0; sweep < 2; ++sweep) {  // 0: L y = P b forwards, 1: U x = y backwards
    const double *Minv = sweep == 0 ? Linv : Uinv;
    for (int kk = 0; kk < nblk; ++kk) {
      const int k = sweep == 0 ? kk : nblk - 1 - kk;
      const int k0 = k * TB, kb = min(TB, n - k0);
      // ---- x_k = inv(T_kk) r_k into the staging vector (r_k = x[k0 ..] is still being read by other CTAs)
      __syncthreads();
      for (int i = tid; i < TB; i += 256) vs[i] = i < kb ? __ldcg(x + k0 + i) : 0.0;
      __syncthreads();
      tri_rows_dot(Minv + (size_t)k * TB * TB, TB, gw, total_warps, kb, TB, vs, lane, [&](int r, double v) { __stcg(xk_g + r, v); });
      grid_barrier(counter, (++nbar) * G);
      // ---- r_i -= T_ik x_k on the rows still to be solved; CTA 0 moves x_k to its place
      for (int i = tid; i < TB; i += 256) vs[i] = i < kb ? __ldcg(xk_g + i) : 0.0;
      __syncthreads();
      if (blockIdx.x == 0)
        for (int i = tid; i < kb; i += 256) __stcg(x + k0 + i, vs[i]);
      const int r_lo = sweep == 0 ? k0 + TB : 0, r_hi = sweep == 0 ? n : k0;
      if (r_lo < r_hi)
        tri_rows_dot(LU + k0, ld, r_lo + gw, total_warps, r_hi, kb, vs, lane, [&](int r, double v) { __stcg(x + r, __ldcg(x + r) - v); });
      grid_barrier(counter, (++nbar) * G);
    }
  }
}

void lu_apply_fast(Context &c, const Context::LuBlock &b, const double *in, double *out, const int *skip) {
  static int max_per_sm = 0;
  if (!max_per_sm) {
    BS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_per_sm, k_lu_apply_coop, 256, 0));
    max_per_sm = std::max(1, std::min(max_per_sm, 2));
  }
  int G = (int)std::min<size_t>((size_t)max_per_sm * c.sm_count, std::max<size_t>(1, (b.n + 63) / 64));
  unsigned int *counter = reinterpret_cast<unsigned int *>(c.wsi("lu.apply_counter", 4));
  BS_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), c.stream));
  double *xk = c.wsd("lu.apply_xk", TB + 2);
  const double *LU = b.LU;
  size_t ld = b.ld;
  int n = (int)b.n;
  const int *perm = b.perm;
  const double *Li = b.LinvT, *Ui = b.UinvT;
  void *args[] = {&LU, &ld, &n, &perm, &Li, &Ui, &in, &out, &xk, &counter, &skip};
  BS_CUDA(cudaLaunchCooperativeKernel((void *)k_lu_apply_coop, dim3(G), dim3(256), args, 0, c.stream));
  count_launch(c);
}

// band filter of assemble_monolithic_preconditioner (ref: bem_stokes.cc:3437-3475): entries whose reference column index
// lies outside [ri - band, ri + band) are dropped; ref_of maps local block index -> reference index
__global__ void k_band_filter(double *LU, size_t ld, int n, const int *ref_of, int band) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || j >= n) return;
  const long long ri = ref_of[i], rj = ref_of[j];
  const long long lo = ri > band ? ri - band : 0, hi = ri + band;
  if (!(rj >= lo && rj < hi)) LU[(size_t)i * ld + j] = 0.0;
}

void precond_factor_blocks(Context &c, const DMat &M, size_t col_off, size_t n_total, size_t max_block, int band) {
  const size_t nblocks = (max_block == 0 || max_block >= n_total) ? 1 : (n_total + max_block - 1) / max_block;
  c.lu_blocks.assign(nblocks, Context::LuBlock());
  size_t lu_total = 0, inv_total = 0, piv_total = 0;
  for (size_t b = 0; b < nblocks; ++b) {
    Context::LuBlock &B = c.lu_blocks[b];
    B.off = n_total * b / nblocks;
    B.n = n_total * (b + 1) / nblocks - B.off;
    B.ld = (B.n + 15) & ~(size_t)15;
    lu_total += B.n * B.ld;
    inv_total += 2 * ((B.n + TB - 1) / TB) * TB * TB;
    piv_total += 2 * ((B.n + 3) & ~(size_t)3);
  }
  c.d_lu.alloc(lu_total + 2);
  c.d_luinv.alloc(inv_total + 2);
  c.d_piv.alloc(piv_total + 2);
  size_t lu_at = 0, inv_at = 0, piv_at = 0;
  for (size_t b = 0; b < nblocks; ++b) {
    Context::LuBlock &B = c.lu_blocks[b];
    const size_t nt = (B.n + TB - 1) / TB, np = (B.n + 3) & ~(size_t)3;
    B.LU = c.d_lu.p + lu_at, lu_at += B.n * B.ld;
    B.LinvT = c.d_luinv.p + inv_at, B.UinvT = B.LinvT + nt * TB * TB, inv_at += 2 * nt * TB * TB;
    B.piv = c.d_piv.p + piv_at, B.perm = B.piv + np, piv_at += 2 * np;
    BS_CUDA(cudaMemsetAsync(B.LU, 0, B.n * B.ld * sizeof(double), c.stream));
    BS_CUDA(cudaMemcpy2DAsync(B.LU, B.ld * sizeof(double), M.p + B.off * M.ld + col_off + B.off, M.ld * sizeof(double),
                              B.n * sizeof(double), B.n, cudaMemcpyDeviceToDevice, c.stream));
    add_rank1_block(c, M, B.off, col_off + B.off, B.n, B.LU, B.ld);  // implicit V correction of M, if any
    if (band > 0) {
      std::vector<int> ref(B.n);
      for (size_t i = 0; i < B.n; ++i) {
        const size_t gi = col_off + B.off + i;
        ref[i] = gi < c.n3() ? (int)((size_t)c.node_of_pos[gi / 3] + (gi % 3) * c.N) : (int)gi;
      }
      int *d_ref = c.wsi("lu.band_ref", B.n + 2);
      BS_CUDA(cudaMemcpyAsync(d_ref, ref.data(), B.n * sizeof(int), cudaMemcpyHostToDevice, c.stream));
      k_band_filter<<<dim3((unsigned)((B.n + 255) / 256), (unsigned)B.n), 256, 0, c.stream>>>(B.LU, B.ld, (int)B.n, d_ref, band);
      BS_CUDA(cudaStreamSynchronize(c.stream));  // `ref` is a host temporary
      count_launch(c);
    }
    lu_factor(c, B.LU, B.n, B.ld, B.piv);
    k_piv_to_perm<<<1, 32, 0, c.stream>>>(B.piv, B.perm, (int)B.n);
    k_tri_inverse<<<(unsigned)nt, TB, 0, c.stream>>>(B.LU, B.ld, (int)B.n, B.LinvT, B.UinvT);
    BS_CUDA(cudaGetLastError());
    count_launch(c, 2);
  }
}

// ---------------------------------------------------------------------------------------------------------
// preconditioner application on the local slice
// ---------------------------------------------------------------------------------------------------------
void apply_precond(Context &c, const double *in_loc, double *out_loc, const int *skip) {
  const size_t mloc = c.local_vec_len(c.prec_which);
  switch (c.prec_kind) {
    case BS_PREC_NONE:
      copy(c, in_loc, out_loc, mloc);
      break;
    case BS_PREC_JACOBI:
      mul_elem(c, in_loc, c.d_prec_diag.p, out_loc, mloc);
      break;
    case BS_PREC_DIRECT:
    case BS_PREC_BLOCK_DIRECT:
    case BS_PREC_BAND: {
      size_t covered = 0;
      for (const Context::LuBlock &b : c.lu_blocks) {
        BS_REQUIRE(b.off + b.n <= mloc, "preconditioner was factorised for a larger system: call bs_precond_setup again");
        if (std::getenv("BS_LU_SUBSTITUTION")) {  // exact substitution (4 launches per 32 columns), for comparison
          copy(c, in_loc + b.off, out_loc + b.off, b.n);
          lu_solve(c, b.LU, b.n, b.ld, b.piv, out_loc + b.off);
        } else {
          lu_apply_fast(c, b, in_loc + b.off, out_loc + b.off, skip);
        }
        covered = b.off + b.n;
      }
      if (covered < mloc) copy(c, in_loc + covered, out_loc + covered, mloc - covered);  // rigid unknowns pass through
      break;
    }
    default:
      throw Error(BS_ERR_INVALID, "unknown preconditioner kind");
  }
}

// ---------------------------------------------------------------------------------------------------------
// GMRES
// ---------------------------------------------------------------------------------------------------------
__global__ void k_scale_inv_norm(const double *s2, const double *src, double *dst, size_t n) {
  const double a = rsqrt(*s2);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i] * a;
}

// Orthogonalise z against basis[0..dim) and leave  coef[0..dim) (+ coef[m+2 .. m+2+dim) second-pass part) and
// coef[dim] = ||z||^2 on the device.  No host synchronisation in CGS2 mode.
static void orthogonalize(Context &c, const double *basis, size_t mloc, int dim, double *z, double *coef, int m, int its,
                          double *nrm_scratch) {
  if (c.gmres_ortho == BS_ORTHO_CGS2) {
    // classical Gram-Schmidt applied twice (one fused multi-dot + one fused multi-axpy per pass, one reduction
    // each): orthogonal to machine precision, so the re-orthogonalisation safeguard of the reference's modified
    // Gram-Schmidt is built in
    multi_dot(c, basis, mloc, dim, z, mloc, coef);
    allreduce(c, coef, dim);
    multi_axpy(c, basis, mloc, dim, coef, -1.0, z, mloc);
    multi_dot(c, basis, mloc, dim, z, mloc, coef + (m + 2));
    allreduce(c, coef + (m + 2), dim);
    multi_axpy(c, basis, mloc, dim, coef + (m + 2), -1.0, z, mloc);
  } else {
    // deal.II SolverGMRES::modified_gram_schmidt verbatim: sequential projections, and at every 5th iteration a
    // second pass if the vector lost more than 10*sqrt(eps) of its norm
    auto norm2 = [&](const double *v) {
      multi_dot(c, v, 0, 1, v, mloc, nrm_scratch);
      allreduce(c, nrm_scratch, 1);
      double s2;
      BS_CUDA(cudaMemcpyAsync(&s2, nrm_scratch, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      BS_CUDA(cudaStreamSynchronize(c.stream));
      return s2;
    };
    const bool check = (its % 5 == 0);
    double norm_start2 = 0.0;
    if (check) norm_start2 = norm2(z);
    BS_CUDA(cudaMemsetAsync(coef + (m + 2), 0, sizeof(double) * (m + 2), c.stream));
    for (int i = 0; i < dim; ++i) {
      multi_dot(c, basis + (size_t)i * mloc, mloc, 1, z, mloc, coef + (m + 2) + i);
      allreduce(c, coef + (m + 2) + i, 1);
      multi_axpy(c, basis + (size_t)i * mloc, mloc, 1, coef + (m + 2) + i, -1.0, z, mloc);
    }
    BS_CUDA(cudaMemcpyAsync(coef, coef + (m + 2), sizeof(double) * dim, cudaMemcpyDeviceToDevice, c.stream));
    BS_CUDA(cudaMemsetAsync(coef + (m + 2), 0, sizeof(double) * (m + 2), c.stream));
    bool reorth = false;
    if (check) {
      const double nv2 = norm2(z);
      reorth = !(std::sqrt(nv2) > 10.0 * std::sqrt(norm_start2) * std::sqrt(2.220446049250313e-16));
    }
    if (reorth)
      for (int i = 0; i < dim; ++i) {
        multi_dot(c, basis + (size_t)i * mloc, mloc, 1, z, mloc, coef + (m + 2) + i);
        allreduce(c, coef + (m + 2) + i, 1);
        multi_axpy(c, basis + (size_t)i * mloc, mloc, 1, coef + (m + 2) + i, -1.0, z, mloc);
      }
  }
  multi_dot(c, z, 0, 1, z, mloc, coef + dim);
  allreduce(c, coef + dim, 1);
}

// nrhs independent GMRES recurrences advanced in lockstep: ONE pass over the matrix per iteration serves all
// right-hand sides (multi-RHS GEMV), everything else is per system.  deal.II SolverGMRES semantics per system
// (left preconditioning, Givens, absolute tolerance on the preconditioned residual, restart max_tmp-2).
int gmres_batched(Context &c, int which, int nrhs, const double *d_B, double *d_X, size_t ldv, double tol, int max_steps,
                  int max_tmp, int *iters, double *final_res) {
  BS_REQUIRE(max_tmp >= 3, "max_n_tmp_vectors must be >= 3");
  BS_REQUIRE(nrhs >= 1, "nrhs must be positive");
  // the identity "preconditioner" has no matrix of its own: it acts on the system being solved (after a solve with another
  // matrix - the DN route solves with V - the vector length of apply_precond would otherwise be that matrix's)
  if (c.prec_kind == BS_PREC_NONE) c.prec_which = which;
  // default: the device-resident iteration (bs_gmres.cu).  This host-driven loop remains for deal.II's modified
  // Gram-Schmidt verbatim (host decisions every 5th iteration) and for callback communicators.
  if (gmres_device_eligible(c, nrhs, max_tmp))
    return gmres_device(c, which, nrhs, d_B, d_X, ldv, tol, max_steps, max_tmp, iters, final_res);
  const int m = max_tmp - 2;  // restart length (deal.II: n_tmp_vectors - 2 inner iterations)
  const size_t mloc = c.local_vec_len(which), mfull = c.full_vec_len(which);
  const size_t ldw = (mloc + 3) & ~(size_t)1, ldx = (mfull + 4) & ~(size_t)3;
  const size_t CS = (size_t)2 * (m + 2) + 2;
  const DMat &M = mat_of(c, which);
  struct P_ { double *p; } basis, w, z, coef, xfull;  // context-owned, grow-only (no cudaMalloc per solve)
  basis.p = c.wsd("gmres.basis", (size_t)nrhs * (m + 1) * mloc + 2);
  w.p = c.wsd("gmres.w", (size_t)nrhs * ldw);
  z.p = c.wsd("gmres.z", (size_t)nrhs * ldw);
  coef.p = c.wsd("gmres.coef", (size_t)nrhs * CS);
  xfull.p = c.wsd("gmres.xfull", (size_t)nrhs * ldx);
  struct Sys {
    std::vector<double> H, gamma, ci, si, h, yk;
    int its = 0, dim = 0;
    double rho = 0;
    bool active = true, converged = false;
  };
  std::vector<Sys> S(nrhs);
  for (auto &sy : S) {
    sy.H.assign((size_t)(m + 1) * m, 0.0);
    sy.gamma.assign(m + 1, 0.0);
    sy.ci.assign(m, 0.0);
    sy.si.assign(m, 0.0);
    sy.h.assign(m + 2, 0.0);
    sy.yk.assign(m, 0.0);
  }
  std::vector<double> hbuf((size_t)nrhs * CS);
  auto Bs = [&](int s) { return basis.p + (size_t)s * (m + 1) * mloc; };
  auto Cf = [&](int s) { return coef.p + (size_t)s * CS; };
  const bool p2p = c.p2p && c.nranks > 1;
  BS_REQUIRE(!p2p || nrhs <= Context::XCHG_SLOTS, "peer exchange supports at most 8 right-hand sides in lockstep");
  double *const xbuf = p2p ? c.d_xchg.p : xfull.p;
  const size_t xld = p2p ? c.xchg_ld : ldx;
  // `published`: the sources already sit in the replicated buffers of all ranks (stored there by the kernel that
  // produced them); otherwise gather them now
  auto matvec_all = [&](const std::vector<int> &act, const std::vector<const double *> &src, bool published) {
    if (p2p) {
      if (!published)
        for (size_t k = 0; k < act.size(); ++k) p2p_scatter(c, which, src[k], (int)k, nullptr, nullptr);
      p2p_wait(c);
    } else {
      for (size_t k = 0; k < act.size(); ++k) exchange(c, which, src[k], xbuf + k * xld);
    }
    if (act.size() == 1) gemv(c, M, xbuf, w.p);
    else gemv_multi(c, M, (int)act.size(), xbuf, xld, w.p, ldw);
  };
  // normalise z by 1/sqrt(*s2) into the local basis vector and (peer mode) into slot `slot` of every rank
  auto normalize_publish = [&](const double *s2, const double *zsrc, double *basis_dst, int slot) {
    if (p2p) {
      p2p_scatter(c, which, zsrc, slot, s2, basis_dst);
    } else {
      k_scale_inv_norm<<<(unsigned)std::min<size_t>((mloc + 255) / 256, 1184), 256, 0, c.stream>>>(s2, zsrc, basis_dst, mloc);
      count_launch(c);
    }
  };
  auto finish = [&](int s) {  // back substitution H y = gamma, x += sum y_i v_i
    Sys &sy = S[s];
    const int dim = sy.dim;
    for (int i = dim - 1; i >= 0; --i) {
      double t = sy.gamma[i];
      for (int j = i + 1; j < dim; ++j) t -= sy.H[(size_t)i * m + j] * sy.yk[j];
      sy.yk[i] = t / sy.H[(size_t)i * m + i];
    }
    if (dim > 0) {
      BS_CUDA(cudaMemcpyAsync(Cf(s), sy.yk.data(), sizeof(double) * dim, cudaMemcpyHostToDevice, c.stream));
      multi_axpy(c, Bs(s), mloc, dim, Cf(s), 1.0, d_X + (size_t)s * ldv, mloc);
      BS_CUDA(cudaStreamSynchronize(c.stream));
    }
    sy.dim = 0;
  };
  const unsigned sgrid = (unsigned)std::min<size_t>((mloc + 255) / 256, 1184);
  const bool trace = std::getenv("BS_TRACE") != nullptr;
  double t_exch = 0, t_mv = 0, t_orth = 0, t_sync = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  auto tsync = [&](double &acc) {  // trace mode: attribute device time to the phase that queued it
    if (!trace) return;
    const double t0 = now();
    cudaStreamSynchronize(c.stream);
    acc += now() - t0;
  };

  while (true) {
    std::vector<int> act;
    for (int s = 0; s < nrhs; ++s)
      if (S[s].active) act.push_back(s);
    if (act.empty()) break;
    // ---- r0 = M^{-1} (b - A x) for every active system
    std::vector<const double *> src;
    for (int s : act) src.push_back(d_X + (size_t)s * ldv);
    matvec_all(act, src, false);
    for (size_t k = 0; k < act.size(); ++k) {
      const int s = act[k];
      sub(c, d_B + (size_t)s * ldv, w.p + k * ldw, w.p + k * ldw, mloc);
      apply_precond(c, w.p + k * ldw, z.p + k * ldw);
      multi_dot(c, z.p + k * ldw, 0, 1, z.p + k * ldw, mloc, Cf(s) + CS - 1);
      allreduce(c, Cf(s) + CS - 1, 1);
    }
    BS_CUDA(cudaMemcpyAsync(hbuf.data(), coef.p, sizeof(double) * nrhs * CS, cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    for (size_t k = 0; k < act.size(); ++k) {
      const int s = act[k];
      Sys &sy = S[s];
      sy.rho = std::sqrt(hbuf[(size_t)s * CS + CS - 1]);
      if (sy.rho <= tol) {
        sy.converged = true;
        sy.active = false;
        continue;
      }
      if (sy.its >= max_steps) {
        sy.active = false;
        continue;
      }
      std::fill(sy.gamma.begin(), sy.gamma.end(), 0.0);
      sy.gamma[0] = sy.rho;
      sy.dim = 0;
    }
    {  // first basis vectors, published into the slots of the systems that stay active
      int slot = 0;
      for (size_t k = 0; k < act.size(); ++k)
        if (S[act[k]].active) normalize_publish(Cf(act[k]) + CS - 1, z.p + k * ldw, Bs(act[k]), slot++);
    }
    // ---- Arnoldi, all still-active systems share the inner index
    for (int inner = 0; inner < m; ++inner) {
      act.clear();
      for (int s = 0; s < nrhs; ++s)
        if (S[s].active) act.push_back(s);
      if (act.empty()) break;
      src.clear();
      for (int s : act) src.push_back(Bs(s) + (size_t)inner * mloc);
      double tt0 = trace ? now() : 0;
      if (trace) {
        if (p2p) p2p_wait(c);
        else
          for (size_t k = 0; k < act.size(); ++k) exchange(c, which, src[k], xbuf + k * xld);
        cudaStreamSynchronize(c.stream);
        t_exch += now() - tt0;
        tt0 = now();
        if (act.size() == 1) gemv(c, M, xbuf, w.p);
        else gemv_multi(c, M, (int)act.size(), xbuf, xld, w.p, ldw);
        cudaStreamSynchronize(c.stream);
        t_mv += now() - tt0;
        tt0 = now();
      } else {
        matvec_all(act, src, true);
      }
      const int dim = inner + 1;
      for (size_t k = 0; k < act.size(); ++k) {
        const int s = act[k];
        ++S[s].its;
        apply_precond(c, w.p + k * ldw, z.p + k * ldw);
        orthogonalize(c, Bs(s), mloc, dim, z.p + k * ldw, Cf(s), m, S[s].its, Cf(s) + CS - 2);
      }
      if (trace) {
        cudaStreamSynchronize(c.stream);
        t_orth += now() - tt0;
        tt0 = now();
      }
      BS_CUDA(cudaMemcpyAsync(hbuf.data(), coef.p, sizeof(double) * nrhs * CS, cudaMemcpyDeviceToHost, c.stream));
      BS_CUDA(cudaStreamSynchronize(c.stream));
      if (trace) t_sync += now() - tt0;
      for (size_t k = 0; k < act.size(); ++k) {
        const int s = act[k];
        Sys &sy = S[s];
        const double *hb = &hbuf[(size_t)s * CS];
        std::vector<double> &h = sy.h;
        for (int i = 0; i < dim; ++i) h[i] = hb[i] + hb[m + 2 + i];
        h[dim] = std::sqrt(hb[dim]);
        // Givens rotations (deal.II SolverGMRES::givens_rotation)
        for (int i = 0; i < inner; ++i) {
          const double t = h[i];
          h[i] = sy.ci[i] * t + sy.si[i] * h[i + 1];
          h[i + 1] = -sy.si[i] * t + sy.ci[i] * h[i + 1];
        }
        const double r = std::hypot(h[inner], h[inner + 1]);
        sy.ci[inner] = h[inner] / r;
        sy.si[inner] = h[inner + 1] / r;
        h[inner] = r;
        sy.gamma[inner + 1] = -sy.si[inner] * sy.gamma[inner];
        sy.gamma[inner] = sy.ci[inner] * sy.gamma[inner];
        for (int i = 0; i < dim; ++i) sy.H[(size_t)i * m + inner] = h[i];
        sy.dim = dim;
        sy.rho = std::fabs(sy.gamma[dim]);
        if (sy.rho <= tol) {
          sy.converged = true;
          sy.active = false;
        } else if (sy.its >= max_steps) {
          sy.active = false;
        }
      }
      {  // next basis vectors of the systems that continue: normalise and publish (fused with the peer exchange)
        int slot = 0;
        for (size_t k = 0; k < act.size(); ++k) {
          const int s = act[k];
          if (S[s].active && inner + 1 < m)
            normalize_publish(Cf(s) + dim, z.p + k * ldw, Bs(s) + (size_t)(inner + 1) * mloc, slot++);
        }
      }
      // finalise the systems that just stopped (stream order keeps the kernels above ahead of the coefficient upload)
      for (int s : act)
        if (!S[s].active) finish(s);
    }
    // restart: systems that are still active fold their Krylov correction into x and start over
    for (int s = 0; s < nrhs; ++s)
      if (S[s].active) finish(s);
  }
  (void)tsync;
  (void)sgrid;
  if (p2p) {  // a rank that timed out in the flag wait reports instead of hanging
    unsigned long long err = 0;
    BS_CUDA(cudaMemcpyAsync(&err, c.d_flags.p + BS_FLAG_ERR, sizeof(err), cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    if (err != 0) throw Error(BS_ERR_COMM, "peer exchange timed out waiting for epoch " + std::to_string(err));
  }
  if (trace)
    fprintf(stderr, "[bs trace rank %d] gmres: exchange %.1f ms, matvec %.1f ms, precond+orthogonalisation %.1f ms, d2h+sync %.1f ms (%d its)\n",
            c.rank, 1e3 * t_exch, 1e3 * t_mv, 1e3 * t_orth, 1e3 * t_sync, S[0].its);
  int rc = BS_OK;
  for (int s = 0; s < nrhs; ++s) {
    if (iters) iters[s] = S[s].its;
    if (final_res) final_res[s] = S[s].rho;
    if (!S[s].converged) rc = BS_ERR_NOT_CONVERGED;
  }
  return rc;
}

int gmres(Context &c, int which, const double *d_b, double *d_x, double tol, int max_steps, int max_tmp, int *iters,
          double *final_res) {
  return gmres_batched(c, which, 1, d_b, d_x, 0, tol, max_steps, max_tmp, iters, final_res);
}

}  // namespace bs
