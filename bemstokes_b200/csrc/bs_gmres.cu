// Device-resident restarted GMRES (deal.II SolverGMRES semantics: left preconditioning, Givens rotations, absolute
// tolerance on the preconditioned residual, restart length max_n_tmp_vectors - 2; ref call sites
// source/bem_stokes.cc:4116, 4332, semantics SURVEY A.7).
//
// Nothing in the iteration returns to the host: the Hessenberg column, the Givens rotations, the residual estimate
// and the convergence decision live in device memory (GmStatus), the kernels of an iteration read the inner index
// from there (so every iteration is the same launch sequence), and kernels queued after convergence return at
// once.  The host only queues iterations and looks at a pinned copy of the status every few iterations, one group
// behind the device.  Per iteration and right-hand side batch:
//     [wait for the peers' slices]  ->  matvec  ->  [preconditioner]
//     ->  k_gm_pass<DOTS>       h1 = V^T z                      (+ cross-rank sum)
//     ->  k_gm_pass<UPD_DOTS>   z -= V h1 ; h2 = V^T z          (+ cross-rank sum)
//     ->  k_gm_pass<UPD_NORM>   z -= V h2 ; |z|^2               (+ cross-rank sum, Givens step, convergence test)
//     ->  k_gm_publish          v_{j+1} = z / |z| into the basis and into every rank's replicated vector buffer
// (classical Gram-Schmidt twice = the re-orthogonalised Gram-Schmidt of the reference in exact arithmetic).
// The cross-rank sums need no collective call: the last CTA of a pass stores this rank's partial sums into every
// peer's reduction buffer over NVLink (CUDA-IPC mapped), raises a per-source flag with a release store, acquires
// the flags of all sources and adds the contributions in rank order, so that every rank holds bit-identical
// coefficients and takes identical decisions.
#include "bs_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace bs {

constexpr int GM_MAXSYS = Context::XCHG_SLOTS;
constexpr int GM_THREADS = 256;
constexpr int GM_CHUNK = 2 * GM_THREADS;  // vector elements per CTA of a pass (one double2 per thread)
constexpr int GM_KT = 64;                 // basis vectors per shared-memory reduction tile

struct GmStatus {
  int j;         // inner index of the running cycle = number of Hessenberg columns finished
  int all_done;  // every system has stopped: kernels queued behind this point return at once
  unsigned int ticket;
  int pad;
  int done[GM_MAXSYS];  // 0 running, 1 converged, 2 max_steps reached, 3 exchange failure
  int its[GM_MAXSYS];
  int dim[GM_MAXSYS];   // columns of the running cycle that belong to system s (frozen when it stops)
  double rho[GM_MAXSYS];
  double inv_norm[GM_MAXSYS];
};

struct GmDev {  // kernel parameter block
  GmStatus *st;
  double *basis;  // [nrhs][(m+1)][ldb]
  size_t ldb, bstride;
  double *z;  // [nrhs][ldw] vector being orthogonalised
  size_t ldw;
  double *H, *gamma, *ci, *si, *c1, *c2, *yk;  // per system: (m+1)*m, m+1, m, m, m+2, m+2, m
  int m, nrhs;
  size_t mloc;
  double *partial;  // [(s*kld + k)*nchunks + chunk]
  int nchunks, kld;
  double tol;
  int max_steps;
  int nranks, rank;
  double *const *peer_x;                  // replicated vector buffers of all ranks (nranks > 1)
  unsigned long long *const *peer_flags;  // flag words of all ranks
  unsigned long long *flags;              // own flag words
  size_t red_off, red_cap;                // reduction area inside every rank's exchange buffer
  double *red_local;                      // nranks == 1: local scratch of red_cap doubles
  double *xloc;                           // this rank's replicated vector buffer
  size_t xld, slice_off;
};

enum { GM_DOTS = 0, GM_UPD_DOTS = 1, GM_UPD_NORM = 2, GM_NORM_BEGIN = 3 };

__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_sys(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= target; false after 20 s (a peer died: report instead of hanging the GPU)
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long target) {
  const unsigned long long t0 = globaltimer();
  while (ld_acquire_sys(flag) < target)
    if (globaltimer() - t0 > 20000000000ull) return false;
  return true;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// peer exchange of the Krylov slices with a device-resident epoch (shared with the host-driven solver)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_p2p_signal_dev(unsigned long long *flags, unsigned long long *const *peer_flags, int nranks, int myrank) {
  __shared__ unsigned long long e;
  if (threadIdx.x == 0) e = ++flags[BS_FLAG_EPOCH];
  __syncthreads();
  __threadfence_system();
  if ((int)threadIdx.x < nranks) st_release_sys(peer_flags[threadIdx.x] + BS_FLAG_XCHG + myrank, e);
}
__global__ void k_p2p_wait_dev(unsigned long long *flags, int nranks, const int *skip, GmStatus *st) {
  if (skip && *skip) return;
  const int r = threadIdx.x;
  if (r >= nranks) return;
  const unsigned long long e = flags[BS_FLAG_EPOCH];
  if (!wait_flag(flags + BS_FLAG_XCHG + r, e)) {
    flags[BS_FLAG_ERR] = e;
    if (st) {
      st->all_done = 1;
      for (int s = 0; s < GM_MAXSYS; ++s)
        if (st->done[s] == 0) st->done[s] = 3;
    }
  }
}
void p2p_signal(Context &c) {
  k_p2p_signal_dev<<<1, 32, 0, c.stream>>>(c.d_flags.p, c.d_peer_flags.p, c.nranks, c.rank);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}
void p2p_wait_only(Context &c, const int *skip, void *st) {
  k_p2p_wait_dev<<<1, 32, 0, c.stream>>>(c.d_flags.p, c.nranks, skip, (GmStatus *)st);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

// ---------------------------------------------------------------------------------------------------------
// cross-rank sum of `total` values held as per-chunk partial sums; executed by the last CTA of a pass.
// out(idx) receives the sum over chunks and ranks (fixed order: bit-identical on every rank).
// ---------------------------------------------------------------------------------------------------------
template <class Out>
__device__ void gm_reduce(const GmDev &P, int cnt, Out out) {
  GmStatus *st = P.st;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int total = P.nrhs * cnt;
  __shared__ unsigned long long s_seq;
  __shared__ int s_fail;
  if (tid == 0) {
    s_seq = P.nranks > 1 ? ++P.flags[BS_FLAG_SEQ] : 0;
    s_fail = 0;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const size_t par = (size_t)(seq & 1);
  for (int idx = wid; idx < total; idx += GM_THREADS / 32) {
    const int s = idx / cnt, k = idx - s * cnt;
    const double *pp = P.partial + ((size_t)s * P.kld + k) * P.nchunks;
    double v = 0.0;
    for (int ch = lane; ch < P.nchunks; ch += 32) v += ld_cg(pp + ch);
    v = warp_sum(v);
    if (st->done[s]) v = 0.0;  // stopped systems leave stale partial sums behind
    if (lane == 0) {
      if (P.nranks == 1) P.red_local[idx] = v;
      else
        for (int r = 0; r < P.nranks; ++r) P.peer_x[r][P.red_off + (par * P.nranks + P.rank) * P.red_cap + idx] = v;
    }
  }
  if (P.nranks > 1) {
    __threadfence_system();
    __syncthreads();
    if (tid < P.nranks) {
      st_release_sys(P.peer_flags[tid] + BS_FLAG_RED + P.rank, seq);
      if (!wait_flag(P.flags + BS_FLAG_RED + tid, seq)) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) {
      if (tid == 0) {
        P.flags[BS_FLAG_ERR] = seq;
        st->all_done = 1;
        for (int s = 0; s < GM_MAXSYS; ++s)
          if (st->done[s] == 0) st->done[s] = 3;
      }
      __syncthreads();
    }
    const double *mine = P.xloc + P.red_off + par * P.nranks * P.red_cap;
    for (int idx = tid; idx < total; idx += GM_THREADS) {
      double v = 0.0;
      for (int r = 0; r < P.nranks; ++r) v += ld_sys(mine + (size_t)r * P.red_cap + idx);
      out(idx, v);
    }
  } else {
    __threadfence_block();
    __syncthreads();
    for (int idx = tid; idx < total; idx += GM_THREADS) out(idx, ((volatile double *)P.red_local)[idx]);
  }
  __syncthreads();
}

// one Gram-Schmidt pass over this rank's slice of z (see the file header); grid (nchunks, nrhs)
template <int MODE>
__global__ void __launch_bounds__(GM_THREADS) k_gm_pass(GmDev P) {
  GmStatus *st = P.st;
  if (st->all_done) return;
  extern __shared__ double dyn[];  // Givens scratch of the last CTA: [nrhs][3][m+2]
  __shared__ double red[GM_THREADS / 32][GM_KT];
  __shared__ double s_nrm[GM_MAXSYS];
  __shared__ int s_last;
  const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = st->j;
  const int nb = (MODE == GM_NORM_BEGIN) ? 0 : j + 1;
  if (st->done[s] == 0) {
    const size_t i0 = (size_t)blockIdx.x * GM_CHUNK + 2 * (size_t)tid;
    const bool v0 = i0 < P.mloc, v1 = i0 + 1 < P.mloc;
    double *zp = P.z + (size_t)s * P.ldw + i0;
    double2 zz = make_double2(0.0, 0.0);
    if (v0) {
      zz = *reinterpret_cast<const double2 *>(zp);
      if (!v1) zz.y = 0.0;
    }
    const double *B = P.basis + (size_t)s * P.bstride + i0;
    if (MODE == GM_UPD_DOTS || MODE == GM_UPD_NORM) {
      const double *cf = (MODE == GM_UPD_DOTS ? P.c1 : P.c2) + (size_t)s * (P.m + 2);
      if (v0) {
        // eight basis vectors per step: all loads are issued before the first FMA needs one (memory-level parallelism;
        // the k loop is a dependent FMA chain per element otherwise paced by one load latency per vector)
        int k = 0;
        for (; k + 8 <= nb; k += 8) {
          double2 b[8];
          double cc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            b[u] = *reinterpret_cast<const double2 *>(B + (size_t)(k + u) * P.ldb);
            cc[u] = cf[k + u];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            zz.x = fma(-cc[u], b[u].x, zz.x);
            zz.y = fma(-cc[u], b[u].y, zz.y);
          }
        }
        for (; k < nb; ++k) {
          const double2 b = *reinterpret_cast<const double2 *>(B + (size_t)k * P.ldb);
          const double cc = cf[k];
          zz.x = fma(-cc, b.x, zz.x);
          zz.y = fma(-cc, b.y, zz.y);
        }
        if (!v1) zz.y = 0.0;
        *reinterpret_cast<double2 *>(zp) = zz;
      }
    }
    if (MODE == GM_DOTS || MODE == GM_UPD_DOTS) {
      for (int k0 = 0; k0 < nb; k0 += GM_KT) {
        const int kt = min(GM_KT, nb - k0);
        for (int kk = 0; kk < kt; kk += 8) {  // eight loads in flight, then eight interleaved warp reductions
          double p[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            p[u] = 0.0;
            if (v0 && kk + u < kt) {
              const double2 b = *reinterpret_cast<const double2 *>(B + (size_t)(k0 + kk + u) * P.ldb);
              p[u] = v1 ? fma(b.y, zz.y, b.x * zz.x) : b.x * zz.x;
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) p[u] = warp_sum(p[u]);
          if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (kk + u < kt) red[wid][kk + u] = p[u];
          }
        }
        __syncthreads();
        if (tid < kt) {
          double v = 0.0;
#pragma unroll
          for (int w = 0; w < GM_THREADS / 32; ++w) v += red[w][tid];
          P.partial[((size_t)s * P.kld + k0 + tid) * P.nchunks + blockIdx.x] = v;
        }
        __syncthreads();
      }
    } else {
      double p = warp_sum(fma(zz.x, zz.x, zz.y * zz.y));
      if (lane == 0) red[wid][0] = p;
      __syncthreads();
      if (tid == 0) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < GM_THREADS / 32; ++w) v += red[w][0];
        P.partial[((size_t)s * P.kld) * P.nchunks + blockIdx.x] = v;
      }
    }
  }
  // ---- last CTA of the grid: sums over chunks and ranks, then the scalar part of the iteration
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) st->ticket = 0;
  if (MODE == GM_DOTS) {
    gm_reduce(P, nb, [&](int idx, double v) { P.c1[(size_t)(idx / nb) * (P.m + 2) + idx % nb] = v; });
    return;
  }
  if (MODE == GM_UPD_DOTS) {
    gm_reduce(P, nb, [&](int idx, double v) { P.c2[(size_t)(idx / nb) * (P.m + 2) + idx % nb] = v; });
    return;
  }
  gm_reduce(P, 1, [&](int idx, double v) { s_nrm[idx] = v; });
  const int m = P.m;
  if (wid < P.nrhs && st->done[wid] == 0) {
    const int sy = wid;
    double *gamma = P.gamma + (size_t)sy * (m + 1);
    if (MODE == GM_NORM_BEGIN) {
      // r0 = M^-1 (b - A x): rho = |r0|, first basis vector r0 / rho
      const double rho = sqrt(fmax(s_nrm[sy], 0.0));
      for (int i = 1 + lane; i <= m; i += 32) gamma[i] = 0.0;
      if (lane == 0) {
        gamma[0] = rho;
        st->rho[sy] = rho;
        st->inv_norm[sy] = rho > 0.0 ? 1.0 / rho : 0.0;
        st->dim[sy] = 0;
        if (rho <= P.tol) st->done[sy] = 1;
        else if (st->its[sy] >= P.max_steps) st->done[sy] = 2;
      }
    } else {
      // Hessenberg column j: h = h1 + h2, h[j+1] = |z|; Givens rotations (deal.II SolverGMRES::givens_rotation)
      double *hs = dyn + (size_t)sy * 3 * (m + 2), *cs = hs + (m + 2), *ss = cs + (m + 2);
      double *ci = P.ci + (size_t)sy * m, *si = P.si + (size_t)sy * m;
      const double *c1 = P.c1 + (size_t)sy * (m + 2), *c2 = P.c2 + (size_t)sy * (m + 2);
      for (int i = lane; i < nb; i += 32) {
        hs[i] = c1[i] + c2[i];
        if (i < j) {
          cs[i] = ci[i];
          ss[i] = si[i];
        }
      }
      __syncwarp();
      if (lane == 0) {
        const double hn = sqrt(fmax(s_nrm[sy], 0.0));
        hs[nb] = hn;
        for (int i = 0; i < j; ++i) {
          const double t = hs[i];
          hs[i] = cs[i] * t + ss[i] * hs[i + 1];
          hs[i + 1] = -ss[i] * t + cs[i] * hs[i + 1];
        }
        const double r = hypot(hs[j], hs[j + 1]);
        const double cj = r > 0.0 ? hs[j] / r : 1.0, sj = r > 0.0 ? hs[j + 1] / r : 0.0;
        ci[j] = cj;
        si[j] = sj;
        hs[j] = r;
        const double g = gamma[j];
        gamma[j + 1] = -sj * g;
        gamma[j] = cj * g;
        const double rho = fabs(sj * g);
        const int its = ++st->its[sy];
        st->rho[sy] = rho;
        st->inv_norm[sy] = hn > 0.0 ? 1.0 / hn : 0.0;  // hn == 0: the Krylov space is exhausted (rho is 0 too)
        st->dim[sy] = nb;
        if (rho <= P.tol) st->done[sy] = 1;
        else if (its >= P.max_steps) st->done[sy] = 2;
      }
      __syncwarp();
      double *H = P.H + (size_t)sy * (m + 1) * m;
      for (int i = lane; i < nb; i += 32) H[(size_t)i * m + j] = hs[i];
    }
  }
  __syncthreads();
  if (tid == 0) {
    int all = 1;
    for (int sy = 0; sy < P.nrhs; ++sy) all &= (st->done[sy] != 0);
    st->all_done = all;
    st->j = (MODE == GM_NORM_BEGIN) ? 0 : j + 1;
  }
}

// v = z * inv_norm (or the iterate x itself, FROM_X) -> basis vector st->j and slot s of every rank's replicated
// vector buffer; the last CTA publishes the stores (per-source flag on every peer).  grid (chunks, nrhs)
template <bool FROM_X>
__global__ void __launch_bounds__(GM_THREADS) k_gm_publish(GmDev P, const double *X, size_t ldv) {
  GmStatus *st = P.st;
  if (st->all_done) return;
  __shared__ int s_last;
  __shared__ unsigned long long s_e;
  const int s = blockIdx.y, tid = threadIdx.x;
  const int jn = st->j;
  if (st->done[s] == 0 && (FROM_X || jn < P.m)) {
    const double a = FROM_X ? 1.0 : st->inv_norm[s];
    const double *src = FROM_X ? X + (size_t)s * ldv : P.z + (size_t)s * P.ldw;
    double *bdst = FROM_X ? nullptr : P.basis + (size_t)s * P.bstride + (size_t)jn * P.ldb;
    const size_t off = (size_t)s * P.xld + P.slice_off;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < P.mloc; i += (size_t)gridDim.x * blockDim.x) {
      const double v = src[i] * a;
      if (!FROM_X) bdst[i] = v;
      if (P.nranks == 1) P.xloc[off + i] = v;
      else
        for (int r = 0; r < P.nranks; ++r) P.peer_x[r][off + i] = v;  // own buffer included; peers over NVLink
    }
  }
  if (P.nranks == 1) return;
  __threadfence_system();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) {
    st->ticket = 0;
    s_e = ++P.flags[BS_FLAG_EPOCH];
  }
  __syncthreads();
  __threadfence_system();
  if (tid < P.nranks) st_release_sys(P.peer_flags[tid] + BS_FLAG_XCHG + P.rank, s_e);
}

// w = b - w on the systems that still run
__global__ void k_gm_residual(GmDev P, const double *B, size_t ldv, double *w) {
  const GmStatus *st = P.st;
  if (st->all_done) return;
  const int s = blockIdx.y;
  if (st->done[s]) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.mloc; i += (size_t)gridDim.x * blockDim.x)
    w[(size_t)s * P.ldw + i] = B[(size_t)s * ldv + i] - w[(size_t)s * P.ldw + i];
}
__global__ void k_gm_jacobi(GmDev P, const double *w, const double *dinv, double *z) {
  const GmStatus *st = P.st;
  if (st->all_done) return;
  const int s = blockIdx.y;
  if (st->done[s]) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.mloc; i += (size_t)gridDim.x * blockDim.x)
    z[(size_t)s * P.ldw + i] = w[(size_t)s * P.ldw + i] * dinv[i];
}

// back substitution H y = gamma of every system that owns columns of the running cycle; one warp per system
__global__ void k_gm_backsolve(GmDev P) {
  __shared__ double y[1024];
  const GmStatus *st = P.st;
  const int s = blockIdx.x, lane = threadIdx.x, m = P.m;
  const int dim = st->dim[s];
  if (dim == 0) return;
  const double *H = P.H + (size_t)s * (m + 1) * m, *gamma = P.gamma + (size_t)s * (m + 1);
  for (int i = dim - 1; i >= 0; --i) {
    double t = 0.0;
    for (int jj = i + 1 + lane; jj < dim; jj += 32) t = fma(H[(size_t)i * m + jj], y[jj], t);
    t = warp_sum(t);
    if (lane == 0) {
      const double d = H[(size_t)i * m + i];
      y[i] = d != 0.0 ? (gamma[i] - t) / d : 0.0;
    }
    __syncwarp();
  }
  for (int i = lane; i < dim; i += 32) P.yk[(size_t)s * m + i] = y[i];
}
// x += sum_k y_k v_k, then the cycle's columns are consumed (dim = 0)
__global__ void __launch_bounds__(GM_THREADS) k_gm_update_x(GmDev P, double *X, size_t ldv) {
  GmStatus *st = P.st;
  __shared__ int s_last;
  const int s = blockIdx.y, tid = threadIdx.x;
  const int dim = st->dim[s];
  if (dim > 0) {
    const double *yk = P.yk + (size_t)s * P.m;
    const double *B = P.basis + (size_t)s * P.bstride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < P.mloc; i += (size_t)gridDim.x * blockDim.x) {
      double v = X[(size_t)s * ldv + i];
      for (int k = 0; k < dim; ++k) v = fma(yk[k], B[(size_t)k * P.ldb + i], v);
      X[(size_t)s * ldv + i] = v;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (s_last && tid == 0) {
    st->ticket = 0;
    for (int sy = 0; sy < GM_MAXSYS; ++sy) st->dim[sy] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side: queue the kernels, look at the status one group of iterations behind the device
// ---------------------------------------------------------------------------------------------------------
struct GmHost {  // per-context pinned status mirror (created on first use)
  GmStatus *pinned = nullptr;  // [2]
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> sweep_ev;  // pool: start / stop events around every sweep over the matrix (timing only)
  cudaEvent_t span[2] = {nullptr, nullptr};
};

bool gmres_device_eligible(Context &c, int nrhs, int max_tmp) {
  const int m = max_tmp - 2;
  if (std::getenv("BS_GMRES_HOST")) return false;
  if (c.gmres_ortho != BS_ORTHO_CGS2) return false;      // deal.II's MGS verbatim takes host decisions
  if (c.nranks > 1 && !c.p2p) return false;              // callback communicators: host-driven path
  if (nrhs > GM_MAXSYS || m + 2 > 1024) return false;
  if ((size_t)nrhs * (m + 2) > BS_RED_CAP) return false;
  if ((size_t)nrhs * 3 * (m + 2) * sizeof(double) > 160 * 1024) return false;
  return true;
}

int gmres_device(Context &c, int which, int nrhs, const double *d_B, double *d_X, size_t ldv, double tol, int max_steps,
                 int max_tmp, int *iters, double *final_res) {
  const int m = max_tmp - 2;
  const size_t mloc = c.local_vec_len(which), mfull = c.full_vec_len(which);
  const size_t ldb = (mloc + 1) & ~(size_t)1, ldw = (mloc + 3) & ~(size_t)1, ldx = (mfull + 4) & ~(size_t)3;
  const DMat &M = which == BS_MAT_V ? c.V : (which == BS_MAT_K ? c.K : c.A);
  BS_REQUIRE(M.valid(), which == BS_MAT_A ? "monolithic matrix not built" : "matrix not assembled");
  const bool p2p = c.nranks > 1;
  if (p2p) BS_REQUIRE(mfull <= c.xchg_ld, "exchange buffer too small (bs_exchange_export max_vec_len)");
  const int nchunks = (int)std::max<size_t>(1, (mloc + GM_CHUNK - 1) / GM_CHUNK);
  const int kld = m + 2;
  struct P_ { double *p; };  // keeps the workspace pointers apart from the grow-only map
  double *basis = c.wsd("gm.basis", (size_t)nrhs * (m + 1) * ldb + 2);
  double *w = c.wsd("gm.w", (size_t)nrhs * ldw + 2);
  double *z = c.prec_kind == BS_PREC_NONE ? w : c.wsd("gm.z", (size_t)nrhs * ldw + 2);
  const size_t small = (size_t)(m + 1) * m + (m + 1) + 2 * (size_t)m + 2 * (size_t)(m + 2) + m;
  double *sm = c.wsd("gm.small", (size_t)nrhs * small + BS_RED_CAP + 8);
  double *partial = c.wsd("gm.partial", (size_t)nrhs * kld * nchunks + 2);
  double *xfull = p2p ? c.d_xchg.p : c.wsd("gm.xfull", (size_t)nrhs * ldx + 2);
  const size_t xld = p2p ? c.xchg_ld : ldx;
  GmStatus *st = reinterpret_cast<GmStatus *>(c.wsd("gm.status", (sizeof(GmStatus) + 7) / 8 + 2));
  if (!c.gm_host) {
    GmHost *h = new GmHost();
    BS_CUDA(cudaMallocHost((void **)&h->pinned, 2 * sizeof(GmStatus)));
    BS_CUDA(cudaEventCreateWithFlags(&h->ev[0], cudaEventDisableTiming));
    BS_CUDA(cudaEventCreateWithFlags(&h->ev[1], cudaEventDisableTiming));
    c.gm_host = h;
  }
  GmHost &gh = *static_cast<GmHost *>(c.gm_host);

  GmDev P{};
  P.st = st;
  P.basis = basis;
  P.ldb = ldb;
  P.bstride = (size_t)(m + 1) * ldb;
  P.z = z;
  P.ldw = ldw;
  double *q = sm;
  P.H = q, q += (size_t)nrhs * (m + 1) * m;
  P.gamma = q, q += (size_t)nrhs * (m + 1);
  P.ci = q, q += (size_t)nrhs * m;
  P.si = q, q += (size_t)nrhs * m;
  P.c1 = q, q += (size_t)nrhs * (m + 2);
  P.c2 = q, q += (size_t)nrhs * (m + 2);
  P.yk = q, q += (size_t)nrhs * m;
  P.red_local = q;
  P.m = m;
  P.nrhs = nrhs;
  P.mloc = mloc;
  P.partial = partial;
  P.nchunks = nchunks;
  P.kld = kld;
  P.tol = tol;
  P.max_steps = max_steps;
  P.nranks = c.nranks;
  P.rank = c.rank;
  P.peer_x = p2p ? c.d_peer_xbuf.p : nullptr;
  P.peer_flags = p2p ? c.d_peer_flags.p : nullptr;
  P.flags = p2p ? c.d_flags.p : nullptr;
  P.red_off = c.red_off;
  P.red_cap = BS_RED_CAP;
  P.xloc = xfull;
  P.xld = xld;
  P.slice_off = c.slice_offset(which);

  BS_CUDA(cudaMemsetAsync(st, 0, sizeof(GmStatus), c.stream));
  const dim3 gpass(nchunks, nrhs), gvec((unsigned)std::min<size_t>(std::max<size_t>((mloc + 255) / 256, 1), 592), nrhs);
  const size_t dyn_givens = (size_t)nrhs * 3 * (m + 2) * sizeof(double);
  if (dyn_givens > 48 * 1024)
    BS_CUDA(cudaFuncSetAttribute(k_gm_pass<GM_UPD_NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const int *skip = &st->all_done;

  // CUDA events around every sweep over the matrix: the solve time on the stream splits into matvec and the rest
  // (bs_stats gmres_*_last; two event records per sweep, no synchronisation)
  const bool trace = std::getenv("BS_TRACE") != nullptr;
  size_t nev = 0;
  auto matvec = [&]() {  // w = A_loc * (replicated vectors of the slots)
    if (p2p) p2p_wait_only(c, skip, st);
    if (gh.sweep_ev.size() < nev + 2) {
      gh.sweep_ev.resize(nev + 2, nullptr);
      cudaEventCreate(&gh.sweep_ev[nev]);
      cudaEventCreate(&gh.sweep_ev[nev + 1]);
    }
    cudaEventRecord(gh.sweep_ev[nev], c.stream);
    if (nrhs == 1) gemv(c, M, xfull, w, skip);
    else gemv_multi(c, M, nrhs, xfull, xld, w, ldw, skip);
    cudaEventRecord(gh.sweep_ev[nev + 1], c.stream);
    nev += 2;
  };
  if (!gh.span[0]) {
    cudaEventCreate(&gh.span[0]);
    cudaEventCreate(&gh.span[1]);
  }
  cudaEventRecord(gh.span[0], c.stream);
  auto precondition = [&]() {  // z = M^-1 w
    switch (c.prec_kind) {
      case BS_PREC_NONE: break;  // z aliases w
      case BS_PREC_JACOBI:
        k_gm_jacobi<<<gvec, 256, 0, c.stream>>>(P, w, c.d_prec_diag.p, z);
        count_launch(c);
        break;
      default:
        for (int s = 0; s < nrhs; ++s) apply_precond(c, w + (size_t)s * ldw, z + (size_t)s * ldw, skip);
    }
  };
  auto begin_cycle = [&]() {  // r0 = M^-1 (b - A x), rho, first basis vector
    k_gm_publish<true><<<gvec, GM_THREADS, 0, c.stream>>>(P, d_X, ldv);
    count_launch(c);
    matvec();
    k_gm_residual<<<gvec, 256, 0, c.stream>>>(P, d_B, ldv, w);
    count_launch(c);
    precondition();
    k_gm_pass<GM_NORM_BEGIN><<<gpass, GM_THREADS, 0, c.stream>>>(P);
    k_gm_publish<false><<<gvec, GM_THREADS, 0, c.stream>>>(P, nullptr, 0);
    count_launch(c, 2);
  };
  auto iteration = [&]() {
    matvec();
    precondition();
    k_gm_pass<GM_DOTS><<<gpass, GM_THREADS, 0, c.stream>>>(P);
    k_gm_pass<GM_UPD_DOTS><<<gpass, GM_THREADS, 0, c.stream>>>(P);
    k_gm_pass<GM_UPD_NORM><<<gpass, GM_THREADS, dyn_givens, c.stream>>>(P);
    k_gm_publish<false><<<gvec, GM_THREADS, 0, c.stream>>>(P, nullptr, 0);
    count_launch(c, 4);
  };
  auto finish_cycle = [&]() {
    k_gm_backsolve<<<nrhs, 32, 0, c.stream>>>(P);
    k_gm_update_x<<<gvec, GM_THREADS, 0, c.stream>>>(P, d_X, ldv);
    count_launch(c, 2);
  };
  auto failed = [&](const GmStatus &h) {
    for (int s = 0; s < nrhs; ++s)
      if (h.done[s] == 3) return true;
    return false;
  };

  constexpr int CHECK = 8;
  GmStatus last{};
  bool stop_all = false;
  while (!stop_all) {
    begin_cycle();
    int pending = -1, grp = 0;
    bool stop = false;
    for (int inner = 0; inner < m && !stop; ++inner) {
      iteration();
      if ((inner + 1) % CHECK == 0) {
        const int slot = grp & 1;
        BS_CUDA(cudaMemcpyAsync(&gh.pinned[slot], st, sizeof(GmStatus), cudaMemcpyDeviceToHost, c.stream));
        BS_CUDA(cudaEventRecord(gh.ev[slot], c.stream));
        if (pending >= 0) {
          BS_CUDA(cudaEventSynchronize(gh.ev[pending]));
          if (gh.pinned[pending].all_done) stop = true;
        }
        pending = slot;
        ++grp;
      }
    }
    finish_cycle();
    BS_CUDA(cudaMemcpyAsync(&gh.pinned[0], st, sizeof(GmStatus), cudaMemcpyDeviceToHost, c.stream));
    BS_CUDA(cudaStreamSynchronize(c.stream));
    last = gh.pinned[0];
    BS_CUDA(cudaGetLastError());
    if (last.all_done) stop_all = true;
  }
  {
    cudaEventRecord(gh.span[1], c.stream);
    cudaEventSynchronize(gh.span[1]);
    float total = 0, mv = 0, mv_max = 0;
    cudaEventElapsedTime(&total, gh.span[0], gh.span[1]);
    int nmv = 0, its_max = 0;
    for (int s = 0; s < nrhs; ++s) its_max = std::max(its_max, last.its[s]);
    const int real_sweeps = its_max + (its_max == 0 ? 1 : (its_max + m - 1) / m);  // one per iteration + one residual per cycle
    for (size_t i = 0; i + 1 < nev && nmv < real_sweeps; i += 2) {  // sweeps queued behind the convergence point returned at once
      float ms = 0;
      cudaEventElapsedTime(&ms, gh.sweep_ev[i], gh.sweep_ev[i + 1]);
      mv += ms;
      ++nmv;
      mv_max = std::max(mv_max, ms);
    }
    c.stats.gmres_stream_ms_last = total;
    c.stats.gmres_matvec_ms_last = mv;
    c.stats.gmres_sweeps_last = nmv;
    if (trace)
      fprintf(stderr, "[bs trace rank %d] device GMRES: %d iterations, %.2f ms on the stream; %d sweeps over the matrix %.2f ms (max %.3f ms); "
                      "everything else %.3f ms = %.4f ms per iteration\n", c.rank, its_max, total, nmv, mv, mv_max, total - mv,
              (total - mv) / std::max(1, its_max));
  }
  if (failed(last)) throw Error(BS_ERR_COMM, "peer exchange timed out during the GMRES iteration (a rank stopped answering)");
  int rc = BS_OK;
  for (int s = 0; s < nrhs; ++s) {
    if (iters) iters[s] = last.its[s];
    if (final_res) final_res[s] = last.rho[s];
    if (last.done[s] != 1) rc = BS_ERR_NOT_CONVERGED;
  }
  return rc;
}

void gm_host_release(Context &c) {
  if (!c.gm_host) return;
  GmHost *h = static_cast<GmHost *>(c.gm_host);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->ev[0]) cudaEventDestroy(h->ev[0]);
  if (h->ev[1]) cudaEventDestroy(h->ev[1]);
  for (cudaEvent_t e : h->sweep_ev)
    if (e) cudaEventDestroy(e);
  if (h->span[0]) cudaEventDestroy(h->span[0]);
  if (h->span[1]) cudaEventDestroy(h->span[1]);
  delete h;
  c.gm_host = nullptr;
}

}  // namespace bs
