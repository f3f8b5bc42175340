// Representation formula at arbitrary points — replaces BEMProblem::evaluate_stokes_bie and
// evaluate_stokes_bie_on_boundary (ref: source/bem_stokes.cc:5366-5451, 5454-5560):
//     u_a(x) = sum_cells sum_q [ G_ab(y_q - x) f_b(y_q) - (W_abk(y_q - x) n_k) u_b(y_q) ] JxW_q
// with f, u the FE interpolants of `forces` and `vel`.  Same device Green functions as the assembly.
//   E0 k_eval_density   f_q, u_q at the regular quadrature points of every cell
//   E1 k_eval_regular   CTA = 128 evaluation points x a chunk of cells (cell records staged in shared memory),
//                       per-chunk partial sums; E2 k_eval_reduce adds the chunks in a fixed order
//   E3 k_eval_singular  on-boundary variant: one warp per point finds the cells having a support point within
//                       1e-3 of it and integrates those with the singular rule of that local index (ref 5494-5520)
#include "bs_internal.h"
#include "bs_green.cuh"

namespace bs {

template <int NV>
__device__ __forceinline__ int eidx(int i, int j) {
  if (NV == 9) return 3 * i + j;
  const int a = i < j ? i : j, b = i < j ? j : i;
  return a == 0 ? b : (a == 1 ? 2 + b : 5);
}

// dens[cell][6][nq_pad] = f(3), u(3) interpolated at the regular quadrature points
__global__ void k_eval_density(int ncell, int nq, int nq_pad, int na, const int *__restrict__ conn_pos,
                               const double *__restrict__ phi /*[nq][na]*/, const double *__restrict__ forces,
                               const double *__restrict__ vel, double *__restrict__ dens) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)ncell * nq) return;
  const int cell = (int)(gid / nq), q = (int)(gid % nq);
  double f[3] = {0, 0, 0}, u[3] = {0, 0, 0};
  for (int a = 0; a < na; ++a) {
    const int p = conn_pos[(size_t)cell * na + a];
    const double ph = phi[(size_t)q * na + a];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      f[d] = fma(ph, forces[(size_t)3 * p + d], f[d]);
      u[d] = fma(ph, vel[(size_t)3 * p + d], u[d]);
    }
  }
  double *o = dens + (size_t)cell * 6 * nq_pad;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    o[d * nq_pad + q] = f[d];
    o[(3 + d) * nq_pad + q] = u[d];
  }
}

constexpr int EV_T = 128;

template <int KT, int NA>
__global__ void __launch_bounds__(EV_T) k_eval_regular(int npts, const double *__restrict__ pts, int ncell, int cells_per_chunk,
                                                       int nq, int nq_pad, const double *__restrict__ cellq,
                                                       const double *__restrict__ dens, const int *__restrict__ conn_pos,
                                                       const double *__restrict__ support, int skip_near, KernelParams kp,
                                                       double *__restrict__ partial /*[chunks][3][npts]*/) {
  constexpr int NV = GreenTraits<KT>::NV;
  extern __shared__ double sm[];  // [13][nq_pad]
  const int t = threadIdx.x;
  const int i = blockIdx.x * EV_T + t;
  const bool ok = i < npts;
  double x[3] = {0, 0, 0};
  if (ok) {
    x[0] = pts[(size_t)3 * i];
    x[1] = pts[(size_t)3 * i + 1];
    x[2] = pts[(size_t)3 * i + 2];
  }
  double xim[3] = {x[0], x[1], x[2]};
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (d == kp.o) xim[d] = x[d] - 2.0 * (x[d] - kp.wall_pos);  // ref: bem_stokes.cc:5419-5420
  const int c0 = blockIdx.y * cells_per_chunk, c1 = min(ncell, c0 + cells_per_chunk);
  double acc[3] = {0, 0, 0};
  for (int cell = c0; cell < c1; ++cell) {
    __syncthreads();
    for (int k = t; k < 7 * nq_pad; k += EV_T) sm[k] = cellq[(size_t)cell * 7 * nq_pad + k];
    for (int k = t; k < 6 * nq_pad; k += EV_T) sm[7 * nq_pad + k] = dens[(size_t)cell * 6 * nq_pad + k];
    __syncthreads();
    bool near = false;
    if (skip_near) {  // on-boundary variant: cells with a support point within 1e-3 are done by k_eval_singular
      for (int a = 0; a < NA; ++a) {
        const int p = conn_pos[(size_t)cell * NA + a];
        const double dx = support[(size_t)3 * p] - x[0], dy = support[(size_t)3 * p + 1] - x[1],
                     dz = support[(size_t)3 * p + 2] - x[2];
        near |= (sqrt(dx * dx + dy * dy + dz * dz) <= 1e-3);
      }
    }
    if (!ok || near) continue;
    for (int q = 0; q < nq; ++q) {
      double R[3], Rim[3], nJ[3], g[NV], k[NV];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double yq = sm[d * nq_pad + q];
        R[d] = yq - x[d];
        Rim[d] = yq - xim[d];
        nJ[d] = sm[(3 + d) * nq_pad + q];
      }
      green_eval<KT>(R, Rim, nJ, sm[6 * nq_pad + q], kp.eps, kp.o, g, k);
      const double f[3] = {sm[7 * nq_pad + q], sm[8 * nq_pad + q], sm[9 * nq_pad + q]};
      const double u[3] = {sm[10 * nq_pad + q], sm[11 * nq_pad + q], sm[12 * nq_pad + q]};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a] = fma(g[eidx<NV>(a, b)], f[b], fma(k[eidx<NV>(a, b)], u[b], acc[a]));
    }
  }
  if (ok)
#pragma unroll
    for (int a = 0; a < 3; ++a) partial[((size_t)blockIdx.y * 3 + a) * npts + i] = acc[a];
}

__global__ void k_eval_reduce(int npts, int nchunks, const double *__restrict__ partial, double *__restrict__ out, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 3 * npts) return;
  double s = accumulate ? out[idx] : 0.0;
  for (int c = 0; c < nchunks; ++c) s += partial[(size_t)c * 3 * npts + idx];
  out[idx] = s;  // out is component-major: idx = a*npts + i  (ref: val_velocities(i + size/dim*idim))
}

// on-boundary singular part (free-space kernel only, as in the reference: exterior_stokes_kernel.value_tens)
template <int NA, int NAM>
__global__ void __launch_bounds__(128) k_eval_singular(int npts, const double *__restrict__ pts, int ncell,
                                                       const int *__restrict__ conn_pos, const int *__restrict__ conn_map,
                                                       const double *__restrict__ support, const double *__restrict__ map_nodes,
                                                       const double *__restrict__ tab, const int *__restrict__ sing_off,
                                                       const int *__restrict__ sing_nq, const double *__restrict__ forces,
                                                       const double *__restrict__ vel, double eps, double *__restrict__ out) {
  constexpr int REC = NA + 3 * NAM + 1;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= npts) return;
  const double x[3] = {pts[(size_t)3 * i], pts[(size_t)3 * i + 1], pts[(size_t)3 * i + 2]};
  double acc[3] = {0, 0, 0};
  for (int base = 0; base < ncell; base += 32) {
    const int cell = base + lane;
    int hit = -1;
    if (cell < ncell) {
      for (int a = 0; a < NA; ++a) {  // first local index within tol, like the `break` of ref 5494-5505
        const int p = conn_pos[(size_t)cell * NA + a];
        const double dx = support[(size_t)3 * p] - x[0], dy = support[(size_t)3 * p + 1] - x[1],
                     dz = support[(size_t)3 * p + 2] - x[2];
        if (sqrt(dx * dx + dy * dy + dz * dz) <= 1e-3) {
          hit = a;
          break;
        }
      }
    }
    unsigned mask = __ballot_sync(0xffffffffu, hit >= 0);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int c = base + src;
      const int al = __shfl_sync(0xffffffffu, hit, src);
      const int off = sing_off[al], nqs = sing_nq[al];
      for (int q = lane; q < nqs; q += 32) {
        const double *rec = tab + (size_t)(off + q) * REC;
        double y[3] = {0, 0, 0}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0}, f[3] = {0, 0, 0}, u[3] = {0, 0, 0};
        for (int a = 0; a < NAM; ++a) {
          const int m = conn_map[(size_t)c * NAM + a];
          const double ph = rec[NA + 3 * a], dx = rec[NA + 3 * a + 1], dy = rec[NA + 3 * a + 2];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double X = map_nodes[(size_t)3 * m + d];
            y[d] = fma(ph, X, y[d]);
            t1[d] = fma(dx, X, t1[d]);
            t2[d] = fma(dy, X, t2[d]);
          }
        }
        for (int a = 0; a < NA; ++a) {
          const int p = conn_pos[(size_t)c * NA + a];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            f[d] = fma(rec[a], forces[(size_t)3 * p + d], f[d]);
            u[d] = fma(rec[a], vel[(size_t)3 * p + d], u[d]);
          }
        }
        const double w = rec[NA + 3 * NAM];
        const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2], nz = t1[0] * t2[1] - t1[1] * t2[0];
        const double JxW = w * sqrt(nx * nx + ny * ny + nz * nz);
        const double nJ[3] = {w * nx, w * ny, w * nz};
        double R[3] = {y[0] - x[0], y[1] - x[1], y[2] - x[2]}, g[6], k[6];
        green_eval<BS_KERNEL_FREE>(R, R, nJ, JxW, eps, 1, g, k);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) acc[a] = fma(g[eidx<6>(a, b)], f[b], fma(k[eidx<6>(a, b)], u[b], acc[a]));
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double s = acc[a];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) out[(size_t)a * npts + i] += s;
  }
}

template <int KT>
static void launch_eval_regular(Context &c, int npts, const double *d_pts, int chunks, int cpc, const double *dens, int skip_near,
                                const KernelParams &kp, double *partial) {
  const size_t smem = (size_t)13 * c.nq_pad * sizeof(double);
  dim3 grid((npts + EV_T - 1) / EV_T, chunks);
  if (c.na == 4) {
    BS_CUDA(cudaFuncSetAttribute(k_eval_regular<KT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eval_regular<KT, 4><<<grid, EV_T, smem, c.stream>>>(npts, d_pts, c.ncell, cpc, c.nq, c.nq_pad, c.d_cellq.p, dens,
                                                          c.d_conn_pos.p, c.d_support.p, skip_near, kp, partial);
  } else {
    BS_CUDA(cudaFuncSetAttribute(k_eval_regular<KT, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eval_regular<KT, 9><<<grid, EV_T, smem, c.stream>>>(npts, d_pts, c.ncell, cpc, c.nq, c.nq_pad, c.d_cellq.p, dens,
                                                          c.d_conn_pos.p, c.d_support.p, skip_near, kp, partial);
  }
  BS_CUDA(cudaGetLastError());
  count_launch(c);
}

// d_vel / d_forces: internal ordering (3N); d_out: 3*npts component-major
void evaluate_bie(Context &c, int npts, const double *d_pts, const double *d_vel, const double *d_forces, double *d_out,
                  bool on_boundary, bool accumulate) {
  BS_REQUIRE(c.have_geometry && c.have_quadrature, "geometry and quadrature must be set");
  if (npts <= 0) return;
  launch_cell_geometry(c);
  struct P_ { double *p; } dens, partial;
  dens.p = c.wsd("eval.dens", (size_t)c.ncell * 6 * c.nq_pad);
  BS_CUDA(cudaMemsetAsync(dens.p, 0, sizeof(double) * (size_t)c.ncell * 6 * c.nq_pad, c.stream));
  const long long total = (long long)c.ncell * c.nq;
  k_eval_density<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(c.ncell, c.nq, c.nq_pad, c.na, c.d_conn_pos.p, c.d_phi_reg.p,
                                                                       d_forces, d_vel, dens.p);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
  // enough chunks to fill the GPU when there are few evaluation points
  const int ptiles = (npts + EV_T - 1) / EV_T;
  int chunks = std::max(1, std::min(c.ncell, (4 * c.sm_count + ptiles - 1) / ptiles));
  chunks = std::min(chunks, 65535);
  const int cpc = (c.ncell + chunks - 1) / chunks;
  chunks = (c.ncell + cpc - 1) / cpc;
  partial.p = c.wsd("eval.partial", (size_t)chunks * 3 * npts);
  KernelParams kp = c.kp;
  if (on_boundary) kp.type = BS_KERNEL_FREE;  // the reference's on-boundary formula uses the free-space kernel only
  switch (kp.type) {
    case BS_KERNEL_FREE: launch_eval_regular<BS_KERNEL_FREE>(c, npts, d_pts, chunks, cpc, dens.p, on_boundary ? 1 : 0, kp, partial.p); break;
    case BS_KERNEL_FREE_SURFACE: launch_eval_regular<BS_KERNEL_FREE_SURFACE>(c, npts, d_pts, chunks, cpc, dens.p, 0, kp, partial.p); break;
    default: launch_eval_regular<BS_KERNEL_NO_SLIP>(c, npts, d_pts, chunks, cpc, dens.p, 0, kp, partial.p); break;
  }
  k_eval_reduce<<<(3 * npts + 255) / 256, 256, 0, c.stream>>>(npts, chunks, partial.p, d_out, accumulate ? 1 : 0);
  BS_CUDA(cudaGetLastError());
  count_launch(c);
  if (on_boundary) {
    BS_REQUIRE(c.have_singular, "singular quadrature not set");
    const int grid = (npts + 3) / 4;
#define BS_EVS(NA_, NAM_)                                                                                              \
  k_eval_singular<NA_, NAM_><<<grid, 128, 0, c.stream>>>(npts, d_pts, c.ncell, c.d_conn_pos.p, c.d_conn_map.p, c.d_support.p, \
                                                         c.d_map_nodes.p, c.d_sing_tab.p, c.d_sing_off.p, c.d_sing_nq.p,     \
                                                         d_forces, d_vel, c.kp.eps, d_out)
    if (c.na == 4 && c.na_map == 4) BS_EVS(4, 4);
    else if (c.na == 4 && c.na_map == 9) BS_EVS(4, 9);
    else if (c.na == 9 && c.na_map == 4) BS_EVS(9, 4);
    else BS_EVS(9, 9);
#undef BS_EVS
    BS_CUDA(cudaGetLastError());
    count_launch(c);
  }
  BS_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace bs
