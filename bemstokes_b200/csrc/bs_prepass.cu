// Device pre-pass: the per-frame inputs of the hot path computed on the GPU from the geometry the context already
// holds.  Same results as the host restatement `host_prepass` (bs_host.cu), which stays as the
// host-only entry point.
//
//   scalar mass matrix  M_ab = sum_q phi_a phi_b JxW                 ref: source/bem_stokes.cc:2499-2517
//   L2-projected normals  M n = int phi n,  n_hat = n/|n| per node    ref: 3945-3998
//   M n_hat,  l2 = n_hat^T M n_hat                                    ref: 4002-4005
//   rigid modes about the pole and their duals M N_r                  ref: 2626-2641, 2773
//
// The mass matrix is never assembled: every cell keeps its na x na local matrix (one thread per (cell, a) row,
// quadrature data from K0), and M x is a deterministic gather over the node's (cell, local index) patch - the same
// CSR patch table the singular pass uses.  The three normal components are solved together by Jacobi-preconditioned
// CG with device-resident step scalars (no host round trip inside an iteration); the host reads the three residual
// norms once per iteration to stop at 1e-15 relative, as the host version does (reference: Trilinos CG + AMG).
#include "bs_internal.h"

namespace bs {

namespace {

constexpr int PRE_T = 256;

// local mass rows and normal moments: thread (cell, a)
__global__ void k_pre_local(int ncell, int na, int nq, int nq_pad, const double *__restrict__ cellq,
                            const double *__restrict__ phi /*[nq][na]*/, double *__restrict__ mloc /*[ncell][na][na]*/,
                            double *__restrict__ bloc /*[ncell][na][3]*/, double *__restrict__ cell_area /*[4][ncell]: area, int y dS*/) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= ncell * na) return;
  const int cell = gid / na, a = gid - cell * na;
  const double *cq = cellq + (size_t)cell * 7 * nq_pad;
  double m[MAX_NA], b[3] = {0, 0, 0}, area = 0, cy[3] = {0, 0, 0};
  for (int k = 0; k < MAX_NA; ++k) m[k] = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double pa = phi[(size_t)q * na + a], jxw = cq[6 * nq_pad + q];
    const double pj = pa * jxw;
#pragma unroll
    for (int k = 0; k < MAX_NA; ++k)
      if (k < na) m[k] = fma(pj, phi[(size_t)q * na + k], m[k]);
#pragma unroll
    for (int d = 0; d < 3; ++d) b[d] = fma(pa, cq[(3 + d) * nq_pad + q], b[d]);  // phi_a * n_d * JxW
    area += jxw;
#pragma unroll
    for (int d = 0; d < 3; ++d) cy[d] = fma(cq[d * nq_pad + q], jxw, cy[d]);  // first moment: centre of mass (ref: 2487-2493)
  }
  for (int k = 0; k < na; ++k) mloc[((size_t)cell * na + a) * na + k] = m[k];
#pragma unroll
  for (int d = 0; d < 3; ++d) bloc[((size_t)cell * na + a) * 3 + d] = b[d];
  if (a == 0) {
    cell_area[cell] = area;
#pragma unroll
    for (int d = 0; d < 3; ++d) cell_area[(size_t)(1 + d) * ncell + cell] = cy[d];
  }
}

// rhs[3p+d] = sum over the node's patch of bloc; dinv[p] = 1 / M_pp
__global__ void k_pre_gather_rhs(int N, int na, const int *__restrict__ patch_ptr, const int *__restrict__ patch_cell,
                                 const int *__restrict__ patch_local, const double *__restrict__ mloc,
                                 const double *__restrict__ bloc, double *__restrict__ rhs, double *__restrict__ dinv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  double b[3] = {0, 0, 0}, dg = 0;
  for (int e = patch_ptr[p]; e < patch_ptr[p + 1]; ++e) {
    const int cell = patch_cell[e], a = patch_local[e];
#pragma unroll
    for (int d = 0; d < 3; ++d) b[d] += bloc[((size_t)cell * na + a) * 3 + d];
    dg += mloc[((size_t)cell * na + a) * na + a];
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) rhs[(size_t)3 * p + d] = b[d];
  dinv[p] = 1.0 / dg;
}

// y = M x applied to each of the three components of nvec internal vectors (x, y: [nvec][3N])
__global__ void k_pre_mass_mult(int N, int na, int nvec, const int *__restrict__ patch_ptr, const int *__restrict__ patch_cell,
                                const int *__restrict__ patch_local, const int *__restrict__ conn_pos,
                                const double *__restrict__ mloc, const double *__restrict__ x, double *__restrict__ y) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (p >= N || v >= nvec) return;
  const double *xv = x + (size_t)v * 3 * N;
  double s[3] = {0, 0, 0};
  for (int e = patch_ptr[p]; e < patch_ptr[p + 1]; ++e) {
    const int cell = patch_cell[e], a = patch_local[e];
    const double *mr = mloc + ((size_t)cell * na + a) * na;
    const int *cn = conn_pos + (size_t)cell * na;
    for (int k = 0; k < na; ++k) {
      const double mk = mr[k];
      const double *xk = xv + (size_t)3 * cn[k];
      s[0] = fma(mk, xk[0], s[0]);
      s[1] = fma(mk, xk[1], s[1]);
      s[2] = fma(mk, xk[2], s[2]);
    }
  }
  double *yv = y + (size_t)v * 3 * N + (size_t)3 * p;
  yv[0] = s[0];
  yv[1] = s[1];
  yv[2] = s[2];
}

// per-component dot products of two internal vectors: partial[block][3], then k_pre_sum3 -> out[3]
__global__ void k_pre_dot3(int N, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ partial) {
  __shared__ double sh[3][PRE_T / 32];
  double s[3] = {0, 0, 0};
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) s[d] = fma(a[(size_t)3 * p + d], b[(size_t)3 * p + d], s[d]);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int m = 16; m > 0; m >>= 1) s[d] += __shfl_xor_sync(0xffffffffu, s[d], m);
    if ((threadIdx.x & 31) == 0) sh[d][threadIdx.x >> 5] = s[d];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int w = 0; w < PRE_T / 32; ++w) t += sh[threadIdx.x][w];
    partial[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
  }
}
__global__ void k_pre_sum3(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
  if (threadIdx.x < 3) {
    double t = 0;
    for (int b = 0; b < nblocks; ++b) t += partial[(size_t)b * 3 + threadIdx.x];
    out[threadIdx.x] = t;
  }
}

// CG state scalars on the device: sc[0..2] = rz, [3..5] = pAp, [6..8] = rn (|r|^2), [9..11] = rz_new, [12..14] = |b|^2
__global__ void k_pre_cg_init(int N, const double *__restrict__ rhs, const double *__restrict__ dinv, double *__restrict__ x,
                              double *__restrict__ r, double *__restrict__ z, double *__restrict__ pv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double b = rhs[(size_t)3 * p + d];
    x[(size_t)3 * p + d] = 0.0;
    r[(size_t)3 * p + d] = b;
    z[(size_t)3 * p + d] = dinv[p] * b;
    pv[(size_t)3 * p + d] = dinv[p] * b;
  }
}
// x += alpha p, r -= alpha Ap, z = D^-1 r   with alpha_d = rz_d / pAp_d (0 once the component has converged)
__global__ void k_pre_cg_update(int N, const double *__restrict__ sc, const double *__restrict__ dinv,
                                const double *__restrict__ pv, const double *__restrict__ Ap, double *__restrict__ x,
                                double *__restrict__ r, double *__restrict__ z) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double pAp = sc[3 + d];
    const double alpha = (pAp > 0.0 && sc[6 + d] > 1e-30 * sc[12 + d]) ? sc[d] / pAp : 0.0;
    const size_t i = (size_t)3 * p + d;
    x[i] = fma(alpha, pv[i], x[i]);
    const double rr = fma(-alpha, Ap[i], r[i]);
    r[i] = rr;
    z[i] = dinv[p] * rr;
  }
}
// p = z + beta p with beta_d = rz_new_d / rz_d; then rz <- rz_new (done by thread 0 of block 0 after everyone read it:
// the scalars are double buffered through sc[9..11], the swap happens in k_pre_cg_swap)
__global__ void k_pre_cg_direction(int N, const double *__restrict__ sc, const double *__restrict__ z, double *__restrict__ pv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double beta = sc[d] > 0.0 ? sc[9 + d] / sc[d] : 0.0;
    const size_t i = (size_t)3 * p + d;
    pv[i] = fma(beta, pv[i], z[i]);
  }
}
__global__ void k_pre_cg_swap(double *__restrict__ sc) {
  if (threadIdx.x < 3) sc[threadIdx.x] = sc[9 + threadIdx.x];
}

// n_hat = sol / |sol| per node
__global__ void k_pre_normalise(int N, const double *__restrict__ sol, double *__restrict__ nhat) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const double a = sol[(size_t)3 * p], b = sol[(size_t)3 * p + 1], c = sol[(size_t)3 * p + 2];
  const double nrm = sqrt(a * a + b * b + c * c);
  nhat[(size_t)3 * p] = a / nrm;
  nhat[(size_t)3 * p + 1] = b / nrm;
  nhat[(size_t)3 * p + 2] = c / nrm;
}

// six rigid modes about the pole as internal vectors [6][3N]: e_x, e_y, e_z, e_x x r, e_y x r, e_z x r
__global__ void k_pre_rigid(int N, const double *__restrict__ support, double px, double py, double pz, double *__restrict__ Nr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const double x = support[(size_t)3 * p] - px, y = support[(size_t)3 * p + 1] - py, z = support[(size_t)3 * p + 2] - pz;
  const size_t n3 = (size_t)3 * N, i = (size_t)3 * p;
  const double m[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -z, y}, {z, 0, -x}, {-y, x, 0}};
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int d = 0; d < 3; ++d) Nr[r * n3 + i + d] = m[r][d];
}

__global__ void k_pre_sum1(int n, const double *__restrict__ v, double *__restrict__ out) {  // single block, fixed order
  __shared__ double sh[PRE_T];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += PRE_T) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int m = PRE_T / 2; m > 0; m >>= 1) {
    if (threadIdx.x < m) sh[threadIdx.x] += sh[threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace

// Results stay on the device in internal ordering: nhat, Mnhat [3N]; Nr, Nrd [6][3N]; scalars[0] = l2, [1] = area.
void device_prepass(Context &c, int pole_kind, const double pole_in[3], double *d_nhat, double *d_Mnhat, double *d_Nr,
                    double *d_Nrd, double *h_l2, double *h_area, double *h_center_of_mass, double *h_pole_used, int *cg_iterations) {
  BS_REQUIRE(c.have_geometry && c.have_quadrature, "geometry and quadrature must be set before the pre-pass");
  const int N = c.N, na = c.na, ncell = c.ncell;
  const size_t n3 = (size_t)3 * N;
  launch_cell_geometry(c);
  double *mloc = c.wsd("pre.mloc", (size_t)ncell * na * na);
  double *bloc = c.wsd("pre.bloc", (size_t)ncell * na * 3);
  double *carea = c.wsd("pre.area", (size_t)4 * ncell);
  double *rhs = c.wsd("pre.rhs", n3), *dinv = c.wsd("pre.dinv", N);
  double *x = c.wsd("pre.x", n3), *r = c.wsd("pre.r", n3), *z = c.wsd("pre.z", n3), *pv = c.wsd("pre.p", n3), *Ap = c.wsd("pre.Ap", n3);
  const int nb_dot = std::min(4 * c.sm_count, (N + PRE_T - 1) / PRE_T);
  double *partial = c.wsd("pre.partial", (size_t)3 * nb_dot);
  double *sc = c.wsd("pre.scalars", 32);
  const dim3 gN((N + PRE_T - 1) / PRE_T);
  cudaStream_t s = c.stream;

  k_pre_local<<<(ncell * na + PRE_T - 1) / PRE_T, PRE_T, 0, s>>>(ncell, na, c.nq, c.nq_pad, c.d_cellq.p, c.d_phi_reg.p, mloc, bloc, carea);
  k_pre_gather_rhs<<<gN, PRE_T, 0, s>>>(N, na, c.d_patch_ptr.p, c.d_patch_cell.p, c.d_patch_local.p, mloc, bloc, rhs, dinv);
  BS_CUDA(cudaGetLastError());
  auto mult = [&](int nvec, const double *in, double *out) {
    k_pre_mass_mult<<<dim3(gN.x, nvec), PRE_T, 0, s>>>(N, na, nvec, c.d_patch_ptr.p, c.d_patch_cell.p, c.d_patch_local.p,
                                                      c.d_conn_pos.p, mloc, in, out);
  };
  auto dot3 = [&](const double *a, const double *b, double *out) {
    k_pre_dot3<<<nb_dot, PRE_T, 0, s>>>(N, a, b, partial);
    k_pre_sum3<<<1, 32, 0, s>>>(nb_dot, partial, out);
  };
  // ---- Jacobi-preconditioned CG on M sol_d = rhs_d, d = 0..2 together
  k_pre_cg_init<<<gN, PRE_T, 0, s>>>(N, rhs, dinv, x, r, z, pv);
  dot3(r, z, sc + 0);        // rz
  dot3(rhs, rhs, sc + 12);   // |b|^2
  BS_CUDA(cudaMemcpyAsync(sc + 6, sc + 12, 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));  // rn = |b|^2
  double h[3], hb[3];
  BS_CUDA(cudaMemcpyAsync(hb, sc + 12, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  BS_CUDA(cudaStreamSynchronize(s));
  int it = 0;
  const int max_it = 10 * N + 100;
  for (; it < max_it; ++it) {
    mult(1, pv, Ap);
    dot3(pv, Ap, sc + 3);                                               // pAp
    k_pre_cg_update<<<gN, PRE_T, 0, s>>>(N, sc, dinv, pv, Ap, x, r, z);  // uses rn of the previous step as the freeze test
    dot3(r, r, sc + 6);                                                 // rn
    dot3(r, z, sc + 9);                                                 // rz_new
    k_pre_cg_direction<<<gN, PRE_T, 0, s>>>(N, sc, z, pv);
    k_pre_cg_swap<<<1, 32, 0, s>>>(sc);
    BS_CUDA(cudaMemcpyAsync(h, sc + 6, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    bool done = true;
    for (int d = 0; d < 3; ++d) done = done && !(h[d] > 1e-30 * hb[d]);
    if (done) {
      ++it;
      break;
    }
  }
  if (cg_iterations) *cg_iterations = it;
  c.stats.kernel_launches += 4 + 12 * it;
  // ---- normals, M n_hat, l2, area
  k_pre_normalise<<<gN, PRE_T, 0, s>>>(N, x, d_nhat);
  mult(1, d_nhat, d_Mnhat);
  dot3(d_nhat, d_Mnhat, sc + 16);
  for (int k = 0; k < 4; ++k) k_pre_sum1<<<1, PRE_T, 0, s>>>(ncell, carea + (size_t)k * ncell, sc + 20 + k);
  double hs[8];
  BS_CUDA(cudaMemcpyAsync(hs, sc + 16, 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
  BS_CUDA(cudaStreamSynchronize(s));
  if (h_l2) *h_l2 = hs[0] + hs[1] + hs[2];
  if (h_area) *h_area = hs[4];
  const double com[3] = {hs[5] / hs[4], hs[6] / hs[4], hs[7] / hs[4]};  // surface centroid (ref: 2540-2545)
  double pole[3] = {0, 0, 0};                                          // BS_POLE_ORIGIN
  for (int d = 0; d < 3; ++d) {
    if (pole_kind == BS_POLE_POINT) pole[d] = pole_in ? pole_in[d] : 0.0;
    if (pole_kind == BS_POLE_BARICENTER) pole[d] = com[d];
    if (h_center_of_mass) h_center_of_mass[d] = com[d];
    if (h_pole_used) h_pole_used[d] = pole[d];
  }
  // ---- rigid modes about the pole and duals
  if (d_Nr && d_Nrd) {
    k_pre_rigid<<<gN, PRE_T, 0, s>>>(N, c.d_support.p, pole[0], pole[1], pole[2], d_Nr);
    mult(6, d_Nr, d_Nrd);
    BS_CUDA(cudaGetLastError());
    BS_CUDA(cudaStreamSynchronize(s));
  }
  c.stats.kernel_launches += 11;
}

}  // namespace bs
