/* CPU ORACLE (C port) — TEST / BASELINE INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * Plain C + OpenMP restatement of the reference's hot loop nest, used (a) as a second, independent check of
 * the NumPy oracle and (b) as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).
 * Parity status: pinned through tests/test_oracle_golden.py (this port is compared entry-by-entry with
 * oracle/bem_oracle.py, which reproduces the reference's golden outputs).
 *
 * It follows, in order:
 *   kernels            source/kernel.cc:61-104, source/free_surface_kernel.cc:19-72,135-209,
 *                      source/no_slip_wall_kernel.cc:23-116,127-199   (full 3x3 G and 3x3x3 W, pow(R,5) as written)
 *   contraction W.n    source/bem_stokes.cc:5071-5083 (compute_singular_kernel)
 *   assembly loop      source/bem_stokes.cc:2871-2998 (cell, node, q; singular rule when the node is in the cell)
 *   vmult              dense row-major GEMV standing in for TrilinosWrappers::SparseMatrix::vmult
 *   GMRES              deal.II SolverGMRES semantics: left preconditioning, modified Gram-Schmidt
 *                      with the every-5th-iteration re-orthogonalisation test, Givens rotations
 * It FLATTERS the reference: no per-entry Epetra SumIntoGlobalValues, no CRS index traffic, threads over rows
 * (the reference's assembly is single-threaded per MPI rank).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

static void G_free(const double *p, double eps, double G[3][3]) {
  double R = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + eps;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double delta = 1.0 * (i == j);
      G[i][j] = (p[i] * p[j] / (R * R * R) + delta / R) / (8 * PI);
    }
}
static void W_free(const double *p, double eps, double W[3][3][3]) {
  double R = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + eps;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) W[i][j][k] = -3 * p[i] * p[j] * p[k] / (pow(R, 5)) / (4 * PI);
}
static void G_fs(const double *p, const double *q, int o, double eps, double G[3][3]) {
  double a[3][3], b[3][3];
  G_free(p, eps, a);
  G_free(q, eps, b);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[i][j] = (i == o) ? a[i][j] - b[i][j] : a[i][j] + b[i][j];
}
static void W_fs(const double *p, const double *q, int o, double eps, double W[3][3][3]) {
  double a[3][3][3], b[3][3][3];
  W_free(p, eps, a);
  W_free(q, eps, b);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) W[i][j][k] = (i == o) ? a[i][j][k] - b[i][j][k] : a[i][j][k] + b[i][j][k];
}
static void G_ns(const double *p, const double *pi_, int o, double eps, double G[3][3]) {
  double h0 = 0.5 * (pi_[o] - p[o]);
  double R = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + eps;
  double Ri = sqrt(pi_[0] * pi_[0] + pi_[1] * pi_[1] + pi_[2] * pi_[2]) + eps;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double d = 1.0 * (i == j), di1 = 1.0 * (i == o), dj1 = 1.0 * (j == o);
      double base = (p[i] * p[j] / (R * R * R) + d / R) - (pi_[i] * pi_[j] / (Ri * Ri * Ri) + d / Ri);
      double t5 = (-3 * pi_[i] * pi_[j] / (Ri * Ri * Ri * Ri * Ri) + d / (Ri * Ri * Ri));
      double t2 = 2. * h0 * h0 * t5;
      double t3 = 2. * h0 * (pi_[o] * t5 + ((di1 * pi_[j] - dj1 * pi_[i]) / (Ri * Ri * Ri)));
      G[i][j] = ((i == o) ? base - t2 + t3 : base + t2 - t3) / (8 * PI);
    }
}
static void W_ns(const double *p, const double *pi_, int o, double eps, double W[3][3][3]) {
  double h0 = 0.5 * (pi_[o] - p[o]);
  double R = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + eps;
  double Ri = sqrt(pi_[0] * pi_[0] + pi_[1] * pi_[1] + pi_[2] * pi_[2]) + eps;
  for (int i = 0; i < 3; ++i) {
    double di1 = 1.0 * (i == o);
    for (int j = 0; j < 3; ++j) {
      double dij = 1.0 * (i == j);
      for (int k = 0; k < 3; ++k) {
        double djk = 1.0 * (k == j), dik = 1.0 * (i == k);
        double w = -1. * p[i] * p[j] * p[k] / pow(R, 5);
        w -= -1. * pi_[i] * pi_[j] * pi_[k] / pow(Ri, 5);
        double brk = -(dik * pi_[j] + dij * pi_[k] * djk * pi_[i]) / pow(Ri, 5) + 5. * (pi_[i] * pi_[j] * pi_[k]) / pow(Ri, 7);
        double t2 = 2 * h0 * h0 * brk;
        double t3 = (-2 * h0) * (pi_[o] * brk + (djk * pi_[i] * pi_[o] - di1 * pi_[j] * pi_[k]) / pow(Ri, 5));
        if (i == o) w = w - t2 - t3;
        else w = w + t2 + t3;
        W[i][j][k] = w * 3 / (4 * PI);
      }
    }
  }
}

typedef struct {
  int type;      /* 0 free, 1 free surface, 2 no slip */
  double eps;
  int o;
  double wall_pos;
} port_kernel;

/* FEValues::reinit: y, n, JxW at the nq points of a rule, from the map-shape tables tab[q][nam][3] */
static void fe_cell(const double *X /*[nam][3]*/, int nam, int nq, const double *tab, const double *w, double *y, double *n,
                    double *jxw) {
  for (int q = 0; q < nq; ++q) {
    double yy[3] = {0, 0, 0}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
    for (int a = 0; a < nam; ++a) {
      const double *t = tab + ((size_t)q * nam + a) * 3;
      for (int d = 0; d < 3; ++d) {
        yy[d] += t[0] * X[3 * a + d];
        t1[d] += t[1] * X[3 * a + d];
        t2[d] += t[2] * X[3 * a + d];
      }
    }
    double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2], nz = t1[0] * t2[1] - t1[1] * t2[0];
    double J = sqrt(nx * nx + ny * ny + nz * nz);
    y[3 * q] = yy[0]; y[3 * q + 1] = yy[1]; y[3 * q + 2] = yy[2];
    n[3 * q] = nx / J; n[3 * q + 1] = ny / J; n[3 * q + 2] = nz / J;
    jxw[q] = w[q] * J;
  }
}

/* one (node, cell) pair: local_single_layer / local_double_layer (3 x 3*na), ref bem_stokes.cc:2915-2951 */
static void pair_block(const double *x, const port_kernel *kp, int nq, const double *y, const double *n, const double *jxw,
                       const double *phi /*[nq][na]*/, int na, double *lv /*[3][3][na]*/, double *lk) {
  memset(lv, 0, sizeof(double) * 9 * na);
  memset(lk, 0, sizeof(double) * 9 * na);
  for (int q = 0; q < nq; ++q) {
    double R[3], Rim[3], xim[3] = {x[0], x[1], x[2]};
    xim[kp->o] -= 2 * (x[kp->o] - kp->wall_pos);
    for (int d = 0; d < 3; ++d) {
      R[d] = y[3 * q + d] - x[d];
      Rim[d] = y[3 * q + d] - xim[d];
    }
    double G[3][3], W[3][3][3], S[3][3];
    if (kp->type == 1) { G_fs(R, Rim, kp->o, kp->eps, G); W_fs(R, Rim, kp->o, kp->eps, W); }
    else if (kp->type == 2) { G_ns(R, Rim, kp->o, kp->eps, G); W_ns(R, Rim, kp->o, kp->eps, W); }
    else { G_free(R, kp->eps, G); W_free(R, kp->eps, W); }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        S[i][j] = 0;
        for (int k = 0; k < 3; ++k) S[i][j] += W[i][j][k] * n[3 * q + k];
      }
    for (int idim = 0; idim < 3; ++idim)
      for (int a = 0; a < na; ++a)
        for (int jdim = 0; jdim < 3; ++jdim) {
          lv[(idim * 3 + jdim) * na + a] += G[idim][jdim] * phi[(size_t)q * na + a] * jxw[q];
          lk[(idim * 3 + jdim) * na + a] -= S[idim][jdim] * phi[(size_t)q * na + a] * jxw[q];
        }
  }
}

/* V, K: (3*nrows) x (3*N), rows component-major over the row subset, columns component-major (j + b*N).
 * returns the number of (node, q-point) pair evaluations */
long long port_assemble(int N, int ncell, int na, int nam, const double *support, const int *conn, const double *map_nodes,
                        const int *conn_map, int nq, const double *phi_reg, const double *tab_reg, const double *w_reg,
                        const int *sing_nq, const int *sing_off, const double *sing_phi, const double *sing_tab,
                        const double *sing_w, int type, double eps, int o, double wall_pos, int row_begin, int row_end,
                        double *V, double *K, int nthreads) {
  port_kernel kp = {type, eps, o, wall_pos};
  const int nr = row_end - row_begin;
  const size_t ncols = (size_t)3 * N;
  long long pairs = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int maxs = nq;
  for (int a = 0; a < na; ++a) if (sing_nq[a] > maxs) maxs = sing_nq[a];
#pragma omp parallel reduction(+ : pairs)
  {
    double *y = malloc(sizeof(double) * 3 * maxs), *n = malloc(sizeof(double) * 3 * maxs), *jxw = malloc(sizeof(double) * maxs);
    double *ys = malloc(sizeof(double) * 3 * maxs), *ns = malloc(sizeof(double) * 3 * maxs), *js = malloc(sizeof(double) * maxs);
    double *lv = malloc(sizeof(double) * 9 * na), *lk = malloc(sizeof(double) * 9 * na);
    double X[9 * 3];
    int tid = 0, nt = 1;
#ifdef _OPENMP
    tid = omp_get_thread_num();
    nt = omp_get_num_threads();
#endif
    const int chunk = (nr + nt - 1) / nt;
    const int i0 = row_begin + tid * chunk, i1 = (i0 + chunk < row_end) ? i0 + chunk : row_end;
    for (int c = 0; c < ncell && i0 < i1; ++c) {
      for (int a = 0; a < nam; ++a)
        for (int d = 0; d < 3; ++d) X[3 * a + d] = map_nodes[(size_t)3 * conn_map[(size_t)c * nam + a] + d];
      fe_cell(X, nam, nq, tab_reg, w_reg, y, n, jxw);
      for (int i = i0; i < i1; ++i) {
        int sing = -1;
        for (int a = 0; a < na; ++a)
          if (conn[(size_t)c * na + a] == i) { sing = a; break; }
        if (sing >= 0) {
          const int m = sing_nq[sing], off = sing_off[sing];
          fe_cell(X, nam, m, sing_tab + (size_t)off * nam * 3, sing_w + off, ys, ns, js);
          pair_block(support + 3 * (size_t)i, &kp, m, ys, ns, js, sing_phi + (size_t)off * na, na, lv, lk);
          pairs += m;
        } else {
          pair_block(support + 3 * (size_t)i, &kp, nq, y, n, jxw, phi_reg, na, lv, lk);
          pairs += nq;
        }
        const int k = i - row_begin;
        for (int idim = 0; idim < 3; ++idim)
          for (int jdim = 0; jdim < 3; ++jdim)
            for (int a = 0; a < na; ++a) {
              const size_t row = (size_t)k + (size_t)idim * nr, col = (size_t)conn[(size_t)c * na + a] + (size_t)jdim * N;
              V[row * ncols + col] += lv[(idim * 3 + jdim) * na + a];
              K[row * ncols + col] += lk[(idim * 3 + jdim) * na + a];
            }
      }
    }
    free(y); free(n); free(jxw); free(ys); free(ns); free(js); free(lv); free(lk);
  }
  return pairs;
}

void port_gemv(const double *A, long long rows, long long cols, const double *x, double *y, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (long long r = 0; r < rows; ++r) {
    const double *a = A + r * cols;
    double s = 0;
    for (long long c = 0; c < cols; ++c) s += a[c] * x[c];
    y[r] = s;
  }
}

static double dotp(const double *a, const double *b, long long n) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long long i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
static void axpyp(double a, const double *x, double *y, long long n) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) y[i] += a * x[i];
}

/* deal.II SolverGMRES semantics; diag_inv == NULL -> identity preconditioner, else Jacobi. returns iterations */
int port_gmres(const double *A, long long n, const double *b, double *x, const double *diag_inv, double tol, int max_steps,
               int max_tmp, double *final_res, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int m = max_tmp - 2;
  double *Vb = malloc(sizeof(double) * (size_t)(m + 1) * n), *w = malloc(sizeof(double) * n);
  double *H = calloc((size_t)(m + 1) * m, sizeof(double)), *gamma = malloc(sizeof(double) * (m + 1));
  double *ci = malloc(sizeof(double) * m), *si = malloc(sizeof(double) * m), *h = malloc(sizeof(double) * (m + 2)),
         *yk = malloc(sizeof(double) * m);
  int its = 0, converged = 0;
  double rho = 0;
  for (;;) {
    port_gemv(A, n, n, x, w, 0);
    for (long long i = 0; i < n; ++i) w[i] = (b[i] - w[i]) * (diag_inv ? diag_inv[i] : 1.0);
    rho = sqrt(dotp(w, w, n));
    if (rho <= tol) { converged = 1; break; }
    if (its >= max_steps) break;
    for (long long i = 0; i < n; ++i) Vb[i] = w[i] / rho;
    memset(gamma, 0, sizeof(double) * (m + 1));
    gamma[0] = rho;
    int dim = 0, stop = 0;
    for (int inner = 0; inner < m && !stop; ++inner) {
      ++its;
      double *vv = Vb + (size_t)(inner + 1) * n;
      port_gemv(A, n, n, Vb + (size_t)inner * n, vv, 0);
      if (diag_inv) for (long long i = 0; i < n; ++i) vv[i] *= diag_inv[i];
      dim = inner + 1;
      double norm_start = 0;
      if (its % 5 == 0) norm_start = sqrt(dotp(vv, vv, n));
      for (int i = 0; i < dim; ++i) {
        h[i] = dotp(vv, Vb + (size_t)i * n, n);
        axpyp(-h[i], Vb + (size_t)i * n, vv, n);
      }
      int reorth = 0;
      if (its % 5 == 0) {
        double nv = sqrt(dotp(vv, vv, n));
        if (!(nv > 10. * norm_start * sqrt(2.220446049250313e-16))) reorth = 1;
      }
      if (reorth)
        for (int i = 0; i < dim; ++i) {
          double ht = dotp(vv, Vb + (size_t)i * n, n);
          h[i] += ht;
          axpyp(-ht, Vb + (size_t)i * n, vv, n);
        }
      double s = sqrt(dotp(vv, vv, n));
      h[dim] = s;
      for (long long i = 0; i < n; ++i) vv[i] /= s;
      for (int i = 0; i < inner; ++i) {
        double t = h[i];
        h[i] = ci[i] * t + si[i] * h[i + 1];
        h[i + 1] = -si[i] * t + ci[i] * h[i + 1];
      }
      double r = hypot(h[inner], h[inner + 1]);
      ci[inner] = h[inner] / r;
      si[inner] = h[inner + 1] / r;
      h[inner] = r;
      gamma[inner + 1] = -si[inner] * gamma[inner];
      gamma[inner] = ci[inner] * gamma[inner];
      for (int i = 0; i < dim; ++i) H[(size_t)i * m + inner] = h[i];
      rho = fabs(gamma[dim]);
      if (rho <= tol) { converged = 1; stop = 1; }
      else if (its >= max_steps) stop = 1;
    }
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int j = i + 1; j < dim; ++j) s -= H[(size_t)i * m + j] * yk[j];
      yk[i] = s / H[(size_t)i * m + i];
    }
    for (int i = 0; i < dim; ++i) axpyp(yk[i], Vb + (size_t)i * n, x, n);
    if (converged || its >= max_steps) break;
  }
  if (final_res) *final_res = rho;
  free(Vb); free(w); free(H); free(gamma); free(ci); free(si); free(h); free(yk);
  return converged ? its : -its;
}

int port_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
