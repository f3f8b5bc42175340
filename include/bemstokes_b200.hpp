// bemstokes_b200.hpp — header-only C++17 host mirror of the reference's hot-path interface over the C-ABI
// (include/bemstokes_b200.h).  Same names, argument meaning and error behaviour as the reference classes, so a
// test written against BEMStokes::BEMProblem<3> reads the same here:
//
//   StokesKernel<3>, FreeSurfaceStokesKernel<3>, NoSlipWallStokesKernel<3>   include/kernel.h, free_surface_kernel.h,
//                                                                            no_slip_wall_kernel.h
//   DeviceMatrix            anything with vmult(dst, src) + operator()(i,j)  TrilinosWrappers::SparseMatrix as used at
//                                                                            the ~45 vmult sites; include/operator.h:22-66
//   DirectPreconditioner    set_up / initialize / vmult                      include/direct_preconditioner.h:27-51
//   SolverControl           max_steps / tolerance / last_step / last_value   deal.II SolverControl
//   BEMProblem<3>           public members + assemble_stokes_system, solve_system, dirichlet_to_neumann_operator,
//                           tangential_projector_body, evaluate_stokes_bie   include/bem_stokes.h:106-660
//
// Errors: the reference throws deal.II exceptions (caught in source/main.cc:48-71); here every non-zero ABI status
// becomes a bemstokes_b200::Error carrying bs_last_error().  There is no CPU fallback.
#pragma once
#include "bemstokes_b200.h"

#include <array>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace bemstokes_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error("libbemstokes_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
  if (rc != 0) throw Error(rc, bs_last_error());
}

using Tensor1 = std::array<double, 3>;
using Tensor2 = std::array<std::array<double, 3>, 3>;
using Tensor3 = std::array<std::array<std::array<double, 3>, 3>, 3>;

// ---------------------------------------------------------------------------------------------------------------
// Green kernels — point evaluation through the same device functions the assembly inlines
// ---------------------------------------------------------------------------------------------------------------
template <int dim>
class StokesKernel {
  static_assert(dim == 3, "the B200 hot path is 3-D");

public:
  explicit StokesKernel(const double eps = 0., int device = 0) : epsilon(eps), device_(device) {}
  virtual ~StokesKernel() = default;
  void set_wall_orientation(const unsigned int o) { wall_orientation = o; }
  Tensor2 value_tens(const Tensor1 &p) const { return eval_G(p, p); }
  Tensor3 value_tens2(const Tensor1 &p) const { return eval_W(p, p); }
  double epsilon;
  unsigned int wall_orientation = 1;

protected:
  virtual int type() const { return BS_KERNEL_FREE; }
  Tensor2 eval_G(const Tensor1 &p, const Tensor1 &q) const {
    double G[9];
    check(bs_kernel_eval(device_, type(), epsilon, (int)wall_orientation, 1, p.data(), q.data(), G, nullptr));
    Tensor2 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r[i][j] = G[3 * i + j];
    return r;
  }
  Tensor3 eval_W(const Tensor1 &p, const Tensor1 &q) const {
    double W[27];
    check(bs_kernel_eval(device_, type(), epsilon, (int)wall_orientation, 1, p.data(), q.data(), nullptr, W));
    Tensor3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) r[i][j][k] = W[9 * i + 3 * j + k];
    return r;
  }
  int device_;
};

template <int dim>
class FreeSurfaceStokesKernel : public StokesKernel<dim> {
public:
  using StokesKernel<dim>::StokesKernel;
  Tensor2 value_tens_image(const Tensor1 &p, const Tensor1 &p_image) const { return this->eval_G(p, p_image); }
  Tensor3 value_tens_image2(const Tensor1 &p, const Tensor1 &p_image) const { return this->eval_W(p, p_image); }

protected:
  int type() const override { return BS_KERNEL_FREE_SURFACE; }
};

template <int dim>
class NoSlipWallStokesKernel : public FreeSurfaceStokesKernel<dim> {
public:
  using FreeSurfaceStokesKernel<dim>::FreeSurfaceStokesKernel;

protected:
  int type() const override { return BS_KERNEL_NO_SLIP; }
};

// ---------------------------------------------------------------------------------------------------------------
struct SolverControl {
  SolverControl(unsigned int max_steps_ = 1000, double tolerance_ = 1e-10) : max_steps(max_steps_), tolerance(tolerance_) {}
  unsigned int max_steps;
  double tolerance;
  unsigned int last_step() const { return last_step_; }
  double last_value() const { return last_value_; }
  unsigned int last_step_ = 0;
  double last_value_ = 0;
};

using Vector = std::vector<double>;

class DeviceMatrix {
public:
  DeviceMatrix(bs_context *ctx = nullptr, int which = BS_MAT_V) : ctx(ctx), which(which) {}
  unsigned int m() const {
    int r = 0;
    check(bs_matrix_size(ctx, which, &r, nullptr));
    return (unsigned int)r;
  }
  void vmult(Vector &dst, const Vector &src) const {
    dst.resize(src.size());
    check(bs_vmult(ctx, which, src.data(), dst.data()));
  }
  double operator()(const unsigned int i, const unsigned int j) const {
    int r = (int)i, c = (int)j;
    double v;
    check(bs_get_entries(ctx, which, 1, &r, &c, &v));
    return v;
  }
  bs_context *ctx;
  int which;
};

class DirectPreconditioner {
public:
  void set_up(SolverControl &, int /*AdditionalData*/ = 0) {}
  void initialize(const DeviceMatrix &matrix) {
    m = matrix;
    check(bs_precond_setup(m.ctx, m.which, BS_PREC_DIRECT, 0));
  }
  void vmult(Vector &dst, const Vector &src) const {
    dst.resize(src.size());
    check(bs_precond_vmult(m.ctx, src.data(), dst.data()));
  }
  DeviceMatrix m;
};

// ---------------------------------------------------------------------------------------------------------------
// Quad meshes: GMSH-v1 .msh and UCD .inp (the formats of read_input_mesh_file, source/bem_stokes.cc:496-523) and
// the synthetic cube-sphere.  Cells in deal.II lexicographic vertex order.
// ---------------------------------------------------------------------------------------------------------------
struct QuadMesh {
  std::vector<double> nodes;  // [N][3]
  std::vector<int> conn;      // [ncell][4]
  int degree = 1;
  int n_nodes() const { return (int)nodes.size() / 3; }
  int n_cells() const { return (int)conn.size() / 4; }
};

inline QuadMesh read_mesh(const std::string &path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open mesh file " + path);
  QuadMesh m;
  std::map<int, int> ids;
  auto add_quad = [&](int a, int b, int c, int d) {
    for (int v : {a, b, d, c}) m.conn.push_back(ids.at(v));  // ccw file order -> lexicographic
  };
  if (path.size() > 4 && path.substr(path.size() - 4) == ".inp") {
    int nv, nc, z0, z1, z2;
    f >> nv >> nc >> z0 >> z1 >> z2;
    for (int k = 0; k < nv; ++k) {
      int id;
      double x, y, z;
      f >> id >> x >> y >> z;
      ids[id] = k;
      m.nodes.insert(m.nodes.end(), {x, y, z});
    }
    for (int k = 0; k < nc; ++k) {
      int id, mat, a, b, c, d;
      std::string type;
      f >> id >> mat >> type;
      if (type == "quad") {
        f >> a >> b >> c >> d;
        add_quad(a, b, c, d);
      } else {
        std::string rest;
        std::getline(f, rest);
      }
    }
  } else {
    std::string tok;
    int nv = 0, ne = 0;
    while (f >> tok && tok != "$NOD") {}
    f >> nv;
    for (int k = 0; k < nv; ++k) {
      int id;
      double x, y, z;
      f >> id >> x >> y >> z;
      ids[id] = k;
      m.nodes.insert(m.nodes.end(), {x, y, z});
    }
    while (f >> tok && tok != "$ELM") {}
    f >> ne;
    for (int k = 0; k < ne; ++k) {
      int id, type, rp, re, nn;
      f >> id >> type >> rp >> re >> nn;
      std::vector<int> v(nn);
      for (auto &x : v) f >> x;
      if (type == 3) add_quad(v[0], v[1], v[2], v[3]);
    }
  }
  return m;
}

// ---------------------------------------------------------------------------------------------------------------
// BEMProblem<3> — the hot-path members of BEMStokes::BEMProblem<3> (all public, as in the reference)
// ---------------------------------------------------------------------------------------------------------------
template <int dim>
class BEMProblem {
  static_assert(dim == 3, "the B200 hot path is 3-D");

public:
  explicit BEMProblem(int device = 0) : device(device) {}
  ~BEMProblem() {
    if (ctx) bs_destroy(ctx);
  }
  BEMProblem(const BEMProblem &) = delete;

  // ---- parameters (names of declare_parameters, source/bem_stokes.cc:207-476) ----
  unsigned int quadrature_order = 8;              // "Internal Quadrature", gauss
  std::string singular_quadrature_type = "Mixed";  // Mixed | Duffy | Telles
  unsigned int singular_quadrature_order = 5;
  bool reflect_kernel = false, no_slip_kernel = false;
  std::array<double, 3> wall_spans_0{{10., 0., 10.}}, wall_position_0{{0., 0., 0.}};
  std::string grid_type = "Real";                  // Real | ImposedForce | ImposedVelocity
  unsigned int imposed_component = 1;
  double assemble_scaling = 1.;
  bool use_internal_alpha = false, monolithic_bool = true, solve_directly = true;
  std::string preconditioner_type = "Direct";      // Direct | ILU | AMG | Jacobi | None
  bool bandwith_preconditioner = false;
  unsigned int bandwith = 100, gmres_restart = 100, num_rigid = 6;
  SolverControl solver_control;
  bool keep_VK = true;

  // ---- state ----
  int device;
  bs_context *ctx = nullptr;
  QuadMesh mesh;
  unsigned int n_dofs = 0, N = 0, kernel_wall_orientation = 1;
  Vector normal_vector_pure, M_normal_vector_pure, V_x_normals_body, monolithic_rhs, monolithic_solution, stokes_forces,
      shape_velocities;
  std::vector<Vector> N_rigid, N_rigid_dual;
  Vector rigid_velocities, rigid_total_forces;
  double l2normGamma_pure = 0, surface = 0;
  DeviceMatrix V_matrix, K_matrix, monolithic_system_matrix;
  DirectPreconditioner direct_trilinos_preconditioner;
  bool reassemble_preconditoner = false;
  std::ostream *pcout = &std::cout;

  // read_domain + reinit: mesh in, context + geometry + quadrature on the device
  void read_domain(const QuadMesh &m) { mesh = m; }
  void reinit() {
    if (ctx) check(bs_destroy(ctx));
    ctx = nullptr;
    check(bs_create(&ctx, device, mesh.degree, mesh.degree));
    N = mesh.n_nodes();
    n_dofs = 3 * N;
    euler_vec.assign(n_dofs, 0.);
    for (unsigned int i = 0; i < N; ++i)
      for (int d = 0; d < 3; ++d) euler_vec[i + d * N] = mesh.nodes[3 * i + d];
    check(bs_set_geometry(ctx, (int)N, euler_vec.data(), mesh.n_cells(), mesh.conn.data(), (int)N, mesh.conn.data(), nullptr));
    check(bs_set_quadrature(ctx, (int)quadrature_order, nullptr, nullptr));
    const int kind = singular_quadrature_type == "Duffy" ? BS_SING_DUFFY : singular_quadrature_type == "Telles" ? BS_SING_TELLES : BS_SING_MIXED;
    check(bs_set_singular_quadrature(ctx, kind, (int)singular_quadrature_order));
    set_kernel();
    V_matrix = DeviceMatrix(ctx, BS_MAT_V);
    K_matrix = DeviceMatrix(ctx, BS_MAT_K);
    monolithic_system_matrix = DeviceMatrix(ctx, BS_MAT_A);
    shape_velocities.assign(n_dofs, 0.);
  }

  // pre-pass (mass matrix, rigid modes, L2 normals) on the device: bem_stokes.cc:2440-2788, 3922-4011
  void compute_center_of_mass_and_rigid_modes(unsigned int /*frame*/ = 0) {
    normal_vector_pure.assign(n_dofs, 0.);
    M_normal_vector_pure.assign(n_dofs, 0.);
    std::vector<double> nr(6 * (size_t)n_dofs), nd(6 * (size_t)n_dofs);
    check(bs_prepass(ctx, BS_POLE_ORIGIN, nullptr, normal_vector_pure.data(), M_normal_vector_pure.data(), &l2normGamma_pure,
                     nr.data(), nd.data(), &surface, nullptr, nullptr, nullptr, nullptr));
    N_rigid.assign(6, Vector());
    N_rigid_dual.assign(6, Vector());
    for (int r = 0; r < 6; ++r) {
      N_rigid[r].assign(nr.begin() + (size_t)r * n_dofs, nr.begin() + (size_t)(r + 1) * n_dofs);
      N_rigid_dual[r].assign(nd.begin() + (size_t)r * n_dofs, nd.begin() + (size_t)(r + 1) * n_dofs);
    }
    *pcout << "The Mass (Surface) of the entire system is : " << surface << std::endl;
  }
  void compute_normal_vector() {}  // computed together with the mass matrix above

  // ref: BEMProblem::assemble_stokes_system (bem_stokes.cc:2840-3435), same log lines
  void assemble_stokes_system(bool correction_on_V = true) {
    set_kernel();
    check(bs_assemble_VK(ctx));
    V_x_normals_body.assign(n_dofs, 0.);
    if (correction_on_V)
      check(bs_correct_V(ctx, normal_vector_pure.data(), M_normal_vector_pure.data(), l2normGamma_pure, V_x_normals_body.data()));
    else
      V_matrix.vmult(V_x_normals_body, normal_vector_pure);
    *pcout << "Check on the V operator Norm (should be zero) pure: " << linfty(V_x_normals_body) << std::endl;
    Vector post;
    V_matrix.vmult(post, normal_vector_pure);
    *pcout << "Check on the V operator Norm post (should be one) pure: " << dot(post, normal_vector_pure) / N << std::endl;
    check(bs_correct_K(ctx, use_internal_alpha ? 1 : 0));
    for (unsigned int k = 0; k < 3; ++k) {
      Vector e(n_dofs, 0.), ke;
      for (unsigned int i = 0; i < N; ++i) e[i + k * N] = 1.;
      K_matrix.vmult(ke, e);
      *pcout << "check with versor vector : " << k << " l_infty : " << linfty(ke) << std::endl;
    }
    if (monolithic_bool) {
      std::vector<double> nr, nd;
      for (unsigned int r = 0; r < num_rigid; ++r) {
        nr.insert(nr.end(), N_rigid[r].begin(), N_rigid[r].end());
        nd.insert(nd.end(), N_rigid_dual[r].begin(), N_rigid_dual[r].end());
      }
      monolithic_rhs.assign(n_dofs + num_rigid, 0.);
      const int gt = grid_type == "ImposedForce" ? BS_GRID_IMPOSED_FORCE : grid_type == "ImposedVelocity" ? BS_GRID_IMPOSED_VELOCITY : BS_GRID_REAL;
      check(bs_build_monolithic(ctx, nullptr, (int)num_rigid, nr.data(), nd.data(), normal_vector_pure.data(),
                                M_normal_vector_pure.data(), l2normGamma_pure, gt, (int)imposed_component, assemble_scaling,
                                shape_velocities.data(), keep_VK ? 1 : 0, monolithic_rhs.data()));
      if (monolithic_solution.size() != monolithic_rhs.size()) monolithic_solution.assign(monolithic_rhs.size(), 0.);
    }
  }

  void tangential_projector_body(const Vector &input_vel, Vector &output_vel) {
    output_vel.resize(input_vel.size());
    check(bs_tangential_projector(ctx, input_vel.data(), output_vel.data()));
  }

  // ref: BEMProblem::solve_system (bem_stokes.cc:4158-4508)
  void solve_system(bool monolithic_booly = true) {
    if (!monolithic_booly) throw Error(BS_ERR_UNSUPPORTED, "use dirichlet_to_neumann_operator for the DN route");
    if (solve_directly) {
      check(bs_direct_solve(ctx, BS_MAT_A, monolithic_rhs.data(), monolithic_solution.data()));
      solver_control.last_step_ = 1;
    } else {
      *pcout << "preconditioner_type = " << preconditioner_type << std::endl;
      if (preconditioner_type == "Jacobi") check(bs_precond_setup(ctx, BS_MAT_A, BS_PREC_JACOBI, 0));
      else if (preconditioner_type == "None") check(bs_precond_setup(ctx, BS_MAT_A, BS_PREC_NONE, 0));
      else if (preconditioner_type == "Direct") {
        if (direct_trilinos_preconditioner.m.ctx == nullptr || reassemble_preconditoner) {
          direct_trilinos_preconditioner.initialize(monolithic_system_matrix);
          reassemble_preconditoner = false;
        }
      } else  // ILU / AMG on the dense pattern = exact LU (optionally of the band copy, bem_stokes.cc:3437-3505)
        check(bs_precond_setup(ctx, BS_MAT_A, bandwith_preconditioner ? BS_PREC_BAND : BS_PREC_DIRECT, (int)bandwith));
      int its = 0;
      double res = 0;
      const int rc = bs_gmres(ctx, BS_MAT_A, monolithic_rhs.data(), monolithic_solution.data(), solver_control.tolerance,
                              (int)solver_control.max_steps, (int)gmres_restart, &its, &res);
      solver_control.last_step_ = (unsigned int)its;
      solver_control.last_value_ = res;
      check(rc);
      *pcout << "   Iterations needed to solve monolithic:         " << its << std::endl;
      if (its > 100) reassemble_preconditoner = true;
    }
    Vector ax;
    monolithic_system_matrix.vmult(ax, monolithic_solution);
    double linf = 0, l2 = 0;
    for (size_t i = 0; i < ax.size(); ++i) {
      const double d = ax[i] - monolithic_rhs[i];
      linf = std::max(linf, std::fabs(d));
      l2 += d * d;
    }
    *pcout << "FINAL CHECK 0 " << linf << " : " << std::sqrt(l2) << std::endl;
    stokes_forces.assign(monolithic_solution.begin(), monolithic_solution.begin() + n_dofs);
    rigid_velocities.assign(num_rigid, 0.);
    rigid_total_forces.assign(num_rigid, 0.);
    for (unsigned int r = 0; r < num_rigid; ++r) {
      rigid_velocities[r] = monolithic_solution[n_dofs + r] * assemble_scaling;
      rigid_total_forces[r] = dot(stokes_forces, N_rigid_dual[r]);
    }
  }

  // DN(u) = P V^{-1} (P K P u)   (bem_stokes.cc:4072-4129)
  void dirichlet_to_neumann_operator(const Vector &input_vel, Vector &output_force) {
    Vector v1, v2, f(n_dofs, 0.);
    tangential_projector_body(input_vel, v1);
    K_matrix.vmult(v2, v1);
    tangential_projector_body(v2, v1);
    if (solve_directly) check(bs_direct_solve(ctx, BS_MAT_V, v1.data(), f.data()));
    else {
      check(bs_precond_setup(ctx, BS_MAT_V, BS_PREC_NONE, 0));
      int its;
      double res;
      check(bs_gmres(ctx, BS_MAT_V, v1.data(), f.data(), solver_control.tolerance, (int)solver_control.max_steps, (int)gmres_restart, &its, &res));
    }
    tangential_projector_body(f, output_force);
  }

  // ref: evaluate_stokes_bie (bem_stokes.cc:5366-5451); val_points [P][3], result component-major
  void evaluate_stokes_bie(const std::vector<Tensor1> &val_points, const Vector &vel, const Vector &forces, Vector &val_velocities) {
    val_velocities.assign(3 * val_points.size(), 0.);
    set_kernel();
    check(bs_evaluate_bie(ctx, (int)val_points.size(), val_points[0].data(), vel.data(), forces.data(), val_velocities.data(), 0));
  }

  static double linfty(const Vector &v) {
    double m = 0;
    for (double x : v) m = std::max(m, std::fabs(x));
    return m;
  }
  static double dot(const Vector &a, const Vector &b) {
    double s = 0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
  }

private:
  Vector euler_vec;
  void set_kernel() {
    kernel_wall_orientation = 1;  // last axis with wall_spans[0][axis]==0 (bem_stokes.cc:2861-2866)
    for (unsigned int i = 0; i < 3; ++i)
      if (wall_spans_0[i] == 0) kernel_wall_orientation = i;
    const int kt = reflect_kernel ? BS_KERNEL_FREE_SURFACE : (no_slip_kernel ? BS_KERNEL_NO_SLIP : BS_KERNEL_FREE);
    check(bs_set_kernel(ctx, kt, 0., (int)kernel_wall_orientation, wall_position_0.data()));
  }
};

}  // namespace bemstokes_b200
