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

#include <algorithm>
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
using Matrix3 = std::array<std::array<double, 3>, 3>;  // FullMatrix<double>(3,3) of the rotation code
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
  bool fused_assembly = false;                     // never store K (bs_assemble_fused): sizes where V and K do not fit together
  std::vector<unsigned char> col_is_K;             // per dof: the unknown is a wall velocity, column -K (index sets of 3194-3245)
  // hanging-node constraints: dof -> (constraining dof, coefficient) (ref: the AffineConstraints of 2970-2995, 3156-3183)
  std::map<unsigned int, std::vector<std::pair<unsigned int, double>>> constraints;
  bool solve_with_torque = false;                  // "Impose a torque on the flagellum" (216, 3252-3256, 3340-3352)
  Vector N_flagellum_torque, N_flagellum_torque_dual;
  double torque_rhs = -2., flagellum_omega = 0.;
  // frame loop (bem_stokes.cc:215-216, 222-229, 282-300, 327-329)
  unsigned int n_frames = 120, delta_frame = 1;
  bool bool_rot = true, bool_dipl = false, bool_dipl_x = false, bool_dipl_y = false, bool_dipl_z = false;
  std::string input_grid_path = "../debug_grids/", input_grid_base_name = "sphere_mesh_3d_", input_grid_format = "msh";
  std::string res_strategy = "Forward", output_dir = ".";
  double time_step = 0.1;
  std::array<double, 4> initial_quaternion{{1., 0., 0., 0.}};

  // ---- state ----
  int device;
  bs_context *ctx = nullptr;
  QuadMesh mesh;
  unsigned int n_dofs = 0, N = 0, kernel_wall_orientation = 1;
  Vector normal_vector_pure, M_normal_vector_pure, V_x_normals_body, monolithic_rhs, monolithic_solution, stokes_forces,
      shape_velocities;
  std::vector<Vector> N_rigid, N_rigid_dual;
  Vector rigid_velocities, rigid_total_forces, baricenter_rigid_velocities;
  std::vector<Vector> DN_N_rigid;   // DN(N_rigid[r]) of the last solve_system(false)
  std::vector<double> final_matrix; // num_rigid x num_rigid, row-major
  Matrix3 rotation_matrix = identity3(), old_rotation_matrix = identity3();
  Vector old_rigid_velocities, old_rigid_displacements_for_sim;
  Vector next_euler_vec, rigid_puntual_velocities, rigid_puntual_translation_velocities, next_rigid_puntual_displacements,
      rigid_puntual_displacements, rigid_displacements_for_sim, total_velocities;
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
    {  // constraints and the torque unknown take effect in the calls below
      std::vector<int> dof, ptr(1, 0), cols;
      std::vector<double> coefs;
      for (const auto &kv : constraints) {
        dof.push_back((int)kv.first);
        for (const auto &e : kv.second) {
          cols.push_back((int)e.first);
          coefs.push_back(e.second);
        }
        ptr.push_back((int)cols.size());
      }
      check(bs_set_constraints(ctx, (int)dof.size(), dof.data(), ptr.data(), cols.data(), coefs.data()));
      if (solve_with_torque) check(bs_set_torque_mode(ctx, N_flagellum_torque.data(), N_flagellum_torque_dual.data(), torque_rhs));
      else check(bs_set_torque_mode(ctx, nullptr, nullptr, 0.));
    }
    const unsigned char *flags = col_is_K.size() == n_dofs ? col_is_K.data() : nullptr;
    if (fused_assembly) {
      std::vector<double> nr0;
      for (unsigned int r = 0; r < num_rigid; ++r) nr0.insert(nr0.end(), N_rigid[r].begin(), N_rigid[r].end());
      check(bs_set_column_flags(ctx, flags));
      check(bs_assemble_fused(ctx, (int)num_rigid, nr0.data(), normal_vector_pure.data(), M_normal_vector_pure.data(), l2normGamma_pure,
                              grid_type == "Real" ? shape_velocities.data() : nullptr));
    } else {
      check(bs_assemble_VK(ctx));
    }
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
    for (unsigned int k = 0; k < 3 && !fused_assembly; ++k) {
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
      monolithic_rhs.assign(n_dofs + num_rigid + (solve_with_torque ? 1 : 0), 0.);
      const int gt = grid_type == "ImposedForce" ? BS_GRID_IMPOSED_FORCE : grid_type == "ImposedVelocity" ? BS_GRID_IMPOSED_VELOCITY : BS_GRID_REAL;
      check(bs_build_monolithic(ctx, flags, (int)num_rigid, nr.data(), nd.data(), normal_vector_pure.data(),
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
    if (!monolithic_booly) {
      solve_system_dn();
      return;
    }
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
    baricenter_rigid_velocities = rigid_velocities;  // this solve's velocities about the pole (bem_stokes.cc:4479-4492)
    if (solve_with_torque) flagellum_omega = monolithic_solution[n_dofs + num_rigid];  // ref 4398-4401
  }

  // DN(u_k) = P V^{-1} (P K P u_k) for up to 8 velocities in one device call (bem_stokes.cc:4072-4129)
  void dirichlet_to_neumann_operator_multi(const std::vector<Vector> &input_vels, std::vector<Vector> &output_forces) {
    const int nvec = (int)input_vels.size();
    Vector U((size_t)nvec * n_dofs), F((size_t)nvec * n_dofs, 0.);
    for (int k = 0; k < nvec; ++k) std::copy(input_vels[k].begin(), input_vels[k].end(), U.begin() + (size_t)k * n_dofs);
    std::vector<int> its(nvec, 0);
    if (!solve_directly) check(bs_precond_setup(ctx, BS_MAT_V, BS_PREC_NONE, 0));
    check(bs_dn_operator_multi(ctx, nvec, U.data(), F.data(), solve_directly ? 1 : 0, solver_control.tolerance,
                               (int)solver_control.max_steps, (int)gmres_restart, its.data()));
    solver_control.last_step_ = (unsigned int)*std::max_element(its.begin(), its.end());
    *pcout << "   Iterations needed to solve DN:         " << solver_control.last_step_ << std::endl;
    output_forces.assign(nvec, Vector(n_dofs));
    for (int k = 0; k < nvec; ++k) std::copy(F.begin() + (size_t)k * n_dofs, F.begin() + (size_t)(k + 1) * n_dofs, output_forces[k].begin());
  }
  void dirichlet_to_neumann_operator(const Vector &input_vel, Vector &output_force) {
    std::vector<Vector> out;
    dirichlet_to_neumann_operator_multi({input_vel}, out);
    output_force = out[0];
  }
  // solve_system(false): the rigid-body problem through the DN operator, the 1 + num_rigid systems as one batch
  // (bem_stokes.cc:4163-4258); the 6 x 6 system, which the reference hands to GMRES, is solved by Gaussian elimination
  void solve_system_dn() {
    std::vector<Vector> in, out;
    in.push_back(shape_velocities);
    for (unsigned int r = 0; r < num_rigid; ++r) in.push_back(N_rigid[r]);
    dirichlet_to_neumann_operator_multi(in, out);
    stokes_forces = out[0];
    DN_N_rigid.assign(out.begin() + 1, out.end());
    const unsigned int nr = num_rigid;
    std::vector<double> Fm((size_t)nr * nr, 0.), rhs(nr, 0.);
    for (unsigned int i = 0; i < nr; ++i) rhs[i] = -dot(N_rigid_dual[i], stokes_forces);
    for (unsigned int i = 0; i < nr; ++i) {
      if (grid_type == "ImposedVelocity") {
        Fm[(size_t)i * nr + i] = 1.;
        rhs[i] = (i == imposed_component) ? 1. : 0.;
      } else {
        if (grid_type == "ImposedForce" && i == imposed_component) rhs[i] += 1.;
        for (unsigned int j = 0; j < nr; ++j) Fm[(size_t)i * nr + j] = dot(N_rigid_dual[i], DN_N_rigid[j]);
      }
    }
    final_matrix = Fm;
    // Gaussian elimination with partial pivoting on the nr x nr system
    std::vector<double> a = Fm, x = rhs;
    for (unsigned int k = 0; k < nr; ++k) {
      unsigned int pv = k;
      for (unsigned int i = k + 1; i < nr; ++i)
        if (std::fabs(a[(size_t)i * nr + k]) > std::fabs(a[(size_t)pv * nr + k])) pv = i;
      if (pv != k) {
        for (unsigned int j = 0; j < nr; ++j) std::swap(a[(size_t)k * nr + j], a[(size_t)pv * nr + j]);
        std::swap(x[k], x[pv]);
      }
      for (unsigned int i = k + 1; i < nr; ++i) {
        const double l = a[(size_t)i * nr + k] / a[(size_t)k * nr + k];
        for (unsigned int j = k; j < nr; ++j) a[(size_t)i * nr + j] -= l * a[(size_t)k * nr + j];
        x[i] -= l * x[k];
      }
    }
    for (int k = (int)nr - 1; k >= 0; --k) {
      for (unsigned int j = k + 1; j < nr; ++j) x[k] -= a[(size_t)k * nr + j] * x[j];
      x[k] /= a[(size_t)k * nr + k];
    }
    rigid_velocities = x;
    baricenter_rigid_velocities = x;
    for (unsigned int r = 0; r < nr; ++r)
      for (unsigned int i = 0; i < n_dofs; ++i) stokes_forces[i] += x[r] * DN_N_rigid[r][i];
    rigid_total_forces.assign(nr, 0.);
    for (unsigned int r = 0; r < nr; ++r) rigid_total_forces[r] = dot(stokes_forces, N_rigid_dual[r]);
    reassemble_preconditoner = true;
  }

  // ref: evaluate_stokes_bie (bem_stokes.cc:5366-5451); val_points [P][3], result component-major
  void evaluate_stokes_bie(const std::vector<Tensor1> &val_points, const Vector &vel, const Vector &forces, Vector &val_velocities) {
    val_velocities.assign(3 * val_points.size(), 0.);
    set_kernel();
    check(bs_evaluate_bie(ctx, (int)val_points.size(), val_points[0].data(), vel.data(), forces.data(), val_velocities.data(), 0));
  }

  // ---- multi-frame workflow (host logic; Forward strategy, body-only swimmer, one mesh file per frame) ----------
  static Matrix3 identity3() {
    Matrix3 I{};
    for (int i = 0; i < 3; ++i) I[i][i] = 1.;
    return I;
  }
  // ref: compute_rotation_matrix_from_quaternion bem_stokes.cc:4512-4525
  static void compute_rotation_matrix_from_quaternion(Matrix3 &R, const std::array<double, 4> &q) {
    R[0] = {{1. - 2 * (q[3] * q[3] + q[2] * q[2]), -2 * q[0] * q[3] + 2 * q[1] * q[2], 2 * q[0] * q[2] + 2 * q[1] * q[3]}};
    R[1] = {{2 * q[0] * q[3] + 2 * q[1] * q[2], 1. - 2 * (q[3] * q[3] + q[1] * q[1]), -2 * q[0] * q[1] + 2 * q[3] * q[2]}};
    R[2] = {{-2 * q[0] * q[2] + 2 * q[1] * q[3], 2 * q[0] * q[1] + 2 * q[3] * q[2], 1. - 2 * (q[1] * q[1] + q[2] * q[2])}};
  }
  // ref: update_rotation_matrix bem_stokes.cc:4527-4720 (forward Euler on the quaternion; the theta scheme's 4x4
  // system, which the reference hands to GMRES, is solved by Gaussian elimination)
  void update_rotation_matrix(Matrix3 &rotation, const std::array<double, 3> &omega, const double dt,
                              const bool forward_euler = true, const double theta = 0.5) const {
    std::array<double, 4> q;
    q[0] = std::sqrt(1. + rotation[0][0] + rotation[1][1] + rotation[2][2]) / 2;
    q[1] = 1 / q[0] * 0.25 * (rotation[2][1] - rotation[1][2]);
    q[2] = 1 / q[0] * 0.25 * (rotation[0][2] - rotation[2][0]);
    q[3] = 1 / q[0] * 0.25 * (rotation[1][0] - rotation[0][1]);
    auto normalise = [](std::array<double, 4> &v) {
      const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
      for (double &x : v) x /= n;
    };
    normalise(q);
    const double op[4] = {0., omega[0], omega[1], omega[2]};
    const double S[4][4] = {{q[0], -q[1], -q[2], -q[3]}, {q[1], q[0], q[3], -q[2]}, {q[2], -q[3], q[0], q[1]}, {q[3], q[2], -q[1], q[0]}};
    std::array<double, 4> qdot{{0, 0, 0, 0}};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) qdot[i] += 0.5 * S[i][j] * op[j];
    if (forward_euler) {
      for (int i = 0; i < 4; ++i) q[i] += dt * qdot[i];
    } else {
      const double h = theta * dt * 0.5;
      double A[4][5] = {{1. + h * op[0], h * op[1], h * op[2], h * op[3], 0}, {-h * op[1], 1. + h * op[0], -h * op[3], h * op[2], 0},
                        {-h * op[2], h * op[3], 1. + h * op[0], -h * op[1], 0}, {-h * op[3], -h * op[2], h * op[1], 1. + h * op[0], 0}};
      for (int i = 0; i < 4; ++i) A[i][4] = q[i] + (1 - theta) * dt * qdot[i];
      for (int c = 0; c < 4; ++c) {  // partial pivoting
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
          if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        for (int k = 0; k < 5; ++k) std::swap(A[c][k], A[piv][k]);
        for (int r = c + 1; r < 4; ++r) {
          const double f = A[r][c] / A[c][c];
          for (int k = c; k < 5; ++k) A[r][k] -= f * A[c][k];
        }
      }
      for (int r = 3; r >= 0; --r) {
        double v = A[r][4];
        for (int k = r + 1; k < 4; ++k) v -= A[r][k] * q[k];
        q[r] = v / A[r][r];
      }
    }
    normalise(q);
    compute_rotation_matrix_from_quaternion(rotation, q);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double d = -(i == j ? 1. : 0.);
        for (int k = 0; k < 3; ++k) d += rotation[k][i] * rotation[k][j];
        if (std::fabs(d) >= 1e-7)
          *pcout << "Something Wrong in Rotations, " << (i == j ? "on" : "out") << " the diagonal " << std::fabs(d) << std::endl;
      }
  }
  // ref: apply_rotation_along_axis bem_stokes.cc:846-878 (Rodrigues)
  static void apply_rotation_along_axis(Tensor1 &out, const Tensor1 &in, const Tensor1 &a, const double angle) {
    const double c = std::cos(angle), s = std::sin(angle);
    const double R[3][3] = {{c + a[0] * a[0] * (1 - c), a[0] * a[1] * (1 - c) - a[2] * s, a[0] * a[2] * (1 - c) + a[1] * s},
                            {a[0] * a[1] * (1 - c) + a[2] * s, c + a[1] * a[1] * (1 - c), a[1] * a[2] * (1 - c) - a[0] * s},
                            {a[0] * a[2] * (1 - c) - a[1] * s, a[1] * a[2] * (1 - c) + a[0] * s, c + a[2] * a[2] * (1 - c)}};
    Tensor1 r{{0, 0, 0}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r[i] += R[i][j] * in[j];
    out = r;
  }
  QuadMesh read_input_mesh_file(unsigned int frame) const {  // bem_stokes.cc:496-523
    return read_mesh(input_grid_path + input_grid_base_name + std::to_string(frame) + "." + input_grid_format);
  }
  // ref: compute_euler_vector bem_stokes.cc:2247-2431 (mesh-file branch): nodes of the frame, rotated, component-major
  void compute_euler_vector(Vector &euler, unsigned int frame, bool consider_displacements = true) const {
    const QuadMesh m = read_input_mesh_file(frame);
    if (m.n_nodes() != (int)N) throw std::runtime_error("frame grid has a different topology");
    euler.assign(n_dofs, 0.);
    for (unsigned int i = 0; i < N; ++i)
      for (int r = 0; r < 3; ++r) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += rotation_matrix[r][k] * m.nodes[3 * i + k];
        euler[i + r * N] = v;
      }
    if (consider_displacements && bool_dipl)
      for (unsigned int i = 0; i < n_dofs; ++i) euler[i] += rigid_displacements_for_sim[i];
  }
  void project_shape_velocities(unsigned int /*frame*/) {  // bem_stokes.cc:2120-2137, isoparametric
    shape_velocities.assign(n_dofs, 0.);
    for (unsigned int i = 0; i < n_dofs; ++i) shape_velocities[i] = (next_euler_vec[i] - euler_vec[i]) / time_step;
  }
  // ref: update_system_state bem_stokes.cc:4725-4846 ("Forward"; Heun's corrector restores the backed-up state and
  // integrates the mean of the two velocities)
  void update_system_state(bool /*compute*/, unsigned int /*frame*/, bool consider_rotations, bool consider_displacements,
                           const std::string &res_system = "Forward") {
    if (res_system == "Heun" && res_strategy == "Heun") {
      rotation_matrix = old_rotation_matrix;
      rigid_displacements_for_sim = old_rigid_displacements_for_sim;
      for (unsigned int r = 0; r < num_rigid; ++r) rigid_velocities[r] = 0.5 * rigid_velocities[r] + 0.5 * old_rigid_velocities[r];
    } else if (res_system == "Forward" && res_strategy == "Heun") {
      old_rigid_velocities = rigid_velocities;
      old_rotation_matrix = rotation_matrix;
      old_rigid_displacements_for_sim = rigid_displacements_for_sim;
    }
    // the punctual velocities come from the LAST solve (baricenter_rigid_velocities, ref 4784-4789); Heun's mean enters
    // only through omega in update_rotation_matrix
    rigid_puntual_velocities.assign(n_dofs, 0.);
    for (unsigned int r = 0; r < 3; ++r)
      for (unsigned int i = 0; i < n_dofs; ++i) rigid_puntual_velocities[i] += assemble_scaling * baricenter_rigid_velocities[r] * N_rigid[r][i];
    rigid_puntual_translation_velocities = rigid_puntual_velocities;
    for (unsigned int r = 3; r < num_rigid; ++r)
      for (unsigned int i = 0; i < n_dofs; ++i) rigid_puntual_velocities[i] += assemble_scaling * baricenter_rigid_velocities[r] * N_rigid[r][i];
    if (consider_rotations) update_rotation_matrix(rotation_matrix, {{rigid_velocities[3], rigid_velocities[4], rigid_velocities[5]}}, time_step);
    next_rigid_puntual_displacements.assign(n_dofs, 0.);
    for (unsigned int i = 0; i < n_dofs; ++i) next_rigid_puntual_displacements[i] = time_step * rigid_puntual_translation_velocities[i];
    if (consider_displacements) {
      const bool flag[3] = {bool_dipl_x, bool_dipl_y, bool_dipl_z};
      for (int c = 0; c < 3; ++c)
        if (flag[c])
          for (unsigned int i = c * N; i < (c + 1) * N; ++i) rigid_displacements_for_sim[i] += next_rigid_puntual_displacements[i];
    }
  }
  // deal.II Vector<double>::block_write: "<size>\n[" raw doubles "]"
  static void block_write(const std::string &path, const Vector &v) {
    std::ofstream f(path, std::ios::binary);
    const std::string head = std::to_string(v.size()) + "\n[";
    f.write(head.data(), (std::streamsize)head.size());
    f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(double)));
    f.write("]", 1);
  }
  void output_save_stokes_results(unsigned int cycle) const {  // the .bin files of bem_stokes.cc:5264-5316
    const std::string d = output_dir + "/", c = std::to_string(cycle);
    block_write(d + "stokes_forces_" + c + ".bin", stokes_forces);
    block_write(d + "shape_velocities_" + c + ".bin", shape_velocities);
    block_write(d + "total_velocities_" + c + ".bin", total_velocities);
    Vector rot;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rot.push_back(rotation_matrix[i][j]);
    block_write(d + "rotation_matrix_" + c + ".bin", rot);
    block_write(d + "4_6_rigid_velocities_" + c + ".bin", rigid_velocities);
    block_write(d + "4_6_overall_forces_" + c + ".bin", rigid_total_forces);
    block_write(d + "stokes_rigid_displ_" + c + ".bin", next_rigid_puntual_displacements);
    block_write(d + "stokes_rigid_vel_" + c + ".bin", rigid_puntual_velocities);
    block_write(d + "euler_vec_" + c + ".bin", euler_vec);
    block_write(d + "normal_vector" + c + ".bin", normal_vector_pure);
  }
  // ref: BEMProblem::run bem_stokes.cc:5636-5888 (Forward and Heun)
  void run(unsigned int start_frame = 0, unsigned int end_frame = 0) {
    if (res_strategy != "Forward" && res_strategy != "Heun") throw Error(BS_ERR_UNSUPPORTED, "unknown time integration " + res_strategy);
    compute_rotation_matrix_from_quaternion(rotation_matrix, initial_quaternion);
    read_domain(read_input_mesh_file(start_frame % n_frames));
    reinit();
    rigid_displacements_for_sim.assign(n_dofs, 0.);
    rigid_puntual_displacements.assign(n_dofs, 0.);
    Vector frame_euler;
    compute_euler_vector(frame_euler, start_frame % n_frames, true);
    reassemble_preconditoner = true;
    auto solve_frame = [&](unsigned int i) {  // geometry of frame_euler -> pre-pass -> shape velocities -> assemble -> solve
      euler_vec = frame_euler;
      check(bs_set_geometry(ctx, (int)N, euler_vec.data(), mesh.n_cells(), mesh.conn.data(), (int)N, mesh.conn.data(), nullptr));
      compute_center_of_mass_and_rigid_modes(i);
      compute_normal_vector();
      if (grid_type != "Real") next_euler_vec = euler_vec;
      project_shape_velocities(i);
      if (grid_type != "Real") shape_velocities.assign(n_dofs, 0.);
      *pcout << "Assembling" << std::endl;
      assemble_stokes_system(true);
      monolithic_solution.assign(n_dofs + num_rigid, 0.);
      solve_system(monolithic_bool);
    };
    for (unsigned int i = start_frame; i <= end_frame; i += delta_frame) {
      *pcout << "Analyzing frame = " << i << " over " << n_frames << std::endl;
      compute_euler_vector(next_euler_vec, (i + 1) % n_frames, true);
      solve_frame(i);
      if (res_strategy == "Forward") {
        update_system_state(true, i, bool_rot, bool_dipl, "Forward");
      } else {  // Heun: predictor state, geometry and solve at the next frame, corrector with the mean velocity
        update_system_state(true, i, bool_rot, bool_dipl, "Forward");
        compute_euler_vector(frame_euler, (i + 1) % n_frames, true);
        compute_euler_vector(next_euler_vec, (i + 2) % n_frames, true);
        solve_frame(i + 1);  // ref 5797: compute_center_of_mass_and_rigid_modes(i+1)
        update_system_state(true, i, bool_rot, bool_dipl, "Heun");
      }
      total_velocities = shape_velocities;
      for (unsigned int k = 0; k < n_dofs; ++k) total_velocities[k] += rigid_puntual_velocities[k];
      rigid_puntual_displacements = next_rigid_puntual_displacements;
      output_save_stokes_results(i);
      *pcout << "preparing for new time" << std::endl;
      compute_euler_vector(frame_euler, (i + delta_frame) % n_frames, true);
    }
    *pcout << "THE END" << std::endl;
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
