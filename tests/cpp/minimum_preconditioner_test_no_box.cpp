// C++ host test over include/bemstokes_b200.hpp, written like the reference's own tests
// (tests/minimum_preconditioner_test_no_box.cc, tests/rigidity_sphere.cc, tests/reflected_kernel_test_G.cc,
// tests/wall_kernel_test_G.cc): public-member overrides, assemble_stokes_system, direct solve as the reference
// solution, then GMRES with different preconditioners; stdout carries the reference's log lines so that the
// harness (tests/test_cpp_host.py) can diff the fingerprints against tests/golden/reference_goldens.json.
#include <bemstokes_b200.hpp>

#include <cstdio>
#include <iomanip>

using namespace bemstokes_b200;

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: " << argv[0] << " <grid_test/sphere_half_refined_0.inp>" << std::endl;
    return 2;
  }
  try {
    const double tol = 1e-8;
    BEMProblem<3> bem_problem_3d;
    bem_problem_3d.pcout = &std::cout;
    std::cout << std::setprecision(6);
    bem_problem_3d.pcout->precision(6);
    std::cout << "Minimum Test for the preconditioner with exterior problem and the monolithic system" << std::endl;
    bem_problem_3d.use_internal_alpha = false;
    bem_problem_3d.quadrature_order = 8;              // parameters_test_alpha_box.prm: Internal Quadrature gauss 8
    bem_problem_3d.singular_quadrature_order = 10;    //                              Singular quadrature order 10
    bem_problem_3d.grid_type = "ImposedForce";
    bem_problem_3d.imposed_component = 1;
    bem_problem_3d.solver_control = SolverControl(1000, 1e-10);
    bem_problem_3d.read_domain(read_mesh(argv[1]));
    std::cout << "We have a tria of " << bem_problem_3d.mesh.n_cells() << " cells." << std::endl;
    bem_problem_3d.reinit();
    std::cout << "There are " << bem_problem_3d.n_dofs << " degrees of freedom" << std::endl;
    bem_problem_3d.compute_center_of_mass_and_rigid_modes(0);
    bem_problem_3d.compute_normal_vector();
    bem_problem_3d.assemble_stokes_system(true);

    std::cout << "Solving directly the monolithic system" << std::endl;
    bem_problem_3d.solve_directly = true;
    bem_problem_3d.solve_system(true);
    const Vector reference_monolithic_solution = bem_problem_3d.monolithic_solution;

    std::cout << "Solving using an iterative sovler combined with a preconditioner" << std::endl;
    bem_problem_3d.solve_directly = false;
    for (const char *prec : {"Direct", "Jacobi", "None"}) {
      std::cout << prec << std::endl;
      bem_problem_3d.preconditioner_type = prec;
      bem_problem_3d.monolithic_solution.assign(bem_problem_3d.monolithic_solution.size(), 0.);
      bem_problem_3d.solve_system(true);
      for (size_t i = 0; i < reference_monolithic_solution.size(); ++i) {
        const double foo = std::abs(reference_monolithic_solution[i] - bem_problem_3d.monolithic_solution[i]);
        if (foo > tol) std::cout << "ERROR, index i : " << i << " : " << foo << " , instead of : " << 0 << std::endl;
      }
    }
    std::cout << "rigid velocity for the unit force : " << std::setprecision(10) << bem_problem_3d.rigid_velocities[1] << std::endl;

    // tests/reflected_kernel_test_G.cc and tests/wall_kernel_test_G.cc
    for (unsigned int i = 0; i < 3; ++i) {
      std::cout << "Testing a perfect slip kernel using a normal along the " << i << " axis" << std::endl;
      FreeSurfaceStokesKernel<3> fs_kernel;
      const double position = 1., ktol = 1e-6;
      fs_kernel.set_wall_orientation(i);
      Tensor1 valuation_point{{0, 0, 0}};
      valuation_point[i] = position;
      valuation_point[(i + 1) % 3] = 3.;
      const Tensor2 G = fs_kernel.value_tens_image(valuation_point, valuation_point);
      for (unsigned int j = 0; j < 3; ++j) std::cout << (std::abs(G[i][j]) < ktol ? "OK" : "ERROR") << std::endl;
      NoSlipWallStokesKernel<3> ns_kernel;
      ns_kernel.set_wall_orientation(i);
      Tensor1 source{{0, 0, 0}}, val{{0, 0, 0}};
      source[i] = 6.67; source[(i + 1) % 3] = 3.234; source[(i + 2) % 3] = 9.234;
      val[i] = position; val[(i + 1) % 3] = 3.667; val[(i + 2) % 3] = 0.214456;
      Tensor1 source_image = source;
      source_image[i] -= 2 * (source[i] - position);
      Tensor1 R, R_image;
      for (int d = 0; d < 3; ++d) { R[d] = val[d] - source[d]; R_image[d] = val[d] - source_image[d]; }
      const Tensor2 Gn = ns_kernel.value_tens_image(R, R_image);
      for (unsigned int a = 0; a < 3; ++a)
        for (unsigned int b = 0; b < 3; ++b) std::cout << (std::abs(Gn[a][b]) < ktol ? "OK" : "ERROR") << std::endl;
    }
  } catch (const std::exception &exc) {  // like source/main.cc:48-71
    std::cerr << std::endl << "----------------------------------------------------" << std::endl
              << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl;
    return 1;
  }
  return 0;
}
