// Written like the reference's tests/sphere_translation.cc on the C++ host mirror: frame 0 and 1 of the translating
// sphere grids, finite-difference shape velocity, Real grid, Direct preconditioner, then the state update and the
// result files; afterwards the same two frames through BEMProblem::run.  argv[1] = directory of the grids,
// argv[2] = output directory.  Prints the reference's lines ("ERROR on rigid translation 0 : U , exact , rel").
#include <bemstokes_b200.hpp>
#include <cmath>

using namespace bemstokes_b200;

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const double tol = 1e-5;
  std::cout << "Test for the Motility tensor of a sphere in free space" << std::endl;
  BEMProblem<3> bem_problem_3d;
  bem_problem_3d.quadrature_order = 8;            // parameters_test_alpha_box.prm
  bem_problem_3d.singular_quadrature_order = 10;
  bem_problem_3d.reflect_kernel = false;
  bem_problem_3d.no_slip_kernel = false;
  bem_problem_3d.use_internal_alpha = false;
  bem_problem_3d.grid_type = "Real";
  bem_problem_3d.monolithic_bool = true;
  bem_problem_3d.solve_directly = false;
  bem_problem_3d.preconditioner_type = "Direct";
  bem_problem_3d.reassemble_preconditoner = true;
  bem_problem_3d.input_grid_path = std::string(argv[1]) + "/";
  bem_problem_3d.input_grid_base_name = "sphere_translation_";
  bem_problem_3d.input_grid_format = "msh";
  bem_problem_3d.output_dir = argv[2];
  bem_problem_3d.n_frames = 2;

  bem_problem_3d.read_domain(bem_problem_3d.read_input_mesh_file(0));
  bem_problem_3d.reinit();
  const double exact_velocity = 1. / 120. / bem_problem_3d.time_step;
  bem_problem_3d.compute_center_of_mass_and_rigid_modes(0);
  bem_problem_3d.compute_normal_vector();
  bem_problem_3d.compute_euler_vector(bem_problem_3d.next_euler_vec, 1, true);
  bem_problem_3d.project_shape_velocities(0);
  bem_problem_3d.assemble_stokes_system(true);
  bem_problem_3d.solve_system(bem_problem_3d.monolithic_bool);
  const Vector &U = bem_problem_3d.rigid_velocities;
  if (std::fabs(U[0] - exact_velocity) / std::fabs(exact_velocity) <= tol)
    std::cout << "OK rigid translation " << 0 << std::endl;
  else
    std::cout << "ERROR on rigid translation " << 0 << " : " << U[0] << " , " << exact_velocity << " , "
              << std::fabs(U[0] - exact_velocity) / std::fabs(exact_velocity) << std::endl;
  for (unsigned int i = 1; i < 3; ++i) {
    if (std::fabs(U[i]) <= tol) std::cout << "OK rigid translation " << i << std::endl;
    else std::cout << "ERROR on rigid traslation " << i << " : " << U[i] << std::endl;
  }
  for (unsigned int i = 3; i < 6; ++i) {
    if (std::fabs(U[i]) <= tol) std::cout << "OK rigid rotation " << i << std::endl;
    else std::cout << "ERROR on rigid rotation " << i << " : " << U[i] << std::endl;
  }
  bem_problem_3d.update_system_state(true, 0, false, false, "Forward");
  bem_problem_3d.total_velocities = bem_problem_3d.shape_velocities;
  bem_problem_3d.output_save_stokes_results(0);
  const double u0 = U[0];

  // the same through the frame loop: frame 0, then frame 1 whose next frame is frame 0 again (reversed stroke)
  BEMProblem<3> looped;
  looped.quadrature_order = 8;
  looped.singular_quadrature_order = 10;
  looped.grid_type = "Real";
  looped.solve_directly = false;
  looped.preconditioner_type = "Direct";
  looped.input_grid_path = bem_problem_3d.input_grid_path;
  looped.input_grid_base_name = "sphere_translation_";
  looped.output_dir = argv[2];
  looped.n_frames = 2;
  std::ostringstream sink;
  looped.pcout = &sink;
  looped.run(0, 1);
  std::cout.precision(12);
  std::cout << "run frame 1 velocity ratio " << looped.rigid_velocities[0] / u0 << std::endl;

  // Heun predictor-corrector (bem_stokes.cc:5780-5830) on the same grids, one frame: result files into <out>/heun
  if (argc > 3) {
    BEMProblem<3> heun;
    heun.quadrature_order = 8;
    heun.singular_quadrature_order = 10;
    heun.grid_type = "Real";
    heun.solve_directly = true;
    heun.res_strategy = "Heun";
    heun.input_grid_path = bem_problem_3d.input_grid_path;
    heun.input_grid_base_name = "sphere_translation_";
    heun.output_dir = argv[3];
    heun.n_frames = 2;
    heun.pcout = &sink;
    heun.run(0, 0);
    std::cout << "heun mean velocity " << heun.rigid_velocities[0] << " predictor " << heun.old_rigid_velocities[0]
              << " last solve " << heun.baricenter_rigid_velocities[0] << std::endl;
  }
  return 0;
}
