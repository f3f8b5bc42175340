// Written like the reference's tests/rotation_test.cc and tests/rotation_test_cranck_nicholson.cc: a point on axis i is
// rotated about axis i+2 with omega(t) = 2 pi cos(2 pi t) by the quaternion integrator and compared with the exact
// Rodrigues rotation by sin(2 pi t); every check line must read "OK : OK : OK : " (tests/rotation_test.output).
// The time step is 1e-4 instead of 1e-6 and the tolerance scaled with it (10 dt, as in the reference).
// Host code only: no GPU call is made.  argv[1] = "cn" selects the theta scheme (Crank-Nicolson).
#include <bemstokes_b200.hpp>
#include <cmath>
#include <cstring>

using namespace bemstokes_b200;

template class bemstokes_b200::BEMProblem<3>;  // compile every member of the mirror (run(), update_system_state(), ...)

int main(int argc, char **argv) {
  const bool forward_euler = !(argc > 1 && std::strcmp(argv[1], "cn") == 0);
  const double dt = 1e-4, tol = 10 * dt, pi = 3.14159265358979323846;
  const unsigned int dim = 3, nsteps = (unsigned int)std::lround(1. / dt);
  BEMProblem<3> bem_problem_3d;
  std::cout << "Minimum Test for the rotation with quaternions" << std::endl;
  for (unsigned int i = 0; i < dim; ++i) {
    Tensor1 P_0{{0, 0, 0}}, axis{{0, 0, 0}};
    P_0[i] = 1.;
    axis[(i + 2) % dim] = 1.;
    Matrix3 rotation_matrix = BEMProblem<3>::identity3();
    std::cout << P_0[0] << " " << P_0[1] << " " << P_0[2] << std::endl;
    for (unsigned int j = 0; j < nsteps; ++j) {
      std::array<double, 3> omega{{0, 0, 0}};
      omega[(i + 2) % dim] = std::cos(2 * pi * j / nsteps) * (2 * pi);
      bem_problem_3d.update_rotation_matrix(rotation_matrix, omega, dt, forward_euler);
      if (j % 1000 == 0) {
        Tensor1 P_test{{0, 0, 0}}, P_ref;
        for (unsigned int r = 0; r < dim; ++r)
          for (unsigned int k = 0; k < dim; ++k) P_test[r] += rotation_matrix[r][k] * P_0[k];
        BEMProblem<3>::apply_rotation_along_axis(P_ref, P_0, axis, std::sin(2 * pi * j / nsteps));
        std::cout << "Testing j = " << j << std::endl;
        for (unsigned int k = 0; k < dim; ++k) {
          if (std::fabs(P_ref[k] - P_test[k]) > tol) {
            std::cout << "ERROR !!!" << std::endl;
            break;
          }
          std::cout << "OK : ";
        }
        std::cout << std::endl;
      }
    }
  }
  return 0;
}
