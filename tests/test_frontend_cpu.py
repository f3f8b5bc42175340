"""Host front-end (bemstokes_b200/frontend.py): parameter files, quaternion integrator, result-file format.
No GPU needed.  The rotation tests restate the reference's tests/rotation_test.cc and
tests/rotation_test_cranck_nicholson.cc (all lines of their .output files read 'OK : OK : OK :') with a larger time
step: 1e-4 instead of 1e-6, and the tolerance scaled with it (the reference's 1e-5 is 10 dt; the first-order
quadrature of omega(t) leaves an O(2 pi dt) phase lag)."""
import math
import os

import numpy as np
import pytest

from bemstokes_b200 import frontend as fe

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_parse_reference_parameter_file():
    tree = fe.parse_prm(os.path.join(GOLDEN, "parameters_test_alpha_box.prm"))
    main = tree["BEMStokes::BEMProblem<3>"]
    assert main["Grid"] == "ImposedForce" and main["Wall 0 spans"] == "80,0,80"
    assert main["Solver"]["Tolerance"] == "1.e-10"
    assert tree["Internal Quadrature"]["Quadrature order"] == "8"
    assert "FE_Q<2,3>(1)" in tree["Finite Element Stokes"]["Finite element space"]


def test_parameters_reach_the_problem():
    import bemstokes_b200 as bb
    p = bb.BEMProblem()
    p.parse_parameters(os.path.join(GOLDEN, "parameters_test_alpha_box.prm"))
    assert (p.grid_type, p.quadrature_order, p.singular_quadrature_order) == ("ImposedForce", 8, 10)
    assert p.wall_spans_0 == (80.0, 0.0, 80.0) and p.wall_position_0 == (0.0, 1.4, 0.0)
    assert p.use_internal_alpha is True and p.monolithic_bool is True and p.reflect_kernel is False
    assert p.input_grid_base_name == "sphere_mesh_3d_" and p.input_grid_format == "msh" and p.n_frames == 120
    assert p.solver_control.tolerance == 1e-10 and p.solver_control.max_steps == 1000000
    assert p.initial_quaternion == [1.0, 0.0, 0.0, 0.0] and p.force_pole == (0.0, 0.0, 0.0)
    assert p.wall_bool[:6] == [True] * 6 and p.wall_bool[6:] == [False, False]
    q = bb.BEMProblem().parse_parameters(os.path.join(GOLDEN, "parameters_test_alpha_box_ref_quadrature.prm"))
    assert q.quadrature_order == 15 and q.singular_quadrature_order == 20
    with pytest.raises(KeyError):
        bb.BEMProblem().parse_parameters("subsection BEMStokes::BEMProblem<3>\n set Cylinder Radius = 2\nend\n", strict=True)
    with pytest.raises(ValueError):
        fe.parse_prm("subsection A\n set x = 1\n")


@pytest.mark.parametrize("forward_euler", [True, False])
def test_rotation_with_quaternions(forward_euler):
    dt = 1e-4
    tol = 10 * dt
    nsteps = int(round(1 / dt))
    msgs = []
    for i in range(3):
        P0 = np.zeros(3)
        P0[i] = 1.0
        R = np.eye(3)
        axis = np.zeros(3)
        axis[(i + 2) % 3] = 1.0
        for j in range(nsteps):
            omega = axis * math.cos(2 * math.pi * j / nsteps) * (2 * math.pi)
            R = fe.update_rotation_matrix(R, omega, dt, forward_euler=forward_euler, log=msgs.append)
            if j % 100 == 0:
                # the reference compares the state after step j with the exact rotation at t_j (tests/rotation_test.cc:79)
                Pref = fe.apply_rotation_along_axis(P0, axis, math.sin(2 * math.pi * j / nsteps))
                assert np.abs(Pref - R @ P0).max() <= tol
        assert np.abs(R @ P0 - P0).max() < 1e-4  # full period: back to the start
    assert msgs == []


def test_rotation_matrix_from_quaternion_is_a_rotation():
    rng = np.random.default_rng(0)
    for _ in range(10):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        R = fe.compute_rotation_matrix_from_quaternion(q)
        assert np.abs(R.T @ R - np.eye(3)).max() < 1e-14 and abs(np.linalg.det(R) - 1) < 1e-14
    ax = np.array([0.0, 0.0, 1.0])
    assert np.allclose(fe.apply_rotation_along_axis([1, 0, 0], ax, math.pi / 2), [0, 1, 0], atol=1e-15)


def test_block_write_format(tmp_path):
    v = np.array([1.5, -2.25, 3.0])
    path = str(tmp_path / "v.bin")
    fe.vector_block_write(path, v)
    raw = open(path, "rb").read()
    assert raw[:3] == b"3\n[" and raw[-1:] == b"]" and len(raw) == 3 + 24 + 1   # deal.II Vector::block_write
    assert np.array_equal(fe.vector_block_read(path), v)
    open(path, "wb").write(b"3\n(" + v.tobytes() + b"]")
    with pytest.raises(ValueError):
        fe.vector_block_read(path)


class _StubProblem(fe.FrameLoop):
    """State container for the host-only parts of the frame loop (no device context)."""

    def __init__(self, N):
        self._init_frontend()
        self.n_dofs, self.num_rigid, self.assemble_scaling = 3 * N, 6, 1.0
        rng = np.random.default_rng(3)
        x = rng.uniform(-1, 1, (N, 3))
        R = np.zeros((6, 3, N))
        for c in range(3):
            R[c, c] = 1.0
        R[3, 1], R[3, 2] = -x[:, 2], x[:, 1]
        R[4, 0], R[4, 2] = x[:, 2], -x[:, 0]
        R[5, 0], R[5, 1] = -x[:, 1], x[:, 0]
        self.N_rigid = R.reshape(6, 3 * N)
        self.rigid_displacements_for_sim = np.zeros(3 * N)
        self.log = lambda *_: None


def test_update_system_state_forward_and_heun():
    """ref: update_system_state (bem_stokes.cc:4725-4846)."""
    N = 7
    s = _StubProblem(N)
    s.time_step, s.bool_dipl_x, s.bool_dipl_z = 0.1, True, True
    U = np.array([0.3, -0.2, 0.5, 0.0, 0.0, 2.0])
    s.rigid_velocities = U.copy()
    s.baricenter_rigid_velocities = U.copy()
    s.update_system_state(True, 0, True, True, "Forward")
    assert np.allclose(s.rigid_puntual_velocities, U @ s.N_rigid)
    assert np.allclose(s.rigid_puntual_translation_velocities, U[:3] @ s.N_rigid[:3])
    assert np.allclose(s.next_rigid_puntual_displacements, 0.1 * (U[:3] @ s.N_rigid[:3]))
    d = s.rigid_displacements_for_sim.reshape(3, N)
    assert np.allclose(d[0], 0.03) and np.allclose(d[1], 0.0) and np.allclose(d[2], 0.05)   # y is switched off
    ang = 2.0 * 0.1   # rotation about z by omega*dt (forward Euler on the quaternion: tan(half angle) = omega dt / 2)
    Rz = s.rotation_matrix
    assert abs(math.atan2(Rz[1, 0], Rz[0, 0]) - 2 * math.atan(ang / 2)) < 1e-14 and abs(Rz[2, 2] - 1) < 1e-14
    # Heun: the predictor backs the state up, the corrector restores it and averages the velocities
    h = _StubProblem(N)
    h.res_strategy, h.time_step = "Heun", 0.1
    h.rigid_velocities = U.copy()
    h.baricenter_rigid_velocities = U.copy()
    h.update_system_state(True, 0, True, False, "Forward")
    R_pred = h.rotation_matrix.copy()
    assert np.array_equal(h.old_rigid_velocities, U) and np.array_equal(h.old_rotation_matrix, np.eye(3))
    U2 = np.array([0.1, 0.0, 0.1, 0.0, 0.0, 1.0])
    h.rigid_velocities = U2.copy()
    h.baricenter_rigid_velocities = U2.copy()
    h.update_system_state(True, 0, True, False, "Heun")
    assert np.allclose(h.rigid_velocities, 0.5 * (U + U2))
    half = fe.update_rotation_matrix(np.eye(3), 0.5 * (U + U2)[3:], 0.1)
    assert np.allclose(h.rotation_matrix, half) and not np.allclose(h.rotation_matrix, R_pred)


def test_result_files_and_main_arguments(tmp_path):
    N = 3
    s = _StubProblem(N)
    s.output_dir = str(tmp_path)
    n = 3 * N
    for name in ("stokes_forces", "shape_velocities", "total_velocities", "next_rigid_puntual_displacements",
                 "rigid_puntual_velocities", "euler_vec", "normal_vector", "rigid_puntual_displacements"):
        setattr(s, name, np.arange(n, dtype=float) + len(name))
    s.rigid_velocities, s.rigid_total_forces = np.arange(6.0), -np.arange(6.0)
    s.output_save_stokes_results(5)
    assert np.array_equal(fe.vector_block_read(str(tmp_path / "stokes_forces_5.bin")), s.stokes_forces)
    assert np.array_equal(fe.vector_block_read(str(tmp_path / "4_6_overall_forces_5.bin")), s.rigid_total_forces)
    assert np.array_equal(s.read_rotation_matrix(5), np.eye(3))
    assert (tmp_path / "point_0_on_proc_0_displacement_frame_5.txt").read_text().startswith("5 ")
    from bemstokes_b200.__main__ import parse_args
    a = parse_args([])
    assert (a.start_frame, a.end_frame, a.prm) == (0, 139, "parameters_3.prm")      # main.cc:14-15, 33
    a = parse_args(["3", "7", "--prm", "x.prm"])
    assert (a.start_frame, a.end_frame, a.prm) == (3, 7, "x.prm")


def test_main_reports_exceptions_like_the_reference(capsys):
    from bemstokes_b200.__main__ import main
    assert main(["--prm", "does_not_exist.prm"]) == 1          # main.cc:48-60: message on stderr, status 1
    err = capsys.readouterr().err
    assert "Exception on processing" in err and "Aborting!" in err
