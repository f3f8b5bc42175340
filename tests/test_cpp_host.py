"""The C++ host mirror (include/bemstokes_b200.hpp): compiles and links against the C-ABI library on the CPU box;
on the GPU the program — written like the reference's tests/minimum_preconditioner_test_no_box.cc — is run and
its stdout fingerprints are compared with the reference's golden outputs."""
import os
import re
import subprocess

import pytest

from conftest import MESHES, ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "minimum_preconditioner_test_no_box.cpp")
LIBDIR = os.path.join(ROOT, "bemstokes_b200")


def build_exe(tmp, src=SRC):
    exe = os.path.join(tmp, os.path.splitext(os.path.basename(src))[0])
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", LIBDIR,
           "-lbemstokes_b200", "-Wl,-rpath," + LIBDIR, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_compiles_and_links(tmp_path):
    import bemstokes_b200  # noqa: F401  (makes sure the library is built)
    exe = build_exe(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_host_reference_style_test(tmp_path, goldens):
    import bemstokes_b200  # noqa: F401
    exe = build_exe(str(tmp_path))
    r = subprocess.run([exe, os.path.join(MESHES, "sphere_half_refined_0.inp")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "We have a tria of 106 cells." in out and "There are 324 degrees of freedom" in out
    assert "The Mass (Surface) of the entire system is : 12.1766" in out          # tests/rigidity_sphere.output:8
    assert "Check on the V operator Norm (should be zero) pure: 0.00245963" in out  # tests/rigidity_sphere.output:15
    assert "Check on the V operator Norm post (should be one) pure: 1" in out
    for k in range(3):
        assert "check with versor vector : %d l_infty : 1" % k in out
    its = [int(x) for x in re.findall(r"Iterations needed to solve monolithic:\s+(\d+)", out)]
    assert its == [1, goldens["gmres_iterations_no_box"]["Jacobi"], 40]
    assert "ERROR" not in out
    assert out.count("OK") == 3 * (3 + 9)
    fc = [float(x) for x in re.findall(r"FINAL CHECK 0 ([-0-9.e+]+)", out)]
    assert len(fc) == 4 and max(fc) < 1e-9


@pytest.mark.parametrize("scheme", ["forward_euler", "cn"])
def test_cpp_rotation_test(tmp_path, scheme):
    """tests/cpp/rotation_test.cpp = the reference's tests/rotation_test.cc / rotation_test_cranck_nicholson.cc on the
    C++ mirror's quaternion integrator (host code, no GPU): every check line reads 'OK : OK : OK : '."""
    import bemstokes_b200  # noqa: F401
    exe = build_exe(str(tmp_path), os.path.join(ROOT, "tests", "cpp", "rotation_test.cpp"))
    r = subprocess.run([exe] + (["cn"] if scheme == "cn" else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert lines[0] == "Minimum Test for the rotation with quaternions"       # tests/rotation_test.output:1
    checks = [lines[k + 1] for k, l in enumerate(lines) if l.startswith("Testing j = ")]
    assert len(checks) == 3 * 10 and all(c == "OK : OK : OK : " for c in checks)
    assert "ERROR" not in r.stdout and "Something Wrong in Rotations" not in r.stdout


@pytest.mark.gpu
def test_cpp_sphere_translation_and_frame_loop(tmp_path, goldens):
    """tests/cpp/sphere_translation.cpp = the reference's tests/sphere_translation.cc on the C++ mirror, then the same
    grids through BEMProblem::run; the printed lines are those of tests/sphere_translation.output."""
    import bemstokes_b200  # noqa: F401
    import numpy as np
    from bemstokes_b200 import frontend as fe
    exe = build_exe(str(tmp_path), os.path.join(ROOT, "tests", "cpp", "sphere_translation.cpp"))
    hdir = os.path.join(str(tmp_path), "heun_cpp")
    os.makedirs(hdir)
    r = subprocess.run([exe, MESHES, str(tmp_path), hdir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    G = goldens["sphere_translation"]
    m = re.search(r"ERROR on rigid translation 0 : ([-0-9.e+]+) , ([-0-9.e+]+) , ([-0-9.e+]+)", out)
    assert m, out[-1500:]
    assert abs(float(m.group(1)) - G["rigid_velocity_0"]) < 6e-7 and abs(float(m.group(3)) - G["rel_error"]) < 6e-8
    assert "Check on the V operator Norm (should be zero) pure: 0.00218351" in out   # tests/sphere_translation.output
    for i in (1, 2):
        assert "OK rigid translation %d" % i in out
    for i in (3, 4, 5):
        assert "OK rigid rotation %d" % i in out
    ratio = float(re.search(r"run frame 1 velocity ratio ([-0-9.e+]+)", out).group(1))
    assert abs(ratio + 1.0) < 2e-2
    d0 = fe.vector_block_read(os.path.join(str(tmp_path), "stokes_rigid_displ_0.bin"))
    N = len(d0) // 3
    assert np.abs(d0[:N] - 0.1 * G["rigid_velocity_0"]).max() < 1e-7
    assert os.path.exists(os.path.join(str(tmp_path), "rotation_matrix_1.bin"))
    # ---- the C++ Heun branch on the device against the Python frame loop (same grids, same parameters), file by file
    import bemstokes_b200 as bb
    pdir = os.path.join(str(tmp_path), "heun_py")
    os.makedirs(pdir)
    ph = bb.BEMProblem()
    ph.quadrature_order, ph.singular_quadrature_order = 8, 10
    ph.grid_type, ph.res_strategy, ph.solve_directly = "Real", "Heun", True
    ph.input_grid_path, ph.input_grid_base_name, ph.input_grid_format = MESHES, "sphere_translation_", "msh"
    ph.n_frames, ph.output_dir = 2, pdir
    ph.log = lambda *_: None
    ph.run(0, 0)
    for name in ("4_6_rigid_velocities_0.bin", "stokes_rigid_vel_0.bin", "stokes_rigid_displ_0.bin", "rotation_matrix_0.bin",
                 "stokes_forces_0.bin", "total_velocities_0.bin"):
        a = fe.vector_block_read(os.path.join(hdir, name))
        b = fe.vector_block_read(os.path.join(pdir, name))
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max()), name
    # the punctual velocities are those of the second (corrector) solve, not of the Heun mean (ref 4784-4789)
    m = re.search(r"heun mean velocity ([-0-9.e+]+) predictor ([-0-9.e+]+) last solve ([-0-9.e+]+)", out)
    mean, pred, last = (float(m.group(k)) for k in (1, 2, 3))
    assert abs(mean - 0.5 * (pred + last)) < 1e-12 and abs(last + pred) < 2e-2 * abs(pred)
    vel = fe.vector_block_read(os.path.join(hdir, "stokes_rigid_vel_0.bin"))
    assert abs(vel[0] - last) < 1e-9
    ph.close()
