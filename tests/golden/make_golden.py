#!/usr/bin/env python
"""Extract the reference's own golden numbers for the hot path into small JSON fixtures.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/reference_goldens.json, tests/golden/singular_quadrature_table.json and copies the few
small meshes the parity tests need into tests/golden/meshes/ (mesh files are input DATA, not source code).
Every value is parsed verbatim from a reference test output; the file:line of each is recorded.
"""
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lines(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read().split("\n")


def mat3(ls, start):
    return [[float(x) for x in ls[start + r].split()] for r in range(3)]


def find(ls, needle, nth=0):
    hits = [i for i, l in enumerate(ls) if needle in l]
    return hits[nth]


def main():
    g = {}
    a = lines("tests/alpha_test.output")
    i0, i1 = find(a, "Test on V", 0), find(a, "Test on V", 1)
    k0, k1 = find(a, "Test on K", 0), find(a, "Test on K", 1)
    g["alpha_test"] = {
        "source": "tests/alpha_test.output:%d-%d,%d-%d" % (i0 + 1, k0 + 4, i1 + 1, k1 + 4),
        "setup": "sphere_half_refined_0.inp, Gauss 8, Mixed singular order 10, node 0 (file vertex 1), free-space",
        "Q1": {"V": mat3(a, i0 + 1), "K": mat3(a, k0 + 1)},
        "Q2": {"V": mat3(a, i1 + 1), "K": mat3(a, k1 + 1)},
    }
    d = lines("tests/dof_renumbering.output")
    dv, dk = find(d, "Test on V", 0), find(d, "Test on K", 0)
    g["dof_renumbering"] = {"source": "tests/dof_renumbering.output:%d-%d" % (dv + 1, dk + 4),
                            "setup": "sphere_coarse_0.inp 6 cells, Q2 (78 DoF), Gauss 8, QIterated(10,2); node = file vertex 2",
                            "V": mat3(d, dv + 1), "K": mat3(d, dk + 1)}

    def vnorm(rel):
        ls = lines(rel)
        i = find(ls, "Check on the V operator Norm (should be zero):")
        surf = find(ls, "The Mass (Surface) of the entire system is")
        return {"source": "%s:%d" % (rel, i + 1), "Vn_linf": float(ls[i].split(":")[1]),
                "surface": float(ls[surf].split(":")[1])}

    g["Vn_free"] = vnorm("tests/rigidity_sphere.output")
    g["Vn_free_surface"] = vnorm("tests/reflected_kernel_test_stresses.output")
    g["Vn_no_slip"] = vnorm("tests/wall_kernel_test_velocity.output")
    vg = lines("tests/V_test_with_Green.output")
    hits = [l for l in vg if "Check on the V operator Norm (should be zero):" in l]
    g["V_test_with_Green"] = {"source": "tests/V_test_with_Green.output", "Vn_linf": [float(h.split(":")[1]) for h in hits]}

    p = lines("tests/minimum_preconditioner_test_no_box.output")
    its = [int(l.split(":")[1]) for l in p if "Iterations needed to solve monolithic" in l]
    g["gmres_iterations_no_box"] = {"source": "tests/minimum_preconditioner_test_no_box.output (last 6 lines)",
                                    "ILU": its[0], "Jacobi": its[1], "AMG": its[2],
                                    "setup": "sphere_half_refined_0.inp Q1, ImposedForce, imposed component 1, tol 1e-10"}
    r = lines("tests/rigidity_sphere.output")
    fc = [l for l in r if l.startswith("FINAL CHECK 0")]
    g["rigidity_sphere_final_check0"] = {"source": "tests/rigidity_sphere.output", "linf": [float(l.split()[3]) for l in fc]}
    st = lines("tests/sphere_translation.output")
    err = [l for l in st if l.startswith("ERROR on rigid translation 0")][0]
    nums = [float(x) for x in re.findall(r"[-0-9.e]+", err.split(":")[1])]
    g["sphere_translation"] = {"source": "tests/sphere_translation.output (ERROR line = expected text) + tests/sphere_translation.cc:60-75",
                               "setup": "grid_test/sphere_translation_{0,1}.msh, Q1, grid Real, shape velocity = (x_1 - x_0)/time_step, time_step 0.1",
                               "rigid_velocity_0": nums[0], "exact": nums[1], "rel_error": nums[2],
                               "Vn_linf": float([l for l in st if "Check on the V operator Norm (should be zero):" in l][0].split(":")[1]),
                               "surface": float([l for l in st if "The Mass (Surface) of the entire system is" in l][0].split(":")[1])}
    sr = lines("tests/sphere_rotation.output")
    g["sphere_rotation"] = {"source": "tests/sphere_rotation.output + tests/sphere_rotation.cc:30,62,99-102",
                            "setup": "grid_test/sphere_rotation_{0,1}.msh, grid Real, exact omega = 2 pi/120/time_step about x, tol 1e-2",
                            "Vn_linf": float([l for l in sr if "Check on the V operator Norm (should be zero):" in l][0].split(":")[1]),
                            "omega_exact": 2 * 3.141592653589793 / 120 / 0.1, "tol": 1e-2,
                            "ok_lines": len([l for l in sr if l.startswith("OK rigid")])}
    # field evaluation: tests/test_bie_4.output prints, for the rotation modes, the double-layer
    # potential at the interior point (0.1, 0.1, 0.1) and the squared distance to the test's naive expectation (the
    # mode's value at node 0); translations and all exterior points (4, 4, 4) print "OK".  The grid is refined twice
    # there, but DL[rigid motion] = that motion inside a closed surface holds on any closed grid.
    tb = lines("tests/test_bie_4.output")
    modes = {}
    for k, l in enumerate(tb):
        if l.startswith("Test on the ") and "th rigid mode" in l:
            i = int(l.split()[3][0])
            blk = tb[k + 1:k + 8]
            mode0 = [float(t) for t in blk[0].split("=")[1].split()]
            e = {"mode_at_node0": mode0, "interior_ok": any(b.startswith("OK interior") for b in blk),
                 "exterior_ok": any(b.startswith("OK exterior") for b in blk)}
            ux = [b for b in blk if b.startswith("ux =")]
            if ux:
                uz = [b for b in blk if b.startswith("uz =")][0].split("=")[1].split()
                e["interior"] = [float(ux[0].split("=")[1]), float([b for b in blk if b.startswith("uy =")][0].split("=")[1]), float(uz[0])]
                e["interior_sq_dist_to_node0_value"] = float(uz[1])
            modes[str(i)] = e
    g["test_bie_4"] = {"source": "tests/test_bie_4.output + tests/test_bie_4.cc:18-20 (tol_int 6e-2, tol_ext 1e-5)",
                       "setup": "sphere_half_refined_0.inp (refined twice in the reference), free-space kernel, u = N_rigid[i], f = 0, "
                                "points (0.1,0.1,0.1) and (4,4,4)", "tol_int": 6e-2, "tol_ext": 1e-5, "modes": modes}
    t2 = lines("tests/test_bie_2.output")
    g["test_bie_2"] = {"source": "tests/test_bie_2.output + tests/test_bie_2.cc:19-20", "tol": 1e-3,
                       "setup": "single layer of f = normal_vector at (0.1,0.1,0.1) and (4,4,4): both below tol",
                       "interior_ok": any(l.startswith("OK interior") for l in t2), "exterior_ok": any(l.startswith("OK exterior") for l in t2)}
    dl = lines("tests/imposed_rotation_test_on_dilated_sphere.output")
    g["dilated_sphere"] = {"source": "tests/imposed_rotation_test_on_dilated_sphere.output + .cc:28-34,56-67,84-90",
                           "setup": "sphere_half_refined_0.inp scaled by L = 10 and shifted by 34.913639 per axis, force pole "
                                    "Baricenter, ImposedForce unit torque i = 3..5, exact omega = 1/(8 pi L^3), tol 3e-2",
                           "L": 10.0, "shift": 34.913639, "tol": 3e-2,
                           "surface": float([l for l in dl if "The Mass (Surface) of the entire system is" in l][0].split(":")[1]),
                           "Vn_linf": float([l for l in dl if "Check on the V operator Norm (should be zero):" in l][0].split(":")[1]),
                           "ok_lines": len([l for l in dl if "OK OMEGA_" in l])}
    vq = lines("tests/V_test_with_Green_Q2.output")
    g["V_test_with_Green_Q2"] = {"source": "tests/V_test_with_Green_Q2.output:13,20,38,45 (parameters_test_alpha_box_ref_quadrature.prm: "
                                           "Gauss 15, singular order 20)",
                                 "surface": [float(l.split(":")[1]) for l in vq if "The Mass (Surface) of the entire system is" in l],
                                 "Vn_linf": [float(l.split(":")[1]) for l in vq if "Check on the V operator Norm (should be zero) pure:" in l]}
    for name in ("V_test_with_Green_cube", "V_test_with_Green_Q2_cube"):
        vc = lines("tests/%s.output" % name)
        g[name] = {"source": "tests/%s.output (grid_test/sphere_0.inp refined globally WITHOUT manifold: a cube of surface 8)" % name,
                   "surface": [float(l.split(":")[1]) for l in vc if "The Mass (Surface) of the entire system is" in l],
                   "Vn_linf": [float(l.split(":")[1]) for l in vc if "Check on the V operator Norm (should be zero) pure:" in l]}
    # rigidity_spiral / rigidity_flagellum: ImposedForce unit loads i = 0..5 on a helical body; for every j != i the test
    # prints "OK" (|U_j/U_i| < 6e-3) or "ratio U_j U_i": the 6x6 mobility matrix to six digits
    import re as _re
    for name, grid in (("rigidity_spiral", "spiral_0.msh"), ("rigidity_flagellum", "flagellum_0.msh")):
        out = lines("tests/%s.output" % name)
        marks = [l.strip() for l in out if l.strip() == "OK" or _re.fullmatch(r"[-0-9.e+]+ [-0-9.e+]+ [-0-9.e+]+", l.strip())]
        assert len(marks) == 30, (name, len(marks))
        cols = []
        for i in range(6):
            col = {}
            js = [j for j in range(6) if j != i]
            for j, mk in zip(js, marks[5 * i:5 * i + 5]):
                if mk == "OK":
                    col[str(j)] = "OK"
                else:
                    ratio, uj, ui = [float(t) for t in mk.split()]
                    col[str(j)] = uj
                    col[str(i)] = ui
            cols.append(col)
        g[name] = {"source": "tests/%s.output + tests/%s.cc (tol 6e-3, grid_test/%s, ImposedForce, pole Origin)" % (name, name, grid),
                   "grid": grid, "tol": 6e-3,
                   "surface": float([l for l in out if "The Mass (Surface) of the entire system is" in l][0].split(":")[1]),
                   "Vn_linf": float([l for l in out if "Check on the V operator Norm (should be zero) pure:" in l][0].split(":")[1]),
                   "mobility_columns": cols}
    # motility_spiral / motility_flagellum: ImposedVelocity unit rigid velocity i; for every j != i "OK" or |F_j/F_i|
    for name, grid in (("motility_spiral", "spiral_0.msh"), ("motility_flagellum", "flagellum_0.msh")):
        out = lines("tests/%s.output" % name)
        marks = [l.strip() for l in out if l.strip() == "OK" or _re.fullmatch(r"[-0-9.e+]+( [-0-9.e+]+ [-0-9.e+]+)?", l.strip())]
        assert len(marks) == 30, (name, len(marks))
        # an entry is "OK", the ratio |F_j/F_i|, or [ratio, F_j, F_i] (the two tests print differently)
        conv = lambda m: m if m == "OK" else ([float(t) for t in m.split()] if " " in m else float(m))
        g[name] = {"source": "tests/%s.output + tests/%s.cc (tol 6e-3, grid_test/%s, ImposedVelocity)" % (name, name, grid),
                   "grid": grid, "tol": 6e-3,
                   "force_ratio_columns": [[conv(m) for m in marks[5 * i:5 * i + 5]] for i in range(6)]}
    bt = lines("tests/baricenter_torus.output")
    g["baricenter_torus"] = {"source": "tests/baricenter_torus.output:9-11 (grid_test/torus_0.inp)",
                             "surface": float([l for l in bt if "The Mass (Surface) of the entire system is" in l][0].split(":")[1]),
                             "center_of_mass": [float(t) for t in [l for l in bt if l.startswith("Center of mass position")][0].split("=")[1].split()]}
    mr = lines("tests/motility_rotation_spiral.output")
    pairs = [l for l in mr if " --- " in l]
    assert len(pairs) == 6
    g["motility_rotation_spiral"] = {
        "source": "tests/motility_rotation_spiral.output (last 6 lines: F_j(frame 30) : R F_j(frame 0) --- U_j : U_j) + .cc",
        "setup": "grid_test/spiral_0.msh and spiral_30.msh, ImposedVelocity unit omega_x (i = 3), solved directly; the frame-0 "
                 "forces are rotated by the quaternion (cos(a/2), sin(a/2), 0, 0), a = -2 pi 30/120",
        "forces_frame30": [float(l.split(":")[0]) for l in pairs],
        "forces_frame0_rotated": [float(l.split(":")[1].split("---")[0]) for l in pairs],
        "surface": [float(l.split(":")[1]) for l in mr if "The Mass (Surface) of the entire system is" in l],
        "Vn_linf": [float(l.split(":")[1]) for l in mr if "Check on the V operator Norm (should be zero) pure:" in l]}
    g["imposed_rotation"] = {"source": "tests/imposed_rotation_test_on_sphere.cc:28-31", "omega": 1.0 / (8 * 3.141592653589793),
                             "tol": 1.2e-3}
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(g, f, indent=1)

    # ---- singular quadrature error table -----------------------------------------------------------
    cc = "\n".join(lines("tests/integrate_one_over_r_Q2.cc"))
    exact = {}
    for m in re.finditer(r"v\[(\d)\]\[(\d)\]\[(\d)\]\s*=\s*([-0-9.eE]+)\s*;", cc):
        exact["%s,%s,%s" % (m.group(1), m.group(2), m.group(3))] = float(m.group(4))
    out = lines("tests/integrate_one_over_r_Q2.output")
    table = []
    order, vertex = None, None
    vid = -1
    for l in out:
        m = re.match(r"\s*=+Quadrature Order: (\d+)", l)
        if m:
            order, vid = int(m.group(1)), -1
            continue
        m = re.match(r"\s*=+Vertex: ([-0-9.e]+) ([-0-9.e]+)", l)
        if m:
            vid += 1
            continue
        m = re.match(r"f\(x,y\) = x\^(\d) y\^(\d), Errors = Telles ([-0-9.e+]+), LWGaussOneR ([-0-9.e+]+), "
                     r"QIteraded\(QGauss\) ([-0-9.e+]+), QDuffy ([-0-9.e+]+), exact value ([-0-9.e+]+)", l)
        if m:
            table.append([order, vid, int(m.group(1)), int(m.group(2))] + [float(m.group(k)) for k in range(3, 8)])
    with open(os.path.join(HERE, "singular_quadrature_table.json"), "w") as f:
        json.dump({"source": "tests/integrate_one_over_r_Q2.output (+ exact constants of integrate_one_over_r_Q2.cc)",
                   "columns": ["order", "support_point", "i", "j", "err_telles", "err_lw", "err_qiterated", "err_duffy",
                               "exact_printed"],
                   "exact": exact, "rows": table}, f)

    # ---- meshes (input data) -------------------------------------------------------------------------
    os.makedirs(os.path.join(HERE, "meshes"), exist_ok=True)
    for rel in ["tests/grid_test/sphere_half_refined_0.inp", "tests/grid_test/sphere_0.inp",
                "tests/grid_test/sphere_coarse_0.inp", "debug_grids/sphere_mesh_3d_0.msh",
                "debug_grids/prolate_spheroid_lambda_2_ref_0.msh", "debug_grids/sphere_very_refined_0.inp",
                "debug_grids/sphere_very_very_refined_0.inp", "debug_grids/sphere_2.inp",
                "tests/grid_test/sphere_translation_0.msh", "tests/grid_test/sphere_translation_1.msh",
                "tests/grid_test/sphere_rotation_0.msh", "tests/grid_test/sphere_rotation_1.msh",
                "tests/grid_test/spiral_0.msh", "tests/grid_test/flagellum_0.msh", "tests/grid_test/torus_0.inp",
                "tests/grid_test/spiral_30.msh"]:
        src = os.path.join(REF, rel)
        if os.path.exists(src):
            dst = os.path.join(HERE, "meshes", os.path.basename(rel))
            shutil.copy(src, dst)
            os.chmod(dst, 0o644)
        else:
            print("missing", rel)
    # ---- parameter files (input data of the front-end tests) ------------------------------------------
    for rel in ["tests/parameters_test_alpha_box.prm", "tests/parameters_test_alpha_box_ref_quadrature.prm"]:
        src = os.path.join(REF, rel)
        if os.path.exists(src):
            dst = os.path.join(HERE, os.path.basename(rel))
            shutil.copy(src, dst)
            os.chmod(dst, 0o644)
        else:
            print("missing", rel)
    print("wrote goldens:", len(table), "quadrature rows,", len(exact), "exact constants")


if __name__ == "__main__":
    main()
