"""Host front-end of the reference without deal.II / deal2lkit: the parameter-file subset the hot
path consumes, the multi-frame swimmer loop, the quaternion integrator and the result files.  Plain host code around
the C-ABI calls of `BEMProblem`; nothing here touches the device directly.

ref (source/bem_stokes.cc unless stated):
    declare_parameters 207-476, ParameterAcceptor::initialize main.cc:37   -> parse_prm, FrameLoop.parse_parameters
    read_input_mesh_file 496-523                                            -> FrameLoop.read_input_mesh_file
    apply_rotation_along_axis 846-878                                       -> apply_rotation_along_axis
    project_shape_velocities 2120-2246 (isoparametric branch)               -> FrameLoop.project_shape_velocities
    compute_euler_vector 2247-2431 (mesh-file branch, body-only)            -> FrameLoop.compute_euler_vector
    compute_rotation_matrix_from_quaternion 4512-4525, update_rotation_matrix 4527-4720
    update_system_state 4725-4846                                           -> FrameLoop.update_system_state
    save/read_rotation_matrix 5106-5131, output_save_stokes_results 5140-5316 (the .bin / .txt files; no .vtu)
    run 5636-5888, reinit_for_new_time 5890-5920                            -> FrameLoop.run
Out of scope here: walls / bounding box / cylinder meshing, IGES, the flagellum handler, squirmer velocity files,
Galerkin assembly (SURVEY §2 marks them outside the hot path)."""
import math
import os
import re

import numpy as np

from .mesh import read_mesh


# ---------------------------------------------------------------------------------------------------------------
# deal.II ParameterHandler text format: `subsection X` ... `set Key = value` ... `end`, `#` comments
# ---------------------------------------------------------------------------------------------------------------
def parse_prm(src):
    """Nested dict {section: {...}, key: "value"} from a .prm file path or its text."""
    if "\n" not in src and os.path.exists(src):
        with open(src) as f:
            src = f.read()
    root = {}
    stack = [root]
    pending = ""
    for raw in src.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.endswith("\\"):  # continuation
            pending += line[:-1]
            continue
        line, pending = pending + line, ""
        if line.lower().startswith("subsection "):
            name = line[len("subsection "):].strip()
            stack.append(stack[-1].setdefault(name, {}))
        elif line.lower() == "end":
            if len(stack) == 1:
                raise ValueError("unbalanced 'end' in parameter file")
            stack.pop()
        elif line.lower().startswith("set "):
            key, _, val = line[4:].partition("=")
            stack[-1][key.strip()] = val.strip()
        else:
            raise ValueError("cannot parse parameter line: %r" % raw)
    if len(stack) != 1:
        raise ValueError("unterminated subsection in parameter file")
    return root


def _bool(s):
    s = s.strip().lower()
    if s in ("true", "yes", "on", "1"):
        return True
    if s in ("false", "no", "off", "0"):
        return False
    raise ValueError("not a bool: %r" % s)


def _floats(s):
    return tuple(float(t) for t in s.split(","))


def _fe_degree(s):
    m = re.search(r"FE_Q(?:<[^>]*>)?\((\d+)\)", s)
    if not m:
        raise ValueError("only FE_Q systems are supported: %r" % s)
    return int(m.group(1))


# reference key -> (attribute, converter); the subset that reaches the hot path of a body-only problem
PARAMETERS = {
    "Total number of frames": ("n_frames", int),
    "Delta between frames": ("delta_frame", int),
    "Gmres restart evert": ("gmres_restart", int),
    "Consider rigid rotations": ("bool_rot", _bool),
    "Consider rigid displacement to move the swimmer": ("bool_dipl", _bool),
    "Consider rigid displacement x to move the swimmer": ("bool_dipl_x", _bool),
    "Consider rigid displacement y to move the swimmer": ("bool_dipl_y", _bool),
    "Consider rigid displacement z to move the swimmer": ("bool_dipl_z", _bool),
    "Monolithic resolurion strategy": ("monolithic_bool", _bool),
    "Use a direct resolution strategy": ("solve_directly", _bool),
    "Grid": ("grid_type", str),
    "Singular quadrature kind": ("singular_quadrature_type", str),
    "Singular quadrature order": ("singular_quadrature_order", int),
    "Force Pole to be used": ("force_pole_kind", str),
    "Force Pole Point Setting": ("force_arbitrary_point", _floats),
    "Type of preconditioner to be used": ("preconditioner_type", str),
    "Use a bandwith preconditioner": ("bandwith_preconditioner", _bool),
    "Bandwith for the preconditioner": ("bandwith", int),
    "Use alpha for the internal problem": ("use_internal_alpha", _bool),
    "Input path to grid": ("input_grid_path", str),
    "Input grid base name": ("input_grid_base_name", str),
    "Input grid format": ("input_grid_format", str),
    "Time Integration": ("res_strategy", str),
    "Imposed Component for Non Real Simulation": ("imposed_component", int),
    "Time interval between frames": ("time_step", float),
    "Reflect the kernel": ("reflect_kernel", _bool),
    "Use no slip kernel": ("no_slip_kernel", _bool),
    "Wall 0 spans": ("wall_spans_0", _floats),
    "Wall center position wall 0": ("wall_position_0", _floats),
    "Use state from previous frame": ("use_previous_state", _bool),
    "Scaling for monolithic assembling": ("assemble_scaling", float),
    "Print extra debug information": ("extra_debug_info", _bool),
    "Create a bounding box": ("create_box_bool", _bool),
}
for _k in range(4):
    PARAMETERS["Initial quaternion value q[%d]" % _k] = ("initial_quaternion_%d" % _k, float)
for _k in range(8):
    PARAMETERS["Wall %d bool" % _k] = ("wall_bool_%d" % _k, _bool)


# ---------------------------------------------------------------------------------------------------------------
# deal.II Vector<double>::block_write / block_read:  "<size>\n[" + raw doubles + "]"
# ---------------------------------------------------------------------------------------------------------------
def vector_block_write(path, v):
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
    with open(path, "wb") as f:
        f.write(("%d\n[" % v.size).encode())
        f.write(v.tobytes())
        f.write(b"]")


def vector_block_read(path):
    with open(path, "rb") as f:
        data = f.read()
    nl = data.index(b"\n")
    n = int(data[:nl])
    if data[nl + 1:nl + 2] != b"[" or data[nl + 2 + 8 * n:nl + 3 + 8 * n] != b"]":
        raise ValueError("%s is not a deal.II block_write file" % path)
    return np.frombuffer(data, dtype=np.float64, count=n, offset=nl + 2).copy()


# ---------------------------------------------------------------------------------------------------------------
# rigid rotations
# ---------------------------------------------------------------------------------------------------------------
def compute_rotation_matrix_from_quaternion(q):
    """ref: 4512-4525."""
    q0, q1, q2, q3 = q
    return np.array([[1. - 2 * (q3 * q3 + q2 * q2), -2 * q0 * q3 + 2 * q1 * q2, 2 * q0 * q2 + 2 * q1 * q3],
                     [2 * q0 * q3 + 2 * q1 * q2, 1. - 2 * (q3 * q3 + q1 * q1), -2 * q0 * q1 + 2 * q3 * q2],
                     [-2 * q0 * q2 + 2 * q1 * q3, 2 * q0 * q1 + 2 * q3 * q2, 1. - 2 * (q1 * q1 + q2 * q2)]])


def update_rotation_matrix(rotation, omega, dt, forward_euler=True, theta=0.5, log=None):
    """One step of the quaternion integrator (ref: 4527-4720): quaternion of `rotation`, qdot = S^-1 (0, omega)/2,
    forward Euler (or the theta scheme, whose 4x4 system the reference hands to GMRES - solved directly here),
    renormalisation, new rotation matrix.  Returns the new matrix; orthogonality defects above 1e-7 are reported
    through `log` with the reference's messages."""
    R = np.asarray(rotation, dtype=float)
    q = np.zeros(4)
    q[0] = math.sqrt(1. + R[0, 0] + R[1, 1] + R[2, 2]) / 2
    q[1] = 1 / q[0] * 0.25 * (R[2, 1] - R[1, 2])
    q[2] = 1 / q[0] * 0.25 * (R[0, 2] - R[2, 0])
    q[3] = 1 / q[0] * 0.25 * (R[1, 0] - R[0, 1])
    q /= math.sqrt(float(q @ q))
    op = np.array([0., omega[0], omega[1], omega[2]])
    S = 0.5 * np.array([[q[0], -q[1], -q[2], -q[3]],
                        [q[1], q[0], q[3], -q[2]],
                        [q[2], -q[3], q[0], q[1]],
                        [q[3], q[2], -q[1], q[0]]])
    qdot = S @ op
    if forward_euler:
        q = q + dt * qdot
    else:
        h = theta * dt * 0.5
        A = np.array([[1. + h * op[0], h * op[1], h * op[2], h * op[3]],
                      [-h * op[1], 1. + h * op[0], -h * op[3], h * op[2]],
                      [-h * op[2], h * op[3], 1. + h * op[0], -h * op[1]],
                      [-h * op[3], -h * op[2], h * op[1], 1. + h * op[0]]])
        q = np.linalg.solve(A, q + (1 - theta) * dt * qdot)
    q /= math.sqrt(float(q @ q))
    Rn = compute_rotation_matrix_from_quaternion(q)
    defect = Rn.T @ Rn
    for i in range(3):
        for j in range(3):
            d = abs(defect[i, j] - (1.0 if i == j else 0.0))
            if d >= 1e-7 and log is not None:
                log("Something Wrong in Rotations, %s the diagonal %g" % ("on" if i == j else "out", d))
    return Rn


def apply_rotation_along_axis(p, axis, angle):
    """Rodrigues rotation of point p about the unit `axis` (ref: 846-878)."""
    a = np.asarray(axis, dtype=float)
    c, s = math.cos(angle), math.sin(angle)
    R = np.array([[c + a[0] * a[0] * (1 - c), a[0] * a[1] * (1 - c) - a[2] * s, a[0] * a[2] * (1 - c) + a[1] * s],
                  [a[0] * a[1] * (1 - c) + a[2] * s, c + a[1] * a[1] * (1 - c), a[1] * a[2] * (1 - c) - a[0] * s],
                  [a[0] * a[2] * (1 - c) - a[1] * s, a[1] * a[2] * (1 - c) + a[0] * s, c + a[2] * a[2] * (1 - c)]])
    return R @ np.asarray(p, dtype=float)


# ---------------------------------------------------------------------------------------------------------------
# frame loop (mixed into BEMProblem)
# ---------------------------------------------------------------------------------------------------------------
class FrameLoop:
    """Multi-frame workflow of BEMProblem::run for a body-only swimmer described by one mesh file per frame."""

    def _init_frontend(self):
        self.n_frames, self.delta_frame = 120, 1
        self.bool_rot, self.bool_dipl = True, False
        self.bool_dipl_x = self.bool_dipl_y = self.bool_dipl_z = False
        self.input_grid_path, self.input_grid_base_name, self.input_grid_format = "../debug_grids/", "sphere_mesh_3d_", "msh"
        self.res_strategy = "Forward"
        self.time_step = 0.1
        self.force_arbitrary_point = (1.0, 0.0, 0.0)
        self.use_previous_state = False
        self.extra_debug_info = False
        self.create_box_bool = False
        for k in range(8):
            setattr(self, "wall_bool_%d" % k, False)
        self.initial_quaternion = [1.0, 0.0, 0.0, 0.0]
        self.rotation_matrix = np.eye(3)
        self.output_dir = "."
        self.log = print
        self.frame_results = []

    # ---- parameters -------------------------------------------------------------------------------------------
    def parse_parameters(self, prm, strict=False):
        """Apply a parsed parameter tree (or a .prm path / text).  Keys outside the hot-path subset are ignored
        unless `strict`; the sections follow ParameterAcceptor's names (main.cc:37, bem_stokes.h:414-419)."""
        tree = prm if isinstance(prm, dict) else parse_prm(prm)
        main = None
        for name, sec in tree.items():
            if isinstance(sec, dict) and name.startswith("BEMStokes::BEMProblem"):
                main = sec
        if main is None:
            raise ValueError("no 'BEMStokes::BEMProblem<3>' subsection in the parameter file")
        for key, val in main.items():
            if isinstance(val, dict):
                if key == "Solver":
                    self.solver_control.max_steps = int(val.get("Max steps", self.solver_control.max_steps))
                    self.solver_control.tolerance = float(val.get("Tolerance", self.solver_control.tolerance))
                continue
            if key in PARAMETERS:
                attr, conv = PARAMETERS[key]
                setattr(self, attr, conv(val))
            elif strict:
                raise KeyError("parameter %r is not part of the B200 hot path" % key)
        self.initial_quaternion = [getattr(self, "initial_quaternion_%d" % k, self.initial_quaternion[k]) for k in range(4)]
        q = tree.get("Internal Quadrature", {})
        if q:
            if q.get("Quadrature to generate", "gauss") != "gauss" or int(q.get("Number of repetitions", 1)) != 1:
                raise ValueError("only plain Gauss rules are supported for 'Internal Quadrature'")
            self.quadrature_order = int(q.get("Quadrature order", self.quadrature_order))
        fs, fm = tree.get("Finite Element Stokes", {}), tree.get("Finite Element Mapping", {})
        if "Finite element space" in fs:
            self.fe_degree = _fe_degree(fs["Finite element space"])
        if "Finite element space" in fm:
            self.map_degree = _fe_degree(fm["Finite element space"])
        self.convert_bool_parameters()
        return self

    def convert_bool_parameters(self):
        """ref: 5564-5583 - collects the wall flags; this front-end runs body-only problems (image kernels model the
        first wall without meshing it)."""
        self.wall_bool = [bool(getattr(self, "wall_bool_%d" % k)) for k in range(8)]
        if self.force_pole_kind == "Origin":
            self.force_pole = (0.0, 0.0, 0.0)
        elif self.force_pole_kind == "Point":
            self.force_pole = tuple(self.force_arbitrary_point)
        elif self.force_pole_kind not in (None, "Baricenter"):   # Baricenter: the pre-pass computes the surface centroid
            raise ValueError("unknown force pole %r" % (self.force_pole_kind,))
        return self

    # ---- geometry per frame -------------------------------------------------------------------------------------
    def read_input_mesh_file(self, frame):
        """ref: 496-523 (material ids forced to 0: body-only)."""
        path = os.path.join(self.input_grid_path, "%s%d.%s" % (self.input_grid_base_name, frame, self.input_grid_format))
        return read_mesh(path)

    def compute_euler_vector(self, frame, consider_displacements=True):
        """Nodes of frame `frame` rotated by the current rotation matrix (+ accumulated rigid displacement), as the
        component-major euler vector (ref: 2247-2431, mesh-file branch; the walls are not meshed here).  The reference
        adds `rigid_displacements_for_sim` inside its per-node loop, i.e. once per body node; that is only reachable
        with 'Consider rigid displacement...' switched on and is mirrored as a single addition."""
        m = self.read_input_mesh_file(frame)
        if m.n_nodes != self.mesh.n_nodes or m.n_cells != self.mesh.n_cells:
            raise ValueError("frame %d has a different topology from the reference grid" % frame)
        pts = m.nodes @ self.rotation_matrix.T
        euler = np.ascontiguousarray(pts.T.reshape(-1))
        if consider_displacements and self.bool_dipl:
            euler = euler + self.rigid_displacements_for_sim
        return euler

    def _set_euler(self, euler):
        nodes = euler.reshape(3, -1).T.copy()
        self.mesh = type(self.mesh)(nodes, self.mesh.conn, self.mesh.degree)
        self.map_mesh = self.mesh
        self.update_geometry()

    def project_shape_velocities(self, frame):
        """Finite-difference shape velocity between two frames (ref: 2120-2137, isoparametric branch)."""
        if self.fe_degree != self.map_degree:
            raise NotImplementedError("L2 projection of shape velocities for non isoparametric spaces")
        self.shape_velocities = (self.next_euler_vec - self.euler_vec) / self.time_step
        return self.shape_velocities

    # ---- state update ---------------------------------------------------------------------------------------------
    def update_system_state(self, compute, frame, consider_rotations, consider_displacements, res_system):
        """ref: 4725-4846 ('Forward'; Heun's corrector swaps in the backed-up state)."""
        n, nr = self.n_dofs, self.num_rigid
        if res_system == "Heun" and self.res_strategy == "Heun":
            self.rotation_matrix = self.old_rotation_matrix.copy()
            self.rigid_displacements_for_sim = self.old_rigid_displacements_for_sim.copy()
            self.rigid_velocities = 0.5 * self.rigid_velocities + 0.5 * self.old_rigid_velocities
        elif res_system == "Forward" and self.res_strategy == "Heun":
            self.old_rigid_velocities = self.rigid_velocities.copy()
            self.old_rotation_matrix = self.rotation_matrix.copy()
            self.old_rigid_displacements_for_sim = self.rigid_displacements_for_sim.copy()
        bary = self.baricenter_rigid_velocities
        self.rigid_puntual_velocities = np.zeros(n)
        for i in range(3):
            self.rigid_puntual_velocities += self.assemble_scaling * bary[i] * self.N_rigid[i]
        self.rigid_puntual_translation_velocities = self.rigid_puntual_velocities.copy()
        for i in range(3, nr):
            self.rigid_puntual_velocities += self.assemble_scaling * bary[i] * self.N_rigid[i]
        if consider_rotations:
            self.rotation_matrix = update_rotation_matrix(self.rotation_matrix, self.rigid_velocities[3:6], self.time_step,
                                                          log=self.log)
        self.next_rigid_puntual_displacements = self.time_step * self.rigid_puntual_translation_velocities
        if consider_displacements:
            N = n // 3
            for flag, c in ((self.bool_dipl_x, 0), (self.bool_dipl_y, 1), (self.bool_dipl_z, 2)):
                if flag:
                    self.rigid_displacements_for_sim[c * N:(c + 1) * N] += self.next_rigid_puntual_displacements[c * N:(c + 1) * N]
        return self

    # ---- files ----------------------------------------------------------------------------------------------------
    def save_rotation_matrix(self, rotation, frame):
        vector_block_write(os.path.join(self.output_dir, "rotation_matrix_%d.bin" % frame), np.asarray(rotation).reshape(-1))

    def read_rotation_matrix(self, frame):
        return vector_block_read(os.path.join(self.output_dir, "rotation_matrix_%d.bin" % frame)).reshape(3, 3)

    def output_save_stokes_results(self, cycle):
        """The .bin / .txt result files of ref 5264-5316 (deal.II block_write format, same names)."""
        d = self.output_dir
        w = lambda name, v: vector_block_write(os.path.join(d, name), v)
        w("stokes_forces_%d.bin" % cycle, self.stokes_forces)
        w("shape_velocities_%d.bin" % cycle, self.shape_velocities)
        w("total_velocities_%d.bin" % cycle, self.total_velocities)
        self.save_rotation_matrix(self.rotation_matrix, cycle)
        w("4_6_rigid_velocities_%d.bin" % cycle, self.rigid_velocities)
        w("4_6_overall_forces_%d.bin" % cycle, self.rigid_total_forces)
        w("stokes_rigid_displ_%d.bin" % cycle, self.next_rigid_puntual_displacements)
        w("stokes_rigid_vel_%d.bin" % cycle, self.rigid_puntual_velocities)
        w("euler_vec_%d.bin" % cycle, self.euler_vec)
        w("normal_vector%d.bin" % cycle, self.normal_vector)
        with open(os.path.join(d, "point_0_on_proc_0_displacement_frame_%d.txt" % cycle), "w") as f:
            f.write("%d " % cycle + "".join("%g " % self.rigid_puntual_displacements[0] for _ in range(3)) + "\n")

    # ---- the loop -------------------------------------------------------------------------------------------------
    def reinit_for_new_time(self, frame):
        """ref: 5890-5920."""
        self.log("preparing for new time")
        self.euler_vec = self.compute_euler_vector(frame % self.n_frames, True)
        self.next_euler_vec = np.zeros_like(self.euler_vec)
        self.rigid_puntual_velocities = np.zeros(self.n_dofs)

    def run(self, start_frame=0, end_frame=0):
        """ref: BEMProblem::run (5636-5888) for a body-only swimmer: per frame geometry -> device pre-pass -> shape
        velocities -> assembly -> (LU of the Direct preconditioner when asked) -> solve -> state update -> files."""
        if self.res_strategy not in ("Forward", "Heun"):
            raise NotImplementedError(self.res_strategy)
        self.convert_bool_parameters()
        self.rotation_matrix = compute_rotation_matrix_from_quaternion(self.initial_quaternion)
        self.set_mesh(self.read_input_mesh_file(start_frame % self.n_frames))   # read_domain
        self.reinit()
        n = self.n_dofs
        self.rigid_displacements_for_sim = np.zeros(n)
        self.rigid_puntual_displacements = np.zeros(n)
        self.wall_velocities = np.zeros(n)
        if start_frame != 0 and self.use_previous_state:
            self.log("getting old stuff")
            self.rigid_puntual_displacements = vector_block_read(
                os.path.join(self.output_dir, "stokes_rigid_displ_%d.bin" % (start_frame - 1)))
            self.rotation_matrix = self.read_rotation_matrix(start_frame - 1)
        self.euler_vec = self.compute_euler_vector(start_frame % self.n_frames, True)
        self.reassemble_preconditoner = True
        self.frame_results = []
        for i in range(start_frame, end_frame + 1, self.delta_frame):
            self.log("Analyzing frame = %d over %d" % (i, self.n_frames))
            self.next_euler_vec = self.compute_euler_vector((i + 1) % self.n_frames, True)
            self._solve_frame(i)
            if self.res_strategy == "Forward":
                self.update_system_state(True, i, self.bool_rot, self.bool_dipl, self.res_strategy)
            else:  # Heun: predictor state, geometry and solve at the next frame, corrector with the averaged velocity
                self.update_system_state(True, i, self.bool_rot, self.bool_dipl, "Forward")
                self.reinit_for_new_time((i + 1) % self.n_frames)
                self.next_euler_vec = self.compute_euler_vector((i + 2) % self.n_frames, True)
                self._solve_frame(i + 1)   # ref 5797: compute_center_of_mass_and_rigid_modes(i+1)
                self.update_system_state(True, i, self.bool_rot, self.bool_dipl, self.res_strategy)
            self.total_velocities = self.shape_velocities + self.rigid_puntual_velocities + self.wall_velocities
            self.rigid_puntual_displacements = self.next_rigid_puntual_displacements
            if self.res_strategy == "Forward" and not self.fused_assembly and self.keep_VK:
                # FINAL CHECK (5846-5869): P K P u_total - V f on the corrected operators
                Pu = self.tangential_projector_body(self.total_velocities)
                r = self.tangential_projector_body(self.K_matrix @ Pu) - self.V_matrix @ self.stokes_forces
                self.final_test = r
                self.log("FINAL CHECK %g : %g" % (np.abs(r).max(), np.linalg.norm(r)))
            self.output_save_stokes_results(i)
            self.frame_results.append({"frame": i, "rigid_velocities": self.rigid_velocities.copy(),
                                       "rigid_total_forces": self.rigid_total_forces.copy(),
                                       "rotation_matrix": self.rotation_matrix.copy(),
                                       "gmres_iterations": self.solver_control.last_step()})
            self.reinit_for_new_time((i + self.delta_frame) % self.n_frames)
        self.log("THE END")
        return self.frame_results

    def _solve_frame(self, i):
        self._set_euler(self.euler_vec)
        self.compute_center_of_mass_and_rigid_modes(i)
        self.compute_normal_vector()
        if self.grid_type != "Real":
            self.next_euler_vec = self.euler_vec.copy()
        self.project_shape_velocities(i)
        if self.grid_type != "Real":
            self.shape_velocities = np.zeros(self.n_dofs)
        self.log("%g" % np.linalg.norm(self.shape_velocities))
        self.log("Assembling")
        self.assemble_stokes_system(True)
        if self.reassemble_preconditoner and not self.solve_directly and self.preconditioner_type == "Direct":
            self.log("refactorizing direct_preconditioner")
            self.direct_trilinos_preconditioner.set_up(self.solver_control)
            self.direct_trilinos_preconditioner.initialize(self.monolithic_system_matrix if self.monolithic_bool else self.V_matrix)
            self.reassemble_preconditoner = False
        self.monolithic_solution = np.zeros(self.n_dofs + self.num_rigid)
        self.solve_system(self.monolithic_bool)
