"""Host-side mirror of the reference's hot-path interface, driving the C-ABI.

`BEMProblem` keeps the member names of BEMStokes::BEMProblem<3> (include/bem_stokes.h:106-660) that the hot
path and the reference's tests touch: assemble_stokes_system, solve_system, dirichlet_to_neumann_operator,
tangential_projector_body, V_matrix / K_matrix / monolithic_system_matrix (vmult + element reads),
monolithic_rhs / monolithic_solution, stokes_forces, rigid_velocities, rigid_total_forces, and the parameter
names of declare_parameters (source/bem_stokes.cc:207-476) as attributes.  `DirectPreconditioner` mirrors
include/direct_preconditioner.h:27-51, the kernel classes mirror include/kernel.h, free_surface_kernel.h and
no_slip_wall_kernel.h.  All arithmetic happens in libbemstokes_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .mesh import QuadMesh
from .frontend import FrameLoop
from .prepass import DevicePrepass, Prepass, gauss_1d


def _dp(a):
    return a.ctypes.data_as(_lib.c_double_p)


def _ip(a):
    return a.ctypes.data_as(_lib.c_int_p)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


_SING = {"Mixed": _lib.SING_MIXED, "Duffy": _lib.SING_DUFFY, "Telles": _lib.SING_TELLES}
_GRID = {"Real": _lib.GRID_REAL, "ImposedForce": _lib.GRID_IMPOSED_FORCE, "ImposedVelocity": _lib.GRID_IMPOSED_VELOCITY}


# ---------------------------------------------------------------------------------------------------------------
# Green kernels (point evaluation on the device through the same functions the assembly kernels inline)
# ---------------------------------------------------------------------------------------------------------------
class StokesKernel:
    """ref: include/kernel.h StokesKernel<3>: value_tens (G, rank 2), value_tens2 (W, rank 3)."""
    _type = _lib.KERNEL_FREE

    def __init__(self, eps=0.0, device=0):
        self.epsilon, self.device, self.wall_orientation = eps, device, 1

    def set_wall_orientation(self, o):
        self.wall_orientation = int(o)

    def _eval(self, p, p_image, want_w):
        p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        q = p if p_image is None else np.ascontiguousarray(np.atleast_2d(p_image), dtype=np.float64)
        n = p.shape[0]
        G = np.zeros((n, 3, 3))
        W = np.zeros((n, 3, 3, 3)) if want_w else None
        check(lib.bs_kernel_eval(self.device, self._type, self.epsilon, self.wall_orientation, n, _dp(p), _dp(q),
                                 _dp(G), _dp(W) if want_w else None))
        return G, W

    def value_tens(self, p):
        G, _ = self._eval(p, None, False)
        return G[0] if np.ndim(p) == 1 else G

    def value_tens2(self, p):
        _, W = self._eval(p, None, True)
        return W[0] if np.ndim(p) == 1 else W


class FreeSurfaceStokesKernel(StokesKernel):
    """ref: include/free_surface_kernel.h: value_tens_image / value_tens_image2."""
    _type = _lib.KERNEL_FREE_SURFACE

    def value_tens_image(self, p, p_image):
        G, _ = self._eval(p, p_image, False)
        return G[0] if np.ndim(p) == 1 else G

    def value_tens_image2(self, p, p_image):
        _, W = self._eval(p, p_image, True)
        return W[0] if np.ndim(p) == 1 else W


class NoSlipWallStokesKernel(FreeSurfaceStokesKernel):
    """ref: include/no_slip_wall_kernel.h."""
    _type = _lib.KERNEL_NO_SLIP


# ---------------------------------------------------------------------------------------------------------------
# matrix / preconditioner objects with the deal.II vmult concept
# ---------------------------------------------------------------------------------------------------------------
class DeviceMatrix:
    """Anything with vmult(dst, src) — what SolverGMRES::solve requires (include/operator.h:22-66)."""

    def __init__(self, problem, which):
        self._p, self.which = problem, which

    def m(self):
        r = C.c_int()
        check(lib.bs_matrix_size(self._p._ctx, self.which, C.byref(r), None))
        return r.value

    n = m

    def vmult(self, dst, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        assert dst.dtype == np.float64 and dst.flags.c_contiguous and dst.shape == src.shape
        if self._p.n_mpi_processes > 1:
            dst[...] = 0.0  # every rank writes its owned rows only; the host sums the slices
        if src.ndim == 1:
            check(lib.bs_vmult(self._p._ctx, self.which, _vp(src), _vp(dst)))
        else:
            check(lib.bs_vmult_multi(self._p._ctx, self.which, src.shape[0], _vp(src), _vp(dst)))
        self._p._allsum(dst)
        return dst

    def __matmul__(self, x):
        return self.vmult(np.zeros(np.shape(x)), x)

    def __call__(self, i, j):
        """Element read, like TrilinosWrappers::SparseMatrix::operator()(i,j)."""
        return float(self.entries([i], [j])[0])

    def entries(self, rows, cols):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        out = np.zeros(len(rows))
        check(lib.bs_get_entries(self._p._ctx, self.which, len(rows), _ip(rows), _ip(cols), _dp(out)))
        return out

    def to_dense(self):
        """Whole matrix on the host in reference ordering (parity tests, 'Save matrices as txt files')."""
        n = self.m()
        r, c = np.meshgrid(np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32), indexing="ij")
        return self.entries(r.reshape(-1), c.reshape(-1)).reshape(n, n)


class DirectPreconditioner:
    """ref: include/direct_preconditioner.h:27-51 — set_up(SolverControl, AdditionalData), initialize(matrix),
    vmult(dst, src).  The LU lives on the device (dense blocked LU with partial pivoting)."""

    def __init__(self):
        self._matrix = None
        self.kind = _lib.PREC_DIRECT

    def set_up(self, solver_control=None, additional_data=None):
        self.solver_control = solver_control

    def initialize(self, matrix, block_only=False):
        self._matrix = matrix
        self.kind = _lib.PREC_BLOCK_DIRECT if block_only else _lib.PREC_DIRECT
        check(lib.bs_precond_setup(matrix._p._ctx, matrix.which, self.kind, 0))

    def vmult(self, dst, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        check(lib.bs_precond_vmult(self._matrix._p._ctx, _vp(src), _vp(dst)))
        return dst


class SolverControl:
    def __init__(self, max_steps=1000, tolerance=1e-10):
        self.max_steps, self.tolerance = max_steps, tolerance
        self._last_step, self._last_value = 0, 0.0

    def last_step(self):
        return self._last_step

    def last_value(self):
        return self._last_value


# ---------------------------------------------------------------------------------------------------------------
class BEMProblem(FrameLoop):
    def __init__(self, device=0, rank=0, nranks=1, comm=None, stream=None):
        self.device, self.this_mpi_process, self.n_mpi_processes = device, rank, nranks
        self.comm, self.stream = comm, stream
        # parameters (names follow declare_parameters, bem_stokes.cc:207-476)
        self.fe_degree = 1
        self.map_degree = None            # default: isoparametric
        self.quadrature_order = 8         # "Internal Quadrature" gauss order
        self.singular_quadrature_type = "Mixed"
        self.singular_quadrature_order = 5
        self.reflect_kernel = False
        self.no_slip_kernel = False
        self.epsilon = 0.0
        self.wall_spans_0 = (10.0, 0.0, 10.0)
        self.wall_position_0 = (0.0, 0.0, 0.0)
        self.grid_type = "Real"
        self.imposed_component = 1
        self.assemble_scaling = 1.0
        self.use_internal_alpha = False
        self.monolithic_bool = True
        self.solve_directly = True
        self.preconditioner_type = "Direct"
        self.bandwith_preconditioner = False
        self.bandwith = 100
        self.preconditioner_block = 0     # "BlockDirect": largest diagonal block in rows (0 = the rank's whole row block)
        self.gmres_restart = 100
        self.gmres_orthogonalization = "CGS2"   # "MGS" = deal.II's modified Gram-Schmidt verbatim
        self.solver_control = SolverControl(1000, 1e-10)
        self.force_pole = (0.0, 0.0, 0.0)   # explicit pole point ("Force Pole Point Setting"; the origin by default)
        self.force_pole_kind = None          # None: use force_pole; "Origin" | "Point" | "Baricenter" as in the reference
        self.host_prepass = False         # True: mass matrix / normals / rigid modes by the host code (bs_host_prepass)
        self.keep_VK = True
        self.fused_assembly = False    # True: never store K (bs_assemble_fused); body-only monolithic systems
        self.col_is_K = None           # per dof: True -> the unknown is a wall velocity, column -K (index-set logic 3194-3245)
        self.use_peer_exchange = True  # multi-GPU: NVLink peer stores fused into the Krylov-vector kernel (else NCCL allgather)
        self.num_rigid = 6
        self.constraints = None        # hanging nodes: {dof: [(dof, coefficient), ...]} (reference ordering; ref 2970-2995)
        self.solve_with_torque = False # "Impose a torque on the flagellum" (ref 216, 3252-3256, 3340-3352)
        self.N_flagellum_torque = None
        self.N_flagellum_torque_dual = None
        self.torque_rhs = -2.0         # the reference's hard-coded motor torque (3352)
        self.flagellum_omega = 0.0
        self._ctx = None
        self.mesh = None
        self.direct_trilinos_preconditioner = DirectPreconditioner()
        self.reassemble_preconditoner = False
        self.shape_velocities = None
        self._init_frontend()   # frame loop / parameter file state (frontend.py)

    # ---- geometry ---------------------------------------------------------------------------------------
    def set_mesh(self, mesh: QuadMesh, map_mesh: QuadMesh = None):
        """read_domain + reinit + euler vector in one step: the unknown-space mesh and (optionally) a separate
        mapping-space mesh (MappingFEField(map_dh, euler_vec), bem_stokes.cc:1851)."""
        self.mesh = mesh
        self.map_mesh = mesh if map_mesh is None else map_mesh
        self.fe_degree = mesh.degree
        self.map_degree = self.map_mesh.degree
        return self

    def reinit(self):
        if self._ctx is not None:
            check(lib.bs_destroy(self._ctx))
        ctx = _lib.ctx_p()
        check(lib.bs_create(C.byref(ctx), self.device, self.fe_degree, self.map_degree))
        self._ctx = ctx
        self._block_prec_ready = False
        if self.stream is not None:
            check(lib.bs_set_stream(ctx, C.c_void_p(self.stream)))
        if self.n_mpi_processes > 1:
            check(lib.bs_set_partition(ctx, self.this_mpi_process, self.n_mpi_processes, None, self.mesh.n_nodes))
            if self.comm is None:
                raise ValueError("nranks > 1 needs a communicator (bemstokes_b200.comm.TorchComm)")
            self.comm.attach(ctx)
            if self.use_peer_exchange:
                self._enable_peer_exchange(ctx)
        mm = self.map_mesh
        euler = np.ascontiguousarray(mm.nodes.T.reshape(-1))  # component-major euler_vec
        check(lib.bs_set_geometry(ctx, mm.n_nodes, _dp(euler), self.mesh.n_cells, _ip(mm.conn), self.mesh.n_nodes,
                                  _ip(self.mesh.conn), None))
        check(lib.bs_set_quadrature(ctx, self.quadrature_order, None, None))
        check(lib.bs_set_singular_quadrature(ctx, _SING[self.singular_quadrature_type], self.singular_quadrature_order))
        self._set_kernel()
        self.N = self.mesh.n_nodes
        self.n_dofs = 3 * self.N
        self.V_matrix = DeviceMatrix(self, _lib.MAT_V)
        self.K_matrix = DeviceMatrix(self, _lib.MAT_K)
        self.monolithic_system_matrix = DeviceMatrix(self, _lib.MAT_A)
        if self.shape_velocities is None or len(self.shape_velocities) != self.n_dofs:
            self.shape_velocities = np.zeros(self.n_dofs)
        return self

    def update_geometry(self, mesh=None, map_mesh=None):
        """New coordinates on the same connectivity (the reference's per-frame compute_euler_vector): keeps the
        context, ordering, cell blocks and matrix storage."""
        if mesh is not None:
            self.mesh = mesh
            self.map_mesh = mesh if map_mesh is None else map_mesh
        mm = self.map_mesh
        euler = np.ascontiguousarray(mm.nodes.T.reshape(-1))
        check(lib.bs_set_geometry(self._ctx, mm.n_nodes, _dp(euler), self.mesh.n_cells, _ip(mm.conn), self.mesh.n_nodes,
                                  _ip(self.mesh.conn), None))
        return self

    def _push_constraints_and_torque(self):
        ctx = self._ctx
        if self.constraints:
            dofs = sorted(self.constraints)
            ptr = np.zeros(len(dofs) + 1, dtype=np.int32)
            cols, coefs = [], []
            for k, d in enumerate(dofs):
                for col, coef in self.constraints[d]:
                    cols.append(col)
                    coefs.append(coef)
                ptr[k + 1] = len(cols)
            dof = np.ascontiguousarray(dofs, dtype=np.int32)
            cols = np.ascontiguousarray(cols if cols else [0], dtype=np.int32)
            coefs = np.ascontiguousarray(coefs if coefs else [0.0], dtype=np.float64)
            check(lib.bs_set_constraints(ctx, len(dofs), _ip(dof), _ip(ptr), _ip(cols), _dp(coefs)))
        else:
            check(lib.bs_set_constraints(ctx, 0, None, None, None, None))
        if self.solve_with_torque:
            nt = np.ascontiguousarray(self.N_flagellum_torque, dtype=np.float64)
            ntd = np.ascontiguousarray(self.N_flagellum_torque_dual, dtype=np.float64)
            check(lib.bs_set_torque_mode(ctx, _dp(nt), _dp(ntd), float(self.torque_rhs)))
        else:
            check(lib.bs_set_torque_mode(ctx, None, None, 0.0))

    def _set_kernel(self):
        # ref: kernel_wall_orientation = last axis with wall_spans[0][axis]==0 (bem_stokes.cc:2861-2866)
        o = 1
        for i in range(3):
            if self.wall_spans_0[i] == 0:
                o = i
        self.kernel_wall_orientation = o
        ktype = _lib.KERNEL_FREE
        if self.reflect_kernel:
            ktype = _lib.KERNEL_FREE_SURFACE
        elif self.no_slip_kernel:
            ktype = _lib.KERNEL_NO_SLIP
        wp = np.asarray(self.wall_position_0, dtype=np.float64)
        check(lib.bs_set_kernel(self._ctx, ktype, self.epsilon, o, _dp(wp)))

    def _enable_peer_exchange(self, ctx):
        """Exchange CUDA-IPC handles of every rank's replicated Krylov buffer / arrival flags (host plumbing over
        torch.distributed) and hand them to the library: from then on the solve makes no allgather call."""
        import torch.distributed as dist
        buf = (C.c_ubyte * 128)()
        check(lib.bs_exchange_export(ctx, 3 * self.mesh.n_nodes + 8, buf))
        allh = [None] * self.n_mpi_processes
        dist.all_gather_object(allh, bytes(buf), group=getattr(self.comm, "group", None))
        cat = b"".join(allh)
        arr = (C.c_ubyte * len(cat)).from_buffer_copy(cat)
        check(lib.bs_exchange_import(ctx, self.n_mpi_processes, arr))

    def _allsum(self, arr):
        """Sum host arrays over ranks (replicated host vectors out of rank-owned slices); no-op on one rank."""
        if self.n_mpi_processes > 1:
            import torch
            import torch.distributed as dist
            dev = self.comm.device if self.comm is not None and self.comm.device is not None else "cpu"
            t = torch.from_numpy(arr).to(dev)
            dist.all_reduce(t, group=getattr(self.comm, "group", None))
            arr[...] = t.cpu().numpy()
        return arr

    def owned_nodes(self):
        n = C.c_int()
        out = np.zeros(self.N, dtype=np.int32)
        check(lib.bs_get_owned_nodes(self._ctx, C.byref(n), _ip(out)))
        return out[:n.value].copy()

    def _owned_mask(self, nr):
        own = np.zeros(self.n_dofs + nr, dtype=bool)
        nodes = self.owned_nodes()
        for c in range(3):
            own[nodes + c * self.N] = True
        if self.this_mpi_process == self.n_mpi_processes - 1:
            own[self.n_dofs:] = True
        return own

    # ---- pre-pass (host) --------------------------------------------------------------------------------
    def compute_center_of_mass_and_rigid_modes(self, frame=0):
        """ref: bem_stokes.cc:2440-2788 (+ compute_normal_vector 3922-4011): on the device (bs_prepass); the host
        restatement (bs_host_prepass) only when `host_prepass` is set."""
        if self.host_prepass:
            if self.force_pole_kind == "Baricenter":
                raise NotImplementedError("the Baricenter pole needs the device pre-pass")
            self.point_force_pole = np.asarray(self.force_pole, dtype=float)
            self._pre = Prepass(self.map_mesh.nodes, self.map_mesh.conn.astype(np.int64), self.map_degree, self.N,
                                self.mesh.conn.astype(np.int64), self.fe_degree, self.quadrature_order, self.force_pole)
        else:
            kind = self.force_pole_kind or "Point"
            self._pre = DevicePrepass(self._ctx, self.N, self.force_pole, kind)
            self.center_of_mass_body = self._pre.center_of_mass_body
            self.point_force_pole = self._pre.point_force_pole   # ref: 2546-2552
        self.N_rigid = self._pre.N_rigid
        self.N_rigid_dual = self._pre.N_rigid_dual
        self.support_points = self._pre.support_points
        self.surface = self._pre.area
        return self

    def compute_normal_vector(self):
        p = self._pre
        self.normal_vector = p.normal_vector
        self.normal_vector_pure = p.normal_vector_pure
        self.M_normal_vector_pure = p.M_normal_vector_pure
        self.l2normGamma_pure = p.l2normGamma_pure
        return self

    # ---- assembly -----------------------------------------------------------------------------------------
    def assemble_stokes_system(self, correction_on_V=True):
        """ref: BEMProblem::assemble_stokes_system (bem_stokes.cc:2840-3435)."""
        ctx = self._ctx
        self._set_kernel()
        self._push_constraints_and_torque()
        nh = np.ascontiguousarray(self.normal_vector_pure)
        mn = np.ascontiguousarray(self.M_normal_vector_pure)
        if self.fused_assembly:
            Nr0 = np.ascontiguousarray(self.N_rigid[:self.num_rigid])
            sv0 = np.ascontiguousarray(self.shape_velocities, dtype=np.float64)
            fl = None if self.col_is_K is None else np.ascontiguousarray(self.col_is_K, dtype=np.uint8)
            check(lib.bs_set_column_flags(ctx, fl.ctypes.data_as(_lib.c_ubyte_p) if fl is not None else None))
            check(lib.bs_assemble_fused(ctx, self.num_rigid, _dp(Nr0), _dp(nh), _dp(mn), self.l2normGamma_pure,
                                        _dp(sv0) if self.grid_type == "Real" else None))
        else:
            check(lib.bs_assemble_VK(ctx))
        self.V_x_normals_body = np.zeros(self.n_dofs)
        if correction_on_V:
            check(lib.bs_correct_V(ctx, _dp(nh), _dp(mn), self.l2normGamma_pure, _dp(self.V_x_normals_body)))
        else:
            self.V_matrix.vmult(self.V_x_normals_body, nh)
        check(lib.bs_correct_K(ctx, 1 if self.use_internal_alpha else 0))
        if self.monolithic_bool:
            nr = self.num_rigid
            nx = nr + (1 if self.solve_with_torque else 0)   # + the flagellum's angular velocity
            self.monolithic_rhs = np.zeros(self.n_dofs + nx)
            Nr = np.ascontiguousarray(self.N_rigid[:nr])
            Nd = np.ascontiguousarray(self.N_rigid_dual[:nr])
            sv = np.ascontiguousarray(self.shape_velocities, dtype=np.float64)
            flags = None
            if self.col_is_K is not None:
                flags = np.ascontiguousarray(self.col_is_K, dtype=np.uint8)
            check(lib.bs_build_monolithic(ctx, flags.ctypes.data_as(_lib.c_ubyte_p) if flags is not None else None,
                                          nr, _dp(Nr), _dp(Nd), _dp(nh), _dp(mn), self.l2normGamma_pure,
                                          _GRID[self.grid_type], self.imposed_component, self.assemble_scaling, _dp(sv),
                                          1 if self.keep_VK else 0, _dp(self.monolithic_rhs)))
            if getattr(self, "monolithic_solution", None) is None or len(self.monolithic_solution) != self.n_dofs + nx:
                self.monolithic_solution = np.zeros(self.n_dofs + nx)
        return self

    def tangential_projector_body(self, input_vel, output_vel=None):
        out = np.zeros(self.n_dofs) if output_vel is None else output_vel
        src = np.ascontiguousarray(input_vel, dtype=np.float64)
        check(lib.bs_tangential_projector(self._ctx, _vp(src), _vp(out)))
        return out

    # ---- solve --------------------------------------------------------------------------------------------
    def _setup_preconditioner(self, which):
        t = self.preconditioner_type
        if t == "Jacobi":
            check(lib.bs_precond_setup(self._ctx, which, _lib.PREC_JACOBI, 0))
        elif t in ("Direct",):
            if self.direct_trilinos_preconditioner._matrix is None or self.reassemble_preconditoner:
                self.direct_trilinos_preconditioner.initialize(DeviceMatrix(self, which))
                self.reassemble_preconditoner = False
        elif t == "BlockDirect":
            # block-Jacobi form of the DirectPreconditioner for row-sharded runs: every rank factorises its own diagonal
            # block once and keeps the LU across frames, like the reference's reuse of direct_trilinos_preconditioner
            # (bem_stokes.cc:5768-5779; refactorised when a solve needed more than 100 iterations, 4336-4339)
            if not getattr(self, "_block_prec_ready", False) or self.reassemble_preconditoner:
                check(lib.bs_precond_setup(self._ctx, which, _lib.PREC_BLOCK_DIRECT, int(self.preconditioner_block)))
                self._block_prec_ready = True
                self.reassemble_preconditoner = False
        elif t in ("ILU", "AMG"):
            # on these dense matrices ILU(0) is the exact LU and ML collapses to a direct coarse solve
            kind = _lib.PREC_BAND if self.bandwith_preconditioner else _lib.PREC_DIRECT
            check(lib.bs_precond_setup(self._ctx, which, kind, int(self.bandwith)))
        elif t in ("None", None):
            check(lib.bs_precond_setup(self._ctx, which, _lib.PREC_NONE, 0))
        else:
            raise ValueError("preconditioner %r is not part of the B200 hot path (SOR/SSOR are Trilinos specific)" % t)

    def gmres(self, which, x, b):
        its, res = C.c_int(), C.c_double()
        check(lib.bs_set_gmres_orthogonalization(
            self._ctx, _lib.ORTHO_MGS if self.gmres_orthogonalization == "MGS" else _lib.ORTHO_CGS2))
        rc = lib.bs_gmres(self._ctx, which, _vp(b), _vp(x), self.solver_control.tolerance, self.solver_control.max_steps,
                          self.gmres_restart, C.byref(its), C.byref(res))
        self.solver_control._last_step, self.solver_control._last_value = its.value, res.value
        check(rc)
        return its.value

    def gmres_multi(self, which, X, B):
        """nrhs systems advanced in lockstep on the device (one multi-RHS sweep over the matrix per iteration)."""
        nrhs = B.shape[0]
        its = (C.c_int * nrhs)()
        res = (C.c_double * nrhs)()
        check(lib.bs_set_gmres_orthogonalization(
            self._ctx, _lib.ORTHO_MGS if self.gmres_orthogonalization == "MGS" else _lib.ORTHO_CGS2))
        rc = lib.bs_gmres_multi(self._ctx, which, nrhs, _vp(B), _vp(X), self.solver_control.tolerance,
                                self.solver_control.max_steps, self.gmres_restart, its, res)
        self.last_steps = [its[k] for k in range(nrhs)]
        self.last_values = [res[k] for k in range(nrhs)]
        self.solver_control._last_step, self.solver_control._last_value = max(self.last_steps), max(self.last_values)
        check(rc)
        return self.last_steps

    def resistance_matrix(self):
        """Full 6x6 rigid-body resistance matrix with the six right-hand sides solved as one batch (BASELINE config
        'prolate spheroid ... 6 batched RHS'; the reference's tests/rigidity_sphere.cc:60-86 does six sequential solves
        of the ImposedVelocity system with rhs = unit vector on rigid row r).  Column r = forces/torques for unit
        rigid velocity r."""
        assert self.grid_type == "ImposedVelocity" and self.n_mpi_processes == 1
        n, nr = self.n_dofs, self.num_rigid
        B = np.zeros((nr, n + nr))
        for r in range(nr):
            B[r, n + r] = 1.0
        X = np.zeros_like(B)
        self._setup_preconditioner(_lib.MAT_A)
        self.gmres_multi(_lib.MAT_A, X, B)
        self.batched_solutions = X
        return np.array([[X[r, :n] @ self.N_rigid_dual[i] for r in range(nr)] for i in range(nr)])

    def solve_system(self, monolithic_booly=True):
        """ref: BEMProblem::solve_system (bem_stokes.cc:4158-4508)."""
        n, nr = self.n_dofs, self.num_rigid
        if monolithic_booly:
            b = np.ascontiguousarray(self.monolithic_rhs)
            x = self.monolithic_solution
            if self.solve_directly:
                check(lib.bs_direct_solve(self._ctx, _lib.MAT_A, _vp(b), _vp(x)))
                self.solver_control._last_step = 1
            else:
                self._setup_preconditioner(_lib.MAT_A)
                if self.n_mpi_processes > 1:
                    own = self._owned_mask(len(x) - n)
                    x[~own] = 0.0
                its = self.gmres(_lib.MAT_A, x, b)
                self._allsum(x)
                if its > 100:
                    self.reassemble_preconditoner = True
            r = self.monolithic_system_matrix @ x - b
            self.final_check_0 = (float(np.abs(r).max()), float(np.linalg.norm(r)))
            # split the solution into tractions and wall velocities (the reference compares A(i,i) with -K(i,i) /
            # V(i,i), bem_stokes.cc:4351-4368; here the column flags say it directly)
            isK = np.zeros(n, dtype=bool) if self.col_is_K is None else np.asarray(self.col_is_K, dtype=bool)
            self.stokes_forces = np.where(isK, 0.0, x[:n])
            self.wall_velocities = np.where(isK, x[:n], 0.0)
            self.rigid_velocities = x[n:n + nr].copy() * self.assemble_scaling
            if self.solve_with_torque:   # ref 4398-4409: the new shape velocity is omega * N_flagellum_torque
                self.flagellum_omega = float(x[n + nr])
        else:
            self.solve_dn()
        self.rigid_total_forces = np.array([self.stokes_forces @ self.N_rigid_dual[r] for r in range(nr)])
        # velocities are solved about the force pole; the reference reports them at the origin (4479-4492)
        self.baricenter_rigid_velocities = self.rigid_velocities.copy()
        pole = np.asarray(getattr(self, "point_force_pole", self.force_pole), dtype=float)
        if nr >= 6 and np.any(pole != 0.0):
            self.rigid_velocities = self.rigid_velocities.copy()
            self.rigid_velocities[:3] += np.cross(self.baricenter_rigid_velocities[3:6], -pole)
        return self

    def dirichlet_to_neumann_operator_multi(self, input_vels):
        """DN(u_k) = P V^{-1} (P K P u_k) for up to 8 velocities in one device call (bs_dn_operator_multi): one
        multi-right-hand-side sweep over K and the V-systems advanced in lockstep (ref: bem_stokes.cc:4072-4129 called
        6 + 1 times by solve_system(false), 4163-4258)."""
        U = np.ascontiguousarray(np.atleast_2d(input_vels), dtype=np.float64)
        F = np.zeros_like(U)
        its = (C.c_int * U.shape[0])()
        if not self.solve_directly:
            self._setup_preconditioner(_lib.MAT_V)
            check(lib.bs_set_gmres_orthogonalization(
                self._ctx, _lib.ORTHO_MGS if self.gmres_orthogonalization == "MGS" else _lib.ORTHO_CGS2))
        rc = lib.bs_dn_operator_multi(self._ctx, U.shape[0], _vp(U), _vp(F), 1 if self.solve_directly else 0,
                                      self.solver_control.tolerance, self.solver_control.max_steps, self.gmres_restart, its)
        self.last_steps = [its[k] for k in range(U.shape[0])]
        self.solver_control._last_step = max(self.last_steps)
        check(rc)
        return F

    def dirichlet_to_neumann_operator(self, input_vel, output_force=None):
        """DN(u) = P V^{-1} (P K P u)   (ref: bem_stokes.cc:4072-4129)."""
        F = self.dirichlet_to_neumann_operator_multi(np.asarray(input_vel, dtype=np.float64)[None, :])[0]
        if output_force is not None:
            output_force[:] = F
            return output_force
        return F

    def solve_dn(self, batched=True):
        """solve_system(false): the rigid-body problem through the DN operator (ref: bem_stokes.cc:4163-4258).  The
        reference calls the operator once for the shape velocities and once per rigid mode; here the 1 + num_rigid
        systems are one batch (`batched=False` keeps the sequential calls, for timing)."""
        nr = self.num_rigid
        inputs = np.vstack([np.asarray(self.shape_velocities, dtype=np.float64)[None, :], self.N_rigid[:nr]])
        if batched:
            out = self.dirichlet_to_neumann_operator_multi(inputs)
        else:
            out = np.vstack([self.dirichlet_to_neumann_operator(u)[None, :] for u in inputs])
        self.stokes_forces = out[0].copy()
        self.DN_N_rigid = out[1:]
        final_rhs = -np.array([self.N_rigid_dual[i] @ self.stokes_forces for i in range(nr)])
        F = np.zeros((nr, nr))
        for i in range(nr):
            if self.grid_type == "ImposedForce":
                if i == self.imposed_component:
                    final_rhs[i] += 1.0
                F[i] = [self.N_rigid_dual[i] @ self.DN_N_rigid[j] for j in range(nr)]
            elif self.grid_type == "ImposedVelocity":
                F[i, i] = 1.0
                final_rhs[i] = 1.0 if i == self.imposed_component else 0.0
            else:
                F[i] = [self.N_rigid_dual[i] @ self.DN_N_rigid[j] for j in range(nr)]
        self.final_matrix, self.final_rhs = F, final_rhs
        # the reference hands the 6 x 6 system to GMRES with the identity preconditioner (4244); it is solved directly here
        self.rigid_velocities = np.linalg.solve(F, final_rhs)
        self.stokes_forces = self.stokes_forces + self.rigid_velocities @ self.DN_N_rigid
        self.reassemble_preconditoner = True

    # ---- field evaluation ---------------------------------------------------------------------------------
    def evaluate_stokes_bie(self, val_points, vel, forces, val_velocities=None):
        """ref: BEMProblem::evaluate_stokes_bie (bem_stokes.cc:5366-5451); output component-major (i + P*idim)."""
        pts = np.ascontiguousarray(val_points, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(3 * len(pts)) if val_velocities is None else val_velocities
        out[:] = 0.0
        self._set_kernel()
        check(lib.bs_evaluate_bie(self._ctx, len(pts), _dp(pts), _dp(np.ascontiguousarray(vel, dtype=np.float64)),
                                  _dp(np.ascontiguousarray(forces, dtype=np.float64)), _dp(out), 0))
        return out

    def evaluate_stokes_bie_on_boundary(self, val_points, vel, forces, val_velocities):
        """ref: evaluate_stokes_bie_on_boundary (bem_stokes.cc:5454-5560): accumulates into val_velocities."""
        pts = np.ascontiguousarray(val_points, dtype=np.float64).reshape(-1, 3)
        check(lib.bs_evaluate_bie(self._ctx, len(pts), _dp(pts), _dp(np.ascontiguousarray(vel, dtype=np.float64)),
                                  _dp(np.ascontiguousarray(forces, dtype=np.float64)), _dp(val_velocities), 1))
        return val_velocities

    def approximate_velocity_gradient(self, val_points, vel, forces, h):
        """ref: approximate_velocity_gradient (bem_stokes.cc:5332-5364), including its one-sided '/h' scaling:
        grad[i][j][k] = (u_j(x_i + h e_k) - u_j(x_i - h e_k)) / h.  All 6*P stencil points go in one device call."""
        pts = np.ascontiguousarray(val_points, dtype=np.float64).reshape(-1, 3)
        P = len(pts)
        sten = np.repeat(pts, 6, axis=0).reshape(P, 6, 3)
        for k in range(3):
            sten[:, 2 * k, k] += h
            sten[:, 2 * k + 1, k] -= h
        u = self.evaluate_stokes_bie(sten.reshape(-1, 3), vel, forces).reshape(3, P, 6)
        grad = np.zeros((P, 3, 3))
        for j in range(3):
            for k in range(3):
                grad[:, j, k] = (u[j, :, 2 * k] - u[j, :, 2 * k + 1]) / h
        return grad

    # ---- stats ------------------------------------------------------------------------------------------------
    def stats(self):
        s = _lib.BsStats()
        check(lib.bs_get_stats(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def reset_stats(self):
        check(lib.bs_reset_stats(self._ctx))

    def close(self):
        if self._ctx is not None:
            check(lib.bs_destroy(self._ctx))
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
