"""Device solvers behind DirectPreconditioner / SolverDirect (ref: source/direct_preconditioner.cc:10-23, call sites
source/bem_stokes.cc:4109-4111, 4264-4266, 4312) and the device-resident GMRES iteration (ref: 4116, 4332): blocked LU with
the trailing update on the FP64 tensor path, the cooperative block-triangular application, diagonal blocks and the band copy
of assemble_monolithic_preconditioner (3437-3475), all against dense NumPy algebra on the matrix read back from the device."""
import ctypes as C
import os

import numpy as np
import pytest

import bemstokes_b200 as bb
from bemstokes_b200 import _lib
from bemstokes_b200._lib import lib, check
from oracle import bem_oracle as bo
from conftest import MESHES

pytestmark = pytest.mark.gpu


def make(mesh, **kw):
    p = bb.BEMProblem()
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = 6, 8
    p.grid_type, p.imposed_component = "ImposedVelocity", 0
    p.solve_directly = False
    for k, v in kw.items():
        setattr(p, k, v)
    p.reinit()
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    p.assemble_stokes_system(True)
    return p


@pytest.fixture(scope="module")
def prob():
    p = make(bb.cubesphere(m=10))   # 602 nodes, 1 812 unknowns: 15 trsv blocks, LU panels of every kind
    A = p.monolithic_system_matrix.to_dense()
    yield p, A
    p.close()


def test_lu_tensor_path_direct_solve(prob):
    p, A = prob
    n = A.shape[0]
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, n)
    x = np.zeros(n)
    check(lib.bs_direct_solve(p._ctx, _lib.MAT_A, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p)))
    xo = np.linalg.solve(A, b)
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()
    assert np.abs(A @ x - b).max() <= 1e-11 * np.abs(b).max() * np.abs(A).sum(axis=1).max()


@pytest.mark.parametrize("block", [0, 500])
def test_block_direct_application_and_gmres(prob, block):
    """BS_PREC_BLOCK_DIRECT: the node rows (optionally cut into diagonal blocks of at most `block` rows) are solved
    exactly, the rigid unknowns pass through; the cooperative application equals the dense block solve."""
    p, A = prob
    n3, n = p.n_dofs, A.shape[0]
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_BLOCK_DIRECT, block))
    rng = np.random.default_rng(2)
    v = rng.uniform(-1, 1, n)
    y = np.zeros(n)
    check(lib.bs_precond_vmult(p._ctx, v.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
    # the blocks are contiguous in the library's own ordering: recover it from the application itself when block > 0
    if block == 0:
        yo = v.copy()
        yo[:n3] = np.linalg.solve(A[:n3, :n3], v[:n3])
        assert np.abs(y - yo).max() <= 1e-9 * np.abs(yo).max()
    else:
        # M^-1 is block diagonal: applying it and multiplying back with A's diagonal blocks is the identity; check the
        # defining property instead of the ordering: M y = v on every block  <=>  y = M^-1 v, with M built from unit vectors
        Minv = np.zeros((n, n))
        for j in range(0, n, 64):   # 64 columns per call keep the test short
            E = np.zeros((min(64, n - j), n))
            E[np.arange(E.shape[0]), j + np.arange(E.shape[0])] = 1.0
            Y = np.zeros_like(E)
            for k in range(E.shape[0]):
                check(lib.bs_precond_vmult(p._ctx, E[k].ctypes.data_as(C.c_void_p), Y[k].ctypes.data_as(C.c_void_p)))
            Minv[:, j:j + E.shape[0]] = Y.T
        M = np.linalg.inv(Minv)
        mask = np.abs(M) > 1e-9 * np.abs(M).max()      # the sparsity pattern of M: diagonal blocks
        assert np.abs(M - A)[mask].max() <= 1e-7 * np.abs(A).max()   # ... and on it M equals A
        nblocks = -(-n3 // block)
        assert mask[:n3, :n3].sum() <= 1.05 * nblocks * (n3 / nblocks) ** 2 + n3
    # preconditioned GMRES converges to the solution of A x = b; with the whole node block solved exactly it needs fewer
    # iterations than without (several diagonal blocks of this first-kind operator do not: measured 50 against 38 here)
    b = p.monolithic_rhs.copy()
    x = np.zeros(n)
    its = p.gmres(_lib.MAT_A, x, b)
    xo = np.linalg.solve(A, b)
    assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_NONE, 0))
    x0 = np.zeros(n)
    its0 = p.gmres(_lib.MAT_A, x0, b)
    if block == 0:
        assert its < its0, (its, its0)


def test_fast_application_equals_substitution(prob, monkeypatch):
    p, A = prob
    n = A.shape[0]
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_DIRECT, 0))
    v = np.random.default_rng(3).uniform(-1, 1, n)
    y1, y2 = np.zeros(n), np.zeros(n)
    check(lib.bs_precond_vmult(p._ctx, v.ctypes.data_as(C.c_void_p), y1.ctypes.data_as(C.c_void_p)))
    monkeypatch.setenv("BS_LU_SUBSTITUTION", "1")
    check(lib.bs_precond_vmult(p._ctx, v.ctypes.data_as(C.c_void_p), y2.ctypes.data_as(C.c_void_p)))
    monkeypatch.delenv("BS_LU_SUBSTITUTION")
    yo = np.linalg.solve(A, v)
    assert np.abs(y2 - yo).max() <= 1e-9 * np.abs(yo).max()
    assert np.abs(y1 - yo).max() <= 1e-9 * np.abs(yo).max()
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_NONE, 0))


def test_band_preconditioner(prob):
    """BS_PREC_BAND = assemble_monolithic_preconditioner with bandwith_preconditioner (ref: bem_stokes.cc:3437-3475):
    entries with reference column index outside [i - band, i + band) are dropped before the factorisation."""
    p, A = prob
    n = A.shape[0]
    band = 400
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    lo = np.where(i > band, i - band, 0)
    B = np.where((j >= lo) & (j < i + band), A, 0.0)
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_BAND, band))
    v = np.random.default_rng(4).uniform(-1, 1, n)
    y = np.zeros(n)
    check(lib.bs_precond_vmult(p._ctx, v.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
    yo = np.linalg.solve(B, v)
    assert np.abs(y - yo).max() <= 1e-8 * np.abs(yo).max()
    # through the reference's parameters: "ILU" with bandwith_preconditioner
    p.preconditioner_type, p.bandwith_preconditioner, p.bandwith = "ILU", True, band
    p.monolithic_solution[:] = 0
    p.solve_system(True)
    xo = np.linalg.solve(A, p.monolithic_rhs)
    assert np.abs(p.monolithic_solution - xo).max() <= 1e-8 * np.abs(xo).max()
    p.preconditioner_type, p.bandwith_preconditioner = "None", False
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_NONE, 0))


def test_band_argument_checked():
    p = make(bb.cubesphere(m=2))
    assert lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_BAND, 0) == -1   # a band needs a positive width
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_BAND, 1))          # the diagonal alone: regular
    p.close()


def test_device_gmres_restart_and_max_steps(prob):
    """Restart cycles and the max_steps exit of the device-resident iteration against the oracle's GMRES."""
    p, A = prob
    n = A.shape[0]
    b = p.monolithic_rhs.copy()
    check(lib.bs_precond_setup(p._ctx, _lib.MAT_A, _lib.PREC_NONE, 0))
    p.gmres_restart = 12   # 10 inner iterations per cycle
    x = np.zeros(n)
    its = p.gmres(_lib.MAT_A, x, b)
    xo, its_o, _, ok = bo.gmres(lambda v: A @ v, b, tol=1e-10, max_n_tmp_vectors=12)
    assert ok and abs(its - its_o) <= 2, (its, its_o)
    assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()
    p.solver_control.max_steps = 7
    x = np.zeros(n)
    with pytest.raises(_lib.BemStokesError) as e:
        p.gmres(_lib.MAT_A, x, b)
    assert e.value.code == _lib.ERR_NOT_CONVERGED and p.solver_control.last_step() == 7
    p.solver_control.max_steps, p.gmres_restart = 1000, 100
    # host-driven loop (deal.II's modified Gram-Schmidt verbatim) and the device-resident loop agree
    x1, x2 = np.zeros(n), np.zeros(n)
    p.gmres_orthogonalization = "MGS"
    i1 = p.gmres(_lib.MAT_A, x1, b)
    p.gmres_orthogonalization = "CGS2"
    i2 = p.gmres(_lib.MAT_A, x2, b)
    assert abs(i1 - i2) <= 1
    assert np.abs(x1 - x2).max() <= 1e-9 * np.abs(x2).max()


def test_identity_preconditioner_follows_the_system(prob):
    """After a solve with V (the DN route, ref: bem_stokes.cc:4072-4129) an unpreconditioned solve of the monolithic system
    must act on all 3N + 6 unknowns: the right-hand sides of the resistance problem live in the rigid rows only, so an
    identity 'preconditioner' sized for V returned x = 0 after 0 iterations.  Host-driven loop (the one callback
    communicators use) and device-resident loop."""
    p, A = prob
    n = A.shape[0]
    B = np.zeros((6, n))
    for r in range(6):
        B[r, n - 6 + r] = 1.0
    Xo = np.linalg.solve(A, B.T).T
    for ortho in ("MGS", "CGS2"):
        p.preconditioner_type = "None"
        p.gmres_orthogonalization = ortho
        p.dirichlet_to_neumann_operator(p.N_rigid[0])
        X = np.zeros_like(B)
        its = p.gmres_multi(_lib.MAT_A, X, B)
        assert min(its) > 5, its
        assert np.abs(X - Xo).max() <= 1e-7 * np.abs(Xo).max()
    p.gmres_orthogonalization = "CGS2"


@pytest.mark.parametrize("meshname,grid_type", [("sphere_half_refined_0.inp", "Real"),
                                                ("prolate_spheroid_lambda_2_ref_0.msh", "ImposedForce")])
def test_dn_route_batched(meshname, grid_type):
    """solve_system(false) with the 1 + 6 DN systems as ONE device batch (bs_dn_operator_multi; ref: bem_stokes.cc:4073-4129,
    4163-4258): final_matrix, rigid velocities and stokes_forces against the oracle's dense algebra, <= 1e-9."""
    from oracle import port
    mesh = bb.read_mesh(os.path.join(MESHES, meshname))
    p = bb.BEMProblem()
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = 8, 10
    p.grid_type, p.imposed_component = grid_type, 2
    p.monolithic_bool, p.solve_directly, p.preconditioner_type = False, False, "None"
    p.solver_control.tolerance = 1e-12
    p.reinit()
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    x = mesh.nodes
    if grid_type == "Real":   # a smooth swimming stroke
        p.shape_velocities = np.concatenate([np.sin(x[:, 0]) * x[:, 1], 0.5 * x[:, 1] * x[:, 2], 0.3 * x[:, 0] * x[:, 1] - 0.1])
    p.assemble_stokes_system(True)
    p.solve_system(False)
    its_batched = list(p.last_steps)
    geo = bo.Geometry(mesh.nodes, mesh.conn.astype(np.int64), 1)
    Vo, Ko, _ = port.assemble_VK(geo, bo.KernelSpec(), 8, "Mixed", 10)
    pre = bo.Prepass(geo, 8)
    Vc, _ = bo.correct_V(Vo, pre)
    Kc = bo.correct_K(Ko, geo.N)
    import scipy.linalg as sla
    lu = sla.lu_factor(Vc)
    DN = lambda u: pre.P(sla.lu_solve(lu, pre.P(Kc @ pre.P(u))))
    f0 = DN(p.shape_velocities)
    DNr = [DN(pre.N_rigid[r]) for r in range(6)]
    Fo = np.array([[pre.N_rigid_dual[i] @ DNr[j] for j in range(6)] for i in range(6)])
    rhs = -np.array([pre.N_rigid_dual[i] @ f0 for i in range(6)])
    if grid_type == "ImposedForce":
        rhs[2] += 1.0
    Uo = np.linalg.solve(Fo, rhs)
    fo = f0 + sum(Uo[r] * DNr[r] for r in range(6))
    assert np.abs(p.final_matrix - Fo).max() <= 1e-9 * np.abs(Fo).max()
    assert np.abs(p.rigid_velocities - Uo).max() <= 1e-9 * max(1e-30, np.abs(Uo).max()) + 1e-13
    assert np.abs(p.stokes_forces - fo).max() <= 1e-9 * np.abs(fo).max()
    # batched == the reference's sequence of single calls
    F1 = p.final_matrix.copy()
    p.solve_dn(batched=False)
    assert np.abs(p.final_matrix - F1).max() <= 1e-9 * np.abs(F1).max()
    assert max(its_batched) >= 1
    # one LU for all seven systems (solve_directly)
    p.solve_directly = True
    p.solve_system(False)
    assert np.abs(p.final_matrix - Fo).max() <= 1e-9 * np.abs(Fo).max()
    assert np.abs(p.stokes_forces - fo).max() <= 1e-9 * np.abs(fo).max()
    p.close()


def test_constrained_rows_and_torque_unknown():
    """The two boundary extras of bs_build_monolithic: hanging-node constraint rows (ref: bem_stokes.cc:2970-2995, 3024-3025,
    3078, 3156-3183) and the flagellum torque unknown of solve_with_torque (ref: 3143-3147, 3191, 3252-3256, 3340-3352), against
    the oracle's restatement: entries <= 1e-12 of the row scale, solution <= 1e-9."""
    mesh = bb.read_mesh(os.path.join(MESHES, "sphere_half_refined_0.inp"))
    N = mesh.n_nodes
    # artificial hanging nodes: node 5 is the mean of nodes 7 and 11 (all three components), node 20 follows node 3 (x only)
    cons = {5 + c * N: [(7 + c * N, 0.5), (11 + c * N, 0.5)] for c in range(3)}
    cons[20] = [(3, 1.0)]
    p = bb.BEMProblem()
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = 8, 10
    p.grid_type, p.solve_directly = "Real", True
    p.constraints = cons
    p.reinit()
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    x = mesh.nodes
    p.shape_velocities = np.concatenate([np.sin(x[:, 0]) * x[:, 1], 0.5 * x[:, 1] * x[:, 2], 0.3 * x[:, 0] * x[:, 1] - 0.1])
    # torque mode: a rotation about z restricted to the "flagellum" half z < 0, dual = mass matrix times it
    geo = bo.Geometry(mesh.nodes, mesh.conn.astype(np.int64), 1)
    pre = bo.Prepass(geo, 8)
    Nt = pre.N_rigid[5] * np.tile(x[:, 2] < 0, 3)
    M = bo.mass_matrix(geo, 8)[0]
    Ntd = np.concatenate([M @ Nt[c * N:(c + 1) * N] for c in range(3)])
    for torque in (False, True):
        p.solve_with_torque = torque
        p.N_flagellum_torque, p.N_flagellum_torque_dual = Nt, Ntd
        p.assemble_stokes_system(True)
        Vo, Ko = bo.assemble_VK(geo, bo.KernelSpec(), 8, "Mixed", 10)
        Vo, Ko = bo.apply_constraints(Vo, Ko, cons)
        Vc, _ = bo.correct_V(Vo, pre, cons)
        Kc = bo.correct_K(Ko, N, False, cons)
        Ao, bvec = bo.monolithic(Vc, Kc, pre, "Real", 1, 1.0, p.shape_velocities, None, cons,
                                 (Nt, Ntd, -2.0) if torque else None)
        scale = lambda B: np.maximum(np.abs(B).max(axis=1, keepdims=True), 1e-300)
        assert (np.abs(p.V_matrix.to_dense() - Vc) / scale(Vc)).max() < 1e-12
        assert (np.abs(p.K_matrix.to_dense() - Kc) / scale(Kc)).max() < 1e-12
        A = p.monolithic_system_matrix.to_dense()
        assert A.shape == Ao.shape == (3 * N + 6 + torque, 3 * N + 6 + torque)
        assert (np.abs(A - Ao) / scale(Ao)).max() < 1e-12
        assert np.abs(p.monolithic_rhs - bvec).max() <= 1e-12 * max(1.0, np.abs(bvec).max())
        p.monolithic_solution[:] = 0
        p.solve_system(True)
        xo = np.linalg.solve(Ao, bvec)
        assert np.abs(p.monolithic_solution - xo).max() <= 1e-9 * np.abs(xo).max()
        for ii, entries in cons.items():     # the constraint equations hold in the solution
            assert abs(p.monolithic_solution[ii] - sum(cf * p.monolithic_solution[cl] for cl, cf in entries)) < 1e-9 * np.abs(xo).max()
        if torque:
            assert abs(p.flagellum_omega - xo[-1]) <= 1e-9 * abs(xo[-1])
            assert abs(Ntd @ p.stokes_forces - (-2.0)) < 1e-8      # the imposed motor torque
    # the fused (no-K) assembly refuses both
    p.fused_assembly = True
    with pytest.raises(_lib.BemStokesError):
        p.assemble_stokes_system(True)
    p.close()
