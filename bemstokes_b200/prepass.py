"""Pre-pass feeding the hot path: scalar mass matrix, L2-projected unit normals, rigid modes and their
duals (ref: compute_center_of_mass_and_rigid_modes source/bem_stokes.cc:2440-2788, compute_normal_vector
3922-4011).  O(N) sparse work that the reference also does on the host; its outputs are
inputs of the C-ABI (bs_correct_V / bs_build_monolithic).  Two implementations behind the C-ABI, shared by the Python
and the C++ host mirrors: `DevicePrepass` (bs_prepass, csrc/bs_prepass.cu: CUDA, from the geometry a context already
holds - what BEMProblem uses) and `Prepass` (bs_host_prepass, csrc/bs_host.cu: pure host code, no GPU needed)."""
import ctypes as C

import numpy as np

from . import _lib


def gauss_1d(n):
    x = np.zeros(n)
    w = np.zeros(n)
    rc = _lib.lib.bs_make_gauss_1d(n, x.ctypes.data_as(_lib.c_double_p), w.ctypes.data_as(_lib.c_double_p))
    if rc != n:
        _lib.check(rc)
    return x, w


class Prepass:
    def __init__(self, map_nodes, map_conn, map_degree, n_nodes, conn, degree, quad_order, pole=(0., 0., 0.)):
        dp, ip = _lib.c_double_p, _lib.c_int_p
        euler = np.ascontiguousarray(np.asarray(map_nodes, dtype=np.float64).T.reshape(-1))
        cm = np.ascontiguousarray(map_conn, dtype=np.int32)
        cs = np.ascontiguousarray(conn, dtype=np.int32)
        n3 = 3 * n_nodes
        self.normal_vector_pure = np.zeros(n3)
        self.M_normal_vector_pure = np.zeros(n3)
        self.N_rigid = np.zeros((6, n3))
        self.N_rigid_dual = np.zeros((6, n3))
        self.support_points = np.zeros((n_nodes, 3))
        l2, area = C.c_double(), C.c_double()
        pl = np.asarray(pole, dtype=np.float64)
        _lib.check(_lib.lib.bs_host_prepass(degree, map_degree, len(map_nodes), euler.ctypes.data_as(dp), len(cs),
                                            cm.ctypes.data_as(ip), n_nodes, cs.ctypes.data_as(ip), quad_order,
                                            pl.ctypes.data_as(dp), self.normal_vector_pure.ctypes.data_as(dp),
                                            self.M_normal_vector_pure.ctypes.data_as(dp), C.byref(l2),
                                            self.N_rigid.ctypes.data_as(dp), self.N_rigid_dual.ctypes.data_as(dp),
                                            C.byref(area), self.support_points.ctypes.data_as(dp)))
        self.normal_vector = self.normal_vector_pure  # body-only meshes: every node belongs to the swimmer
        self.l2normGamma_pure = l2.value
        self.area = area.value


class DevicePrepass:
    """Same attributes as `Prepass`, computed on the GPU by bs_prepass from the context's geometry and quadrature."""

    POLE_KINDS = {"Origin": 0, "Point": 1, "Baricenter": 2}

    def __init__(self, ctx, n_nodes, pole=(0., 0., 0.), pole_kind="Point"):
        dp = _lib.c_double_p
        n3 = 3 * n_nodes
        self.normal_vector_pure = np.zeros(n3)
        self.M_normal_vector_pure = np.zeros(n3)
        self.N_rigid = np.zeros((6, n3))
        self.N_rigid_dual = np.zeros((6, n3))
        self.support_points = np.zeros((n_nodes, 3))
        l2, area, its = C.c_double(), C.c_double(), C.c_int()
        pl = np.asarray(pole, dtype=np.float64)
        self.center_of_mass_body = np.zeros(3)
        self.point_force_pole = np.zeros(3)
        _lib.check(_lib.lib.bs_prepass(ctx, self.POLE_KINDS[pole_kind], pl.ctypes.data_as(dp),
                                       self.normal_vector_pure.ctypes.data_as(dp),
                                       self.M_normal_vector_pure.ctypes.data_as(dp), C.byref(l2),
                                       self.N_rigid.ctypes.data_as(dp), self.N_rigid_dual.ctypes.data_as(dp), C.byref(area),
                                       self.support_points.ctypes.data_as(dp), self.center_of_mass_body.ctypes.data_as(dp),
                                       self.point_force_pole.ctypes.data_as(dp), C.byref(its)))
        self.normal_vector = self.normal_vector_pure
        self.l2normGamma_pure = l2.value
        self.area = area.value
        self.cg_iterations = its.value
