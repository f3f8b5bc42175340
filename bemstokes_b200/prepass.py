"""Host pre-pass feeding the hot path: scalar mass matrix, L2-projected unit normals, rigid modes and their
duals (ref: compute_center_of_mass_and_rigid_modes source/bem_stokes.cc:2440-2788, compute_normal_vector
3922-4011).  O(N) sparse work that the reference also does on the host; its outputs are
inputs of the C-ABI (bs_correct_V / bs_build_monolithic)."""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import _lib
from .mesh import Q1_UNIT, Q2_UNIT


def gauss_1d(n):
    x = np.zeros(n)
    w = np.zeros(n)
    rc = _lib.lib.bs_make_gauss_1d(n, x.ctypes.data_as(_lib.c_double_p), w.ctypes.data_as(_lib.c_double_p))
    if rc != n:
        _lib.check(rc)
    return x, w


def _lag(degree, t):
    if degree == 1:
        return np.stack([1 - t, t], -1), np.stack([-np.ones_like(t), np.ones_like(t)], -1)
    return (np.stack([2 * (t - .5) * (t - 1), -4 * t * (t - 1), 2 * t * (t - .5)], -1),
            np.stack([4 * t - 3, -8 * t + 4, 4 * t - 1], -1))


def shape_table(degree, xi):
    """phi[nq,na], dphi[nq,na,2] of FE_Q(degree) in deal.II dof order."""
    unit = Q1_UNIT if degree == 1 else Q2_UNIT
    lx, dx = _lag(degree, xi[:, 0])
    ly, dy = _lag(degree, xi[:, 1])
    ix = np.rint(unit[:, 0] * degree).astype(int)
    iy = np.rint(unit[:, 1] * degree).astype(int)
    phi = lx[:, ix] * ly[:, iy]
    dphi = np.stack([dx[:, ix] * ly[:, iy], lx[:, ix] * dy[:, iy]], -1)
    return phi, dphi


class Prepass:
    def __init__(self, map_nodes, map_conn, map_degree, n_nodes, conn, degree, quad_order, pole=(0., 0., 0.),
                 body_nodes=None):
        x1, w1 = gauss_1d(quad_order)
        xi = np.array([[a, b] for b in x1 for a in x1])
        w = np.array([wa * wb for wb in w1 for wa in w1])
        pm, dpm = shape_table(map_degree, xi)
        ph, _ = shape_table(degree, xi)
        X = map_nodes[map_conn]                                   # [nc, nam, 3]
        t1 = np.einsum("qa,cad->cqd", dpm[:, :, 0], X)
        t2 = np.einsum("qa,cad->cqd", dpm[:, :, 1], X)
        nn = np.cross(t1, t2)
        J = np.linalg.norm(nn, axis=2)
        jxw = J * w[None, :]
        nrm = nn / J[:, :, None]
        self.area = float(jxw.sum())
        Mloc = np.einsum("qa,qb,cq->cab", ph, ph, jxw)
        na = conn.shape[1]
        rows = np.repeat(conn, na, axis=1).reshape(-1)
        cols = np.tile(conn, (1, na)).reshape(-1)
        M = sp.coo_matrix((Mloc.reshape(-1), (rows, cols)), shape=(n_nodes, n_nodes)).tocsc()
        self.M = M
        rhs = np.zeros((n_nodes, 3))
        np.add.at(rhs, conn.reshape(-1), np.einsum("qa,cqd,cq->cad", ph, nrm, jxw).reshape(-1, 3))
        lu = spla.splu(M)
        nt = lu.solve(rhs)
        nt /= np.linalg.norm(nt, axis=1)[:, None]
        body = np.ones(n_nodes, dtype=bool) if body_nodes is None else body_nodes
        self.normal_vector = nt.T.reshape(-1).copy()
        npure = nt * body[:, None]
        self.normal_vector_pure = npure.T.reshape(-1).copy()
        self.M_normal_vector_pure = (M @ npure).T.reshape(-1).copy()
        self.l2normGamma_pure = float(self.normal_vector_pure @ self.M_normal_vector_pure)
        # support points of the unknown space
        unit = Q1_UNIT if degree == 1 else Q2_UNIT
        pu, _ = shape_table(map_degree, unit)
        sup = np.zeros((n_nodes, 3))
        sup[conn.reshape(-1)] = np.einsum("ab,cbd->cad", pu, X).reshape(-1, 3)
        self.support_points = sup
        x = (sup - np.asarray(pole)[None, :]) * body[:, None]
        R = np.zeros((6, 3, n_nodes))
        for c in range(3):
            R[c, c] = 1.0 * body
        R[3, 1], R[3, 2] = -x[:, 2], x[:, 1]
        R[4, 0], R[4, 2] = x[:, 2], -x[:, 0]
        R[5, 0], R[5, 1] = -x[:, 1], x[:, 0]
        self.N_rigid = R.reshape(6, 3 * n_nodes)
        self.N_rigid_dual = np.stack([(M @ R[r].T).T.reshape(-1) for r in range(6)], 0)
