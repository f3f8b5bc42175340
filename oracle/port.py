"""ctypes wrapper of oracle/libbem_port.so (C + OpenMP restatement of the reference loop nest).
TEST / BASELINE INFRASTRUCTURE ONLY — see the header of bem_port.c."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import bem_oracle as bo

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libbem_port.so")


def _load():
    if not os.path.exists(_SO):
        # -march=native binaries do not travel between hosts: build on first use on this machine
        subprocess.check_call(["make", "-s", "-C", HERE])
    lib = C.CDLL(_SO)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.port_assemble.restype = C.c_longlong
    lib.port_assemble.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, ip, dp, ip, C.c_int, dp, dp, dp, ip, ip, dp, dp, dp,
                                  C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, dp, dp, C.c_int]
    lib.port_gemv.argtypes = [dp, C.c_longlong, C.c_longlong, dp, dp, C.c_int]
    lib.port_gemv.restype = None
    lib.port_gmres.argtypes = [dp, C.c_longlong, dp, dp, dp, C.c_double, C.c_int, C.c_int, dp, C.c_int]
    lib.port_gmres.restype = C.c_int
    lib.port_max_threads.restype = C.c_int
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = _load()
        except OSError:
            subprocess.check_call(["make", "-s", "-B", "-C", HERE])
            _lib = _load()
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def max_threads():
    return lib().port_max_threads()


def assemble_VK(geo, kernel, quad_order=8, sing_kind="Mixed", sing_order=5, row_begin=0, row_end=None, nthreads=0):
    """Same contract as bem_oracle.assemble_VK for the contiguous row range [row_begin,row_end)."""
    N, na, nam = geo.N, geo.na, geo.map_conn.shape[1]
    row_end = N if row_end is None else row_end
    xi, w = bo.gauss2(quad_order)
    phi, _ = bo.shape(geo.degree, xi)
    pm, dpm = bo.shape(geo.map_degree, xi)
    tab = np.ascontiguousarray(np.stack([pm, dpm[:, :, 0], dpm[:, :, 1]], -1))
    s_nq, s_off, s_phi, s_tab, s_w = [], [], [], [], []
    off = 0
    for a in range(na):
        sx, sw = bo.singular_rule(sing_kind, sing_order, geo.degree, a)
        s_nq.append(len(sw))
        s_off.append(off)
        off += len(sw)
        s_phi.append(bo.shape(geo.degree, sx)[0])
        p, d = bo.shape(geo.map_degree, sx)
        s_tab.append(np.stack([p, d[:, :, 0], d[:, :, 1]], -1))
        s_w.append(sw)
    s_nq, s_off = np.array(s_nq, dtype=np.int32), np.array(s_off, dtype=np.int32)
    s_phi = np.ascontiguousarray(np.concatenate(s_phi, 0))
    s_tab = np.ascontiguousarray(np.concatenate(s_tab, 0))
    s_w = np.ascontiguousarray(np.concatenate(s_w, 0))
    nr = row_end - row_begin
    V = np.zeros((3 * nr, 3 * N))
    K = np.zeros((3 * nr, 3 * N))
    sup = np.ascontiguousarray(geo.support)
    conn = np.ascontiguousarray(geo.conn, dtype=np.int32)
    mn = np.ascontiguousarray(geo.map_nodes)
    cm = np.ascontiguousarray(geo.map_conn, dtype=np.int32)
    phi = np.ascontiguousarray(phi)
    w = np.ascontiguousarray(w)
    pairs = lib().port_assemble(N, geo.ncell, na, nam, _dp(sup), _ip(conn), _dp(mn), _ip(cm), len(w), _dp(phi), _dp(tab), _dp(w),
                                _ip(s_nq), _ip(s_off), _dp(s_phi), _dp(s_tab), _dp(s_w), kernel.kind, kernel.eps, kernel.o,
                                float(kernel.wall_position[kernel.o]), row_begin, row_end, _dp(V), _dp(K), nthreads)
    return V, K, pairs


def gemv(A, x, nthreads=0):
    y = np.zeros(A.shape[0])
    lib().port_gemv(_dp(A), A.shape[0], A.shape[1], _dp(np.ascontiguousarray(x)), _dp(y), nthreads)
    return y


def gmres(A, b, x0=None, diag_inv=None, tol=1e-10, max_steps=1000, max_n_tmp_vectors=100, nthreads=0):
    x = np.zeros(len(b)) if x0 is None else np.array(x0, dtype=float)
    res = C.c_double()
    di = _dp(np.ascontiguousarray(diag_inv)) if diag_inv is not None else None
    its = lib().port_gmres(_dp(np.ascontiguousarray(A)), len(b), _dp(np.ascontiguousarray(b)), _dp(x), di, tol, max_steps,
                           max_n_tmp_vectors, C.byref(res), nthreads)
    return x, abs(its), res.value, its > 0
