"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares, the
host helpers (rules, meshes, pre-pass) agree with the oracle, and compute entry points fail LOUDLY without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bemstokes_b200 as bb
from bemstokes_b200 import _lib
from bemstokes_b200.prepass import Prepass
from oracle import bem_oracle as bo
from conftest import MESHES, ROOT


def header_symbols():
    with open(os.path.join(ROOT, "include", "bemstokes_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bs_[a-z0-9_]+)\s*\(", src)) - {"bs_allgatherv_fn", "bs_allreduce_sum_fn"})


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(_lib.lib, s), "symbol %s declared in include/bemstokes_b200.h is not exported" % s
        assert s in _lib.SIGNATURES, "symbol %s has no ctypes signature" % s
    assert _lib.lib.bs_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = _lib.ctx_p()
    rc = _lib.lib.bs_create(C.byref(ctx), 0, 1, 1)
    assert rc == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.lib.bs_last_error()
    with pytest.raises(bb.BemStokesError):
        bb.StokesKernel().value_tens(np.array([1.0, 0.0, 0.0]))


def test_mesh_and_prepass_match_oracle():
    path = os.path.join(MESHES, "sphere_half_refined_0.inp")
    m = bb.read_mesh(path)
    v, q = bo.read_inp(path)
    assert np.array_equal(m.conn, q) and np.allclose(m.nodes, v)
    p = Prepass(m.nodes, m.conn, 1, m.n_nodes, m.conn, 1, 8)
    po = bo.Prepass(bo.Geometry(v, q, 1), 8)
    assert abs(p.area - po.area) < 1e-12
    assert np.abs(p.normal_vector_pure - po.nhat).max() < 1e-13
    assert np.abs(p.M_normal_vector_pure - po.Mnhat).max() < 1e-14
    assert np.abs(p.N_rigid_dual - po.N_rigid_dual).max() < 1e-14 and np.abs(p.N_rigid - po.N_rigid).max() < 1e-15
    assert abs(p.l2normGamma_pure - po.l2) < 1e-13
    q2 = bb.cubesphere(2, 2)
    p2 = Prepass(q2.nodes, q2.conn, 2, q2.n_nodes, q2.conn, 2, 8)
    po2 = bo.Prepass(bo.Geometry(q2.nodes, q2.conn.astype(np.int64), 2), 8)
    assert np.abs(p2.normal_vector_pure - po2.nhat).max() < 1e-13 and np.abs(p2.M_normal_vector_pure - po2.Mnhat).max() < 1e-14
    msh = bb.read_mesh(os.path.join(MESHES, "sphere_mesh_3d_0.msh"))
    v2, q2 = bo.read_msh(os.path.join(MESHES, "sphere_mesh_3d_0.msh"))
    assert msh.n_nodes == 386 and msh.n_cells == 384 and np.array_equal(msh.conn, q2)


def test_cubesphere_and_q2():
    for deg in (1, 2):
        cs = bb.cubesphere(2, deg)
        n0, c0 = bo.cubesphere(2, deg)
        assert cs.n_cells == 96 and cs.n_nodes == len(n0) == (98 if deg == 1 else 386)
        assert np.allclose(np.linalg.norm(cs.nodes, axis=1), 1.0)
        # same point set, outward orientation
        geo = bo.Geometry(cs.nodes, cs.conn.astype(np.int64), deg)
        y, n, j = bo.fe_cell(geo.map_nodes[geo.map_conn[0]], deg, np.array([[.5, .5]]), np.array([1.0]))
        assert (y[0] @ n[0]) > 0
    m = bb.read_mesh(os.path.join(MESHES, "sphere_coarse_0.inp"))
    q2 = bb.to_q2(m, 1.0)
    assert q2.n_nodes == 26 and q2.conn.shape == (6, 9)


@pytest.mark.parametrize("kind,name", [(0, "Mixed"), (1, "Duffy"), (2, "Telles")])
def test_singular_rules_match_oracle(kind, name):
    for deg in (1, 2):
        for a in range(4 if deg == 1 else 9):
            for order in (4, 7, 10):
                n = _lib.lib.bs_make_singular_rule(kind, order, deg, a, 0, None, None)
                xi, w = np.zeros((n, 2)), np.zeros(n)
                _lib.lib.bs_make_singular_rule(kind, order, deg, a, n, xi.ctypes.data_as(_lib.c_double_p),
                                               w.ctypes.data_as(_lib.c_double_p))
                X, W = bo.singular_rule(name, order, deg, a)
                assert n == len(W)
                assert np.abs(xi - X).max() < 1e-14 and np.abs(w - W).max() < 1e-15


def test_gauss_rule():
    for n in (1, 2, 8, 15, 20):
        x, w = np.zeros(n), np.zeros(n)
        assert _lib.lib.bs_make_gauss_1d(n, x.ctypes.data_as(_lib.c_double_p), w.ctypes.data_as(_lib.c_double_p)) == n
        xo, wo = bo.gauss1(n)
        assert np.abs(x - xo).max() < 1e-15 and np.abs(w - wo).max() < 1e-15
