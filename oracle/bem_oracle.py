"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

A plain NumPy restatement of the BEMStokes hot path (collocation assembly of the single/double-layer
Stokes matrices V, K and the monolithic GMRES solve).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product
(``bemstokes_b200``) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this file against the reference's own golden
outputs (tests/golden/*.json, generated from /root/reference/tests/*.output by tests/golden/make_golden.py):
alpha_test 3x3 V/K sums (Q1), dof_renumbering sums (Q2), the ||V n||_inf fingerprints of the three Green
kernels, GMRES iteration counts 46 (Jacobi) / 10 (exact block), the rotation mobility 1/(8 pi) and the
1 578-line singular-quadrature error table.

The reference itself (deal.II + Trilinos + deal2lkit + MPI) cannot be built in this image, so the arithmetic
that lives in those un-vendored dependencies (QGauss, QGaussOneOverR, QTelles, QDuffy/QSplit, QIterated,
FE_Q shape functions, FEValues JxW/normals, SolverGMRES) is restated from their published algorithms.

All ``ref:`` citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------------------------------------
# Mesh I/O (ref: source/bem_stokes.cc:496-523 read_input_mesh_file; formats SURVEY Appendix C)
# --------------------------------------------------------------------------------------------------------


def _lexi(q):
    """file order v1 v2 v3 v4 (counter-clockwise) -> deal.II lexicographic (v1, v2, v4, v3)."""
    return [q[0], q[1], q[3], q[2]]


def read_inp(path):
    """UCD .inp reader: returns (verts[nv,3], quads[nc,4]) 0-based, deal.II lexicographic vertex order."""
    with open(path) as f:
        toks = f.read().split("\n")
    lines = [l for l in toks if l.strip() and not l.strip().startswith("#")]
    nv, nc = int(lines[0].split()[0]), int(lines[0].split()[1])
    ids, verts = {}, np.zeros((nv, 3))
    for k in range(nv):
        p = lines[1 + k].split()
        ids[int(p[0])] = k
        verts[k] = [float(p[1]), float(p[2]), float(p[3])]
    quads = []
    for k in range(nc):
        p = lines[1 + nv + k].split()
        if p[2] != "quad":
            continue
        quads.append(_lexi([ids[int(v)] for v in p[3:7]]))
    return verts, np.array(quads, dtype=np.int64)


def read_msh(path):
    """GMSH v1 ($NOD/$ELM) reader, quads only (element type 3)."""
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip()]
    i = lines.index("$NOD")
    nv = int(lines[i + 1])
    ids, verts = {}, np.zeros((nv, 3))
    for k in range(nv):
        p = lines[i + 2 + k].split()
        ids[int(p[0])] = k
        verts[k] = [float(p[1]), float(p[2]), float(p[3])]
    i = lines.index("$ELM")
    ne = int(lines[i + 1])
    quads = []
    for k in range(ne):
        p = lines[i + 2 + k].split()
        if int(p[1]) == 3:
            quads.append(_lexi([ids[int(v)] for v in p[5:9]]))
    return verts, np.array(quads, dtype=np.int64)


def read_mesh(path):
    return read_inp(path) if path.endswith(".inp") else read_msh(path)


# FE_Q(2) unit support points in deal.II order (ref: tests/integrate_one_over_r_Q2.output:6-14)
UNIT_SUPPORT = {
    1: np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=float),
    2: np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0, .5], [1, .5], [.5, 0], [.5, 1], [.5, .5]], dtype=float),
}


def cubesphere(r=None, degree=1, scale=(1.0, 1.0, 1.0), m=None):
    """Synthetic quad sphere: each face of [-1,1]^3 split into m x m quads (m = 2^r unless given),
    outward orientation, every Q1/Q2 node projected radially to the unit sphere (then scaled per axis)."""
    m = 2 ** r if m is None else int(m)
    sub = m * degree  # node lattice intervals per face edge
    key2id, pts = {}, []

    def nid(p):
        k = tuple(int(round(c)) for c in p)
        if k not in key2id:
            key2id[k] = len(pts)
            pts.append(k)
        return key2id[k]

    # faces: (origin, u-axis, v-axis) with u x v pointing outward
    faces = []
    for ax in range(3):
        for sgn in (-1, 1):
            u, v = (ax + 1) % 3, (ax + 2) % 3
            if sgn < 0:
                u, v = v, u
            faces.append((ax, sgn, u, v))
    conn = []
    us = UNIT_SUPPORT[degree]
    for (ax, sgn, u, v) in faces:
        for j in range(m):
            for i in range(m):
                cell = []
                for (a, b) in us:
                    p = [0, 0, 0]
                    p[ax] = sgn * sub
                    p[u] = -sub + 2 * (i * degree + int(round(a * degree)))
                    p[v] = -sub + 2 * (j * degree + int(round(b * degree)))
                    cell.append(nid(p))
                conn.append(cell)
    P = np.array(pts, dtype=float)
    P /= np.linalg.norm(P, axis=1)[:, None]
    P *= np.asarray(scale, dtype=float)[None, :]
    return P, np.array(conn, dtype=np.int64)


def q2_from_q1(verts, quads, project_radius=None):
    """Build isoparametric Q2 nodes (edge mid-points, cell centres) from a Q1 quad mesh; optional radial
    projection to a sphere of the given radius (what deal.II's SphericalManifold does up to O(1e-7))."""
    nodes = [tuple(v) for v in verts]
    edge = {}
    conn = []

    def proj(p):
        p = np.asarray(p, dtype=float)
        if project_radius is not None:
            p = p * (project_radius / np.linalg.norm(p))
        return p

    def mid(a, b):
        k = (min(a, b), max(a, b))
        if k not in edge:
            edge[k] = len(nodes)
            nodes.append(tuple(proj(0.5 * (verts[a] + verts[b]))))
        return edge[k]

    for q in quads:
        v0, v1, v2, v3 = (int(x) for x in q)
        c = len(nodes)
        nodes.append(tuple(proj(0.25 * (verts[v0] + verts[v1] + verts[v2] + verts[v3]))))
        conn.append([v0, v1, v2, v3, mid(v0, v2), mid(v1, v3), mid(v0, v1), mid(v2, v3), c])
    return np.array(nodes, dtype=float), np.array(conn, dtype=np.int64)


# --------------------------------------------------------------------------------------------------------
# FE_Q shape functions on [0,1]^2 (deal.II FE_Q<2,3>(p), tensor-product Lagrange)
# --------------------------------------------------------------------------------------------------------


def _lagrange1(degree, x):
    x = np.asarray(x, dtype=float)
    if degree == 1:
        return np.stack([1 - x, x], -1), np.stack([-np.ones_like(x), np.ones_like(x)], -1)
    # nodes 0, 1/2, 1
    l0 = 2 * (x - .5) * (x - 1)
    l1 = -4 * x * (x - 1)
    l2 = 2 * x * (x - .5)
    d0 = 4 * x - 3
    d1 = -8 * x + 4
    d2 = 4 * x - 1
    return np.stack([l0, l1, l2], -1), np.stack([d0, d1, d2], -1)


def shape(degree, xi):
    """phi[nq,n_a], dphi[nq,n_a,2] for the scalar FE_Q(degree) in deal.II dof order."""
    xi = np.atleast_2d(np.asarray(xi, dtype=float))
    lx, dx = _lagrange1(degree, xi[:, 0])
    ly, dy = _lagrange1(degree, xi[:, 1])
    us = UNIT_SUPPORT[degree]
    na = len(us)
    phi = np.zeros((len(xi), na))
    dphi = np.zeros((len(xi), na, 2))
    for a, (sx, sy) in enumerate(us):
        ix, iy = int(round(sx * degree)), int(round(sy * degree))
        phi[:, a] = lx[:, ix] * ly[:, iy]
        dphi[:, a, 0] = dx[:, ix] * ly[:, iy]
        dphi[:, a, 1] = lx[:, ix] * dy[:, iy]
    return phi, dphi


# --------------------------------------------------------------------------------------------------------
# Quadrature rules on [0,1]^2 with deal.II semantics (ref call sites bem_stokes.cc:4929-4953)
# --------------------------------------------------------------------------------------------------------


def gauss1(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return (x + 1) / 2, w / 2


def tensor2(x1, w1, x2, w2):
    """deal.II tensor product: first coordinate fastest."""
    X = np.array([[a, b] for b in x2 for a in x1], dtype=float).reshape(-1, 2)
    W = np.array([wa * wb for wb in w2 for wa in w1], dtype=float)
    return X, W


def gauss2(n):
    x, w = gauss1(n)
    return tensor2(x, w, x, w)


def qiterated2(n, k):
    """QIterated<2>(QGauss<1>(n), k): composite Gauss on k equal sub-intervals per direction."""
    x, w = gauss1(n)
    xs = np.concatenate([(x + i) / k for i in range(k)])
    ws = np.concatenate([w / k for _ in range(k)])
    return tensor2(xs, ws, xs, ws)


def lw_vertex(n, v, factor_out=True):
    """QGaussOneOverR<2>(n, vertex_index, factor_out) — Lachat-Watson rule, 2 n^2 points."""
    gp, gw = gauss2(n)
    pi4 = math.pi / 4
    p1 = np.stack([gp[:, 0], gp[:, 0] * np.tan(pi4 * gp[:, 1])], 1)
    w1 = gw * pi4 / np.cos(pi4 * gp[:, 1])
    if factor_out:
        w1 = w1 * np.linalg.norm(p1, axis=1)
    P = np.concatenate([p1, p1[:, ::-1]], 0)
    W = np.concatenate([w1, w1], 0)
    theta = {0: 0.0, 1: math.pi / 2, 2: -math.pi / 2, 3: math.pi}[v]
    if v != 0:
        c, s = math.cos(theta), math.sin(theta)
        x, y = P[:, 0] - .5, P[:, 1] - .5
        P = np.stack([c * x - s * y + .5, s * x + c * y + .5], 1)
    return P, W


def lw_point(n, s, factor_out=True):
    """QGaussOneOverR<2>(n, Point<2> singularity, factor_out): 4 boxes, degenerate ones skipped."""
    s = np.asarray(s, dtype=float)
    quads = [lw_vertex(n, 3, factor_out), lw_vertex(n, 2, factor_out), lw_vertex(n, 1, factor_out),
             lw_vertex(n, 0, factor_out)]
    origins = [np.array([0., 0.]), np.array([s[0], 0.]), np.array([0., s[1]]), s.copy()]
    unit_v = UNIT_SUPPORT[1]
    Ps, Ws = [], []
    for b in range(4):
        d = np.abs(s - unit_v[b])
        area = d[0] * d[1]
        if area > 1e-8:
            P, W = quads[b]
            Ps.append(origins[b][None, :] + P * d[None, :])
            Ws.append(W * area)
    return np.concatenate(Ps, 0), np.concatenate(Ws, 0)


def telles1(n, s):
    x, w = gauss1(n)
    keep = np.abs(x - s) > 1e-10
    x, w = x[keep], w[keep]
    eb = 2 * s - 1
    es = eb * eb - 1
    a, b = eb * es + abs(es), eb * es - abs(es)
    gb = np.cbrt(a) + np.cbrt(b) + eb
    g = 2 * x - 1
    eta = ((g - gb) ** 3 + gb * (gb * gb + 3)) / (1 + 3 * gb * gb)
    J = 3 * (g - gb) ** 2 / (1 + 3 * gb * gb)
    return (eta + 1) / 2, J * w


def telles2(n, s):
    x1, w1 = telles1(n, s[0])
    x2, w2 = telles1(n, s[1])
    return tensor2(x1, w1, x2, w2)


def duffy(n, beta=1.0):
    gp, gw = gauss2(n)
    xh, yh = gp[:, 0], gp[:, 1]
    P = np.stack([xh ** beta * (1 - yh), xh ** beta * yh], 1)
    return P, gw * beta * xh ** (2 * beta - 1)


def qsplit_duffy(n, s, beta=1.0):
    """QSplit<2>(QDuffy(n, beta), s): up to 4 triangles (s, v1, v2), degenerate ones skipped."""
    s = np.asarray(s, dtype=float)
    bp, bw = duffy(n, beta)
    uv = UNIT_SUPPORT[1]
    Ps, Ws = [], []
    for (i0, i1) in [(0, 2), (1, 3), (0, 1), (2, 3)]:
        B = np.stack([uv[i0] - s, uv[i1] - s], 1)  # columns
        J = abs(np.linalg.det(B))
        if J < 1e-12:
            continue
        Ps.append(s[None, :] + bp @ B.T)
        Ws.append(bw * J)
    return np.concatenate(Ps, 0), np.concatenate(Ws, 0)


def singular_rule(kind, order, degree, a):
    """ref: bem_stokes.cc:4912-4957 get_singular_quadrature, for scalar local index a."""
    s = UNIT_SUPPORT[degree][a]
    if kind == "Mixed":
        return qiterated2(order, degree) if degree > 1 else lw_point(order, s, True)
    if kind == "Duffy":
        return qsplit_duffy(order, s, 1.0)
    if kind == "Telles":
        return telles2(order, s)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------------------
# Surface map / FEValues
# --------------------------------------------------------------------------------------------------------


def fe_cell(X, map_degree, xi, w):
    """y[nq,3], n[nq,3], JxW[nq] of one cell with mapping nodes X[n_a_map,3]."""
    phi, dphi = shape(map_degree, xi)
    y = phi @ X
    t1 = dphi[:, :, 0] @ X
    t2 = dphi[:, :, 1] @ X
    nn = np.cross(t1, t2)
    nrm = np.linalg.norm(nn, axis=1)
    return y, nn / nrm[:, None], w * nrm


# --------------------------------------------------------------------------------------------------------
# Green kernels, literal restatements (vectorised over the leading axes of p)
# --------------------------------------------------------------------------------------------------------

FREE, FREE_SURFACE, NO_SLIP = 0, 1, 2


def _G0(p, r):
    """p_i p_j / r^3 + delta_ij / r   (ref: source/kernel.cc:61-83, before the /(8 pi))."""
    G = p[..., :, None] * p[..., None, :] / (r ** 3)[..., None, None]
    for i in range(3):
        G[..., i, i] += 1.0 / r
    return G


def G_free(p, eps=0.0):
    r = np.sqrt((p * p).sum(-1)) + eps
    return _G0(p, r) / (8 * math.pi)


def W_free(p, eps=0.0):
    """ref: source/kernel.cc:85-104  W_ijk = -3 p_i p_j p_k / r^5 / (4 pi)."""
    r = np.sqrt((p * p).sum(-1)) + eps
    return -3.0 * p[..., :, None, None] * p[..., None, :, None] * p[..., None, None, :] \
        / (r ** 5)[..., None, None, None] / (4 * math.pi)


def G_fs(p, pim, o, eps=0.0):
    """ref: source/free_surface_kernel.cc:19-72."""
    a, b = G_free(p, eps), G_free(pim, eps)
    G = a + b
    G[..., o, :] = a[..., o, :] - b[..., o, :]
    return G


def W_fs(p, pim, o, eps=0.0):
    """ref: source/free_surface_kernel.cc:135-209."""
    a, b = W_free(p, eps), W_free(pim, eps)
    W = a + b
    W[..., o, :, :] = a[..., o, :, :] - b[..., o, :, :]
    return W


def G_ns(p, pim, o, eps=0.0):
    """ref: source/no_slip_wall_kernel.cc:23-116 (3-D branch)."""
    h0 = 0.5 * (pim[..., o] - p[..., o])
    R = np.sqrt((p * p).sum(-1)) + eps
    Ri = np.sqrt((pim * pim).sum(-1)) + eps
    G = np.zeros(p.shape[:-1] + (3, 3))
    for i in range(3):
        for j in range(3):
            d = 1.0 * (i == j)
            di1 = 1.0 * (i == o)
            dj1 = 1.0 * (j == o)
            base = (p[..., i] * p[..., j] / (R * R * R) + d / R) - (pim[..., i] * pim[..., j] / (Ri * Ri * Ri) + d / Ri)
            t5 = (-3 * pim[..., i] * pim[..., j] / (Ri * Ri * Ri * Ri * Ri) + d / (Ri * Ri * Ri))
            t2 = 2. * h0 * h0 * t5
            t3 = 2. * h0 * (pim[..., o] * t5 + ((di1 * pim[..., j] - dj1 * pim[..., i]) / (Ri * Ri * Ri)))
            if i == o:
                G[..., i, j] = base - t2 + t3
            else:
                G[..., i, j] = base + t2 - t3
    return G / (8 * math.pi)


def W_ns(p, pim, o, eps=0.0):
    """ref: source/no_slip_wall_kernel.cc:127-199, including the delta_ij*p_image[k]*delta_jk*p_image[i]
    product (lines 168-169, 188-190) reproduced literally."""
    h0 = 0.5 * (pim[..., o] - p[..., o])
    R = np.sqrt((p * p).sum(-1)) + eps
    Ri = np.sqrt((pim * pim).sum(-1)) + eps
    W = np.zeros(p.shape[:-1] + (3, 3, 3))
    R5, Ri5, Ri7 = R ** 5, Ri ** 5, Ri ** 7
    for i in range(3):
        di1 = 1.0 * (i == o)
        for j in range(3):
            dij = 1.0 * (i == j)
            for k in range(3):
                djk = 1.0 * (k == j)
                dik = 1.0 * (i == k)
                w = -1. * p[..., i] * p[..., j] * p[..., k] / R5
                w = w - (-1. * pim[..., i] * pim[..., j] * pim[..., k] / Ri5)
                brk = (-(dik * pim[..., j] + dij * pim[..., k] * djk * pim[..., i]) / Ri5
                       + 5. * (pim[..., i] * pim[..., j] * pim[..., k]) / Ri7)
                t2 = 2 * h0 * h0 * brk
                t3 = (-2 * h0) * (pim[..., o] * brk + (djk * pim[..., i] * pim[..., o] - di1 * pim[..., j] * pim[..., k]) / Ri5)
                if i == o:
                    w = w - t2 - t3
                else:
                    w = w + t2 + t3
                W[..., i, j, k] = w * 3 / (4 * math.pi)
    return W


class KernelSpec:
    """Mirrors the reflect/no_slip dispatch of compute_G_kernel / compute_W_kernel (bem_stokes.cc:5027-5069)
    and the image-point construction of bem_stokes.cc:2917-2920."""

    def __init__(self, kind=FREE, eps=0.0, wall_orientation=1, wall_position=(0.0, 0.0, 0.0)):
        self.kind, self.eps, self.o = kind, eps, wall_orientation
        self.wall_position = np.asarray(wall_position, dtype=float)

    def GS(self, y, n, x):
        """y[...,3] quadrature points, n[...,3] normals, x[...,3] collocation points (broadcastable).
        Returns G[...,3,3] and S = W.n [...,3,3]."""
        R = y - x
        if self.kind == FREE:
            G, W = G_free(R, self.eps), W_free(R, self.eps)
        else:
            xim = np.array(np.broadcast_to(x, R.shape))
            xim[..., self.o] = xim[..., self.o] - 2 * (xim[..., self.o] - self.wall_position[self.o])
            Rim = y - xim
            if self.kind == FREE_SURFACE:
                G, W = G_fs(R, Rim, self.o, self.eps), W_fs(R, Rim, self.o, self.eps)
            else:
                G, W = G_ns(R, Rim, self.o, self.eps), W_ns(R, Rim, self.o, self.eps)
        S = (W * np.broadcast_to(n, R.shape)[..., None, None, :]).sum(-1)
        return G, S


# --------------------------------------------------------------------------------------------------------
# Geometry container
# --------------------------------------------------------------------------------------------------------


class Geometry:
    """nodes/conn of the unknown space FE_Q(degree)^3 and of the mapping space FE_Q(map_degree)^3."""

    def __init__(self, nodes, conn, degree=1, map_nodes=None, map_conn=None, map_degree=None):
        self.degree = degree
        self.conn = np.asarray(conn, dtype=np.int64)
        self.N = int(self.conn.max()) + 1 if nodes is None else len(nodes)
        self.map_degree = degree if map_degree is None else map_degree
        self.map_nodes = np.asarray(nodes if map_nodes is None else map_nodes, dtype=float)
        self.map_conn = self.conn if map_conn is None else np.asarray(map_conn, dtype=np.int64)
        self.ncell = len(self.conn)
        self.na = self.conn.shape[1]
        # support points = mapped unit support points (bem_stokes.cc:2855-2856)
        self.support = np.zeros((self.N, 3))
        us = UNIT_SUPPORT[degree]
        phi, _ = shape(self.map_degree, us)
        for c in range(self.ncell):
            self.support[self.conn[c]] = phi @ self.map_nodes[self.map_conn[c]]


# --------------------------------------------------------------------------------------------------------
# Assembly (ref: bem_stokes.cc:2871-2998)
# --------------------------------------------------------------------------------------------------------


def assemble_VK(geo, kernel, quad_order=8, sing_kind="Mixed", sing_order=5, rows=None, reg_rule=None):
    """Returns V, K with rows = 3*len(rows) (component-major over the row subset: r = k + a*len(rows))
    and columns component-major over all N nodes (j + b*N)."""
    N, na = geo.N, geo.na
    rows = np.arange(N) if rows is None else np.asarray(rows, dtype=np.int64)
    nr = len(rows)
    rowpos = -np.ones(N, dtype=np.int64)
    rowpos[rows] = np.arange(nr)
    V = np.zeros((3 * nr, 3 * N))
    K = np.zeros((3 * nr, 3 * N))
    xi, w = gauss2(quad_order) if reg_rule is None else reg_rule
    phi_reg, _ = shape(geo.degree, xi)
    srules = [singular_rule(sing_kind, sing_order, geo.degree, a) for a in range(na)]
    sphi = [shape(geo.degree, r[0])[0] for r in srules]
    X = geo.support[rows]  # [nr,3]
    for c in range(geo.ncell):
        Xc = geo.map_nodes[geo.map_conn[c]]
        y, n, jxw = fe_cell(Xc, geo.map_degree, xi, w)
        G, S = kernel.GS(y[None, :, :], n[None, :, :], X[:, None, :])  # [nr,nq,3,3]
        pw = phi_reg * jxw[:, None]  # [nq,na]
        lv = np.einsum("iqab,qj->iabj", G, pw)
        lk = -np.einsum("iqab,qj->iabj", S, pw)
        # singular nodes: first local index a with conn[c,a]==i (bem_stokes.cc:2885-2895)
        for a in range(na):
            i = geo.conn[c, a]
            k = rowpos[i]
            if k < 0:
                continue
            if a != int(np.nonzero(geo.conn[c] == i)[0][0]):
                continue
            sx, sw = srules[a]
            ys, ns, js = fe_cell(Xc, geo.map_degree, sx, sw)
            Gs, Ss = kernel.GS(ys, ns, geo.support[i][None, :])
            pws = sphi[a] * js[:, None]
            lv[k] = np.einsum("qab,qj->abj", Gs, pws)
            lk[k] = -np.einsum("qab,qj->abj", Ss, pws)
        cols = geo.conn[c]
        for a in range(3):
            for b in range(3):
                # np.add.at for repeated nodes inside one cell is not needed (quads have distinct nodes)
                V[a * nr:(a + 1) * nr, b * N + cols] += lv[:, a, b, :]
                K[a * nr:(a + 1) * nr, b * N + cols] += lk[:, a, b, :]
    return V, K


def alpha_sums(geo, kernel, node, quad_order=8, sing_kind="Mixed", sing_order=5):
    """test_V / test_K of tests/alpha_test.cc:104-113: sum_cells sum_q G JxW and sum S JxW at one node."""
    xi, w = gauss2(quad_order)
    tV, tK = np.zeros((3, 3)), np.zeros((3, 3))
    x = geo.support[node]
    for c in range(geo.ncell):
        Xc = geo.map_nodes[geo.map_conn[c]]
        hit = np.nonzero(geo.conn[c] == node)[0]
        if len(hit):
            sx, sw = singular_rule(sing_kind, sing_order, geo.degree, int(hit[0]))
            y, n, jxw = fe_cell(Xc, geo.map_degree, sx, sw)
        else:
            y, n, jxw = fe_cell(Xc, geo.map_degree, xi, w)
        G, S = kernel.GS(y, n, x[None, :])
        tV += np.einsum("qab,q->ab", G, jxw)
        tK += np.einsum("qab,q->ab", S, jxw)
    return tV, tK


# --------------------------------------------------------------------------------------------------------
# Pre-pass (ref: bem_stokes.cc:2440-2788, 3922-4011)
# --------------------------------------------------------------------------------------------------------


def mass_matrix(geo, quad_order=8):
    """Scalar mass matrix M_ab = sum_q phi_a phi_b JxW (block diagonal over components, 2499-2517) and
    the normal right-hand side int phi_a n (3945-3978).  Dense; fine for oracle sizes."""
    xi, w = gauss2(quad_order)
    phi, _ = shape(geo.degree, xi)
    M = np.zeros((geo.N, geo.N))
    rhs = np.zeros((geo.N, 3))
    area = 0.0
    for c in range(geo.ncell):
        y, n, jxw = fe_cell(geo.map_nodes[geo.map_conn[c]], geo.map_degree, xi, w)
        idx = geo.conn[c]
        M[np.ix_(idx, idx)] += np.einsum("qa,qb,q->ab", phi, phi, jxw)
        rhs[idx] += np.einsum("qa,qk,q->ak", phi, n, jxw)
        area += jxw.sum()
    return M, rhs, area


class Prepass:
    """normal_vector(_pure), M_normal_vector_pure, l2normGamma_pure, N_rigid, N_rigid_dual for a body-only
    problem (every node has material_id 0).  Vectors are component-major (i + c*N)."""

    def __init__(self, geo, quad_order=8, pole=(0.0, 0.0, 0.0)):
        N = geo.N
        M, rhs, area = mass_matrix(geo, quad_order)
        self.M, self.area = M, area
        nt = np.linalg.solve(M, rhs)
        nt /= np.linalg.norm(nt, axis=1)[:, None]
        self.nhat = nt.T.reshape(-1).copy()  # component-major
        self.Mnhat = (M @ nt).T.reshape(-1).copy()
        self.l2 = float(self.nhat @ self.Mnhat)
        x = geo.support - np.asarray(pole)[None, :]
        R = np.zeros((6, 3, N))
        for c in range(3):
            R[c, c] = 1.0
        R[3, 1], R[3, 2] = -x[:, 2], x[:, 1]
        R[4, 0], R[4, 2] = x[:, 2], -x[:, 0]
        R[5, 0], R[5, 1] = -x[:, 1], x[:, 0]
        self.N_rigid = R.reshape(6, 3 * N)
        self.N_rigid_dual = np.einsum("ij,rcj->rci", M, R).reshape(6, 3 * N)

    def P(self, v):
        """tangential_projector_body (bem_stokes.cc:4142-4151)."""
        return v - (self.Mnhat @ v) / self.l2 * self.nhat


# --------------------------------------------------------------------------------------------------------
# Corrections + monolithic system (ref: bem_stokes.cc:3004-3098, 3120-3357)
# --------------------------------------------------------------------------------------------------------


def apply_constraints(V, K, constraints):
    """Constrained (hanging-node) rows of V and K as the assembly loop leaves them (bem_stokes.cc:2970-2995): the row of a
    constrained dof ii is not integrated; it holds 1 on the diagonal and -coefficient at the constraining dofs.
    constraints: {dof ii (reference ordering): [(dof, coefficient), ...]}."""
    V, K = V.copy(), K.copy()
    for ii, entries in constraints.items():
        for M in (V, K):
            M[ii, :] = 0.0
            M[ii, ii] = 1.0
            for col, coef in entries:
                M[ii, col] = -coef
    return V, K


def correct_V(V, pre, constraints=None):
    """V correction on the unconstrained rows only (bem_stokes.cc:3017-3032, "We correct only if we don't have constraints")."""
    Vn = V @ pre.nhat
    u = pre.nhat - Vn
    if constraints:
        u[list(constraints.keys())] = 0.0
    return V + np.outer(u, pre.Mnhat) / pre.l2, Vn


def correct_K(K, N, use_internal_alpha=False, constraints=None):
    """K correction; nodes whose x-component dof is constrained are skipped (bem_stokes.cc:3078, is_constrained(i))."""
    K = K.copy()
    C = np.stack([K[:, k * N:(k + 1) * N].sum(1) for k in range(3)], 0)  # C[k] = K e_k
    idx = np.arange(N)
    if constraints:
        idx = np.array([i for i in range(N) if i not in constraints], dtype=np.int64)
    for j in range(3):
        for k in range(3):
            K[idx + j * N, idx + k * N] -= C[k][idx + j * N]
            if j == k and not use_internal_alpha:
                K[idx + j * N, idx + k * N] += 1.0
    return K


def monolithic(V, K, pre, grid_type="ImposedForce", imposed_component=1, scaling=1.0, shape_vel=None, col_is_K=None,
               constraints=None, torque=None):
    """Monolithic matrix / rhs.  col_is_K[j] set: column j belongs to a wall unknown whose
    velocity is solved for, A(:,j) = -K(:,j) (neumann / free-surface tangential sets, bem_stokes.cc:3194-3245);
    otherwise A(:,j) = V(:,j) (body, no-slip, dirichlet sets).
    constraints {dof: [(dof, coef)]}: the row of a constrained dof is 1 on the diagonal, -coef at the constraining dofs,
    nothing in the rigid columns, rhs 0 (bem_stokes.cc:3156-3183).
    torque = (N_torque, N_torque_dual, rhs_value): solve_with_torque (bem_stokes.cc:3191, 3252-3256, 3340-3352): one more
    unknown (the flagellum's angular velocity) with the column -scaling P K P N_torque, the row scaling N_torque_dual,
    right-hand side rhs_value (-2 in the reference), and ZERO right-hand side on every node row."""
    n = V.shape[0]
    nx = 6 + (1 if torque is not None else 0)
    A = np.zeros((n + nx, n + nx))
    b = np.zeros(n + nx)
    A[:n, :n] = V
    if col_is_K is not None:
        f = np.asarray(col_is_K, dtype=bool)
        A[:n, :n][:, f] = -K[:, f]
    for r in range(6):
        A[:n, n + r] = -scaling * pre.P(K @ pre.P(pre.N_rigid[r]))
    if grid_type == "Real" and shape_vel is not None:
        b[:n] = pre.P(K @ pre.P(shape_vel))
    if torque is not None:
        Nt, Ntd, tval = torque
        A[:n, n + 6] = -scaling * pre.P(K @ pre.P(Nt))
        A[n + 6, :n] = scaling * np.asarray(Ntd)
        b[:n] = 0.0
        b[n + 6] = tval
    if constraints:
        for ii, entries in constraints.items():
            A[ii, :] = 0.0
            A[ii, ii] = 1.0
            for col, coef in entries:
                A[ii, col] = -coef
            b[ii] = 0.0
    for r in range(6):
        if grid_type != "Real":
            b[n + r] = 1.0 if r == imposed_component else 0.0
            if grid_type == "ImposedVelocity":
                A[n + r, n + r] = scaling
            else:
                A[n + r, :n] = pre.N_rigid_dual[r]
        else:
            A[n + r, :n] = scaling * pre.N_rigid_dual[r]
    return A, b


# --------------------------------------------------------------------------------------------------------
# deal.II SolverGMRES semantics: left preconditioning, MGS + conditional re-orthogonalisation,
# Givens rotations, absolute tolerance on the preconditioned residual.
# --------------------------------------------------------------------------------------------------------


def gmres(matvec, b, x0=None, prec=None, tol=1e-10, max_steps=1000, max_n_tmp_vectors=100):
    n = len(b)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    prec = (lambda v: v) if prec is None else prec
    m = max_n_tmp_vectors - 2
    its = 0
    hist = []
    while True:
        v = prec(b - matvec(x))
        rho = float(np.linalg.norm(v))
        hist.append(rho)
        if rho <= tol or its >= max_steps:
            return x, its, hist, rho <= tol
        Vb = [v / rho]
        gamma = np.zeros(m + 1)
        gamma[0] = rho
        H = np.zeros((m + 1, m))
        ci, si = np.zeros(m), np.zeros(m)
        done, dim = False, 0
        for inner in range(m):
            its += 1
            vv = prec(matvec(Vb[inner]))
            dim = inner + 1
            h = np.zeros(dim + 1)
            norm_start = 0.0
            if its % 5 == 0:
                norm_start = float(np.linalg.norm(vv))
            for i in range(dim):
                h[i] = vv @ Vb[i]
                vv = vv - h[i] * Vb[i]
            reorth = False
            if its % 5 == 0:
                nv = float(np.linalg.norm(vv))
                if not (nv > 10. * norm_start * math.sqrt(np.finfo(float).eps)):
                    reorth = True
            if reorth:
                for i in range(dim):
                    ht = vv @ Vb[i]
                    h[i] += ht
                    vv = vv - ht * Vb[i]
            s = float(np.linalg.norm(vv))
            h[dim] = s
            Vb.append(vv / s if s != 0 else vv)
            for i in range(inner):
                t = h[i]
                h[i] = ci[i] * t + si[i] * h[i + 1]
                h[i + 1] = -si[i] * t + ci[i] * h[i + 1]
            r = math.hypot(h[inner], h[inner + 1])
            ci[inner], si[inner] = h[inner] / r, h[inner + 1] / r
            h[inner] = r
            gamma[inner + 1] = -si[inner] * gamma[inner]
            gamma[inner] = ci[inner] * gamma[inner]
            H[:dim, inner] = h[:dim]
            rho = abs(gamma[dim])
            hist.append(rho)
            if rho <= tol or its >= max_steps:
                done = True
                break
        yk = np.linalg.solve(np.triu(H[:dim, :dim]), gamma[:dim]) if dim else np.zeros(0)
        for i in range(dim):
            x = x + yk[i] * Vb[i]
        if done:
            return x, its, hist, rho <= tol


def lu_solve_factory(A):
    """Dense LU with partial pivoting (what Amesos KLU / ILU(0) on a dense pattern amount to)."""
    import scipy.linalg as sla
    lu = sla.lu_factor(A)
    return lambda v: sla.lu_solve(lu, v)


# --------------------------------------------------------------------------------------------------------
# Field evaluation (ref: bem_stokes.cc:5366-5451 evaluate_stokes_bie, 5454-5560 _on_boundary)
# --------------------------------------------------------------------------------------------------------


def evaluate_bie(geo, kernel, pts, vel, forces, quad_order=8, on_boundary=False, sing_kind="Mixed", sing_order=5, out=None):
    """u_a(x_i) = sum_cells sum_q [G_ab f_b - S_ab u_b] JxW; vel/forces component-major (i + c*N); result
    component-major over the points (i + P*a).  on_boundary: free kernel, singular rule for cells with a support
    point within 1e-3 of x_i, accumulated into `out`."""
    pts = np.asarray(pts, dtype=float).reshape(-1, 3)
    P, N = len(pts), geo.N
    res = np.zeros(3 * P) if (out is None or not on_boundary) else out
    xi, w = gauss2(quad_order)
    phi, _ = shape(geo.degree, xi)
    F = np.asarray(forces).reshape(3, N).T
    U = np.asarray(vel).reshape(3, N).T
    ker = KernelSpec(FREE, kernel.eps) if on_boundary else kernel
    srules = [singular_rule(sing_kind, sing_order, geo.degree, a) for a in range(geo.na)] if on_boundary else None
    for c in range(geo.ncell):
        Xc = geo.map_nodes[geo.map_conn[c]]
        y, n, jxw = fe_cell(Xc, geo.map_degree, xi, w)
        fq, uq = phi @ F[geo.conn[c]], phi @ U[geo.conn[c]]
        near = np.zeros(P, dtype=int) - 1
        if on_boundary:
            d = np.linalg.norm(pts[:, None, :] - geo.support[geo.conn[c]][None, :, :], axis=2)  # [P, na]
            hit = d <= 1e-3
            near = np.where(hit.any(1), hit.argmax(1), -1)
        reg = np.nonzero(near < 0)[0]
        if len(reg):
            G, S = ker.GS(y[None, :, :], n[None, :, :], pts[reg][:, None, :])
            val = np.einsum("iqab,qb,q->ia", G, fq, jxw) - np.einsum("iqab,qb,q->ia", S, uq, jxw)
            for a in range(3):
                res[reg + a * P] += val[:, a]
        for i in np.nonzero(near >= 0)[0]:
            sx, sw = srules[near[i]]
            ys, ns, js = fe_cell(Xc, geo.map_degree, sx, sw)
            ph = shape(geo.degree, sx)[0]
            G, S = ker.GS(ys, ns, pts[i][None, :])
            val = np.einsum("qab,qb,q->a", G, ph @ F[geo.conn[c]], js) - np.einsum("qab,qb,q->a", S, ph @ U[geo.conn[c]], js)
            for a in range(3):
                res[i + a * P] += val[a]
    return res
