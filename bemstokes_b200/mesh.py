"""Quad surface meshes for the stand-alone host: GMSH-v1 .msh / UCD .inp readers (the two formats the
reference's read_input_mesh_file accepts, source/bem_stokes.cc:496-523) and the synthetic cube-sphere of
SURVEY §8d.  Cells are returned in deal.II local order (Q1: lexicographic v1 v2 v4 v3 of the file order)."""
import numpy as np

Q2_UNIT = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0, .5], [1, .5], [.5, 0], [.5, 1], [.5, .5]], dtype=float)
Q1_UNIT = Q2_UNIT[:4]


class QuadMesh:
    """nodes[N,3], conn[ncell, n_a] (n_a = 4 for degree 1, 9 for degree 2)."""

    def __init__(self, nodes, conn, degree):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.degree = degree

    @property
    def n_nodes(self):
        return self.nodes.shape[0]

    @property
    def n_cells(self):
        return self.conn.shape[0]


def _finish(verts, quads):
    q = np.asarray(quads, dtype=np.int32)
    q = q[:, [0, 1, 3, 2]]  # counter-clockwise file order -> lexicographic
    return QuadMesh(np.asarray(verts, dtype=float), q, 1)


def read_inp(path):
    with open(path) as f:
        rows = [l.split() for l in f if l.strip() and not l.lstrip().startswith("#")]
    nv, nc = int(rows[0][0]), int(rows[0][1])
    ids = {int(r[0]): k for k, r in enumerate(rows[1:1 + nv])}
    verts = [[float(x) for x in r[1:4]] for r in rows[1:1 + nv]]
    quads = [[ids[int(v)] for v in r[3:7]] for r in rows[1 + nv:1 + nv + nc] if r[2] == "quad"]
    return _finish(verts, quads)


def read_msh(path):
    with open(path) as f:
        rows = [l.split() for l in f if l.strip()]
    i = next(k for k, r in enumerate(rows) if r[0] == "$NOD")
    nv = int(rows[i + 1][0])
    ids = {int(r[0]): k for k, r in enumerate(rows[i + 2:i + 2 + nv])}
    verts = [[float(x) for x in r[1:4]] for r in rows[i + 2:i + 2 + nv]]
    i = next(k for k, r in enumerate(rows) if r[0] == "$ELM")
    ne = int(rows[i + 1][0])
    quads = [[ids[int(v)] for v in r[5:9]] for r in rows[i + 2:i + 2 + ne] if int(r[1]) == 3]
    return _finish(verts, quads)


def read_mesh(path):
    return read_inp(path) if str(path).endswith(".inp") else read_msh(path)


def cubesphere(r=None, degree=1, scale=(1.0, 1.0, 1.0), m=None):
    """6 * m^2 quads (m = 2^r unless given); every node projected radially to the unit sphere, then scaled per
    axis (prolate: x*2)."""
    m = 2 ** r if m is None else int(m)
    sub = m * degree
    L = sub + 1
    unit = Q1_UNIT if degree == 1 else Q2_UNIT
    offs = np.rint(unit * degree).astype(np.int64)  # lattice offsets of the local nodes
    jj, ii = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    keys, cells = [], []
    for ax in range(3):
        for sgn in (-1, 1):
            u, v = (ax + 1) % 3, (ax + 2) % 3
            if sgn < 0:
                u, v = v, u
            P = np.zeros((ii.size, len(unit), 3), dtype=np.int64)
            P[:, :, ax] = sgn * sub
            P[:, :, u] = -sub + 2 * (ii[:, None] * degree + offs[None, :, 0])
            P[:, :, v] = -sub + 2 * (jj[:, None] * degree + offs[None, :, 1])
            cells.append(P)
    P = np.concatenate(cells, 0)  # [ncell, na, 3] integer lattice coordinates in [-sub, sub]
    flat = P.reshape(-1, 3) + sub
    code = (flat[:, 0] * (2 * sub + 1) + flat[:, 1]) * (2 * sub + 1) + flat[:, 2]
    uniq, first, inv = np.unique(code, return_index=True, return_inverse=True)
    # number nodes by first appearance to keep the natural cell-walk locality
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    conn = rank[inv].reshape(P.shape[0], P.shape[1])
    pts = (flat[first[order]] - sub).astype(float)
    pts /= np.linalg.norm(pts, axis=1)[:, None]
    pts *= np.asarray(scale, dtype=float)[None, :]
    del L
    return QuadMesh(pts, conn, degree)


def to_q2(mesh, project_radius=None):
    """Isoparametric Q2 nodes from a Q1 mesh: edge mid-points and cell centres (optionally on a sphere)."""
    assert mesh.degree == 1
    v, q = mesh.nodes, mesh.conn.astype(np.int64)
    nv = len(v)
    e_pairs = [(0, 2), (1, 3), (0, 1), (2, 3)]  # local Q2 dofs 4..7
    a = np.stack([q[:, i] for i, _ in e_pairs], 1).reshape(-1)
    b = np.stack([q[:, j] for _, j in e_pairs], 1).reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    code = lo * nv + hi
    uniq, first, inv = np.unique(code, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    eid = rank[inv].reshape(-1, 4)
    emid = 0.5 * (v[lo[first[order]]] + v[hi[first[order]]])
    ctr = 0.25 * (v[q[:, 0]] + v[q[:, 1]] + v[q[:, 2]] + v[q[:, 3]])
    ne = len(emid)
    nodes = np.concatenate([v, emid, ctr], 0)
    if project_radius is not None:
        nodes[nv:] *= (project_radius / np.linalg.norm(nodes[nv:], axis=1))[:, None]
    conn = np.concatenate([q, nv + eid, (nv + ne + np.arange(len(q)))[:, None]], 1)
    return QuadMesh(nodes, conn, 2)
