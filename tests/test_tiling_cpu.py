"""Host logic of the regular assembly pass (K1): the cell blocks that tile the (cell, node, q) loop nest of
BEMProblem::assemble_stokes_system (ref: source/bem_stokes.cc:2871-2998), their colours and - in the cell-split mode - the
pairs of cells the two thread sets of a CTA integrate concurrently.  The invariants the CUDA kernel relies on are checked
here on the CPU through the GPU-free entry point bs_host_cell_blocks (same code path as bs_set_geometry)."""
import ctypes as C
import os

import numpy as np
import pytest

import bemstokes_b200 as bb
from bemstokes_b200._lib import lib, check
from conftest import MESHES


def tiling(mesh, kernel_type, n1d):
    N, nc = mesh.n_nodes, mesh.n_cells
    nodes = np.ascontiguousarray(mesh.nodes, dtype=np.float64)
    conn = np.ascontiguousarray(mesh.conn, dtype=np.int32)
    sizes = np.zeros(6, dtype=np.int32)
    cell_ptr = np.zeros(nc + 1, dtype=np.int32)
    cells = np.full(2 * nc, -7, dtype=np.int32)
    sync = np.zeros(nc, dtype=np.uint32)
    bnodes = np.full(32 * nc, -7, dtype=np.int32)
    first = np.zeros(32 * nc, dtype=np.uint8)
    cstart = np.zeros(65, dtype=np.int32)
    pos = np.zeros(N, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    check(lib.bs_host_cell_blocks(N, nodes.ctypes.data_as(C.POINTER(C.c_double)), nc, ip(conn), kernel_type, n1d, ip(sizes),
                                  ip(cell_ptr), ip(cells), sync.ctypes.data_as(C.POINTER(C.c_uint)), ip(bnodes),
                                  first.ctypes.data_as(C.POINTER(C.c_ubyte)), ip(cstart), ip(pos)))
    nb, tj, cs, ncol, ncells_len, unpaired = [int(v) for v in sizes]
    return dict(nb=nb, tj=tj, cs=cs, ncol=ncol, unpaired=unpaired, cell_ptr=cell_ptr[:nb + 1], cells=cells[:ncells_len],
                sync=sync[:nb], nodes=bnodes[:nb * tj].reshape(nb, tj), first=first[:nb * tj].reshape(nb, tj),
                cstart=cstart[:ncol + 1], pos=pos, cpos=pos[conn])


def check_tiling(mesh, T, expect_cs, expect_tj=None):
    nc, N = mesh.n_cells, mesh.n_nodes
    assert T["cs"] == expect_cs
    if expect_tj is not None:
        assert T["tj"] == expect_tj
    assert sorted(T["pos"].tolist()) == list(range(N))                      # the node order is a permutation
    real = T["cells"][T["cells"] >= 0]
    assert sorted(real.tolist()) == list(range(nc))                          # every cell in exactly one block
    assert T["cell_ptr"][0] == 0 and T["cell_ptr"][-1] == len(T["cells"])
    assert T["cstart"][0] == 0 and T["cstart"][-1] == T["nb"] and np.all(np.diff(T["cstart"]) > 0)
    colour = np.repeat(np.arange(T["ncol"]), np.diff(T["cstart"]))
    lowest = np.full(N, 10 ** 9)
    touching = [[] for _ in range(N)]
    unpaired = 0
    for b in range(T["nb"]):
        cl = T["cells"][T["cell_ptr"][b]:T["cell_ptr"][b + 1]]
        nodes = T["nodes"][b]
        used = nodes[nodes >= 0]
        assert len(used) <= T["tj"] and np.all(np.diff(used) > 0)           # ascending positions, within the tile
        want = np.unique(T["cpos"][cl[cl >= 0]])
        assert np.array_equal(used, want)                                    # exactly the nodes of the block's cells
        for p in used:
            touching[p].append(b)
            lowest[p] = min(lowest[p], colour[b])
        if T["cs"] == 2:
            assert len(cl) % 2 == 0 and len(cl) <= 32
            steps = cl.reshape(-1, 2)
            assert np.all(steps[:, 0] >= 0)                                  # the first thread set always has a cell
            unpaired += int((steps[:, 1] < 0).sum())
            sets = [[set(T["cpos"][c].tolist()) if c >= 0 else set() for c in st] for st in steps]
            for s, (a, bb_) in enumerate(sets):
                assert not (a & bb_), "the two cells of a step share a node"
                if s > 0 and not (int(T["sync"][b]) >> s) & 1:               # the sets may be one step apart
                    pa, pb = sets[s - 1]
                    assert not (a & pb) and not (bb_ & pa), "cells of consecutive steps collide without a barrier"
            assert int(T["sync"][b]) >> len(steps) == 0 and not int(T["sync"][b]) & 1
        else:
            assert np.all(cl >= 0)
    assert unpaired == T["unpaired"]
    for p in range(N):                                                       # blocks sharing a node differ in colour
        cols = [colour[b] for b in touching[p]]
        assert len(set(cols)) == len(cols) and len(cols) >= 1
    for b in range(T["nb"]):                                                 # first touch = lowest colour at the node
        for sl, p in enumerate(T["nodes"][b]):
            if p >= 0:
                assert bool(T["first"][b, sl]) == (colour[b] == lowest[p])


@pytest.mark.parametrize("name", ["cubesphere6", "cubesphere19", "sphere_half_refined_0.inp", "sphere_mesh_3d_0.msh", "torus_0.inp",
                                  "spiral_0.msh", "sphere_0.inp"])
def test_cell_split_tiling_invariants(name):
    mesh = bb.cubesphere(m=int(name[10:])) if name.startswith("cubesphere") else bb.read_mesh(os.path.join(MESHES, name))
    check_tiling(mesh, tiling(mesh, 0, 8), 2, 15)      # free space: 2 x 4 strips
    check_tiling(mesh, tiling(mesh, 1, 8), 2, 20)      # image kernels: 3 x 4 patches
    check_tiling(mesh, tiling(mesh, 2, 8), 2, 20)
    check_tiling(mesh, tiling(mesh, 0, 6), 1)          # other rules: thread pairs, no pairing of cells


def test_cell_split_tiling_quality():
    """Structured mesh of the benchmark family: (almost) every cell has a partner, about one barrier step per strip, and the
    node columns are touched by fewer than two blocks on average."""
    mesh = bb.cubesphere(m=24)
    T = tiling(mesh, 0, 8)
    steps = len(T["cells"]) // 2
    assert steps <= 1.03 * mesh.n_cells / 2
    assert sum(bin(int(m)).count("1") for m in T["sync"]) <= 1.2 * T["nb"]
    assert (T["nodes"] >= 0).sum() / mesh.n_nodes < 2.0
    assert T["ncol"] <= 8
