"""CPU checks of the oracle restatements that have no golden in the reference: hanging-node constraint rows
(ref: source/bem_stokes.cc:2970-2995, 3024-3025, 3078, 3156-3183) and the flagellum torque unknown of solve_with_torque
(ref: 3143-3147, 3191, 3252-3256, 3340-3352).  The properties the reference relies on must hold in the restatement: the
constraint equations are rows of the system, the constrained rows carry no correction, the torque row imposes the motor
torque, and without constraints / torque the extended functions reduce to the pinned ones."""
import os

import numpy as np
import pytest

from oracle import bem_oracle as bo
from conftest import MESHES


@pytest.fixture(scope="module")
def setup():
    nodes, conn = bo.read_mesh(os.path.join(MESHES, "sphere_half_refined_0.inp"))
    geo = bo.Geometry(nodes, conn, 1)
    V, K = bo.assemble_VK(geo, bo.KernelSpec(), 6, "Mixed", 8)
    pre = bo.Prepass(geo, 6)
    return geo, V, K, pre


def test_constraint_rows_and_torque_unknown(setup):
    geo, V, K, pre = setup
    N = geo.N
    cons = {5 + c * N: [(7 + c * N, 0.5), (11 + c * N, 0.5)] for c in range(3)}
    cons[20] = [(3, 1.0)]
    Vk, Kk = bo.apply_constraints(V, K, cons)
    for ii, entries in cons.items():
        for M in (Vk, Kk):
            row = M[ii].copy()
            assert row[ii] == 1.0
            for col, coef in entries:
                assert row[col] == -coef
                row[col] = 0.0
            row[ii] = 0.0
            assert not row.any()
    free = np.array([i for i in range(3 * N) if i not in cons])
    assert np.array_equal(Vk[free], V[free]) and np.array_equal(Kk[free], K[free])
    # corrections leave the constrained rows alone and act on the others as before
    Vc, Vn = bo.correct_V(Vk, pre, cons)
    assert np.array_equal(Vc[list(cons)], Vk[list(cons)])
    assert np.abs((Vc @ pre.nhat)[free] - pre.nhat[free]).max() < 1e-12    # "post (should be one)" on the corrected rows
    Kc = bo.correct_K(Kk, N, False, cons)
    assert np.array_equal(Kc[[5, 5 + N, 5 + 2 * N]], Kk[[5, 5 + N, 5 + 2 * N]])       # node 5: x-component dof constrained -> skipped
    # node 20: only its x-component dof is constrained, but the reference tests is_constrained(i) with i the node (= that dof),
    # so the whole node is skipped; its neighbour 21 is corrected
    assert np.array_equal(Kc[[20, 20 + N, 20 + 2 * N]], Kk[[20, 20 + N, 20 + 2 * N]])
    assert not np.array_equal(Kc[21 + N], Kk[21 + N])
    # ---- monolithic system: constraint equations and the torque row
    x = geo.support
    sv = np.concatenate([np.sin(x[:, 0]) * x[:, 1], 0.5 * x[:, 1] * x[:, 2], 0.3 * x[:, 0] * x[:, 1] - 0.1])
    Nt = pre.N_rigid[5] * np.tile(x[:, 2] < 0, 3)
    M = bo.mass_matrix(geo, 6)[0]
    Ntd = np.concatenate([M @ Nt[c * N:(c + 1) * N] for c in range(3)])
    A, b = bo.monolithic(Vc, Kc, pre, "Real", 1, 1.0, sv, None, cons, (Nt, Ntd, -2.0))
    n = 3 * N
    assert A.shape == (n + 7, n + 7) and not b[:n].any() and b[n + 6] == -2.0 and not b[n:n + 6].any()
    for ii, entries in cons.items():
        assert A[ii, ii] == 1.0 and not A[ii, n:].any() and abs(A[ii].sum() - (1.0 - sum(cf for _, cf in entries))) < 1e-15
    sol = np.linalg.solve(A, b)
    for ii, entries in cons.items():
        assert abs(sol[ii] - sum(cf * sol[cl] for cl, cf in entries)) < 1e-12 * np.abs(sol).max()
    assert abs(Ntd @ sol[:n] + 2.0) < 1e-10                    # the imposed motor torque
    assert abs(sol[n + 6]) > 1e-3                               # ... drives the flagellum unknown
    # without constraints and torque the extended functions are the pinned ones
    A0, b0 = bo.monolithic(bo.correct_V(V, pre)[0], bo.correct_K(K, N), pre, "Real", 1, 1.0, sv)
    A1, b1 = bo.monolithic(bo.correct_V(V, pre, {})[0], bo.correct_K(K, N, False, {}), pre, "Real", 1, 1.0, sv, None, {}, None)
    assert np.array_equal(A0, A1) and np.array_equal(b0, b1)


def test_no_slip_coefficient_tensor_form():
    """The device's cell-split no-slip kernel (bemstokes_b200/csrc/bs_assembly.cu, integrate_no_slip) sums scalar
    coefficients times the tensors R(x)R, Q(x)Q, n(x)Q instead of evaluating the nine entries one by one; this restates that
    form in NumPy and compares it with the literal transcription of the reference (ref: source/no_slip_wall_kernel.cc:23-116,
    127-199 -> oracle G_ns / W_ns) for the three wall orientations."""
    import math
    rng = np.random.default_rng(1)
    for o in range(3):
        x = rng.normal(size=3)
        xim = x.copy()
        xim[o] = x[o] - 2 * (x[o] + 2.0)
        y, n = rng.normal(size=(64, 3)), rng.normal(size=(64, 3))
        R, Q = y - x, y - xim
        G, S = bo.G_ns(R, Q, o), (bo.W_ns(R, Q, o) * n[:, None, None, :]).sum(-1)
        ri, qi = 1 / np.sqrt((R * R).sum(-1)), 1 / np.sqrt((Q * Q).sum(-1))
        h0, Qo = 0.5 * (x[o] - xim[o]), Q[:, o]
        s = 2 * h0 * (h0 - Qo)
        e = np.ones(3)
        e[o] = -1
        PR, PQ, NQ = R[:, :, None] * R[:, None, :], Q[:, :, None] * Q[:, None, :], n[:, :, None] * Q[:, None, :]
        Rn, Qn = (R * n).sum(-1), (Q * n).sum(-1)
        G2, S2 = np.zeros_like(G), np.zeros_like(S)
        for i in range(3):
            for j in range(3):
                anti = (i == o) * Q[:, j] - (j == o) * Q[:, i]
                G2[:, i, j] = (ri ** 3 * PR[:, i, j] + (-qi ** 3 - 3 * e[i] * s * qi ** 5) * PQ[:, i, j]
                               + (i == j) * (ri - qi + e[i] * s * qi ** 3) - 2 * h0 * qi ** 3 * e[i] * anti) / (8 * math.pi)
                br = (-2 * h0 * h0 * NQ[:, i, j] + 2 * h0 * Qo * (NQ[:, i, j] - NQ[:, j, i]) - (i == j) * s * Q[:, i] ** 2 * n[:, i]
                      + (i == o) * 2 * h0 * Qn * Q[:, j])
                S2[:, i, j] = (-Rn * ri ** 5 * PR[:, i, j] + Qn * qi ** 5 * (1 + 5 * e[i] * s * qi ** 2) * PQ[:, i, j]
                               + e[i] * qi ** 5 * br) * 3 / (4 * math.pi)
        assert np.abs(G - G2).max() <= 1e-14 * np.abs(G).max()
        assert np.abs(S - S2).max() <= 1e-14 * np.abs(S).max()
