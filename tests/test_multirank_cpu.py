"""world_size-2 gloo tests of the N>1 plumbing (no GPU): the torch.distributed callbacks handed to bs_set_comm
(uneven allgatherv, allreduce) and the row-sharded GMRES data flow they serve — one allgather of the Krylov
vector per matvec, one allreduce per Gram-Schmidt pass — checked against the serial oracle."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bemstokes_b200.comm import TorchComm, partition_ranges
        from oracle import bem_oracle as bo
        comm = TorchComm(device="cpu")
        # ---- callbacks exactly as the library calls them: raw pointers, uneven counts --------------------
        N, nr = 11, 6
        part = partition_ranges(N, world)
        counts = [3 * (part[r + 1] - part[r]) + (nr if r == world - 1 else 0) for r in range(world)]
        displs = [3 * part[r] for r in range(world)]
        full_ref = np.arange(3 * N + nr, dtype=np.float64) * 1.5 + 1
        send = np.ascontiguousarray(full_ref[displs[rank]:displs[rank] + counts[rank]])
        recv = np.zeros(3 * N + nr)
        ci = (C.c_int * world)(*counts)
        di = (C.c_int * world)(*displs)
        rc = comm.allgatherv_cb(None, send.ctypes.data, counts[rank], recv.ctypes.data, ci, di, None)
        assert rc == 0 and np.array_equal(recv, full_ref)
        buf = np.full(5, float(rank + 1))
        assert comm.allreduce_cb(None, buf.ctypes.data, 5, None) == 0
        assert np.array_equal(buf, np.full(5, sum(range(1, world + 1)), dtype=float))
        # ---- row-sharded GMRES (CGS2) through the same callbacks == serial oracle GMRES ------------------
        rng = np.random.default_rng(3)
        n = 3 * N + nr
        A = rng.uniform(-1, 1, (n, n)) + 6 * np.eye(n)
        b = rng.uniform(-1, 1, n)
        lo, hi = displs[rank], displs[rank] + counts[rank]
        A_loc, b_loc = A[lo:hi].copy(), b[lo:hi].copy()

        def gather(y_loc):
            full = np.zeros(n)
            y = np.ascontiguousarray(y_loc)
            assert comm.allgatherv_cb(None, y.ctypes.data, len(y), full.ctypes.data, ci, di, None) == 0
            return full

        def allsum(v):
            v = np.ascontiguousarray(np.atleast_1d(v).astype(float))
            assert comm.allreduce_cb(None, v.ctypes.data, len(v), None) == 0
            return v

        x_loc = np.zeros(hi - lo)
        r_loc = b_loc - A_loc @ gather(x_loc)
        rho = float(np.sqrt(allsum(r_loc @ r_loc)[0]))
        V = [r_loc / rho]
        gamma, H, ci_, si_ = [rho], [], [], []
        its = 0
        for inner in range(60):
            its += 1
            w = A_loc @ gather(V[inner])
            B = np.stack(V, 0)
            h1 = allsum(B @ w)
            w = w - B.T @ h1
            h2 = allsum(B @ w)
            w = w - B.T @ h2
            s = float(np.sqrt(allsum(w @ w)[0]))
            h = list(h1 + h2) + [s]
            V.append(w / s)
            for i in range(inner):
                t = h[i]
                h[i] = ci_[i] * t + si_[i] * h[i + 1]
                h[i + 1] = -si_[i] * t + ci_[i] * h[i + 1]
            r = np.hypot(h[inner], h[inner + 1])
            ci_.append(h[inner] / r)
            si_.append(h[inner + 1] / r)
            h[inner] = r
            gamma.append(-si_[inner] * gamma[inner])
            gamma[inner] = ci_[inner] * gamma[inner]
            H.append(h[:inner + 1])
            if abs(gamma[inner + 1]) <= 1e-10:
                break
        k = len(H)
        Hm = np.zeros((k, k))
        for j, col in enumerate(H):
            Hm[:j + 1, j] = col
        y = np.linalg.solve(np.triu(Hm), np.array(gamma[:k]))
        x_loc = x_loc + np.stack(V[:k], 1) @ y
        x = gather(x_loc)
        xs, its_s, _, ok = bo.gmres(lambda v: A @ v, b, tol=1e-10)
        assert ok and its == its_s, (its, its_s)
        assert np.abs(x - xs).max() <= 1e-10 * np.abs(xs).max()
        assert np.abs(A @ x - b).max() < 1e-9
        assert comm.n_allgather >= its and comm.n_allreduce >= 3 * its
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_callbacks_and_sharded_gmres():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    for p in procs:
        if p.is_alive():
            p.terminate()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(out) == [1] * world


def test_partition_ranges():
    from bemstokes_b200.comm import partition_ranges
    for N in (7, 98, 24578):
        for P in (1, 2, 3, 8):
            r = partition_ranges(N, P)
            assert r[0] == 0 and r[-1] == N and all(r[i] <= r[i + 1] for i in range(P))
            sizes = [r[i + 1] - r[i] for i in range(P)]
            assert max(sizes) - min(sizes) <= 1
