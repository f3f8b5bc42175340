"""Launched by torchrun (one rank per GPU): row-sharded assembly + GMRES on a small sphere, checked against the
CPU oracle on every rank.  `pytest -m gpu` runs it through tests/test_gpu_multi.py when >= 2 GPUs are visible."""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    import bemstokes_b200 as bb
    from bemstokes_b200.comm import TorchComm
    from oracle import bem_oracle as bo
    comm = TorchComm(device=dev)
    mesh = bb.cubesphere(m=6)  # 218 nodes
    # no manual stream plumbing: the callbacks order torch's collectives on the context's own stream
    p = bb.BEMProblem(device=local, rank=rank, nranks=world, comm=comm)
    p.use_peer_exchange = os.environ.get("BS_PEER_EXCHANGE", "1") == "1"
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = 6, 8
    p.grid_type, p.imposed_component = "ImposedVelocity", 0
    p.solve_directly, p.preconditioner_type = False, "None"
    p.reinit()
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    own = p.owned_nodes()
    allown = [None] * world
    dist.all_gather_object(allown, own.tolist())
    assert sorted(sum(allown, [])) == list(range(mesh.n_nodes)), "row partition does not cover every node once"
    p.assemble_stokes_system(True)
    p.solve_system(True)
    # oracle
    geo = bo.Geometry(mesh.nodes, mesh.conn.astype(np.int64), 1)
    Vo, Ko = bo.assemble_VK(geo, bo.KernelSpec(), 6, "Mixed", 8)
    pre = bo.Prepass(geo, 6)
    Vc, _ = bo.correct_V(Vo, pre)
    Kc = bo.correct_K(Ko, geo.N)
    A, b = bo.monolithic(Vc, Kc, pre, "ImposedVelocity", 0)
    xg, its, _, ok = bo.gmres(lambda v: A @ v, b, tol=1e-10)
    # entries of the owned rows
    N = mesh.n_nodes
    rows = np.concatenate([own + c * N for c in range(3)]).astype(np.int32)
    cols = np.arange(3 * N, dtype=np.int32)
    rr, cc = np.meshgrid(rows, cols, indexing="ij")
    Kg = p.K_matrix.entries(rr.reshape(-1), cc.reshape(-1)).reshape(len(rows), 3 * N)
    ek = np.abs(Kg - Kc[rows]).max() / np.abs(Kc).max()
    Ag = p.monolithic_system_matrix.entries(rr.reshape(-1), cc.reshape(-1)).reshape(len(rows), 3 * N)
    ea = np.abs(Ag - A[rows][:, :3 * N]).max() / np.abs(A).max()
    es = np.abs(p.monolithic_solution - xg).max() / np.abs(xg).max()
    assert ek < 1e-12 and ea < 1e-12, (ek, ea)
    assert abs(p.solver_control.last_step() - its) <= 1, (p.solver_control.last_step(), its)
    assert es < 1e-8, es
    assert p.final_check_0[0] < 1e-9
    n_ar_solve = comm.n_allreduce
    if p.use_peer_exchange:
        # the solve itself made no collective call (Krylov slices and Gram-Schmidt sums go through peer memory): only the
        # monolithic build and the host-side checks did
        assert comm.n_allgather <= 10, comm.n_allgather
    else:
        assert comm.n_allgather >= its
    # block-Jacobi DirectPreconditioner: every rank factorises its own diagonal block (ref: direct_preconditioner.cc:10-23)
    its_plain = p.solver_control.last_step()
    p.preconditioner_type = "BlockDirect"
    ar0 = comm.n_allreduce
    p.monolithic_solution[:] = 0
    p.solve_system(True)
    esp = np.abs(p.monolithic_solution - xg).max() / np.abs(xg).max()
    assert esp < 1e-8, esp
    if p.use_peer_exchange:
        assert comm.n_allreduce - ar0 <= 2, "the preconditioned solve made allreduce calls"   # _allsum of the solution only
    its_block = p.solver_control.last_step()
    p.preconditioner_type = "None"
    from bemstokes_b200._lib import lib, check
    check(lib.bs_precond_setup(p._ctx, 2, 0, 0))
    # DN operator of a rigid mode on the sharded V and K (bs_dn_operator_multi: sweep over K, gather, lockstep V-solve, gather)
    dn = p.dirichlet_to_neumann_operator(pre.N_rigid[0])
    dn_o = pre.P(np.linalg.solve(Vc, pre.P(Kc @ pre.P(pre.N_rigid[0]))))
    edn = np.abs(dn - dn_o).max() / np.abs(dn_o).max()
    assert edn < 1e-8, edn
    # six right-hand sides in lockstep through the same exchange path
    if world >= 1:
        nr = 6
        Bm = np.zeros((nr, 3 * N + 6))
        for r in range(nr):
            Bm[r, 3 * N + r] = 1.0
        Xm = np.zeros_like(Bm)
        own_mask = p._owned_mask(6)
        p.gmres_multi(2, Xm, Bm)
        Xm[:, ~own_mask] = 0.0
        p._allsum(Xm)
        Xo = np.linalg.solve(A, Bm.T).T
        eb = np.abs(Xm - Xo).max() / np.abs(Xo).max()
        assert eb < 1e-7, eb
    print("rank %d/%d ok: owned %d nodes, entry err K %.1e A %.1e, GMRES its %d (oracle %d; block-Jacobi %d), solution err %.1e, "
          "allgathers %d allreduces %d (%d up to the first solve)" % (rank, world, len(own), ek, ea, its_plain, its, its_block, es,
                                           comm.n_allgather, comm.n_allreduce, n_ar_solve), flush=True)
    p.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
