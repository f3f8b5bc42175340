#!/usr/bin/env python
"""bench.py — headline benchmark of the BEMStokes hot path on B200 (contract: see the task statement).

One step = one pass of the hot path on a synthetic quad sphere:
    assemble V,K (K0 geometry + K1 regular + K2 singular)  ->  V / K corrections  ->  monolithic build
    ->  GMRES to 1e-10 (absolute, left-preconditioned, deal.II semantics)
Headline `value` = assembly Gentries/s (2*(3N)^2 entries of V and K / assembly time, whole job, inputs resident
in HBM); the line also carries the GMRES matvec HBM GB/s and the time-to-solution of the step — the three parts
of BASELINE.json's metric.  `e2e` = the same assembly metric through the public BEMProblem / C-ABI call with
HOST buffers (geometry H2D inside the timed region, V*n check vector D2H), plus the host-to-host time to solution,
both means over `--steps` repetitions.

Workloads (config.workload): the default is the BASELINE config-4 family — synthetic cube-sphere, Q1, Gauss 8 /
Lachat-Watson 10, translating sphere — with m = round(128 * (N/8)^(1/4)) subdivisions per face edge, i.e. the same
matrix bytes per GPU (87 GB of FP64 A) at every N (weak scaling, rows sharded, no data-path collective in
assembly).  At N=8 this is BASELINE config 4 exactly (98 306 nodes, 294 918 DoF, 696 GB matrix); at N=1 it is the
largest member one 180 GB B200 holds (34 658 nodes, 103 974 DoF).  These sizes use the fused no-K assembly
(bs_assemble_fused: the double-layer tile is consumed in the tile epilogue, V and K would not fit together).
--workload vk keeps both matrices (m = 64 * N^(1/4), 43.5 GB each per GPU); --workload q2 runs BASELINE config 3
(Q2, Gauss 15 / singular 20); --workload c5 / c5ns the image kernels of BASELINE config 5.

`parity_at_scale`: at every N each rank re-assembles the raw operators and compares rows of the stored V (and K
when it is stored) of collocation nodes it owns with the C port of the oracle (same mesh, same rules), entry by
entry; the run fails (exit 3) above 1e-12 of the row scale.

The default 1-GPU run also measures `secondary_workloads` (free-surface and no-slip image kernels, Q2 / config 3,
6 batched right-hand sides / config 2) on smaller meshes, each with its own roofline fraction.

`--impl reference` times the reference's CPU path (the C/OpenMP restatement under oracle/, all host threads
regardless of OMP_NUM_THREADS) on a bounded row sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "assembly Gentries/s + GMRES matvec HBM GB/s; time-to-solution at 3N DoF"
PARITY_TOL_K_PATCH = 1e-8
PARITY_TOL = 1e-12


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def usable_cores():
    """Host threads this process may use: the affinity mask, NOT OMP_NUM_THREADS (torchrun exports 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload(args, nranks):
    kind = args.workload
    if kind == "q2":
        r = args.refine if args.refine is not None else 4
        return dict(kind="cubesphere", m=2 ** r, degree=2, quad=15, sing=20, fused=False,
                    name="BASELINE config 3: cube-sphere (sphere_2.inp topology) refined %dx, Q2, Gauss 15 / Mixed 20" % r)
    if kind in ("c5", "c5ns"):
        m = args.m if args.m else int(round(64 * nranks ** 0.25))
        kern = "free_surface" if kind == "c5" else "no_slip"
        fused = bool(getattr(args, "fused", False))
        if fused and not args.m:
            m = int(round(128 * (nranks / 8.0) ** 0.25))   # the config-4 size: one 87 GB matrix per GPU
        return dict(kind="cubesphere", m=m, degree=1, quad=8, sing=10, fused=fused, kernel=kern, mixed=True,
                    name="BASELINE config 5: synthetic cube-sphere m=%d, Q1, Gauss 8 / Lachat-Watson 10, %s (image system, wall "
                         "y = 1.4), mixed velocity / traction unknowns (tangential velocities on the nodes facing the wall), %s"
                         % (m, "FreeSurfaceStokesKernel" if kind == "c5" else "NoSlipWallStokesKernel",
                            "fused no-K assembly (-K kept for the flagged columns only)" if fused else "V and K both stored"))
    if kind == "vk":
        m = args.m if args.m else int(round(64 * nranks ** 0.25))
        return dict(kind="cubesphere", m=m, degree=1, quad=8, sing=10, fused=False,
                    name="synthetic cube-sphere m=%d (6*m^2 quads), Q1 collocation, Gauss 8 / Lachat-Watson 10, translating "
                         "sphere ImposedVelocity e_x, V and K both stored" % m)
    m = args.m if args.m else int(round(128 * (nranks / 8.0) ** 0.25))
    return dict(kind="cubesphere", m=m, degree=1, quad=8, sing=10, fused=not args.no_fused,
                name="BASELINE config 4 family: synthetic cube-sphere m=%d (6*m^2 quads; m=128 at 8 GPUs = 98 306 nodes), Q1 "
                     "collocation, Gauss 8 / Lachat-Watson 10, translating sphere ImposedVelocity e_x, fused no-K assembly" % m)


def mesh_sizes(wl):
    """(nodes, cells) of the closed cube-sphere without building it."""
    ncell = 6 * wl["m"] ** 2
    return (ncell + 2 if wl["degree"] == 1 else 4 * ncell + 2), ncell


def config_of(wl, world):
    """Workload description, identical on both arms (the driver compares it)."""
    N, ncell = mesh_sizes(wl)
    n = 3 * N
    return {"workload": wl["name"], "nodes": N, "cells": ncell, "dofs": n, "matrix_gb_each": 8.0 * n * n / 1e9,
            "sharding": "rows (collocation nodes) over %d GPU(s), contiguous ranges of the locality order" % world,
            "l2": "working set (%.1f GB of matrix per GPU) far larger than the 126 MB L2, no explicit flush needed"
                  % (8.0 * n * n / 1e9 / world * (1 if wl.get("fused") else 2)),
            "value_definition": "2*(3N)^2 entries of V and K evaluated / (K0+K1+K2 device time, CUDA events on the launching "
                                "stream); ms_per_step is the whole step (assembly + corrections + monolithic build + GMRES)"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def load_traffic(wl, world):
    """ncu-measured DRAM traffic of the two dominant kernels for exactly this workload (profiles/*_traffic.json,
    newest round first)."""
    key = "c4:m=%d:%s:n_gpus=%d" % (wl["m"], "fused" if wl.get("fused") else "vk", world)
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            if t.get("workload_key") == key:
                return t["k_gemv<2>"]["dram_bytes_per_launch"], t["k_assemble_regular"]["dram_bytes_per_assembly"], name, t["k_assemble_regular"]
        except Exception:
            pass
    return None, None, None, {}


def f_pair_of(wl):
    """Algorithmic flops per (node, quadrature point) pair, SURVEY §8d."""
    na = 4 if wl["degree"] == 1 else 9
    k = wl.get("kernel")
    if k == "free_surface":
        return 95 + 36 * na
    if k == "no_slip":
        return 190 + 36 * na
    return 50 + 24 * na


def pairs_count(lib, _lib, wl):
    N, ncell = mesh_sizes(wl)
    na = 4 if wl["degree"] == 1 else 9
    sing_pts = sum(lib.bs_make_singular_rule(_lib.SING_MIXED, wl["sing"], wl["degree"], a, 0, None, None) for a in range(na))
    # every (row node, cell) pair uses nq points, except the na singular pairs of each cell, which use their rule
    return N * ncell * wl["quad"] ** 2, sing_pts * ncell


def oracle_kernel_spec(bo, wl):
    k = wl.get("kernel")
    if k == "free_surface":
        return bo.KernelSpec(bo.FREE_SURFACE, 0.0, 1, (0.0, 1.4, 0.0))
    if k == "no_slip":
        return bo.KernelSpec(bo.NO_SLIP, 0.0, 1, (0.0, 1.4, 0.0))
    return bo.KernelSpec()


def set_problem_kernel(p, wl):
    # tests/parameters_test_alpha_box.prm: wall 0 spans 80,0,80 at y = 1.4
    if wl.get("kernel") == "free_surface":
        p.reflect_kernel, p.wall_spans_0, p.wall_position_0 = True, (80.0, 0.0, 80.0), (0.0, 1.4, 0.0)
    elif wl.get("kernel") == "no_slip":
        p.no_slip_kernel, p.wall_spans_0, p.wall_position_0 = True, (80.0, 0.0, 80.0), (0.0, 1.4, 0.0)


# ---------------------------------------------------------------------------------------------------------------
def parity_at_scale(p, wl, geo, nodes_per_rank, extra_rows=None):
    """Rows of the stored raw V (and K) of nodes this rank owns against the C port (the checker, not the product).
    Returns (rows compared, max row-relative error of V, of K or None)."""
    from oracle import bem_oracle as bo, port
    from bemstokes_b200._lib import lib, check
    from bemstokes_b200 import _lib
    N = geo.N
    n = 3 * N
    if p.fused_assembly:   # raw operators again (the timed steps corrected V in place)
        nh, mn = np.ascontiguousarray(p.normal_vector_pure), np.ascontiguousarray(p.M_normal_vector_pure)
        Nr0 = np.ascontiguousarray(p.N_rigid[:p.num_rigid])
        dp = _lib.c_double_p
        check(lib.bs_assemble_fused(p._ctx, p.num_rigid, Nr0.ctypes.data_as(dp), nh.ctypes.data_as(dp), mn.ctypes.data_as(dp),
                                    p.l2normGamma_pure, None))
    else:
        check(lib.bs_assemble_VK(p._ctx))
    own = p.owned_nodes()
    pick = own[np.unique(np.linspace(0, len(own) - 1, max(1, nodes_per_rank)).astype(np.int64))]
    kspec = oracle_kernel_spec(bo, wl)
    cols = np.arange(n, dtype=np.int32)
    ev = ek = eks = 0.0
    have_k = not p.fused_assembly
    count = 0
    conn = np.asarray(geo.conn)

    def compare(i, Vp, Kp):
        nonlocal ev, ek, eks, count
        # columns of the cells that contain node i: their double-layer entries integrate (R.n) / r^5 with R.n = O(r^2) at
        # Lachat-Watson points next to the node, so rounding is amplified by 1 / r_min; the NumPy oracle and the C port
        # themselves differ by 1e-10 of the row scale there (1e-15 everywhere else): reported separately
        patch = np.unique(conn[(conn == i).any(axis=1)])
        sing = np.zeros(n, dtype=bool)
        for c in range(3):
            sing[patch + c * N] = True
        for c in range(3):
            r = np.full(n, i + c * N, dtype=np.int32)
            vg = p.V_matrix.entries(r, cols)
            ev = max(ev, float(np.abs(vg - Vp[c]).max() / np.abs(Vp[c]).max()))
            if have_k:
                kg = p.K_matrix.entries(r, cols)
                d = np.abs(kg - Kp[c]) / np.abs(Kp[c]).max()
                ek = max(ek, float(d[~sing].max()))
                eks = max(eks, float(d[sing].max()))
            count += 1

    for i in pick:
        Vp, Kp, _ = port.assemble_VK(geo, kspec, wl["quad"], "Mixed", wl["sing"], int(i), int(i) + 1, 1)
        compare(int(i), Vp, Kp)
    if extra_rows is not None:   # the rows the cpu_baseline leg assembled anyway (single rank: all owned)
        r0, Vp, Kp = extra_rows
        nr = Vp.shape[0] // 3
        ownset = set(int(x) for x in own)
        for k in range(0, nr, max(1, nr // 16)):
            if r0 + k in ownset:
                compare(r0 + k, Vp[[k, k + nr, k + 2 * nr]], Kp[[k, k + nr, k + 2 * nr]])
    return count, ev, (ek if have_k else None), (eks if have_k else None)


# ---------------------------------------------------------------------------------------------------------------
def secondary_workloads(device, fp64_peak, args):
    """Smaller members of BASELINE configs 5, 3 and 2 measured in the same process (1 GPU), each with its own
    roofline fraction (pairs x F_pair / K0+K1+K2 time / FP64 peak)."""
    import bemstokes_b200 as bb
    from bemstokes_b200 import _lib
    from bemstokes_b200._lib import lib, check
    import torch
    out = {}

    def assembly_case(tag, wl, reps=3):
        mesh = bb.cubesphere(degree=wl["degree"], m=wl["m"])
        p = bb.BEMProblem(device=device)
        p.set_mesh(mesh)
        p.quadrature_order, p.singular_quadrature_order = wl["quad"], wl["sing"]
        set_problem_kernel(p, wl)
        p.reinit()
        check(lib.bs_assemble_VK(p._ctx))   # warm-up
        p.reset_stats()
        for _ in range(reps):
            check(lib.bs_assemble_VK(p._ctx))
        st = p.stats()
        ms = (st["geometry_ms"] + st["assemble_regular_ms"] + st["assemble_singular_ms"]) / reps
        n = 3 * mesh.n_nodes
        pr, ps = pairs_count(lib, _lib, wl)
        fp = f_pair_of(wl)
        tf = (pr + ps) * fp / (ms * 1e-3) / 1e12
        out[tag] = {"workload": wl["name"], "nodes": mesh.n_nodes, "value": 2.0 * n * n / (ms * 1e-3) / 1e9, "unit": "Gentries/s",
                    "assembly_ms": ms, "flops_per_pair": fp, "achieved_tflops": tf, "frac": tf / fp64_peak}
        p.close()

    # ---- config 1: the reference's own CPU-runnable case (debug_grids/sphere_mesh_3d_0.msh, 386 nodes): whole step
    try:
        m1 = bb.read_mesh(os.path.join(ROOT, "tests", "golden", "meshes", "sphere_mesh_3d_0.msh"))
        p = bb.BEMProblem(device=device)
        p.set_mesh(m1)
        p.quadrature_order, p.singular_quadrature_order = 8, 10
        p.grid_type, p.imposed_component = "ImposedVelocity", 0
        p.solve_directly, p.preconditioner_type = False, "None"
        p.solver_control.tolerance, p.gmres_restart = 1e-10, 200
        p.reinit()

        def c1_step():
            p.compute_center_of_mass_and_rigid_modes()
            p.compute_normal_vector()
            p.assemble_stokes_system(True)
            p.monolithic_solution[:] = 0
            p.solve_system(True)
        c1_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            c1_step()
        torch.cuda.synchronize()
        out["c1_sphere_mesh_3d"] = {"workload": "BASELINE config 1: debug_grids/sphere_mesh_3d_0.msh, Q1, Gauss 8 / Lachat-Watson 10, "
                                                "translating body, pre-pass + assembly + GMRES through the public API",
                                    "nodes": m1.n_nodes, "time_to_solution_ms": 1e3 * (time.perf_counter() - t0) / 5,
                                    "gmres_iterations": p.solver_control.last_step(),
                                    "drag_over_6pi_a_eq": float(p.rigid_total_forces[0] / (6 * math.pi * 0.861)),
                                    "note": "node radii 0.849-0.874: equivalent radius 0.861 (SURVEY §8)"}
        p.close()
    except Exception as e:
        out["c1_sphere_mesh_3d"] = {"error": repr(e)}
    a = argparse.Namespace(workload="c5", m=args.secondary_m, refine=None, no_fused=False)
    assembly_case("c5_free_surface", workload(a, 1))
    a.workload = "c5ns"
    assembly_case("c5_no_slip", workload(a, 1))
    a.workload, a.refine = "q2", args.secondary_refine
    assembly_case("q2_config3", workload(a, 1))
    # ---- config 2: prolate spheroid (x scaled by 2), full 6x6 resistance matrix, 6 batched right-hand sides
    m2 = args.secondary_m
    mesh = bb.cubesphere(degree=1, m=m2, scale=(2.0, 1.0, 1.0))
    p = bb.BEMProblem(device=device)
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = 8, 10
    p.grid_type, p.imposed_component = "ImposedVelocity", 0
    p.solve_directly, p.preconditioner_type, p.keep_VK = False, "None", False
    p.solver_control.tolerance, p.solver_control.max_steps, p.gmres_restart = 1e-10, 1000, 200
    p.reinit()
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    p.assemble_stokes_system(True)
    n = p.n_dofs
    p.resistance_matrix()   # warm-up (workspaces)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Rm = p.resistance_matrix()
    torch.cuda.synchronize()
    t_b = time.perf_counter() - t0
    its_b = list(p.last_steps)
    t0 = time.perf_counter()
    for r in range(6):
        p.monolithic_rhs[:] = 0
        p.monolithic_rhs[n + r] = 1
        p.monolithic_solution[:] = 0
        p.solve_system(True)
    torch.cuda.synchronize()
    t_s = time.perf_counter() - t0
    out["resistance_6rhs"] = {"workload": "BASELINE config 2: synthetic prolate spheroid (cube-sphere m=%d, x scaled by 2), Q1, "
                                          "6 x 6 resistance matrix, 6 batched right-hand sides" % m2,
                              "nodes": mesh.n_nodes, "batched_s": t_b, "sequential_s": t_s, "speedup": t_s / t_b,
                              "iterations": its_b, "R_diag": [float(Rm[i, i]) for i in range(6)],
                              "R_offdiag_over_diag_max": float(np.abs(Rm - np.diag(np.diag(Rm))).max() / np.abs(np.diag(Rm)).min())}
    p.close()
    # ---- DN-operator route: 1 + 6 V-systems of solve_system(false) as one device batch against the
    # reference's sequence of single calls; V and K both stored
    p = bb.BEMProblem(device=device)
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = 8, 10
    p.grid_type, p.monolithic_bool = "Real", False
    p.solve_directly, p.preconditioner_type = False, "None"
    p.solver_control.tolerance, p.solver_control.max_steps, p.gmres_restart = 1e-10, 1000, 200
    p.reinit()
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    x = mesh.nodes
    p.shape_velocities = np.concatenate([np.sin(x[:, 0]) * x[:, 1], 0.5 * x[:, 1] * x[:, 2], 0.3 * x[:, 0] * x[:, 1] - 0.1])
    p.assemble_stokes_system(True)
    p.solve_dn(batched=True)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p.solve_dn(batched=True)
    torch.cuda.synchronize()
    t_b = time.perf_counter() - t0
    its_b = list(p.last_steps)
    Ub = p.rigid_velocities.copy()
    t0 = time.perf_counter()
    p.solve_dn(batched=False)
    torch.cuda.synchronize()
    t_s = time.perf_counter() - t0
    out["dn_route_7rhs"] = {"workload": "solve_system(false): DN operator of the swimming stroke and of the 6 rigid modes on the same "
                                        "prolate mesh (V and K stored)", "nodes": mesh.n_nodes, "batched_s": t_b, "sequential_s": t_s,
                            "speedup": t_s / t_b, "iterations": its_b,
                            "rigid_velocities_batched_vs_sequential": float(np.abs(Ub - p.rigid_velocities).max())}
    p.close()
    return out


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import bemstokes_b200 as bb
    from bemstokes_b200 import _lib
    from bemstokes_b200._lib import lib, check
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
        from bemstokes_b200.comm import TorchComm
        comm = TorchComm(device=dev)
    wl = workload(args, world)
    mesh = bb.cubesphere(degree=wl["degree"], m=wl["m"])
    N, ncell = mesh.n_nodes, mesh.n_cells
    assert (N, ncell) == mesh_sizes(wl)
    n = 3 * N
    p = bb.BEMProblem(device=local, rank=rank, nranks=world, comm=comm)
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = wl["quad"], wl["sing"]
    set_problem_kernel(p, wl)
    if wl.get("mixed"):   # the unknowns facing the wall are tangential velocities: -K columns (ref: bem_stokes.cc:3194-3245)
        flags = np.zeros(n, dtype=bool)
        top = mesh.nodes[:, 1] > 0.6
        for cmp in (0, 2):
            flags[cmp * N:(cmp + 1) * N] = top
        p.col_is_K = flags
    p.grid_type, p.imposed_component = "ImposedVelocity", 0
    p.solve_directly, p.preconditioner_type = False, args.preconditioner
    p.keep_VK = False  # A aliases V's storage
    p.fused_assembly = bool(wl.get("fused"))
    p.use_peer_exchange = not args.no_peer_exchange
    p.solver_control.tolerance, p.solver_control.max_steps = 1e-10, 1000
    p.gmres_restart = 200
    p.preconditioner_block = args.prec_block
    p.reinit()
    # pre-pass (mass matrix, L2 normals, rigid modes; bs_prepass on the device): input of the hot path
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    n_own = len(p.owned_nodes())

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        """one pass of the hot path through the public API; geometry already resident"""
        p.assemble_stokes_system(True)
        p.monolithic_solution[:] = 0
        p.solve_system(True)

    def step_e2e():
        """host buffers in (geometry, quadrature, normals, rigid modes), host result out"""
        t0 = time.perf_counter()
        p.update_geometry()              # bs_set_geometry: host euler vector -> device (per-frame flow of the reference)
        t1 = time.perf_counter()
        p.compute_center_of_mass_and_rigid_modes()   # bs_prepass: device pre-pass, results to host arrays
        p.compute_normal_vector()
        torch.cuda.synchronize()
        t_pre = time.perf_counter() - t1
        vn = np.zeros(n)
        nh = np.ascontiguousarray(p.normal_vector_pure)
        mn = np.ascontiguousarray(p.M_normal_vector_pure)
        fl = None if p.col_is_K is None else np.ascontiguousarray(p.col_is_K, dtype=np.uint8)
        flp = fl.ctypes.data_as(_lib.c_ubyte_p) if fl is not None else None
        if p.fused_assembly:
            Nr0 = np.ascontiguousarray(p.N_rigid[:p.num_rigid])
            check(lib.bs_set_column_flags(p._ctx, flp))
            check(lib.bs_assemble_fused(p._ctx, p.num_rigid, Nr0.ctypes.data_as(_lib.c_double_p), nh.ctypes.data_as(_lib.c_double_p),
                                        mn.ctypes.data_as(_lib.c_double_p), p.l2normGamma_pure, None))
        else:
            check(lib.bs_assemble_VK(p._ctx))
        check(lib.bs_correct_V(p._ctx, nh.ctypes.data_as(_lib.c_double_p), mn.ctypes.data_as(_lib.c_double_p),
                               p.l2normGamma_pure, vn.ctypes.data_as(_lib.c_double_p)))  # D2H of V*n
        torch.cuda.synchronize()
        t_asm = time.perf_counter() - t0 - t_pre
        check(lib.bs_correct_K(p._ctx, 0))
        # finish the step (monolithic + GMRES) to get the host-to-host time to solution
        nr = p.num_rigid
        p.monolithic_rhs = np.zeros(n + nr)
        Nr, Nd = np.ascontiguousarray(p.N_rigid[:nr]), np.ascontiguousarray(p.N_rigid_dual[:nr])
        sv = np.zeros(n)
        dp = _lib.c_double_p
        check(lib.bs_build_monolithic(p._ctx, flp, nr, Nr.ctypes.data_as(dp), Nd.ctypes.data_as(dp), nh.ctypes.data_as(dp),
                                      mn.ctypes.data_as(dp), p.l2normGamma_pure, _lib.GRID_IMPOSED_VELOCITY, 0, 1.0,
                                      sv.ctypes.data_as(dp), 0, p.monolithic_rhs.ctypes.data_as(dp)))
        p.monolithic_solution = np.zeros(n + nr)
        p.solve_system(True)
        torch.cuda.synchronize()
        return t_asm, time.perf_counter() - t0, t_pre

    # ---- warm-up ----
    for _ in range(args.warmup):
        step()
    sync()
    # ---- timed: K steps, device-event timers inside the library + wall clock around the whole loop ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    p.reset_stats()
    if comm is not None:
        comm.n_allgather = comm.n_allreduce = 0
    sync()
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        step()
        its = p.solver_control.last_step()
    sync()
    wall = time.perf_counter() - t0
    st = p.stats()
    n_ag, n_ar = (comm.n_allgather, comm.n_allreduce) if comm else (0, 0)
    clocks = sampler.stop() if rank == 0 else None
    drag = float(p.rigid_total_forces[0])  # every rank holds the gathered solution
    final_check = getattr(p, "final_check_0", (None, None))
    asm_ms = (st["geometry_ms"] + st["assemble_regular_ms"] + st["assemble_singular_ms"]) / args.steps
    # ---- matvec bandwidth: device-resident GEMV loop on the monolithic matrix (L2 flushed by its own 87 GB) ----
    ms_mv = C.c_double()
    check(lib.bs_bench_vmult(p._ctx, _lib.MAT_A, 10, C.byref(ms_mv)))
    # ---- optional: frames with the frame-0 block LU reused as the preconditioner (ref: bem_stokes.cc:5768-5779) ----
    frames = None
    if args.frames > 0:
        frames = run_frames(p, mesh, args, sync, torch)
    # ---- e2e through host buffers: mean over `steps` ----
    e2e_asm, e2e_tts, e2e_pre = [], [], []
    for _ in range(max(1, args.steps)):
        a, b, c_ = step_e2e()
        e2e_asm.append(a)
        e2e_tts.append(b)
        e2e_pre.append(c_)
    sync()
    # ---- parity at this size against the C port (checker) ----
    cb = None
    extra = None
    geo = None
    if not args.no_parity or (not args.no_cpu_baseline and world == 1 and rank == 0):
        from oracle import bem_oracle as bo
        nodes_o, conn_o = bo.cubesphere(degree=wl["degree"], m=wl["m"])   # the oracle's own mesh generator
        geo = bo.Geometry(nodes_o, conn_o, wl["degree"])
    if not args.no_cpu_baseline and world == 1:
        cb, extra = cpu_baseline(wl, sample_seconds=args.cpu_seconds, geo=geo, keep_rows=True, iterations=its)
    par = (0, 0.0, None, None)
    if not args.no_parity:
        par = parity_at_scale(p, wl, geo, args.parity_nodes, extra)
    sync()
    # max / sum over ranks
    vals = torch.tensor([wall, asm_ms, ms_mv.value, float(np.mean(e2e_asm)), float(np.mean(e2e_tts)), par[1],
                         par[2] if par[2] is not None else 0.0, st["solve_ms"] / args.steps,
                         par[3] if par[3] is not None else 0.0], dtype=torch.float64, device=dev)
    cnt = torch.tensor([par[0]], dtype=torch.float64, device=dev)
    # rest of the GMRES iteration besides the sweeps over the matrix, per rank: a rank's figure contains its wait for the
    # slowest rank's sweep, so the minimum over ranks is the overhead proper and the maximum the skew on top of it
    nmv = (st["gmres_stream_ms_last"] - st["gmres_matvec_ms_last"]) / max(1, its)
    nmv_mm = torch.tensor([nmv, -nmv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(nmv_mm, op=dist.ReduceOp.MIN)
    nmv_min, nmv_max = float(nmv_mm[0]), -float(nmv_mm[1])
    wall, asm_ms, mv_ms, e2e_asm_s, e2e_tts_s, par_v, par_k, solve_ms, par_ks = [float(v) for v in vals.cpu()]
    par_rows = int(cnt.item())
    entries = 2.0 * n * n
    rows_loc = 3 * n_own + (6 if rank == world - 1 else 0)
    parity_ok = True
    if rank == 0:
        peaks, peak_src = load_peaks()
        fp64 = C.c_double()
        check(lib.bs_bench_fp64_sustained(local, 1.0, C.byref(fp64)))  # K1 runs for 100s of ms under the power cap
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp64_silicon = 148 * 64 * 2 * sm_max * 1e6 / 1e12
        pr, ps = pairs_count(lib, _lib, wl)
        f_pair = f_pair_of(wl)
        asm_tflops = (pr + ps) * f_pair / (asm_ms * 1e-3) / 1e12   # whole job (all ranks assemble concurrently)
        mv_bytes = 8.0 * (n + 6) * (n + 6)                          # whole job: every rank streams its row block
        mv_gbs_job = mv_bytes / (mv_ms * 1e-3) / 1e9
        mv_gbs_gpu = mv_gbs_job / world
        traffic_mv, traffic_asm, traffic_src, asm_ncu = load_traffic(wl, world)
        roof_mv = {"kernel": "k_gemv<2>", "bound": "hbm", "achieved": mv_gbs_gpu, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": mv_gbs_gpu / peaks["hbm_gbs"], "traffic": traffic_mv, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                   "algorithmic_bytes_per_launch": 8.0 * rows_loc * (n + 6), "launch_ms": mv_ms}
        roof_asm = {"kernel": "k_assemble_regular", "bound": "fp64", "achieved": asm_tflops / world, "peak": fp64.value,
                    "unit": "TFLOP/s", "frac": asm_tflops / world / fp64.value, "traffic": traffic_asm,
                    "traffic_note": "DRAM bytes of one assembly (all colour launches); profiles/%s" % traffic_src,
                    "peak_source": "measured in this run: 8-chain DFMA microbenchmark sustained for 1 s under the power cap "
                                   "(bs_bench_fp64_sustained); silicon figure 148 SM x 64 DFMA x 2 x max clock in frac_of_silicon_peak",
                    "peak_silicon": fp64_silicon, "frac_of_silicon_peak": asm_tflops / world / fp64_silicon,
                    "algorithmic_flops_per_pair": f_pair, "pairs_regular": pr, "pairs_singular": ps, "launch_ms": asm_ms,
                    # what the kernel really executes (ncu, profiles/): the moment formulation needs fewer FP64 instructions
                    # than the convention's 73 FMA-equivalents per pair, so `frac` (algorithmic flops / peak) can exceed the
                    # FP64-pipe utilisation - and 1
                    "executed_fp64_instructions_per_pair": asm_ncu.get("fp64_instructions_per_pair"),
                    "fp64_pipe_active_pct_ncu": asm_ncu.get("fp64_pipe_active_pct")}
        dominant_is_asm = asm_ms >= solve_ms
        parity_ok = par_v <= PARITY_TOL and (par_k or 0.0) <= PARITY_TOL and (par_ks or 0.0) <= PARITY_TOL_K_PATCH
        parity = {"rows": par_rows, "max_row_rel_err_V": par_v, "max_row_rel_err_K": (par_k if not wl.get("fused") else None),
                  # K columns of the cells containing the row's node (cancellation in R.n at the singular rule's points: the
                  # oracle and its C port differ by 1e-10 there themselves), own tolerance
                  "max_row_rel_err_K_own_cells": (par_ks if not wl.get("fused") else None), "tolerance_K_own_cells": PARITY_TOL_K_PATCH,
                  "tolerance": PARITY_TOL, "ok": bool(parity_ok), "checker": "oracle/bem_port.c rows of nodes owned by every rank"}
        if args.no_parity:
            parity = None
        gm_it = solve_ms / max(1, its)
        summary = {"time_to_solution_s": wall / args.steps, "gmres_iterations": its, "gmres_ms_per_iteration": gm_it,
                   # CUDA events around every sweep over the matrix inside the device-resident solve (last step): the rest
                   # (Gram-Schmidt passes with their cross-rank sums, Givens, publish, waits) per iteration
                   "gmres_non_matvec_ms_per_iteration": nmv_min, "gmres_non_matvec_ms_per_iteration_max_over_ranks": nmv_max,
                   "gmres_matvec_ms_in_solve": st["gmres_matvec_ms_last"] / max(1, st["gmres_sweeps_last"]),
                   "ortho": p.gmres_orthogonalization,
                   "preconditioner": args.preconditioner, "drag_over_6pi": drag / (6 * math.pi),
                   "final_check": {"linf": final_check[0], "l2": final_check[1]},
                   "phases_ms": {"assembly": asm_ms, "corrections": st["correct_ms"] / args.steps,
                                 "monolithic": st["monolithic_ms"] / args.steps, "precond_setup": st["precond_setup_ms"] / args.steps,
                                 "gmres": solve_ms},
                   "matvec_hbm_gbs": mv_gbs_job, "matvec_hbm_gbs_per_gpu": mv_gbs_gpu,
                   "collective_calls_in_timed_region": {"allgather": n_ag, "allreduce": n_ar},
                   "exchange": ("none (1 GPU)" if world == 1 else (
                       "NVLink peer stores (Krylov slices and Gram-Schmidt partial sums) from the producing kernels"
                       if p.use_peer_exchange else "NCCL allgather + allreduce callbacks")),
                   "parity_at_scale": parity}
        line = {
            "metric": METRIC, "value": entries / (asm_ms * 1e-3) / 1e9, "unit": "Gentries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, world),
            "summary": summary,
            "e2e": {"value": entries / e2e_asm_s / 1e9, "unit": "Gentries/s", "time_to_solution_s": e2e_tts_s,
                    "repetitions": len(e2e_asm), "statistic": "mean",
                    "prepass_s": float(np.mean(e2e_pre)), "prepass_cg_iterations": int(p._pre.cg_iterations),
                    "gmres_iterations": its, "drag_over_6pi": drag / (6 * math.pi), "parity_at_scale": parity,
                    "note": "time_to_solution_s is host to host per frame: geometry H2D, device pre-pass (mass matrix, L2 "
                            "normals, rigid modes), assembly, corrections, monolithic build, GMRES, solution D2H; value "
                            "excludes the pre-pass",
                    "h2d_bytes_per_step": int(8 * 3 * N + 4 * 2 * ncell * (4 if wl["degree"] == 1 else 9) + 8 * 2 * n + 8 * 12 * n + 8 * n),
                    "d2h_bytes_per_step": int(8 * n + 8 * (n + 6))},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": roof_asm if dominant_is_asm else roof_mv,
            "roofline_secondary": roof_mv if dominant_is_asm else roof_asm,
            "tiling": {"cell_blocks": int(st["n_cell_blocks"]), "colours": int(st["n_colours"]),
                       "node_touch_ratio": st["node_touch_ratio"], "cell_sets": int(st["cell_sets"]),
                       "cell_steps": int(st["cell_steps"]), "unpaired_cells": int(st["unpaired_cells"]),
                       "steps_behind_a_barrier": int(st["sync_steps"])},
            "fp64_peak_tflops_measured": fp64.value, "fp64_peak_tflops_silicon": fp64_silicon,
            "frames": frames,
            "clocks": clocks,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
    p.close()
    if rank == 0:
        if world == 1 and not args.no_secondary and args.workload == "c4":
            try:
                line["secondary_workloads"] = secondary_workloads(local, line["fp64_peak_tflops_measured"], args)
            except Exception as e:  # a secondary workload must never cost the headline line
                line["secondary_workloads"] = {"error": repr(e)}
        # the driver keeps the tail of a long line: repeat the figures that prove parity and the solve at the end
        line["tail_summary"] = {"value": line["value"], "tts_s": summary["time_to_solution_s"], "gmres_its": its,
                                "ms_per_it": summary["gmres_ms_per_iteration"], "non_matvec_ms_per_it": summary["gmres_non_matvec_ms_per_iteration"],
                                "drag_over_6pi": summary["drag_over_6pi"], "parity_rows": par_rows, "parity_err_V": par_v,
                                "parity_ok": bool(parity_ok), "allreduce_calls": n_ar, "allgather_calls": n_ag,
                                "e2e": line["e2e"]["value"], "assembly_ms": asm_ms, "gemv_tbs_per_gpu": mv_gbs_gpu / 1e3,
                                "gemv_frac_of_measured_copy_peak": roof_mv["frac"], "k1_frac_algorithmic": roof_asm["frac"],
                                "k1_fp64_instr_per_pair_ncu": roof_asm["executed_fp64_instructions_per_pair"],
                                "k1_fp64_pipe_active_pct_ncu": roof_asm["fp64_pipe_active_pct_ncu"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        sys.exit(3)


def run_frames(p, mesh, args, sync, torch):
    """`--frames K`: K geometry-perturbed frames solved with the LU of frame 0 as the (block-Jacobi) preconditioner,
    the reference's reuse logic (bem_stokes.cc:5768-5779, 4336-4339): time to solution per frame."""
    import bemstokes_b200 as bb
    base = mesh.nodes.copy()
    p.preconditioner_type = "BlockDirect"
    p.preconditioner_block = args.prec_block
    p.reassemble_preconditoner = True
    out = []
    for f in range(args.frames):
        # rigid rotation about z by 0.02 rad per frame plus a 1 % breathing mode: same mesh, new coordinates
        a = 0.02 * f
        R = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1.0]])
        nodes = (base * (1.0 + 0.01 * math.sin(0.7 * f) * base[:, 2:3] ** 2)) @ R.T
        sync()
        t0 = time.perf_counter()
        p.update_geometry(bb.QuadMesh(nodes, mesh.conn, mesh.degree))
        p.compute_center_of_mass_and_rigid_modes()
        p.compute_normal_vector()
        p.reset_stats()
        p.assemble_stokes_system(True)
        p.monolithic_solution[:] = 0
        p.solve_system(True)
        sync()
        st = p.stats()
        out.append({"frame": f, "tts_s": time.perf_counter() - t0, "gmres_iterations": p.solver_control.last_step(),
                    "lu_ms": st["precond_setup_ms"], "gmres_ms": st["solve_ms"]})
    p.preconditioner_type = args.preconditioner
    p.update_geometry(mesh)   # back to the benchmark geometry (the parity sample compares against it)
    p.compute_center_of_mass_and_rigid_modes()
    p.compute_normal_vector()
    return out


# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline(wl, sample_seconds=15.0, geo=None, keep_rows=False, iterations=None):
    """The reference's CPU path (oracle/bem_port.c) on a bounded row sample of the same workload: all usable host
    threads (affinity mask; OMP_NUM_THREADS is ignored) and, as BASELINE.md §3 asks, one thread = one reference MPI rank."""
    from oracle import bem_oracle as bo, port
    if geo is None:
        nodes, conn = bo.cubesphere(degree=wl["degree"], m=wl["m"])   # the oracle's own mesh generator: no product code here
        geo = bo.Geometry(nodes, conn, wl["degree"])
    cores = usable_cores()
    N = geo.N
    kspec = oracle_kernel_spec(bo, wl)
    # calibrate on one row per thread, then size the sample
    t0 = time.perf_counter()
    port.assemble_VK(geo, kspec, wl["quad"], "Mixed", wl["sing"], 0, cores, cores)
    t_cal = time.perf_counter() - t0
    rows = int(max(cores, min(N, cores * max(1, int(sample_seconds / max(t_cal, 1e-3))))))
    t0 = time.perf_counter()
    V, K, pairs = port.assemble_VK(geo, kspec, wl["quad"], "Mixed", wl["sing"], 0, rows, cores)
    t = time.perf_counter() - t0
    entries = 2.0 * 3 * rows * 3 * N
    # one thread: what one reference MPI rank does (the reference has no threading in assembly)
    rows1 = max(1, int(round(rows / cores * min(1.0, 4.0 / max(t, 1e-3)))))
    t0 = time.perf_counter()
    _, _, pairs1 = port.assemble_VK(geo, kspec, wl["quad"], "Mixed", wl["sing"], 0, rows1, 1)
    t1 = time.perf_counter() - t0
    # matvec sample: the assembled row block itself
    x = np.random.default_rng(0).uniform(-1, 1, 3 * N)
    port.gemv(V, x, cores)
    tm = time.perf_counter()
    reps = 5
    for _ in range(reps):
        port.gemv(V, x, cores)
    t_mv = (time.perf_counter() - tm) / reps
    rate = entries / t / 1e9
    mv_gbs = V.nbytes / t_mv / 1e9
    out = {"value": rate, "unit": "Gentries/s", "cores": cores, "kind": "port",
           "sample": "%d of %d collocation nodes (all cells, all quadrature points), %.1f s; matvec on the %.2f GB row block"
                     % (rows, N, t, V.nbytes / 1e9),
           "pairs_per_s": pairs / t, "matvec_gbs": mv_gbs,
           "one_thread": {"value": 2.0 * 3 * rows1 * 3 * N / t1 / 1e9, "unit": "Gentries/s", "cores": 1,
                          "sample": "%d nodes, %.1f s" % (rows1, t1)},
           "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
           "note": "C/OpenMP restatement of bem_stokes.cc:2871-2998; flatters the reference (no Epetra per-entry "
                   "insertion, threads over rows; the reference assembles single-threaded per MPI rank)"}
    if iterations:
        n = 3.0 * N
        out["time_to_solution_extrapolated_s"] = {
            "assembly": 2.0 * n * n / 1e9 / rate, "gmres": iterations * 8.0 * n * n / 1e9 / mv_gbs,
            "note": "extrapolated from the sample rates to the full 3N x 3N system and %d GMRES iterations "
                    "(the full matrix is not assembled on the host)" % iterations}
    if keep_rows:
        return out, (0, V, K)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import bem_oracle as bo
    wl = workload(args, world)
    nodes, conn = bo.cubesphere(degree=wl["degree"], m=wl["m"])
    geo = bo.Geometry(nodes, conn, wl["degree"])
    vals = []
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    cb = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cb = cpu_baseline(wl, sample_seconds=per_step, geo=geo)
        if i >= args.warmup:
            vals.append((cb["value"], time.perf_counter() - t0))
    v = float(np.mean([a for a, _ in vals]))
    line = {"impl": "reference", "metric": METRIC,
            "value": v, "unit": "Gentries/s", "matvec_hbm_gbs": cb["matvec_gbs"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([b for _, b in vals])), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(wl, world),
            "cpu_baseline": dict(cb, value=v),
            "e2e": {"value": v, "unit": "Gentries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "vk", "q2", "c5", "c5ns"])
    ap.add_argument("--no-peer-exchange", action="store_true", help="multi-GPU: NCCL allgather callbacks instead of NVLink peer stores")
    ap.add_argument("--no-fused", action="store_true", help="c4 family with V and K both stored (needs 2x the memory)")
    ap.add_argument("--fused", action="store_true", help="c5 / c5ns: fused no-K assembly at the config-4 size (mixed boundary conditions)")
    ap.add_argument("--subdiv", dest="m", type=int, default=0, help="cube-sphere subdivisions per face edge (default 128*(N/8)^(1/4); 64*N^(1/4) for --workload vk)")
    ap.add_argument("--refine", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_at_scale row sample")
    ap.add_argument("--parity-nodes", type=int, default=8, help="collocation nodes per rank compared with the C port")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (configs 5, 3, 2)")
    ap.add_argument("--secondary-m", type=int, default=48)
    ap.add_argument("--secondary-refine", type=int, default=4)
    ap.add_argument("--preconditioner", default="None", choices=["None", "Jacobi", "BlockDirect"])
    ap.add_argument("--frames", type=int, default=0, help="also solve K perturbed frames with the frame-0 block LU as preconditioner")
    ap.add_argument("--prec-block", type=int, default=26112, help="BlockDirect: largest diagonal block in rows (0 = the rank's whole row block)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
