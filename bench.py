#!/usr/bin/env python
"""bench.py — headline benchmark of the BEMStokes hot path on B200 (contract: see the task statement).

One step = one pass of the hot path on a synthetic quad sphere:
    assemble V,K (K0 geometry + K1 regular + K2 singular)  ->  V / K corrections  ->  monolithic build
    ->  GMRES to 1e-10 (absolute, left-preconditioned, deal.II semantics)
Headline `value` = assembly Gentries/s (2*(3N)^2 entries of V and K / assembly time, whole job, inputs resident
in HBM); the line also carries the GMRES matvec HBM GB/s and the time-to-solution of the step — the three parts
of BASELINE.json's metric.  `e2e` = the same assembly metric through the public BEMProblem / C-ABI call with
HOST buffers (geometry H2D inside the timed region, V*n check vector D2H), plus the host-to-host time to solution.

Workloads (config.workload): the default is the BASELINE config-4 family — synthetic cube-sphere, Q1, Gauss 8 /
Lachat-Watson 10, translating sphere — with m = round(128 * (N/8)^(1/4)) subdivisions per face edge, i.e. the same
matrix bytes per GPU (87 GB of FP64 A) at every N (weak scaling, rows sharded, no data-path collective in
assembly).  At N=8 this is BASELINE config 4 exactly (98 306 nodes, 294 918 DoF, 696 GB matrix); at N=1 it is the
largest member one 180 GB B200 holds (34 658 nodes, 103 974 DoF).  These sizes use the fused no-K assembly
(bs_assemble_fused: the double-layer tile is consumed in the tile epilogue, V and K would not fit together).
--workload vk keeps both matrices (m = 64 * N^(1/4), 43.5 GB each per GPU); --workload q2 runs BASELINE config 3
(Q2, Gauss 15 / singular 20).

`--impl reference` times the reference's CPU path (the C/OpenMP restatement under oracle/, all host threads)
on a bounded row sample of the same workload.
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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def workload(args, nranks):
    if args.workload == "q2":
        r = args.refine if args.refine is not None else 4
        return dict(kind="cubesphere", m=2 ** r, degree=2, quad=15, sing=20, fused=False,
                    name="BASELINE config 3: cube-sphere (sphere_2.inp topology) refined %dx, Q2, Gauss 15 / Mixed 20" % r)
    if args.workload == "c5":
        m = args.m if args.m else int(round(64 * nranks ** 0.25))
        return dict(kind="cubesphere", m=m, degree=1, quad=8, sing=10, fused=False, kernel="free_surface",
                    name="BASELINE config 5 kernel: synthetic cube-sphere m=%d, Q1, Gauss 8 / Lachat-Watson 10, "
                         "FreeSurfaceStokesKernel (image system, wall y = 1.4), V and K both stored" % m)
    if args.workload == "vk":
        m = args.m if args.m else int(round(64 * nranks ** 0.25))
        return dict(kind="cubesphere", m=m, degree=1, quad=8, sing=10, fused=False,
                    name="synthetic cube-sphere m=%d (6*m^2 quads), Q1 collocation, Gauss 8 / Lachat-Watson 10, translating "
                         "sphere ImposedVelocity e_x, V and K both stored" % m)
    m = args.m if args.m else int(round(128 * (nranks / 8.0) ** 0.25))
    return dict(kind="cubesphere", m=m, degree=1, quad=8, sing=10, fused=not args.no_fused,
                name="BASELINE config 4 family: synthetic cube-sphere m=%d (6*m^2 quads; m=128 at 8 GPUs = 98 306 nodes), Q1 "
                     "collocation, Gauss 8 / Lachat-Watson 10, translating sphere ImposedVelocity e_x, fused no-K assembly" % m)


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
    """ncu-measured DRAM traffic of the two dominant kernels for exactly this workload (profiles/r01_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    key = "c4:m=%d:%s:n_gpus=%d" % (wl["m"], "fused" if wl.get("fused") else "vk", world)
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("workload_key") == key:
            return t["k_gemv<2>"]["dram_bytes_per_launch"], t["k_assemble_regular"]["dram_bytes_per_assembly"]
    except Exception:
        pass
    return None, None


def pairs_count(n_rows_nodes, ncell, nq, sing_pts_per_cell):
    """(node, q-point) pairs: every (row node, cell) pair uses nq points, except the na singular pairs of each
    cell, which use their singular rule."""
    return n_rows_nodes * ncell * nq, sing_pts_per_cell * ncell


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
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
        from bemstokes_b200.comm import TorchComm
        comm = TorchComm(device=dev)
        # one explicit stream for the library's kernels AND torch's collectives (the legacy default stream is
        # not ordered against the library's non-blocking stream)
        side = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(side)
    wl = workload(args, world)
    mesh = bb.cubesphere(degree=wl["degree"], m=wl["m"])
    N, ncell = mesh.n_nodes, mesh.n_cells
    n = 3 * N
    stream = torch.cuda.current_stream().cuda_stream if world > 1 else None
    p = bb.BEMProblem(device=local, rank=rank, nranks=world, comm=comm, stream=stream)
    p.set_mesh(mesh)
    p.quadrature_order, p.singular_quadrature_order = wl["quad"], wl["sing"]
    if wl.get("kernel") == "free_surface":   # tests/parameters_test_alpha_box.prm: wall 0 spans 80,0,80 at y = 1.4
        p.reflect_kernel, p.wall_spans_0, p.wall_position_0 = True, (80.0, 0.0, 80.0), (0.0, 1.4, 0.0)
    p.grid_type, p.imposed_component = "ImposedVelocity", 0
    p.solve_directly, p.preconditioner_type = False, "None"
    p.keep_VK = False  # A aliases V's storage
    p.fused_assembly = bool(wl.get("fused"))
    p.use_peer_exchange = not args.no_peer_exchange
    p.solver_control.tolerance, p.solver_control.max_steps = 1e-10, 1000
    p.gmres_restart = 200
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
        if p.fused_assembly:
            Nr0 = np.ascontiguousarray(p.N_rigid[:p.num_rigid])
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
        check(lib.bs_build_monolithic(p._ctx, None, nr, Nr.ctypes.data_as(dp), Nd.ctypes.data_as(dp), nh.ctypes.data_as(dp),
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
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        step()
        its = p.solver_control.last_step()
    sync()
    wall = time.perf_counter() - t0
    st = p.stats()
    clocks = sampler.stop() if rank == 0 else None
    drag = float(p.rigid_total_forces[0])  # every rank holds the gathered solution
    asm_ms = (st["geometry_ms"] + st["assemble_regular_ms"] + st["assemble_singular_ms"]) / args.steps
    # ---- matvec bandwidth: device-resident GEMV loop on the monolithic matrix (L2 flushed by its own 43 GB) ----
    ms_mv = C.c_double()
    check(lib.bs_bench_vmult(p._ctx, _lib.MAT_A, 10, C.byref(ms_mv)))
    # ---- optional: BASELINE config-2 style batched solve (6 rigid-body right-hand sides in lockstep) ----
    res6 = None
    if args.resistance and world == 1:
        torch.cuda.synchronize()
        t6 = time.perf_counter()
        Rm = p.resistance_matrix()
        torch.cuda.synchronize()
        t_batched = time.perf_counter() - t6
        t6 = time.perf_counter()
        for r in range(6):
            p.monolithic_rhs[:] = 0
            p.monolithic_rhs[n + r] = 1
            p.monolithic_solution[:] = 0
            p.solve_system(True)
        torch.cuda.synchronize()
        t_seq = time.perf_counter() - t6
        res6 = {"batched_s": t_batched, "sequential_s": t_seq, "iterations": p.last_steps,
                "R_diag_over_6pi_8pi": [float(Rm[i, i] / (6 * math.pi if i < 3 else 8 * math.pi)) for i in range(6)]}
    # ---- e2e through host buffers ----
    e2e_asm, e2e_tts, e2e_pre = [], [], []
    for _ in range(max(1, min(args.steps, 2))):
        a, b, c_ = step_e2e()
        e2e_asm.append(a)
        e2e_tts.append(b)
        e2e_pre.append(c_)
    sync()
    # max over ranks
    vals = torch.tensor([wall, asm_ms, ms_mv.value, min(e2e_asm), min(e2e_tts)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    wall, asm_ms, mv_ms, e2e_asm_s, e2e_tts_s = [float(v) for v in vals.cpu()]
    entries = 2.0 * n * n
    rows_loc = 3 * n_own + (6 if rank == world - 1 else 0)
    if rank == 0:
        peaks, peak_src = load_peaks()
        fp64, fp64_burst = C.c_double(), C.c_double()
        check(lib.bs_bench_fp64_peak(local, C.byref(fp64_burst)))
        check(lib.bs_bench_fp64_sustained(local, 1.0, C.byref(fp64)))  # K1 runs for 100s of ms under the power cap
        na = 4 if wl["degree"] == 1 else 9
        nq = wl["quad"] ** 2
        sing_pts = sum(lib.bs_make_singular_rule(_lib.SING_MIXED, wl["sing"], wl["degree"], a, 0, None, None) for a in range(na))
        pr, ps = pairs_count(N, ncell, nq, sing_pts)
        f_pair = (95 + 36 * na) if wl.get("kernel") == "free_surface" else (50 + 24 * na)
        asm_tflops = (pr + ps) * f_pair / (asm_ms * 1e-3) / 1e12   # whole job (all ranks assemble concurrently)
        mv_bytes = 8.0 * (n + 6) * (n + 6)                          # whole job: every rank streams its row block
        mv_gbs_job = mv_bytes / (mv_ms * 1e-3) / 1e9
        mv_gbs_gpu = mv_gbs_job / world
        solve_ms = st["solve_ms"] / args.steps
        traffic_mv, traffic_asm = load_traffic(wl, world)
        roof_mv = {"kernel": "k_gemv<2>", "bound": "hbm", "achieved": mv_gbs_gpu, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": mv_gbs_gpu / peaks["hbm_gbs"], "traffic": traffic_mv, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                   "algorithmic_bytes_per_launch": 8.0 * rows_loc * (n + 6), "launch_ms": mv_ms}
        roof_asm = {"kernel": "k_assemble_regular", "bound": "fp64", "achieved": asm_tflops / world, "peak": fp64.value,
                    "unit": "TFLOP/s", "frac": asm_tflops / world / fp64.value, "traffic": traffic_asm,
                    "traffic_note": "DRAM bytes of one assembly (all colour launches); see profiles/r01_traffic.json",
                    "peak_source": "measured in this run: 8-chain DFMA microbenchmark sustained for 1 s under the power cap "
                                   "(bs_bench_fp64_sustained); burst figure in fp64_peak_tflops_burst",
                    "algorithmic_flops_per_pair": f_pair, "pairs_regular": pr, "pairs_singular": ps, "launch_ms": asm_ms}
        dominant_is_asm = asm_ms >= solve_ms
        line = {
            "metric": "assembly Gentries/s + GMRES matvec HBM GB/s; time-to-solution at 3N DoF",
            "value": entries / (asm_ms * 1e-3) / 1e9, "unit": "Gentries/s",
            "matvec_hbm_gbs": mv_gbs_job, "matvec_hbm_gbs_per_gpu": mv_gbs_gpu,
            "time_to_solution_s": wall / args.steps, "gmres_iterations": its,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "nodes": N, "cells": ncell, "dofs": n, "matrix_gb_each": 8.0 * n * n / 1e9,
                       "sharding": "rows (collocation nodes) over %d GPU(s), contiguous ranges of the locality order" % world,
                       "exchange": ("none (1 GPU)" if world == 1 else ("NVLink peer stores fused into the Krylov-vector kernel + "
                                    "NCCL allreduce of the dots" if p.use_peer_exchange else "NCCL allgather + allreduce")),
                       "collective_calls_in_timed_region": ({"allgather": comm.n_allgather, "allreduce": comm.n_allreduce} if comm else None),
                       "timing": "CUDA events on the launching stream inside the library; working set (%.1f GB matrix per GPU) "
                                 "far larger than the 126 MB L2, no explicit flush needed" % (8.0 * n * n / 1e9 / world),
                       "value_definition": "2*(3N)^2 entries of V and K evaluated / (K0+K1+K2 device time); ms_per_step is the "
                                           "whole step (assembly + corrections + monolithic build + GMRES)"
                                           + ("; fused mode stores V only and consumes the K tile in the epilogue" if wl.get("fused") else "")},
            "phases_ms": {"assembly": asm_ms, "assemble_regular": st["assemble_regular_ms"] / args.steps,
                          "assemble_singular": st["assemble_singular_ms"] / args.steps, "cell_geometry": st["geometry_ms"] / args.steps,
                          "corrections": st["correct_ms"] / args.steps, "monolithic": st["monolithic_ms"] / args.steps,
                          "gmres": solve_ms},
            "tiling": {"cell_blocks": int(st["n_cell_blocks"]), "colours": int(st["n_colours"]),
                       "node_touch_ratio": st["node_touch_ratio"]},
            "resistance_6rhs": res6,
            "drag_over_6pi": (drag / (6 * math.pi)) if drag is not None else None,
            "clocks": clocks,
            "e2e": {"value": entries / e2e_asm_s / 1e9, "unit": "Gentries/s", "time_to_solution_s": e2e_tts_s,
                    "prepass_s": min(e2e_pre), "prepass_cg_iterations": int(p._pre.cg_iterations),
                    "note": "time_to_solution_s is host to host per frame: geometry H2D, device pre-pass (mass matrix, L2 "
                            "normals, rigid modes), assembly, corrections, monolithic build, GMRES, solution D2H; value "
                            "excludes the pre-pass",
                    "h2d_bytes_per_step": int(8 * 3 * N + 4 * 2 * ncell * na + 8 * 2 * n + 8 * 12 * n + 8 * n),
                    "d2h_bytes_per_step": int(8 * n + 8 * (n + 6))},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": roof_asm if dominant_is_asm else roof_mv,
            "roofline_secondary": roof_mv if dominant_is_asm else roof_asm,
            "fp64_peak_tflops_measured": fp64.value, "fp64_peak_tflops_burst": fp64_burst.value,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, sample_seconds=args.cpu_seconds)
        print(json.dumps(line), flush=True)
    p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline(wl, sample_seconds=15.0, nthreads=0):
    """The reference's CPU path (oracle/bem_port.c) on a bounded row sample of the same workload."""
    from oracle import bem_oracle as bo, port
    nodes, conn = bo.cubesphere(degree=wl["degree"], m=wl["m"])   # the oracle's own mesh generator: no product code here
    geo = bo.Geometry(nodes, conn, wl["degree"])
    cores = port.max_threads() if nthreads == 0 else nthreads
    N = geo.N
    kspec = bo.KernelSpec(bo.FREE_SURFACE, 0.0, 1, (0.0, 1.4, 0.0)) if wl.get("kernel") == "free_surface" else bo.KernelSpec()
    # calibrate on one row per thread, then size the sample
    t0 = time.perf_counter()
    _, _, pairs = port.assemble_VK(geo, kspec, wl["quad"], "Mixed", wl["sing"], 0, cores, nthreads)
    t_cal = time.perf_counter() - t0
    rows = int(max(cores, min(N, cores * max(1, int(sample_seconds / max(t_cal, 1e-3))))))
    t0 = time.perf_counter()
    V, K, pairs = port.assemble_VK(geo, kspec, wl["quad"], "Mixed", wl["sing"], 0, rows, nthreads)
    t = time.perf_counter() - t0
    entries = 2.0 * 3 * rows * 3 * N
    # matvec sample: the assembled row block itself
    x = np.random.default_rng(0).uniform(-1, 1, 3 * N)
    port.gemv(V, x, nthreads)
    t1 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        port.gemv(V, x, nthreads)
    t_mv = (time.perf_counter() - t1) / reps
    return {"value": entries / t / 1e9, "unit": "Gentries/s", "cores": cores, "kind": "port",
            "sample": "%d of %d collocation nodes (all cells, all quadrature points), %.1f s; matvec on the %.2f GB row block"
                      % (rows, N, t, V.nbytes / 1e9),
            "pairs_per_s": pairs / t, "matvec_gbs": V.nbytes / t_mv / 1e9,
            "note": "C/OpenMP restatement of bem_stokes.cc:2871-2998; flatters the reference (no Epetra per-entry "
                    "insertion, threads over rows; the reference assembles single-threaded per MPI rank)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = workload(args, world)
    vals = []
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    cb = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cb = cpu_baseline(wl, sample_seconds=per_step)
        if i >= args.warmup:
            vals.append((cb["value"], time.perf_counter() - t0))
    v = float(np.mean([a for a, _ in vals]))
    line = {"impl": "reference", "metric": "assembly Gentries/s + GMRES matvec HBM GB/s; time-to-solution at 3N DoF",
            "value": v, "unit": "Gentries/s", "matvec_hbm_gbs": cb["matvec_gbs"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([b for _, b in vals])), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl["name"]},
            "cpu_baseline": dict(cb, value=v),
            "e2e": {"value": v, "unit": "Gentries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "vk", "q2", "c5"])
    ap.add_argument("--no-peer-exchange", action="store_true", help="multi-GPU: NCCL allgather callbacks instead of NVLink peer stores")
    ap.add_argument("--no-fused", action="store_true", help="c4 family with V and K both stored (needs 2x the memory)")
    ap.add_argument("--subdiv", dest="m", type=int, default=0, help="cube-sphere subdivisions per face edge (default 128*(N/8)^(1/4); 64*N^(1/4) for --workload vk)")
    ap.add_argument("--refine", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resistance", action="store_true", help="also time the 6-RHS batched resistance solve")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
