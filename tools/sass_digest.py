#!/usr/bin/env python
"""Opcode histogram of the hot kernels of libbemstokes_b200.so (cuobjdump -sass): the evidence that the bulk-async
copies (UBLKCP), mbarriers (SYNCS), FP64 FMA / tensor instructions (DFMA, DMMA), the hardware rsqrt seed (MUFU.RSQ64H),
fire-and-forget reductions (RED.E.ADD.F64) and async copies (LDGSTS) are in the machine code.  Writes profiles/r02_sass_digest.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bemstokes_b200", "libbemstokes_b200.so")
HOT = [
    ("K1 free space, Q1, Gauss 8, cell-split, 2-D moments, fused (default bench)", r"k_assemble_regularILi4ELi0ELi0ELi1ELi1ELb0ELb1ELin8ELi2E"),
    ("K1 free space, Q1, Gauss 8, cell-split, 2-D moments, V and K stored", r"k_assemble_regularILi4ELi0ELi0ELi1ELi1ELb0ELb0ELin8ELi2E"),
    ("K1 free space, Q1, Gauss 8, thread pairs, linear rows, fused (BS_NO_CELLSPLIT=1: the kernel before the cell-split mode)", r"k_assemble_regularILi4ELi0ELi0ELi2ELi1ELb0ELb1ELin8ELi1E"),
    ("K1 free space, Q2 (V warps / K warps)", r"k_assemble_regularILi9ELi0ELi0ELi1ELi2ELb0ELb0ELi8E"),
    ("K1 free surface, Q1, cell-split, 2-D moments, single layer launch", r"k_assemble_regularILi4ELi1ELi1ELi1ELi1ELb0ELb0ELin8ELi2E"),
    ("K1 free surface, Q1, cell-split, 2-D moments, double layer launch", r"k_assemble_regularILi4ELi1ELi2ELi1ELi1ELb0ELb0ELin8ELi2E"),
    ("K1 no-slip, Q1, cell-split, coefficient x tensor sums, single layer launch", r"k_assemble_regularILi4ELi2ELi1ELi1ELi1ELb0ELb0ELi0ELi2E"),
    ("K1 no-slip, Q1, cell-split, coefficient x tensor sums, double layer launch", r"k_assemble_regularILi4ELi2ELi2ELi1ELi1ELb0ELb0ELi0ELi2E"),
    ("K5 k_gemv<2>", r"k_gemvILi2E"),
    ("K6 k_gemm_dmma<2,4> (multi-RHS sweep)", r"k_gemm_dmmaILi2ELi4E"),
    ("K8 k_lu_gemm_dmma (LU trailing update)", r"k_lu_gemm_dmma"),
    ("K8 k_lu_apply_coop (block-triangular application)", r"k_lu_apply_coop"),
    ("K7 k_gm_pass<UPD_DOTS> (Gram-Schmidt pass + cross-rank sum)", r"k_gm_passILi1E"),
    ("K7 k_gm_pass<UPD_NORM> (+ Givens, convergence test)", r"k_gm_passILi2E"),
    ("K7 k_gm_publish<false> (normalise + peer stores + flags)", r"k_gm_publishILb0E"),
]
SHOW = ["UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "DMMA", "MUFU.RSQ64H", "REDG.E.ADD.F64", "ATOMG", "SHFL", "LDS", "STS",
        "LDG", "STG", "BAR", "MEMBAR", "CCTL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur][op] += 1
    out = ["# SASS digest of the hot kernels (round 2)", "",
           "`python tools/sass_digest.py` on `bemstokes_b200/libbemstokes_b200.so` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`); "
           "static instruction counts per kernel, opcodes matched by prefix.", "",
           "| kernel | total | " + " | ".join(SHOW) + " |", "|---|---|" + "---|" * len(SHOW)]
    for label, pat in HOT:
        hits = [f for f in funcs if re.search(pat, f)]
        if not hits:
            out.append("| %s | not found | " % label + " | " * len(SHOW))
            continue
        c = funcs[hits[0]]
        tot = sum(c.values())
        cells = []
        for s in SHOW:
            cells.append(str(sum(v for k, v in c.items() if k == s or k.startswith(s + ".") or (s.endswith(".F64") and k.startswith(s)))))
        out.append("| %s | %d | " % (label, tot) + " | ".join(cells) + " |")
    out += ["", "Reading: `UBLKCP` = `cp.async.bulk` (TMA engine) cell-record ring of K1 with `SYNCS` mbarrier arrive / try_wait; "
            "`MUFU.RSQ64H` = hardware seed of 1/r; `REDG.E.ADD.F64` = the later colours' fire-and-forget tile reductions; `DMMA` = FP64 tensor "
            "instructions (`mma.sync.m8n8k4.f64`) of the multi-RHS sweep and of the LU trailing update; `LDGSTS` = `cp.async` staging of the "
            "LU tiles; `MEMBAR` / system-scope loads and stores in the Gram-Schmidt kernels are the peer-memory reductions.", ""]
    path = os.path.join(ROOT, "profiles", "r02_sass_digest.md")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("\n".join(out))


if __name__ == "__main__":
    main()
