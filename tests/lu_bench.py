"""LU micro-benchmark (builder tool, run on the GPU box): python tests/lu_bench.py 18432 36864"""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bemstokes_b200 as bb
from bemstokes_b200._lib import lib, check

p = bb.BEMProblem()
p.set_mesh(bb.cubesphere(m=2))
p.reinit()
for n in [int(a) for a in sys.argv[1:]] or [4096]:
    f, a, r = C.c_double(), C.c_double(), C.c_double()
    check(lib.bs_bench_lu(p._ctx, n, 5, C.byref(f), C.byref(a), C.byref(r)))
    print("n=%d  factor %.1f ms = %.2f TFLOP/s   apply %.3f ms = %.0f GB/s   |Ay-b|_inf %.2e"
          % (n, f.value, 2.0 / 3.0 * n ** 3 / f.value / 1e9, a.value, 8.0 * n * n / a.value / 1e6, r.value), flush=True)
p.close()
