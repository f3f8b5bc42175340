"""ctypes binding of libbemstokes_b200.so (the C-ABI of include/bemstokes_b200.h).

The product has no CPU fallback: if the shared library is missing this module raises at import time, and
every entry point that computes returns BS_ERR_NO_DEVICE without a B200.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BEMSTOKES_B200_LIB") or os.path.join(HERE, "libbemstokes_b200.so")  # override: kernel-variant A/B runs

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libbemstokes_b200.so not found at %s - build it with `python bemstokes_b200/build.py` "
        "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)
ctx_p = C.c_void_p


class BsStats(C.Structure):
    _fields_ = [("assemble_regular_ms", C.c_double), ("assemble_singular_ms", C.c_double), ("geometry_ms", C.c_double),
                ("correct_ms", C.c_double), ("monolithic_ms", C.c_double), ("precond_setup_ms", C.c_double),
                ("solve_ms", C.c_double), ("vmult_ms_last", C.c_double), ("kernel_launches", C.c_longlong),
                ("pairs_regular", C.c_longlong), ("pairs_singular", C.c_longlong),
                ("n_cell_blocks", C.c_longlong), ("n_colours", C.c_longlong), ("node_touch_ratio", C.c_double),
                ("gmres_stream_ms_last", C.c_double), ("gmres_matvec_ms_last", C.c_double), ("gmres_sweeps_last", C.c_longlong),
                ("cell_sets", C.c_longlong), ("cell_steps", C.c_longlong), ("unpaired_cells", C.c_longlong),
                ("sync_steps", C.c_longlong)]


ALLGATHERV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, c_int_p, c_int_p, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)

# name -> (restype, argtypes); the non-gpu test checks that every symbol declared in the header is exported
SIGNATURES = {
    "bs_last_error": (C.c_char_p, []),
    "bs_version": (C.c_int, []),
    "bs_create": (C.c_int, [C.POINTER(ctx_p), C.c_int, C.c_int, C.c_int]),
    "bs_destroy": (C.c_int, [ctx_p]),
    "bs_set_pointer_mode": (C.c_int, [ctx_p, C.c_int]),
    "bs_set_stream": (C.c_int, [ctx_p, C.c_void_p]),
    "bs_set_partition": (C.c_int, [ctx_p, C.c_int, C.c_int, c_int_p, C.c_int]),
    "bs_get_owned_nodes": (C.c_int, [ctx_p, c_int_p, c_int_p]),
    "bs_set_geometry": (C.c_int, [ctx_p, C.c_int, c_double_p, C.c_int, c_int_p, C.c_int, c_int_p, c_int_p]),
    "bs_set_quadrature": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p]),
    "bs_set_singular_quadrature": (C.c_int, [ctx_p, C.c_int, C.c_int]),
    "bs_set_singular_rule": (C.c_int, [ctx_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "bs_set_kernel": (C.c_int, [ctx_p, C.c_int, C.c_double, C.c_int, c_double_p]),
    "bs_make_gauss_1d": (C.c_int, [C.c_int, c_double_p, c_double_p]),
    "bs_make_singular_rule": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
    "bs_prepass": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_double), c_double_p, c_double_p,
                             C.POINTER(C.c_double), c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int)]),
    "bs_host_prepass": (C.c_int, [C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, c_int_p, C.c_int, c_int_p, C.c_int, c_double_p,
                                  c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "bs_host_cell_blocks": (C.c_int, [C.c_int, c_double_p, C.c_int, c_int_p, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p,
                                      C.POINTER(C.c_uint), c_int_p, C.POINTER(C.c_ubyte), c_int_p, c_int_p]),
    "bs_set_constraints": (C.c_int, [ctx_p, C.c_int, c_int_p, c_int_p, c_int_p, c_double_p]),
    "bs_set_torque_mode": (C.c_int, [ctx_p, c_double_p, c_double_p, C.c_double]),
    "bs_assemble_VK": (C.c_int, [ctx_p]),
    "bs_set_column_flags": (C.c_int, [ctx_p, c_ubyte_p]),
    "bs_assemble_fused": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double, c_double_p]),
    "bs_correct_V": (C.c_int, [ctx_p, c_double_p, c_double_p, C.c_double, c_double_p]),
    "bs_correct_K": (C.c_int, [ctx_p, C.c_int]),
    "bs_build_monolithic": (C.c_int, [ctx_p, c_ubyte_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double,
                                      C.c_int, C.c_int, C.c_double, c_double_p, C.c_int, c_double_p]),
    "bs_matrix_size": (C.c_int, [ctx_p, C.c_int, c_int_p, c_int_p]),
    "bs_vmult": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bs_vmult_multi": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bs_get_entries": (C.c_int, [ctx_p, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p]),
    "bs_tangential_projector": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p]),
    "bs_precond_setup": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int]),
    "bs_precond_vmult": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p]),
    "bs_gmres": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, c_int_p, c_double_p]),
    "bs_set_gmres_orthogonalization": (C.c_int, [ctx_p, C.c_int]),
    "bs_gmres_multi": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, c_int_p,
                                 c_double_p]),
    "bs_dn_operator_multi": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, c_int_p]),
    "bs_direct_solve": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bs_evaluate_bie": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int]),
    "bs_kernel_eval": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                 c_double_p]),
    "bs_set_comm": (C.c_int, [ctx_p, ALLGATHERV_FN, ALLREDUCE_FN, C.c_void_p]),
    "bs_exchange_export": (C.c_int, [ctx_p, C.c_size_t, C.c_void_p]),
    "bs_exchange_import": (C.c_int, [ctx_p, C.c_int, C.c_void_p]),
    "bs_get_stats": (C.c_int, [ctx_p, C.POINTER(BsStats)]),
    "bs_reset_stats": (C.c_int, [ctx_p]),
    "bs_bench_vmult": (C.c_int, [ctx_p, C.c_int, C.c_int, c_double_p]),
    "bs_bench_vmult_multi": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "bs_bench_lu": (C.c_int, [ctx_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "bs_bench_fp64_peak": (C.c_int, [C.c_int, c_double_p]),
    "bs_bench_fp64_sustained": (C.c_int, [C.c_int, C.c_double, c_double_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

# enums of the header
KERNEL_FREE, KERNEL_FREE_SURFACE, KERNEL_NO_SLIP = 0, 1, 2
SING_MIXED, SING_DUFFY, SING_TELLES = 0, 1, 2
MAT_V, MAT_K, MAT_A = 0, 1, 2
PREC_NONE, PREC_JACOBI, PREC_DIRECT, PREC_BLOCK_DIRECT, PREC_BAND = 0, 1, 2, 3, 4
GRID_REAL, GRID_IMPOSED_FORCE, GRID_IMPOSED_VELOCITY = 0, 1, 2
PTR_HOST, PTR_DEVICE = 0, 1
ORTHO_CGS2, ORTHO_MGS = 0, 1
ERR_NOT_CONVERGED = -4
ERR_NO_DEVICE = -3


class BemStokesError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbemstokes_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise BemStokesError(rc, (lib.bs_last_error() or b"").decode())
    return rc
