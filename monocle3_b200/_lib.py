"""ctypes binding of libm3b.so -- the C ABI declared in include/m3b.h.

There is deliberately no fallback: if the shared library is missing or no B200
is visible, the functions raise.  (The R package would bind the same symbols
through Rcpp; see INTEGRATION.md.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm3b.so")

M3B_OK = 0
M3B_E_DIMS = -1
M3B_W_NOT_CONVERGED = -2
M3B_E_NOMEM = -3
M3B_E_NULLSPACE = -4
M3B_E_LINDEP = -5
M3B_E_CUDA = -10
M3B_E_NCCL = -11
M3B_E_STATE = -12
M3B_E_NODEVICE = -13
M3B_E_ABORTED = -14

NORM_LOG2, NORM_SIZE_ONLY, NORM_NONE, NORM_LOG10, NORM_BINARY = 0, 1, 2, 3, 4
SF_TOTAL, SF_LOG_TOTAL = 0, 1
FLAG_TIME_SPMV = 1
FLAG_JACOBI_ONLY = 2
FLAG_NO_UNIT_SPLIT = 4  # keep count-1 entries in the 10 B/entry stream (A/B switch)
COMM_ID_BYTES = 128

RNORM_FN = C.CFUNCTYPE(C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_void_p)
SMALL_SVD_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
SHOULD_ABORT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("spmv_cells_out", C.c_int64), ("spmv_genes_out", C.c_int64),
                ("spmv_cells_out_ms", C.c_double), ("spmv_genes_out_ms", C.c_double), ("irlba_ms", C.c_double),
                ("prepare_ms", C.c_double), ("spmv_bytes_cells_out", C.c_int64), ("spmv_bytes_genes_out", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("phase_ms", C.c_double * 8),
                ("svd_sweeps", C.c_int64), ("host_ms", C.c_double * 4),
                ("svd_fallbacks", C.c_int64), ("k1_ms", C.c_double), ("k2_ms", C.c_double),
                ("svd_fallback_reasons", C.c_int64), ("spmv_stream_bytes_cells_out", C.c_int64),
                ("spmv_stream_bytes_genes_out", C.c_int64), ("select_ms", C.c_double * 4), ("k10_ms", C.c_double),
                ("xchg_diag", C.c_double * 8)]


# name -> (restype, argtypes); every symbol include/m3b.h declares
_P = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
SIGNATURES = {
    "m3b_abi_version": (C.c_int, []),
    "m3b_status_string": (C.c_char_p, [C.c_int]),
    "m3b_device_count": (C.c_int, []),
    "m3b_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "m3b_destroy": (None, [_P]),
    "m3b_last_error": (C.c_char_p, [_P]),
    "m3b_set_small_svd": (C.c_int, [_P, SMALL_SVD_FN, _P]),
    "m3b_set_rnorm": (C.c_int, [_P, RNORM_FN, _P]),
    "m3b_set_should_abort": (C.c_int, [_P, SHOULD_ABORT_FN, _P]),
    "m3b_comm_unique_id": (C.c_int, [_P, C.c_char_p]),
    "m3b_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p]),
    "m3b_comm_world": (C.c_int, [_P]),
    "m3b_comm_p2p_handle": (C.c_int, [_P, C.c_char_p]),
    "m3b_comm_p2p_open": (C.c_int, [_P, C.c_char_p]),
    "m3b_matrix_begin": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "m3b_matrix_append_csc": (C.c_int, [_P, C.c_int64, _ip, _ip, _dp]),
    "m3b_matrix_end": (C.c_int, [_P]),
    "m3b_matrix_free": (None, [_P]),
    "m3b_matrix_dims": (C.c_int, [_P, _ip, _lp, _lp, _lp]),
    "m3b_cell_totals": (C.c_int, [_P, C.c_int, _dp]),
    "m3b_size_factors_from_totals": (C.c_int, [_dp, C.c_int64, C.c_int, _dp, C.POINTER(C.c_int)]),
    "m3b_size_factors": (C.c_int, [_P, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
    "m3b_normalize": (C.c_int, [_P, _dp, C.c_int, C.c_double]),
    "m3b_export_values": (C.c_int, [_P, _dp]),
    "m3b_select_genes": (C.c_int, [_P, _ip, C.c_int32, C.POINTER(C.c_uint8), _ip]),
    "m3b_select_genes_for_transform": (C.c_int, [_P, _ip, C.c_int32, C.POINTER(C.c_uint8), _ip]),
    "m3b_gene_nnz": (C.c_int, [_P, _lp]),
    "m3b_gene_moments": (C.c_int, [_P, _dp, _dp]),
    "m3b_knn_self": (C.c_int, [_P, _dp, C.c_int64, C.c_int32, C.c_int32, _ip, _dp]),
    "m3b_jaccard_coeff": (C.c_int, [_P, _dp, C.c_int64, C.c_int32, C.c_int, _dp]),
    "m3b_detect_genes": (C.c_int, [_P, C.c_double, _lp, _lp]),
    "m3b_aggregate_expression": (C.c_int, [_P, _ip, C.c_int32, C.c_int, _ip, C.c_int32, C.c_int, _dp, C.c_int64]),
    "m3b_align_residuals": (C.c_int, [_P, _dp, C.c_int64, C.c_int, C.c_int64, _dp, C.c_int, _dp]),
    "m3b_align_apply_beta": (C.c_int, [_P, _dp, C.c_int64, C.c_int, C.c_int64, _dp, C.c_int, _dp]),
    "m3b_tfidf": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _dp, C.c_double, _dp, _dp]),
    "m3b_pca_irlba": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp, C.c_int, C.c_int,
                                _dp, _dp, _dp, C.c_int64, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "m3b_pca_irlba_device": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp, C.c_int, C.c_int,
                                       _dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "m3b_project": (C.c_int, [_P, _ip, C.c_int32, _dp, _dp, _dp, C.c_int, _dp, C.c_int64]),
    "m3b_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "m3b_measure_fp64_peak": (C.c_int, [_P, _dp]),
    "m3b_reset_stats": (C.c_int, [_P]),
    "m3b_set_flags": (C.c_int, [_P, C.c_int]),
    "m3b_timer_start": (C.c_int, [_P]),
    "m3b_timer_stop": (C.c_int, [_P, _dp]),
}

_lib = None


class M3BError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("m3b status %d: %s" % (status, message))
        self.status = status


def load():
    """Load libm3b.so and set the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libm3b.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C monocle3_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
