"""ctypes binding of libcafe_b200.so (the C ABI in include/cafe_b200.h).

There is no fallback: if the CUDA library is missing or cannot be loaded this module raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcafe_b200.so")

# every symbol include/cafe_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "cafe_b200_create", "cafe_b200_destroy", "cafe_b200_last_error", "cafe_b200_set_prior",
    "cafe_b200_set_error_model", "cafe_b200_eval_base", "cafe_b200_eval_gamma", "cafe_b200_reconstruct",
    "cafe_b200_get_matrix", "cafe_b200_matrix_size", "cafe_b200_root_vectors", "cafe_b200_enqueue_eval",
    "cafe_b200_fetch_result", "cafe_b200_stream", "cafe_b200_last_stats", "cafe_b200_unique_families",
    "cafe_b200_measure_fp64_peak", "cafe_b200_describe", "cafe_b200_discrete_gamma", "cafe_b200_minimize", "cafe_b200_fit", "cafe_b200_simulate", "cafe_b200_pvalues", "cafe_b200_io_last_error", "cafe_b200_io_parse_tree", "cafe_b200_io_read_families",
    "cafe_b200_io_read_error_model", "cafe_b200_io_derive_sizes", "cafe_b200_io_format_results", "cafe_b200_io_format_family_likelihoods", "cafe_b200_io_format_reconstruction", "cafe_b200_branch_probabilities",
    "cafe_b200_create_multi", "cafe_b200_create_bucketed", "cafe_b200_plan_shards", "cafe_b200_n_devices", "cafe_b200_node_columns", "cafe_b200_result_device", "cafe_b200_io_make_prior", "cafe_b200_fit_poisson_prior",
    "cafe_b200_host_alloc", "cafe_b200_host_free", "cafe_b200_debug_read_probe", "cafe_b200_io_format_report", "cafe_b200_io_format_simulation", "cafe_b200_io_format_error_model",
]

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_fp = C.POINTER(C.c_float)
c_up = C.POINTER(C.c_uint8)


class FitOptions(C.Structure):
    _fields_ = [("n_cat", C.c_int32), ("optimize_epsilon", C.c_int32), ("fixed_alpha", C.c_double), ("fixed_lambdas", c_dp),
                ("start", c_dp), ("seed", C.c_uint32), ("max_iterations", C.c_int32)]


class FitResult(C.Structure):
    _fields_ = [("values", C.c_double * 16), ("n_values", C.c_int32), ("neg_lnl", C.c_double), ("iterations", C.c_int32),
                ("evaluations", C.c_int32), ("status", C.c_int32), ("seconds", C.c_double)]


OBJECTIVE = C.CFUNCTYPE(C.c_double, c_dp, C.c_void_p)


class CTree(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("parent", c_ip), ("branch_length", c_dp), ("leaf_col", c_ip),
                ("lambda_class", c_ip)]


_lib = None


def load():
    """Load libcafe_b200.so; raises RuntimeError when it is absent (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libcafe_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is the only implementation; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.cafe_b200_create.restype = C.c_int
    L.cafe_b200_create.argtypes = [C.POINTER(CTree), c_ip, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.cafe_b200_create_multi.restype = C.c_int
    L.cafe_b200_create_multi.argtypes = [C.POINTER(CTree), c_ip, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_ip, C.c_int32, C.POINTER(vp)]
    L.cafe_b200_node_columns.argtypes = [vp, C.POINTER(C.c_int64)]
    L.cafe_b200_create_bucketed.restype = C.c_int
    L.cafe_b200_create_bucketed.argtypes = [C.POINTER(CTree), c_ip, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_ip, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.cafe_b200_plan_shards.argtypes = [C.POINTER(CTree), c_ip, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.cafe_b200_n_devices.restype = C.c_int32
    L.cafe_b200_n_devices.argtypes = [vp]
    L.cafe_b200_destroy.argtypes = [vp]
    L.cafe_b200_last_error.restype = C.c_char_p
    L.cafe_b200_last_error.argtypes = [vp]
    L.cafe_b200_set_prior.argtypes = [vp, c_fp, C.c_int32]
    L.cafe_b200_set_error_model.argtypes = [vp, c_dp, C.c_int32, C.c_int32]
    L.cafe_b200_eval_base.argtypes = [vp, c_dp, C.c_int32, c_dp, c_dp]
    L.cafe_b200_eval_gamma.argtypes = [vp, c_dp, C.c_int32, C.c_double, c_dp, c_dp, C.c_int32,
                                       c_dp, c_dp, c_dp, c_dp, c_up, c_up, C.POINTER(C.c_int64)]
    L.cafe_b200_reconstruct.argtypes = [vp, c_dp, C.c_int32, c_dp, c_dp, C.c_int32, c_ip, c_ip, c_dp]
    L.cafe_b200_get_matrix.argtypes = [vp, C.c_double, C.c_double, c_dp]
    L.cafe_b200_matrix_size.restype = C.c_int32
    L.cafe_b200_matrix_size.argtypes = [vp]
    L.cafe_b200_root_vectors.argtypes = [vp, c_dp, C.c_int32, C.c_double, c_dp]
    L.cafe_b200_enqueue_eval.argtypes = [vp, c_dp, C.c_int32, C.c_double, c_dp, c_dp, C.c_int32]
    L.cafe_b200_fetch_result.argtypes = [vp, c_dp, C.POINTER(C.c_int64)]
    L.cafe_b200_result_device.restype = vp
    L.cafe_b200_result_device.argtypes = [vp]
    L.cafe_b200_stream.restype = vp
    L.cafe_b200_stream.argtypes = [vp]
    L.cafe_b200_last_stats.argtypes = [vp, c_ip, c_ip, c_fp, c_fp]
    L.cafe_b200_unique_families.restype = C.c_int64
    L.cafe_b200_unique_families.argtypes = [vp]
    L.cafe_b200_measure_fp64_peak.argtypes = [C.c_int32, C.c_int32, c_dp]
    L.cafe_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.cafe_b200_host_free.argtypes = [C.c_void_p]
    L.cafe_b200_debug_read_probe.argtypes = [vp, C.POINTER(C.c_int64), C.c_int64]
    L.cafe_b200_describe.argtypes = [vp, C.POINTER(C.c_int64), c_ip, c_ip, c_ip, c_ip, c_dp]
    L.cafe_b200_discrete_gamma.argtypes = [C.c_int32, C.c_double, c_dp, c_dp]
    L.cafe_b200_minimize.argtypes = [OBJECTIVE, C.c_void_p, C.c_int32, c_dp, C.c_int32, c_dp, c_dp, c_ip]
    L.cafe_b200_fit.argtypes = [vp, C.POINTER(FitOptions), C.POINTER(FitResult)]
    L.cafe_b200_pvalues.argtypes = [vp, c_dp, C.c_int32, C.c_int32, C.c_uint64, c_dp]
    L.cafe_b200_simulate.argtypes = [vp, c_dp, C.c_int32, c_dp, c_dp, C.c_int32, C.c_int32, C.c_int32, c_ip, C.c_int64, C.c_uint64,
                                     c_ip, c_ip, c_ip, C.POINTER(C.c_int64)]
    cs = C.c_char_p
    L.cafe_b200_io_last_error.restype = cs
    L.cafe_b200_io_parse_tree.argtypes = [cs, cs, C.c_int32, c_ip, c_ip, c_dp, c_ip, c_ip, c_ip, C.c_char_p, C.c_int64]
    L.cafe_b200_io_read_families.argtypes = [cs, cs, C.POINTER(C.c_int64), c_ip, c_ip, C.c_int64, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
    L.cafe_b200_io_read_error_model.argtypes = [cs, c_dp, C.c_int32, c_ip, c_ip]
    L.cafe_b200_io_derive_sizes.argtypes = [c_ip, C.c_int64, c_ip, c_ip]
    L.cafe_b200_io_format_results.argtypes = [cs, C.c_double, c_dp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double,
                                              C.c_char_p, C.c_int64]
    L.cafe_b200_io_format_family_likelihoods.argtypes = [cs, C.c_int64, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_up, C.c_int32, C.c_char_p, C.c_int64]
    L.cafe_b200_io_format_reconstruction.argtypes = [cs, cs, C.c_int64, c_ip, c_dp, C.c_double, c_dp, C.c_int32, c_dp, C.c_int32, C.c_char_p, C.c_int64]
    L.cafe_b200_io_format_report.argtypes = [cs, cs, c_dp, C.c_int32, cs, C.c_int64, c_ip, c_dp, c_dp, C.c_char_p, C.c_int64]
    L.cafe_b200_io_format_simulation.argtypes = [cs, C.c_int64, c_ip, c_dp, C.c_int32, C.c_char_p, C.c_int64]
    L.cafe_b200_io_format_error_model.argtypes = [c_dp, C.c_int32, C.c_char_p, C.c_int64]
    L.cafe_b200_branch_probabilities.argtypes = [vp, c_dp, C.c_int32, c_ip, c_up, c_dp]
    _lib = L
    return L


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def fp(a):
    return None if a is None else a.ctypes.data_as(c_fp)


def up(a):
    return None if a is None else a.ctypes.data_as(c_up)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
