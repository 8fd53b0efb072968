"""ctypes binding of the C ABI declared in include/probit_b200.h.

The library is the product: if `libprobit_b200.so` is missing or a call fails, this module raises —
there is no CPU fallback (BASELINE.json north_star).  Build it with `python __graft_entry__.py`
(or probit_b200/csrc/build.sh).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprobit_b200.so")

PB_OK, PB_ERR_INVALID, PB_ERR_CUDA, PB_ERR_UNSUPPORTED, PB_ERR_NUMERIC = 0, -1, -2, -3, -4
PB_BASE_EQ, PB_BASE_EXP = 0, 1
PB_LIK_ORDINAL_PROBIT, PB_LIK_GAUSSIAN, PB_LIK_ORDINAL_PROBIT_SAFE = 0, 1, 2


class KernelSpec(C.Structure):
    _fields_ = [("base", C.c_int32), ("periodic", C.c_int32), ("scale", C.c_double),
                ("stretch_in", C.c_double), ("period", C.c_double), ("stretch_out", C.c_double),
                ("distance_form", C.c_int32), ("_pad", C.c_int32)]


class Options(C.Structure):
    """pb_options: the tunables of the drivers, passed explicitly with every call (no global configuration)."""
    _fields_ = [("laplace_pcg_min_n", C.c_int64), ("laplace_nystrom_rank", C.c_int64), ("laplace_cg_tol", C.c_double),
                ("negative_curvature_tol", C.c_double), ("potrf_block", C.c_int32), ("potrf_lookahead", C.c_int32),
                ("potrf_graph", C.c_int32), ("dist_block", C.c_int32), ("potrf_ozaki", C.c_int32), ("ozaki_tile", C.c_int32)]


class LikelihoodSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("J", C.c_int32), ("sigma", C.c_double), ("eps", C.c_double),
                ("cutpoints", C.c_void_p), ("safe_single_precision", C.c_int32), ("_pad", C.c_int32)]


class FitResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("info", C.c_int32), ("error", C.c_double),
                ("sum_ll", C.c_double), ("ftw", C.c_double), ("logdet", C.c_double),
                ("factorizations", C.c_int32), ("pcg_iterations", C.c_int32)]


class Problem(C.Structure):
    _fields_ = [("X", C.c_void_p), ("y", C.c_void_p), ("n", C.c_int64), ("D", C.c_int32), ("_pad", C.c_int32),
                ("kernel", KernelSpec), ("lik", LikelihoodSpec)]


class ProbitB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"probit_b200 error {code}: {message}")
        self.code = code


class NumericError(ProbitB200Error):
    """Non-SPD matrix / NaN detected on the device (PB_ERR_NUMERIC)."""


_p, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_SPEC, _LIK, _PROB, _OPT = C.POINTER(KernelSpec), C.POINTER(LikelihoodSpec), C.POINTER(Problem), C.POINTER(Options)

# name -> (restype, argtypes); mirrors include/probit_b200.h one to one
SIGNATURES = {
    "pb_version": (_i32, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_options_default": (_i32, [_OPT]),
    "pb_launch_count": (C.c_longlong, []),
    "pb_profile_begin": (_i32, []),
    "pb_profile_end": (_i32, [C.POINTER(C.c_longlong), C.POINTER(_f64), C.POINTER(_f64)]),
    "pb_profile_int8": (_i32, [C.POINTER(C.c_longlong), C.POINTER(_f64), C.POINTER(_f64)]),
    "pb_measure_fp64_tensor_peak": (_i32, [C.POINTER(_f64)]),
    "pb_likelihood": (_i32, [_p, _LIK, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "pb_predictive_distributions": (_i32, [_p, _LIK, _p, _p, _i64, _p]),
    "pb_feature_dim": (_i32, [_SPEC, _i32]),
    "pb_features": (_i32, [_p, _SPEC, _p, _i64, _i32, _i64, _p, _i64]),
    "pb_gram_sym": (_i32, [_p, _SPEC, _p, _i64, _i32, _i64, _p, _i64, _p, _f64]),
    "pb_gram_cross": (_i32, [_p, _SPEC, _p, _i64, _p, _i64, _i32, _i64, _i64, _p, _i64]),
    "pb_scale_sym_plus_identity": (_i32, [_p, _p, _i64, _i64, _p, _f64, _p, _i64]),
    "pb_copy_lower_add_diag": (_i32, [_p, _p, _i64, _i64, _f64, _p, _i64]),
    "pb_potrf_workspace_bytes": (_i64, [_i64]),
    "pb_potrf": (_i32, [_p, _p, _i64, _i64, _p, _i64, _p, _OPT]),
    "pb_rebuild_solve_workspace": (_i32, [_p, _p, _i64, _i64, _p, _i64]),
    "pb_transform_block": (_i32, [_p, _p, _i64, _p, _f64, _f64, _i64, _i64, _i64, _i64, _p, _i64]),
    "pb_gemm_nt": (_i32, [_p, _i64, _i64, _i64, _f64, _p, _i64, _p, _i64, _f64, _p, _i64, _i32]),
    "pb_ozaki_scratch_bytes": (_i64, [_i64, _i64, _i64]),
    "pb_ozaki_gemm_nt": (_i32, [_p, _i64, _i64, _i64, _f64, _p, _i64, _p, _i64, _p, _i64, _i32, _p, _i64]),
    "pb_symv": (_i32, [_p, _p, _i64, _i64, _p, _p]),
    "pb_symv_lower_scratch_bytes": (_i64, [_i64]),
    "pb_gemv": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "pb_symv_lower": (_i32, [_p, _p, _i64, _i64, _p, _p, _p, _i64]),
    "pb_trsv": (_i32, [_p, _p, _i64, _i64, _p, _i32, _p, _p]),
    "pb_logdet_chol": (_i32, [_p, _p, _i64, _i64, _p]),
    "pb_trsm_right_lt": (_i32, [_p, _p, _i64, _i64, _p, _p, _i64, _i64]),
    "pb_fit_workspace_bytes": (_i64, [_i64, _i32]),
    "pb_build_gram": (_i32, [_p, _PROB, _p, _i64]),
    "pb_build_features": (_i32, [_p, _PROB, _p, _i64]),
    "pb_workspace_gram": (_i32, [_p, _i64, _i32, C.POINTER(_p), C.POINTER(_i64)]),
    "pb_laplace_fit": (_i32, [_p, _PROB, _f64, _i32, _f64, _i32, _p, _i64, _p, _p, _p, C.POINTER(FitResult), _OPT]),
    "pb_vb_fit": (_i32, [_p, _PROB, _f64, _i32, _p, _i64, _p, _p, _p, C.POINTER(FitResult), _OPT]),
    "pb_gradient_scratch_bytes": (_i64, [_i64]),
    "pb_laplace_gradient": (_i32, [_p, _PROB, _p, _i64, _p, _p, _p, _i64, C.POINTER(_f64), _i32, _OPT]),
    "pb_vb_gradient": (_i32, [_p, _PROB, _p, _i64, _p, _p, _i64, C.POINTER(_f64), _i32, _OPT]),
    "pb_predict_prepare": (_i32, [_p, _PROB, _p, _i32, _p, _i64, C.POINTER(_i32), _OPT]),
    "pb_predict_scratch_bytes": (_i64, [_i64, _i32, _i64]),
    "pb_predict_covariance": (_i32, [_p, _PROB, _p, _p, _i64, _p, _i64, _p, _i64]),
    "pb_predict": (_i32, [_p, _PROB, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p]),
    "pb_comm_unique_id": (_i32, [_p]),
    "pb_comm_create": (_i32, [_p, _i32, _i32, C.POINTER(_p)]),
    "pb_comm_destroy": (_i32, [_p]),
    "pb_comm_rank": (_i32, [_p]),
    "pb_comm_size": (_i32, [_p]),
    "pb_dist_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _OPT]),
    "pb_dist_laplace_fit": (_i32, [_p, _p, _PROB, _f64, _i32, _p, _i64, _p, _p, _p, C.POINTER(FitResult), _OPT]),
    "pb_dist_predict_scratch_bytes": (_i64, [_i64, _i32, _i64]),
    "pb_dist_predict": (_i32, [_p, _p, _PROB, _p, _p, _p, _i64, _f64, _i32, _p, _i64, _i64, _p, _i64, _p, _p,
                               C.POINTER(_f64), C.POINTER(_i32), _OPT]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the signatures.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built: run `python __graft_entry__.py` (nvcc, sm_100a). "
            "probit_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def default_options(**overrides):
    """A pb_options filled with the library defaults, then `overrides` (field name -> value)."""
    o = Options()
    check(load().pb_options_default(C.byref(o)))
    for name, value in overrides.items():
        if name not in dict((f[0], 1) for f in Options._fields_):
            raise KeyError(f"unknown option {name!r}")
        setattr(o, name, type(getattr(o, name))(value))
    return o


def check(status):
    if status == PB_OK:
        return
    msg = load().pb_last_error().decode(errors="replace")
    raise (NumericError if status == PB_ERR_NUMERIC else ProbitB200Error)(status, msg)
