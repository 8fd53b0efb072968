"""Thin torch-tensor wrappers over the C ABI (device memory and streams come from torch).

Every function takes/returns float64 CUDA tensors and enqueues on torch's current stream.
Nothing here computes on the CPU; a missing library or a non-CUDA tensor raises.
"""
import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x, dtype=torch.float64):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.dtype != dtype:
        x = x.to(dtype)
    if not x.is_cuda:
        x = x.cuda()
    return x.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _mat2(x):
    x = _dev(x)
    if x.ndim == 0:
        x = x.reshape(1, 1)
    elif x.ndim == 1:
        x = x[:, None].contiguous()      # mlkernels up-ranks 1-D inputs to (N, 1)
    return x


def padded_ld(n):
    """Leading dimension used for internal N x N matrices (rows 128-byte aligned for TMA)."""
    return (max(int(n), 1) + 15) // 16 * 16


def empty_matrix(n_rows, n_cols):
    """(n_rows, n_cols) view into a buffer whose leading dimension is padded_ld(n_cols)."""
    ld = padded_ld(n_cols)
    buf = torch.empty((n_rows, ld), dtype=torch.float64, device="cuda")
    return buf[:, :n_cols]


def _ld(t):
    assert t.stride(1) == 1, "matrix must be row-major"
    return t.stride(0)


def features(spec, X):
    lib = _lib.load()
    X = _mat2(X)
    n, D = X.shape
    Df = lib.pb_feature_dim(C.byref(spec), D)
    Z = torch.empty((Df, n), dtype=torch.float64, device="cuda")      # feature-major: Z[d, i]
    _lib.check(lib.pb_features(_stream(), C.byref(spec), _ptr(X), n, D, D, _ptr(Z), n))
    return Z


def gram(spec, X, Y=None, diag_add=0.0, diag_vec=None):
    """K(X, X) (+ diag_add I + diag(diag_vec)) or K(X, Y); returns an (n, m) view with padded ld."""
    lib = _lib.load()
    Zx = features(spec, X)
    Df, n = Zx.shape
    if Y is None:
        K = empty_matrix(n, n)
        dv = _dev(diag_vec) if diag_vec is not None else None
        _lib.check(lib.pb_gram_sym(_stream(), C.byref(spec), _ptr(Zx), n, Df, n, _ptr(K), _ld(K), _ptr(dv),
                                   float(diag_add)))
        return K
    Zy = features(spec, Y)
    m = Zy.shape[1]
    K = empty_matrix(n, m)
    _lib.check(lib.pb_gram_cross(_stream(), C.byref(spec), _ptr(Zx), n, _ptr(Zy), m, Df, n, m, _ptr(K), _ld(K)))
    return K


def gram_elwise(spec, X, Y=None):
    """k(x_i, y_i) as an (n, 1) tensor (mlkernels' `elwise`)."""
    X = _mat2(X)
    if Y is None:
        return torch.full((X.shape[0], 1), float(spec.scale), dtype=torch.float64, device="cuda")
    Y = _mat2(Y)
    out = torch.empty((X.shape[0], 1), dtype=torch.float64, device="cuda")
    for i in range(X.shape[0]):     # rarely used: one 1x1 cross Gram per pair would be wasteful for large n
        out[i, 0] = gram(spec, X[i:i + 1], Y[i:i + 1])[0, 0]
    return out


def gemm_nt(A, B, C_out=None, alpha=1.0, beta=0.0, lower_only=False):
    """C = alpha * A @ B.T + beta * C on the FP64 tensor cores."""
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    if C_out is None:
        C_out = empty_matrix(M, N)
    _lib.check(lib.pb_gemm_nt(_stream(), M, N, K, float(alpha), _ptr(A), _ld(A), _ptr(B), _ld(B), float(beta),
                              _ptr(C_out), _ld(C_out), int(lower_only)))
    return C_out


def ozaki_gemm_nt(A, B, C_out, alpha=1.0, lower_only=False):
    """C += alpha * A @ B.T on the INT8 tensor cores (error-free slicing, csrc/ozaki.cu); K must be a multiple of 64."""
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    nbytes = lib.pb_ozaki_scratch_bytes(M, N, K)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(lib.pb_ozaki_gemm_nt(_stream(), M, N, K, float(alpha), _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(C_out),
                                    _ld(C_out), int(lower_only), _ptr(scratch), nbytes))
    return C_out


class Factor:
    """In-place lower Cholesky factor plus the leaf-inverse workspace the solves need."""

    def __init__(self, L, workspace, info):
        self.L, self.workspace, self.info = L, workspace, info


def potrf_(A, check=True, options=None):
    """In-place lower Cholesky of the row-major lower triangle of A. Returns a Factor."""
    lib = _lib.load()
    n = A.shape[0]
    ws_bytes = lib.pb_potrf_workspace_bytes(n)
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.pb_potrf(_stream(), _ptr(A), n, _ld(A), _ptr(ws), ws_bytes, _ptr(info),
                            C.byref(options) if options is not None else None))
    if check:
        i = int(info.item())
        if i != 0:
            raise _lib.NumericError(_lib.PB_ERR_NUMERIC, f"potrf: leading minor of order {i} is not positive definite")
    return Factor(A, ws, info)


def trsv(factor, b, trans=False):
    lib = _lib.load()
    n = factor.L.shape[0]
    rhs = _dev(b).clone()
    x = torch.empty_like(rhs)
    _lib.check(lib.pb_trsv(_stream(), _ptr(factor.L), n, _ld(factor.L), _ptr(factor.workspace), int(trans),
                           _ptr(rhs), _ptr(x)))
    return x


def cholesky_solve(factor, b):
    return trsv(factor, trsv(factor, b, False), True)


def trsm_right_lt_(factor, X):
    """X <- X L^{-T} in place (X is (m, n) row-major)."""
    lib = _lib.load()
    n = factor.L.shape[0]
    _lib.check(lib.pb_trsm_right_lt(_stream(), _ptr(factor.L), n, _ld(factor.L), _ptr(factor.workspace), _ptr(X),
                                    X.shape[0], _ld(X)))
    return X


def logdet_chol(factor):
    lib = _lib.load()
    out = torch.empty(1, dtype=torch.float64, device="cuda")
    n = factor.L.shape[0]
    _lib.check(lib.pb_logdet_chol(_stream(), _ptr(factor.L), n, _ld(factor.L), _ptr(out)))
    return out


def symv(K, x):
    lib = _lib.load()
    n = K.shape[0]
    x = _dev(x)
    y = torch.empty_like(x)
    _lib.check(lib.pb_symv(_stream(), _ptr(K), n, _ld(K), _ptr(x), _ptr(y)))
    return y


def symv_lower(K, x):
    """K @ x for a symmetric K stored in full, reading only the lower triangle's tiles (half the traffic)."""
    lib = _lib.load()
    n = K.shape[0]
    if _ld(K) < (n + 63) // 64 * 64:                      # the kernel reads whole 64-column tiles of the padded rows
        Kp = torch.zeros((n, (n + 63) // 64 * 64), dtype=torch.float64, device="cuda")
        Kp[:, :n] = K
        K = Kp[:, :n]
    x = _dev(x)
    y = torch.empty_like(x)
    nbytes = lib.pb_symv_lower_scratch_bytes(n)
    scratch = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    _lib.check(lib.pb_symv_lower(_stream(), _ptr(K), n, _ld(K), _ptr(x), _ptr(y), _ptr(scratch), nbytes))
    return y


def trmv_lower(A, z):
    """L @ z for the lower triangle of A (used only by the synthetic-data generator)."""
    return torch.tril(A) @ z
