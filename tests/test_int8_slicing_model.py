"""CPU model of the error-free slicing behind the INT8-sliced FP64 contraction (probit_b200/csrc/ozaki.cu): the same
digits, levels and scalings in NumPy integer arithmetic, against exact rational arithmetic.  Pins the NUMERICAL CONTRACT
of the scheme (what the GPU tests then hold the kernel to): digits in [-64, 64], reconstruction to 49 bits below the row's
leading bit (2^-49 .. 2^-48 of its largest entry), int32 accumulators cannot overflow for K <= 65536, and the truncated level sum (levels p + q <= 8) is
within 2^-48 of the exact product relative to the row scales.  CPU only."""
from fractions import Fraction

import numpy as np

S = 7                      # OZ_S: digit planes per value


def slice_rows(P):
    """oz_slice_kernel: per row one exponent e with max|x| 2^-e in [1/4, 1/2); t = x 2^(7 - e) in (-64, 64);
    digits by repeated round-to-nearest, t <- (t - d) * 128 (exact in FP64)."""
    P = np.asarray(P, dtype=np.float64)
    mx = np.abs(P).max(axis=1)
    e = np.zeros(P.shape[0], dtype=np.int64)
    nz = mx > 0
    _, q = np.frexp(mx[nz])                    # mx = m 2^q, m in [1/2, 1)
    e[nz] = q + 1
    t = P * np.exp2(7.0 - e)[:, None]
    digits = np.empty((S,) + P.shape, dtype=np.int64)
    for s in range(S):
        d = np.rint(t)
        assert np.all(np.abs(d) <= 64)
        t = (t - d) * 128.0                   # exact: |t - d| <= 1/2
        digits[s] = d.astype(np.int64)
    return digits, e


def sliced_product(A, B):
    """oz_gemm_kernel: level t = p + q accumulates the exact integer products A_p B_q^T (p, q = 1 .. 7, t <= 8); the
    epilogue adds the levels least-significant first with the exact scales 2^(-7 t) and the row / column exponents."""
    da, ea = slice_rows(A)
    db, eb = slice_rows(B)
    acc = np.zeros((A.shape[0], B.shape[0]), dtype=np.float64)
    worst = 0
    for t in range(S + 1, 1, -1):
        lvl = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
        for p in range(1, t):
            q = t - p
            if p <= S and q <= S:
                lvl += da[p - 1] @ db[q - 1].T
        worst = max(worst, int(np.abs(lvl).max()))
        acc += lvl.astype(np.float64) * 2.0 ** (-7 * t)
    # x = 2^(e - 7) sum_p d_p 128^-(p - 1) = 2^e sum_p d_p 128^-p  ->  product scale 2^(ea + eb), levels carry 128^-(p + q)
    return acc * np.exp2(ea.astype(np.float64))[:, None] * np.exp2(eb.astype(np.float64))[None, :], worst


def test_digits_reconstruct_each_value_to_49_bits_below_the_rows_leading_bit():
    rng = np.random.default_rng(0)
    P = rng.standard_normal((40, 64)) * np.exp(3 * rng.standard_normal((40, 64)).clip(-1, 1)) * 10.0 ** rng.integers(-6, 7, size=(40, 1))
    P[7] = 0.0
    P[9, 3] = 0.0
    d, e = slice_rows(P)
    rec = sum(d[s].astype(np.float64) * 128.0 ** -(s + 1) for s in range(S)) * np.exp2(e.astype(np.float64))[:, None]
    err = np.abs(rec - P)
    # residual <= 1/2 ulp of the last digit = 2^(e - 50); the row maximum lies in [2^(e - 2), 2^(e - 1)): 2^-49 of the
    # row's leading bit position, i.e. between 2^-49 and 2^-48 of its largest entry (+ the rounding of this check's own sum)
    bound = 2.0 ** -50 * np.exp2(e.astype(np.float64))[:, None] * (1 + 1e-3) + 4 * np.spacing(np.abs(P))
    assert np.all(err <= bound)
    assert (err / np.maximum(np.abs(P).max(axis=1)[:, None], 1e-300)).max() <= 2.0 ** -48 * (1 + 1e-3)
    assert np.all(d[:, 7, :] == 0) and e[7] == 0          # an all-zero row slices to zeros


def test_level_sums_fit_int32_for_the_largest_supported_k():
    # worst case per product: 64 * 64 per term; level t holds at most 7 products: K * 7 * 4096 < 2^31  <=>  K <= 74898
    assert 65536 * 7 * 64 * 64 < 2 ** 31
    rng = np.random.default_rng(1)
    A = rng.choice([-1.0, 1.0], size=(3, 4096)) * 0.49          # digits near +-63
    _, worst = sliced_product(A, A)
    assert worst < 2 ** 31


def test_truncated_level_sum_is_within_2_to_minus_48_of_the_exact_product():
    rng = np.random.default_rng(2)
    K = 256
    A = rng.standard_normal((6, K)) * 10.0 ** rng.integers(-5, 6, size=(6, 1))
    B = rng.standard_normal((5, K)) * 10.0 ** rng.integers(-3, 4, size=(5, 1))
    got, _ = sliced_product(A, B)
    for i in range(A.shape[0]):
        for j in range(B.shape[0]):
            exact = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(A[i], B[j]))
            scale = np.abs(A[i]).max() * np.abs(B[j]).max() * K
            # per term: slicing 2^-49 on each factor + the dropped levels (p + q >= 9: below 128^-7 of the leading product)
            assert abs(Fraction(float(got[i, j])) - exact) <= Fraction(2.0 ** -47) * Fraction(float(scale))
    # and against the plain FP64 product it is rounding-level
    ref = A @ B.T
    sc = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :] * np.sqrt(K)
    assert (np.abs(got - ref) / sc).max() < 32 * 2.0 ** -49       # the bound tests/test_gpu_kernels.py holds the kernel to
