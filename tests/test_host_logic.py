"""CPU-only tests of the host logic and of the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from probit_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "probit_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/probit_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pb_version() >= 100
    assert lib.pb_last_error() is not None


def test_struct_layouts_match_the_header():
    from probit_b200 import _lib
    assert ctypes.sizeof(_lib.KernelSpec) == 48
    assert ctypes.sizeof(_lib.LikelihoodSpec) == 40
    assert ctypes.sizeof(_lib.FitResult) == 48
    assert ctypes.sizeof(_lib.Options) == 56
    assert ctypes.sizeof(_lib.Problem) == 32 + 48 + 40
    assert _lib.Problem.kernel.offset == 32 and _lib.Problem.lik.offset == 80


def test_options_are_per_call_not_global():
    """pb_options replaces the process-global tunables: defaults come from the library, overrides live in the caller's
    struct, and there is no setter left to mutate shared state (SURVEY.md §8b re-entrancy contract)."""
    from probit_b200 import _lib
    lib = _lib.load()
    o = _lib.default_options()
    assert (o.laplace_pcg_min_n, o.laplace_nystrom_rank, o.potrf_block, o.potrf_lookahead) == (24576, -1, 0, 1)
    assert o.laplace_cg_tol == 1e-2 and o.negative_curvature_tol == 1e-6
    o2 = _lib.default_options(laplace_pcg_min_n=0, dist_block=256)
    assert o2.laplace_pcg_min_n == 0 and o2.dist_block == 256 and _lib.default_options().laplace_pcg_min_n == 24576
    assert not hasattr(lib, "pb_set_option") and not hasattr(lib, "pb_set_factor_callback")
    with pytest.raises(KeyError):
        _lib.default_options(no_such_option=1)
    # the panel width is part of the workspace layout, so the size query takes the same options
    a = lib.pb_dist_workspace_bytes(8192, 4, 2, 0, ctypes.byref(o2))
    b = lib.pb_dist_workspace_bytes(8192, 4, 2, 0, None)
    assert a > 0 and b > 0


def test_workspace_queries_do_not_need_a_gpu():
    from probit_b200 import _lib
    lib = _lib.load()
    n = 65536
    ws = lib.pb_fit_workspace_bytes(n, 4)
    assert 2 * n * n * 8 < ws < 2 * n * n * 8 + (3 << 29)          # K + factor (64 GiB of the 180 GB) + two int8 slicing buffers (0.9 GiB)
    inverses = ((n // 64) * 64 * 64 + 2 * (n // 256) * 256 * 256) * 8          # leaf inverses + 256-block inverses and transposes
    slicing = 2 * (7 * n * 1024 + 4 * n)                                        # int8 digit planes + row exponents of two panels (look-ahead)
    assert lib.pb_potrf_workspace_bytes(n) == inverses + slicing
    assert lib.pb_potrf_workspace_bytes(1024) == ((1024 // 64) * 64 * 64 + 2 * 4 * 256 * 256) * 8     # below 4096: no slicing scratch
    assert lib.pb_predict_scratch_bytes(n, 4, 4096) >= 4096 * n * 8
    assert lib.pb_launch_count() == 0


def test_kernel_spec_lowering():
    from probit_b200 import _lib, kernels as PK
    s = (2.0 * PK.EQ().stretch(0.7).periodic(0.5)).lower()          # examples/regression.py:123
    assert (s.base, s.periodic, s.scale, s.stretch_in, s.period, s.stretch_out) == (_lib.PB_BASE_EQ, 1, 2.0, 1.0, 0.5, 0.7)
    s = (1.5 * PK.Matern12().stretch(1.2)).lower()                  # examples/classification.py:375
    assert (s.base, s.periodic, s.scale, s.stretch_out) == (_lib.PB_BASE_EXP, 0, 1.5, 1.2)
    s = (PK.EQ().stretch(2.0).stretch(3.0) * 4.0).lower()
    assert (s.scale, s.stretch_out) == (4.0, 6.0)
    s = PK.EQ().periodic(2.0).stretch(3.0).lower()
    assert (s.periodic, s.stretch_in, s.stretch_out, s.period) == (1, 3.0, 1.0, 2.0)
    assert PK.Matern12 is PK.Exp
    with pytest.raises(NotImplementedError):
        PK.EQ() + PK.EQ()
    with pytest.raises(NotImplementedError):
        PK.EQ() * PK.EQ()
    with pytest.raises(NotImplementedError):
        PK.EQ().periodic(1.0).periodic(2.0).lower()


def test_lowered_specs_mean_what_the_oracle_kernels_compute():
    """The flat pb_kernel_spec of an expression tree, evaluated by its documented formula (include/probit_b200.h:
    k = scale * base(phi(x), phi(y)), phi(x) = T(x / stretch_in) / stretch_out), equals the oracle's (mlkernels-style,
    node by node) evaluation of the same tree — for every order of stretch / periodic / scale the reference API allows."""
    import numpy as np
    from oracle import kernels as OK
    from probit_b200 import _lib, kernels as PK
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 2, size=(17, 3))

    def from_spec(s, X):
        U = X / s.stretch_in
        if s.periodic:
            U = np.concatenate([np.sin(2 * np.pi * U / s.period), np.cos(2 * np.pi * U / s.period)], axis=1)
        U = U / s.stretch_out
        d2 = ((U[:, None, :] - U[None, :, :]) ** 2).sum(-1)
        return s.scale * (np.exp(-0.5 * d2) if s.base == _lib.PB_BASE_EQ else np.exp(-np.sqrt(d2)))

    trees = [lambda M: 1.7 * M.EQ().stretch(0.6),
             lambda M: M.Matern12().stretch(1.3) * 0.4,
             lambda M: 2.0 * M.EQ().stretch(0.7).periodic(0.5),
             lambda M: M.EQ().periodic(1.5).stretch(2.0),
             lambda M: 0.9 * M.EQ().stretch(0.35).periodic(0.5).stretch(1.1),
             lambda M: 3.0 * (0.5 * M.Matern12().stretch(0.8).stretch(1.5)).periodic(0.75)]
    for tree in trees:
        ref = tree(OK)(X)
        got = from_spec(tree(PK).lower(), X)
        assert np.abs(got - ref).max() < 1e-14 * max(1.0, np.abs(ref).max())


def test_likelihood_recognition_has_no_fallback():
    from probit_b200 import _lib, approximators as PA, utilities as PU
    k = PA._likelihood_kind
    assert k(PU.log_probit_likelihood, None, None) == _lib.PB_LIK_ORDINAL_PROBIT
    assert k(PU.log_gaussian_likelihood, None, None) == _lib.PB_LIK_GAUSSIAN
    assert k(PU.log_probit_likelihood, PU.grad_log_probit_likelihood, PU.hessian_log_probit_likelihood) == _lib.PB_LIK_ORDINAL_PROBIT_SAFE
    with pytest.raises(NotImplementedError):
        k(lambda f, y, p: 0.0, None, None)
    with pytest.raises(NotImplementedError):
        k(PU.log_probit_likelihood, lambda f, y, p: 0.0, None)


def test_check_cutpoints_follows_the_reference():
    from probit_b200.utilities import CutpointValueError, check_cutpoints
    inf = math.inf
    assert check_cutpoints([-0.5, 0.5], 3).tolist() == [-inf, -0.5, 0.5, inf]
    assert check_cutpoints([-inf, -0.5, 0.5], 3).tolist() == [-inf, -0.5, 0.5, inf]
    assert check_cutpoints([-0.5, 0.5, inf], 3).tolist() == [-inf, -0.5, 0.5, inf]
    assert check_cutpoints([-inf, -0.5, 0.5, inf], 3).tolist() == [-inf, -0.5, 0.5, inf]
    with pytest.raises(CutpointValueError):
        check_cutpoints([-inf, 0.5, -0.5, inf], 3)
    with pytest.raises(ValueError):
        check_cutpoints([0.0, -0.5, 0.5, inf], 3)
    with pytest.raises(ValueError):
        check_cutpoints([-inf, -0.5, 0.5, 1.0], 3)
    with pytest.raises(ValueError):
        check_cutpoints([0.1], 5)


def test_ordinal_generator_follows_the_recipe():
    from probit_b200.datasets import generate_ordinal_data
    X, g, y, cut = generate_ordinal_data(3, 103, 2, 5, 0.4, lambda X, z: 0.5 * z)
    assert X.shape == (103, 2) and y.dtype == np.int64 and cut.shape == (6,)
    assert cut[0] == -np.inf and cut[-1] == np.inf and np.all(np.diff(cut[1:-1]) > 0)
    counts = np.bincount(y, minlength=5)
    assert counts.sum() == 103 and counts.max() - counts.min() <= 1
    # every datum lies inside its own bin: cut[y] <= g <= cut[y+1]
    assert np.all(g >= cut[y]) and np.all(g <= cut[y + 1])
    X2, *_ = generate_ordinal_data(3, 103, 2, 5, 0.4, lambda X, z: 0.5 * z)
    assert np.array_equal(X, X2)


def test_shard_range_partitions_exactly():
    from probit_b200.distributed import shard_range
    for n in (0, 1, 7, 10_000_000):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in parts]
            assert max(sizes) - min(sizes) <= 1


def test_product_fails_loudly_without_a_gpu():
    """No CPU fallback: constructing an approximator needs device memory and must raise without CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from probit_b200 import approximators as PA, kernels as PK, utilities as PU
    with pytest.raises(Exception):
        PA.LaplaceGP((np.zeros((4, 1)), np.zeros(4, dtype=np.int64)), lambda l: PK.EQ().stretch(l), PU.log_probit_likelihood)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "probit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
