"""GPU tests of the factorisation schedules (factorize_device in csrc/gptb200.cu): the spine schedule (diagonal tile -> spine kernel ->
diagonal tile on a stream of its own, one flag-chained back-substitution launch) against the round-1 schedule (panel -> look-ahead
column -> diagonal tile, one back-substitution launch per block) and against a CPU Cholesky (scipy), over sizes that cover one tile,
several tiles, padding, the regime where the trailing kernel keeps eight SMs free for the spine kernel (T <= 36) and the one where it
keeps one (T = 40: the spine kernel's CTAs then start at different times -- the case that exposed an in-place read/write race between
them), and several handles factorising at once."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def problem(N, seed=0, d=3, p=3):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = 0.05 * np.sin(6 * X[:, :1] + np.arange(p)) + 0.01 * rng.standard_normal((N, p))
    return X, Y


def cpu_alpha(X, Y, c, ell, s2, jitter):
    import scipy.linalg as sla
    Z = X / ell
    sq = np.sum(Z * Z, axis=1)
    K = c * np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2.0 * Z @ Z.T, 0.0))
    K[np.diag_indices_from(K)] = c + s2 + jitter
    L = sla.cholesky(K, lower=True)
    alpha = sla.cho_solve((L, True), Y)
    lml = -0.5 * np.sum(Y * alpha) - Y.shape[1] * np.sum(np.log(np.diag(L))) - 0.5 * Y.shape[1] * len(X) * np.log(2 * np.pi)
    return alpha, lml


@pytest.mark.parametrize("N", [100, 128, 129, 834, 2048, 4608, 5120])
def test_schedules_agree_with_each_other_and_with_cpu_cholesky(N):
    from gaussian_process_transportation_b200 import _lib
    X, Y = problem(N)
    c, ell, s2, jitter = 0.1, np.array([0.1, 0.15, 0.2]), 1e-4, 1e-10
    ref_alpha, ref_lml = cpu_alpha(X, Y, c, ell, s2, jitter)
    eng = _lib.Engine(0)
    try:
        eng.set_train(X, Y)
        got = {}
        for spine, back in ((0, 0), (1, 1), (1, 0), (0, 1)):
            eng.set_debug_option("spine_variant", spine)
            eng.set_debug_option("back_substitution_variant", back)
            info, lml = eng.factorize(c, ell, s2, jitter, want_lml=True)
            assert info == 0
            got[(spine, back)] = (eng.export_alpha().copy(), lml, eng.export_L().copy())
            info, lml2, grad = eng.lml(c, ell, s2, jitter, want_grad=True)
            assert info == 0 and abs(lml2 - lml) <= 1e-12 * abs(lml)
        scale = np.max(np.abs(ref_alpha))
        for key, (alpha, lml, L) in got.items():
            # conditioning c / s2 = 1e3: a backward-stable factorisation reproduces alpha to ~1e-16 * 1e3 * growth
            assert np.max(np.abs(alpha - ref_alpha)) / scale < 1e-9, key
            assert abs(lml - ref_lml) < 1e-10 * abs(ref_lml), key
        base = got[(0, 0)]
        for key, (alpha, lml, L) in got.items():
            assert np.max(np.abs(alpha - base[0])) / scale < 1e-11, key
            assert np.max(np.abs(np.tril(L) - np.tril(base[2]))) < 1e-12 * np.max(np.abs(base[2])), key
        # the default schedule is deterministic
        eng.set_debug_option("spine_variant", 1)
        eng.set_debug_option("back_substitution_variant", 1)
        eng.factorize(c, ell, s2, jitter, want_lml=False)
        a1 = eng.export_alpha().copy()
        eng.factorize(c, ell, s2, jitter, want_lml=False)
        assert np.array_equal(a1, eng.export_alpha())
        assert np.array_equal(a1, got[(1, 1)][0])
    finally:
        eng.close()


def test_not_positive_definite_is_reported_by_both_schedules():
    from gaussian_process_transportation_b200 import _lib
    X, Y = problem(700)
    X[650] = X[10]                                   # a duplicated point and no noise: singular to working precision
    eng = _lib.Engine(0)
    try:
        eng.set_train(X, Y)
        for spine in (0, 1):
            eng.set_debug_option("spine_variant", spine)
            info, _ = eng.factorize(1.0, np.full(3, 0.5), 0.0, 0.0, want_lml=False)
            assert info > 0
        info, _ = eng.factorize(1.0, np.full(3, 0.5), 1e-4, 0.0, want_lml=False)
        assert info == 0
    finally:
        eng.close()


def test_handles_factorising_concurrently_agree_with_a_lone_one():
    """Several handles at once on one GPU: their spine kernels (eight CTAs that meet at a barrier) and flag-chained back substitutions
    share the SMs; every result must be the one a lone handle computes, bit for bit."""
    from gaussian_process_transportation_b200 import _lib
    N, c, ell, s2, jitter = 5120, 0.1, np.full(3, 0.1), 1e-4, 1e-10
    X, Y = problem(N, seed=3)
    lone = _lib.Engine(0)
    lone.set_train(X, Y)
    lone.factorize(c, ell, s2, jitter, want_lml=False)
    want = lone.export_alpha().copy()
    lone.close()
    engines = [_lib.Engine(0) for _ in range(6)]
    out, errs = [None] * len(engines), []

    def run(i):
        try:
            e = engines[i]
            e.set_train(X, Y)
            for _ in range(5):
                info, _ = e.factorize(c, ell, s2, jitter, want_lml=False)
                assert info == 0
            out[i] = e.export_alpha().copy()
        except Exception as exc:  # pragma: no cover
            errs.append(exc)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(engines))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for e in engines:
        e.close()
    assert not errs, errs
    for a in out:
        assert a is not None and np.array_equal(a, want)


@pytest.mark.parametrize("N", [300, 834, 4096])
def test_small_batch_variance_paths_agree_with_the_tile_path(N):
    """FP64 variance product: batches of at most eight right-hand-side rows take the matrix-vector kernels, batches too small to fill
    the machine the split-k kernels (csrc/query.cuh); both must reproduce what the one-CTA-per-tile kernel gives for the same points
    (switch "variance_splitk" off, and the same points inside a large batch)."""
    from gaussian_process_transportation_b200 import _lib as L
    X, Y = problem(N, seed=5)
    rng = np.random.default_rng(9)
    eng = L.Engine(0)
    try:
        eng.set_train(X, Y)
        assert eng.factorize(0.1, np.array([0.1, 0.15, 0.2]), 1e-4, 1e-10, want_lml=False)[0] == 0
        eng.prepare_variance()
        big = rng.random((3000, 3))
        flags_all = L.MEAN | L.STD | L.JAC | L.JACVAR | L.DVAR
        ref_big = eng.query(big, flags_all)
        for M, flags in ((1, L.MEAN | L.STD), (1, L.MEAN | L.STD | L.DVAR), (2, L.MEAN | L.STD | L.JAC | L.JACVAR), (8, L.MEAN | L.STD), (9, L.MEAN | L.STD),
                         (100, L.MEAN | L.STD | L.JAC), (100, flags_all), (400, L.MEAN | L.STD)):
            xq = big[:M]
            eng.set_debug_option("variance_splitk", 1)
            a = eng.query(xq, flags)
            eng.set_debug_option("variance_splitk", 0)
            b = eng.query(xq, flags)
            eng.set_debug_option("variance_splitk", 1)
            for key in a:
                scale = max(np.max(np.abs(b[key])), 1e-300)
                assert np.max(np.abs(a[key] - b[key])) <= 1e-11 * scale + 1e-13, (M, key)
                want = ref_big[key][:, :M] if key == "dvar" else ref_big[key][:M]
                assert np.max(np.abs(a[key] - want)) <= 1e-11 * scale + 1e-13, (M, key, "vs large batch")
            # deterministic
            a2 = eng.query(xq, flags)
            for key in a:
                assert np.array_equal(a[key], a2[key]), (M, key)
    finally:
        eng.close()


def test_tiny_batches_in_int8_mode_take_the_exact_matrix_vector_path():
    """A call with at most eight right-hand-side rows runs through the FP64 matrix-vector product whatever the handle's variance mode
    (gptb_query_dev): its std / dvar equal the "fp64" handle's bit for bit, while a large batch in INT8 mode differs from it at 1e-9."""
    from gaussian_process_transportation_b200 import _lib as L
    N = 2048
    X, Y = problem(N, seed=7)
    rng = np.random.default_rng(11)
    xq = rng.random((600, 3))
    out = {}
    for mode in ("fp64", "int8w5"):
        eng = L.Engine(0)
        try:
            eng.set_variance_mode(*L.parse_variance_mode(mode))
            eng.set_train(X, Y)
            assert eng.factorize(0.1, np.full(3, 0.1), 1e-4, 1e-10, want_lml=False)[0] == 0
            eng.prepare_variance()
            out[mode] = (eng.query(xq[:1], L.MEAN | L.STD | L.DVAR), eng.query(xq[:8], L.MEAN | L.STD), eng.query(xq, L.MEAN | L.STD))
        finally:
            eng.close()
    for i in (0, 1):
        for key in out["fp64"][i]:
            assert np.array_equal(out["fp64"][i][key], out["int8w5"][i][key]), (i, key)
    big = np.max(np.abs(out["fp64"][2]["std"] - out["int8w5"][2]["std"]))
    assert 0.0 < big < 1e-7 * np.sqrt(0.1 + 1e-4)
