"""CPU: the per-path arithmetic shared by all kernels (csrc/optmc_math.cuh, host-compiled by g++) against the
numpy oracle and the Random123 known-answer vectors.  The harness is test infrastructure, not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import lsm_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
DP = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostsim") / "libhostsim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(HERE, "hostsim", "host_math.cpp"), "-o", out], check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(DP)


def test_philox_kat(hs):
    """Random123 Philox4x32-10 vectors."""
    kats = [([0] * 4, [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
            ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
            ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
             [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1])]
    for ctr, key, exp in kats:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        hs.hs_philox(c, k, o)
        assert list(o) == exp


@pytest.mark.parametrize("scheme", [2, 3, 4, 5])
def test_heston_step_functions_vs_oracle(hs, scheme):
    rng = np.random.default_rng(4)
    M, N = 256, 40
    Z1, Z2 = orc.draw_heston_normals(rng, N, M)
    S = np.zeros((N + 1, M)); V = np.zeros((N + 1, M))
    hs.hs_heston_paths.argtypes = [C.c_int] + [C.c_double] * 8 + [C.c_long, C.c_int, DP, DP, DP, DP]
    hs.hs_heston_paths(scheme, 100.0, HP["v0"], 0.05, 1.0, HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N,
                       _p(Z1), _p(Z2), _p(S), _p(V))
    if scheme == 2:
        ref = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
    elif scheme == 3:
        ref = orc.heston_paths_full_truncation(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
    elif scheme == 5:
        ref, Vref = orc.heston_paths_qe(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2,
                                        return_v=True)
        np.testing.assert_allclose(V, Vref, rtol=1e-9, atol=1e-15)
        psi_hi = (V[:-1] * 0 + 1).sum()  # both branches are exercised with the Feller-violating parameters
        assert (V == 0).any() and (V > 0.05).any() and psi_hi > 0
    else:  # calibrator: path-major, Z given path-major (n_sim, N)
        Sp, _ = orc.hc_simulate_paths(HP["kappa"], HP["theta"], HP["xi"], HP["rho"], HP["v0"], 100.0, 1.0, 0.05, M, N,
                                      np.ascontiguousarray(Z1.T), np.ascontiguousarray(Z2.T))
        ref = Sp.T
    np.testing.assert_allclose(S, ref, rtol=1e-9 if scheme == 5 else 1e-12)
    if scheme == 2:
        assert (V >= 0).all() and (V == 0).any()  # the absorption at zero is exercised (Feller violated)


def test_gbm_step_vs_oracle(hs):
    rng = np.random.default_rng(5)
    M, N = 128, 50
    Z = orc.draw_gbm_normals(rng, N, M)
    S = np.zeros((N + 1, M))
    hs.hs_gbm_paths.argtypes = [C.c_double] * 4 + [C.c_long, C.c_int, DP, DP]
    hs.hs_gbm_paths(100.0, 0.05, 0.2, 1.0, M, N, _p(Z), _p(S))
    np.testing.assert_allclose(S, orc.gbm_paths_antithetic(100.0, 0.05, 0.2, 1.0, M, N, Z), rtol=1e-13)


@pytest.mark.parametrize("deg", [2, 3])
def test_moments_and_guarded_solve_vs_oracle(hs, deg):
    rng = np.random.default_rng(6)
    n = 20000
    x = rng.uniform(0.55, 1.0, n); y = np.maximum(1 - x, 0) * 100 * rng.uniform(0.8, 1.2, n)
    Q = 3 * deg + 2
    mom = np.zeros(Q); beta = np.zeros(deg + 1)
    hs.hs_fit.argtypes = [C.c_int, C.c_long, DP, DP, DP, DP]
    assert hs.hs_fit(deg, n, _p(x), _p(y), _p(mom), _p(beta)) == 1
    Phi = np.column_stack([x**i for i in range(deg + 1)])
    G = Phi.T @ Phi
    for i in range(deg + 1):
        for j in range(deg + 1):
            assert mom[i + j] == pytest.approx(G[i, j], rel=1e-12)
    np.testing.assert_allclose(mom[2 * deg + 1:], Phi.T @ y, rtol=1e-12)
    ref = orc.cholesky_solve_guarded(G, Phi.T @ y)
    np.testing.assert_allclose(Phi @ beta, Phi @ ref, rtol=1e-8, atol=1e-8)
    hs.hs_poly_eval.restype = C.c_double
    hs.hs_poly_eval.argtypes = [C.c_int, DP, C.c_double]
    assert hs.hs_poly_eval(deg, _p(beta), 0.8) == pytest.approx(float(np.polyval(beta[::-1], 0.8)), rel=1e-13)
    # guard: fewer rows than columns, and a rank-deficient design
    assert hs.hs_fit(deg, deg, _p(x), _p(y), _p(mom), _p(beta)) == 0
    xc = np.full(100, 0.9)
    assert hs.hs_fit(deg, 100, _p(xc), _p(y), _p(mom), _p(beta)) == 0


def test_box_muller_transforms(hs):
    rng = np.random.default_rng(7)
    w = rng.integers(0, 2**32, size=(200000, 2), dtype=np.uint64)
    o32 = (C.c_float * 2)(); o64 = (C.c_double * 2)()
    hs.hs_normal2_f32.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    hs.hs_normal2_f64.argtypes = [C.c_uint32, C.c_uint32, DP]
    z32 = np.empty((20000, 2)); z64 = np.empty((20000, 2))
    for i in range(20000):
        hs.hs_normal2_f32(int(w[i, 0]), int(w[i, 1]), o32); z32[i] = o32[0], o32[1]
        hs.hs_normal2_f64(int(w[i, 0]), int(w[i, 1]), o64); z64[i] = o64[0], o64[1]
    for z in (z32, z64):
        assert abs(z.mean()) < 0.03 and abs(z.var() - 1) < 0.04 and np.isfinite(z).all()
    np.testing.assert_allclose(z32, z64, atol=2e-3)  # same words -> same normals up to the 2^-23 uniform grid
    # extremes stay finite: u = 1 (a = 0) gives 0, the smallest u gives |z| < 6
    hs.hs_normal2_f32(0, 0, o32); assert o32[0] == 0.0
    hs.hs_normal2_f32(0xFFFFFFFF, 0, o32); assert 5 < o32[0] < 6


def test_features_ref7(hs):
    f = np.zeros(7)
    hs.hs_features.argtypes = [C.c_double] * 3 + [DP]
    hs.hs_features(130.0, 100.0, np.sqrt(0.7), _p(f))
    np.testing.assert_allclose(f, orc.features_ref7(np.array([130.0]), 100.0, 0.05, 1.0, 0.3)[0], rtol=1e-15)


def test_fixed_point_exchange_is_exact_and_order_independent(hs):
    """The persistent sweep sums CTA totals as two 48-bit fixed-point chunks with integer adds
    (optmc_math.cuh: fx_encode / fx_decode): any arrival order gives the same bits, the arrival count sits
    in the top byte, and the decoded total is within n * 2^-52 of the exact sum."""
    hs.hs_fx_sum.argtypes = [C.c_long, DP, DP, C.POINTER(C.c_uint64)]
    hs.hs_fx_sum.restype = C.c_int
    rng = np.random.default_rng(3)
    from fractions import Fraction

    for n, scale in ((1, 1.0), (148, 1.0), (148, 1e6), (160, 1e9), (148, 1e-6), (37, 3.0)):
        v = rng.standard_normal(n) * scale
        if scale == 3.0:
            v = np.abs(v).round()  # integers (row counts) are exact
        out = C.c_double()
        words = (C.c_uint64 * 2)()
        assert hs.hs_fx_sum(n, _p(v), C.byref(out), words) == 1
        assert words[0] >> 56 == n % 256 and words[1] >> 56 == n % 256
        exact = sum(Fraction(float(x)) for x in v)
        assert abs(Fraction(out.value) - exact) <= n * Fraction(1, 2**52) + abs(exact) * Fraction(1, 2**52)
        if scale == 3.0:
            assert out.value == float(exact)
        perm = rng.permutation(n)
        out2 = C.c_double()
        words2 = (C.c_uint64 * 2)()
        hs.hs_fx_sum(n, _p(np.ascontiguousarray(v[perm])), C.byref(out2), words2)
        assert (words2[0], words2[1]) == (words[0], words[1]) and out2.value == out.value
    # out of range / non-finite values are reported, never silently wrapped
    for bad in (1e13, -1e13, np.inf, np.nan):
        v = np.array([1.0, bad])
        assert hs.hs_fx_sum(2, _p(v), C.byref(out), words) == 0


def test_brownian_bridge_schedule_matches_oracle_and_reproduces_brownian_covariance():
    """optmc_qmc_bridge_schedule (host code of csrc/qmc.cu, no GPU needed) == the oracle's restatement of Jaeckel's
    construction; as a linear map W = A z it must reproduce Cov(W_s, W_t) = min(s, t) exactly, for any N."""
    import ctypes as C

    from options_model_b200 import _lib as L
    from oracle import lsm_oracle as orc

    lib = L.load_library()
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    for N in (1, 2, 3, 7, 50, 64, 252):
        idx, left, right = (np.zeros(N, dtype=np.int32) for _ in range(3))
        wl, wr, sd = (np.zeros(N) for _ in range(3))
        assert lib.optmc_qmc_bridge_schedule(N, idx.ctypes.data_as(ip), left.ctypes.data_as(ip), right.ctypes.data_as(ip),
                                             wl.ctypes.data_as(dp), wr.ctypes.data_as(dp), sd.ctypes.data_as(dp)) == 0
        o = orc.brownian_bridge_schedule(N)
        for got, want in zip((idx, left, right, wl, wr, sd), o):
            np.testing.assert_allclose(got, want, rtol=1e-15)
        assert sorted(idx.tolist()) == list(range(N))
        A = np.zeros((N, N))  # W = A z
        A[idx[0], 0] = sd[0]
        for s in range(1, N):
            lo = A[left[s] - 1] if left[s] > 0 else 0.0
            A[idx[s]] = wl[s] * lo + wr[s] * A[right[s]]
            A[idx[s], s] += sd[s]
        t = np.arange(1, N + 1)
        np.testing.assert_allclose(A @ A.T, np.minimum.outer(t, t), rtol=1e-12, atol=1e-12)


def test_engine_stream_restatements_are_well_formed():
    """oracle/engine_streams.py (paired network check): the Feistel shuffle is a bijection of [0, n) that depends on the
    key, the dropout mask keeps (256 - round(256 p)) / 256 of the units and differs between layers / steps."""
    from oracle import engine_streams as es

    for n in (1, 2, 3, 1000, 65_537):
        p = es.feistel_perm(n, es.perm_key(42, 0))
        assert np.array_equal(np.sort(p), np.arange(n))
    a, b = es.feistel_perm(5000, es.perm_key(42, 0)), es.feistel_perm(5000, es.perm_key(42, 1))
    assert (a != b).mean() > 0.99 and (a != np.arange(5000)).mean() > 0.99
    rows = np.arange(4096, dtype=np.uint32)
    k0 = es.keep_mask(es.train_drop_key(42, 1), rows, 0, 0.1)
    k1 = es.keep_mask(es.train_drop_key(42, 1), rows, 1, 0.1)
    k2 = es.keep_mask(es.train_drop_key(42, 2), rows, 0, 0.1)
    assert k0.shape == (4096, 128)
    assert abs(k0.mean() - 230 / 256) < 2e-3 and es.keep_scale(0.1) == 256 / 230
    assert 0.1 < (k0 != k1).mean() < 0.25 and 0.1 < (k0 != k2).mean() < 0.25
    assert es.keep_mask(7, rows, 0, 0.0).all()
