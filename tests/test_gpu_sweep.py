"""GPU parity, polynomial LSM sweeps: persistent and streaming implementations against the oracle, edge cases, batches, path sharding, out-of-sample policy.

All calls go through the C ABI (ctypes).  Tolerances (north star): fed identical draws, prices / betas / boundary within
1e-5 relative in fp64 and 1e-4 in fp32 -- the fp64 assertions are far tighter; integer outputs are compared exactly.
"""
import os  # noqa: F401

import numpy as np
import pytest

from gpu_common import _dev, _slab, _check_sweep, _oracle_price_philox, HP  # noqa: F401

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("impl", ["resident", "split"])
@pytest.mark.parametrize("semantics", ["reference", "textbook"])
@pytest.mark.parametrize("basis", ["poly2", "poly3"])
def test_sweep_fp64_vs_oracle_small(eng, mods, golden_dir, impl, semantics, basis):
    L, E, orc = mods
    res_o, S, _ = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 4096, 20, orc.RNGManager(1),
                                         heston_params=HP, return_paths=True)
    if basis == "poly2" and semantics == "reference":  # the committed golden of the restatement
        g = np.load(os.path.join(golden_dir, "oracle_heston_poly2_small.npz"))
        assert res_o.price == pytest.approx(float(g["price"]), rel=1e-13)
    ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", basis=basis, semantics=semantics)
    res = eng.lsm(_slab(eng, S, torch.float64), 100.0, 0.05, 1.0, "put", basis, semantics, impl)
    assert res.impl_used == {"resident": L.SWEEP_RESIDENT, "split": L.SWEEP_SPLIT}[impl]
    # the cubic normal equations are worse conditioned (cond ~1e9): looser check on the fitted curve only
    _check_sweep(res, ref, 1e-11, beta_rtol=1e-6 if basis == "poly2" else 1e-3)


@pytest.mark.parametrize("impl", ["resident", "split"])
def test_sweep_call_and_gbm(eng, mods, impl):
    L, E, orc = mods
    rng = np.random.default_rng(5)
    M, N = 6000, 15
    S = orc.gbm_paths_antithetic(100.0, 0.05, 0.3, 1.0, M, N, orc.draw_gbm_normals(rng, N, M))
    for ot, K in (("call", 95.0), ("put", 110.0)):
        ref = orc.lsm_sweep(S, K, 0.05, 1.0, ot)
        res = eng.lsm(_slab(eng, S, torch.float64), K, 0.05, 1.0, ot, impl=impl)
        _check_sweep(res, ref, 1e-11)


def test_config1_full_size_fp64_and_pins(eng, mods, golden_meta):
    """BASELINE config 1: GBM put 100k x 50, RNGManager(42) child-0 draws; survey pins (6.542437 / 6.042107)."""
    L, E, orc = mods
    mgr = orc.RNGManager(42)
    rng = mgr.get_child_rng()
    M, N = 100_000, 50
    Zh = orc.draw_gbm_normals(rng, N, M)
    S_gpu = eng.paths(E.gbm(100.0, 0.05, 1.0, 0.2), M, N, "f64", E.RngSpec(z1=_dev(Zh)))
    S = orc.gbm_paths_antithetic(100.0, 0.05, 0.2, 1.0, M, N, Zh)
    np.testing.assert_allclose(S_gpu.cpu().numpy(), S, rtol=1e-12)
    pins = golden_meta["poly_pins"]["c1_gbm_put_100k_50"]
    for sem in ("reference", "textbook"):
        ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", semantics=sem)
        assert ref.price == pytest.approx(pins[sem]["price"], rel=1e-12)
        for impl in ("resident", "split"):
            res = eng.lsm(S_gpu, 100.0, 0.05, 1.0, "put", "poly2", sem, impl)
            _check_sweep(res, ref, 1e-10, boundary_rtol=1e-12)  # S_gpu vs numpy S differ in the last ulp
            assert res.boundary[25] == pytest.approx(pins[sem]["boundary_25"], rel=1e-12)
            assert res.boundary[45] == pytest.approx(pins[sem]["boundary_45"], rel=1e-12)


@pytest.mark.parametrize("impl", ["resident", "split"])
def test_sweep_fp32_vs_oracle(eng, mods, impl):
    """fp32 storage: the same fp32 path values go to the kernel and (widened) to the oracle; tolerance 1e-4."""
    L, E, orc = mods
    rng = np.random.default_rng(11)
    M, N = 50_000, 40
    Z1, Z2 = orc.draw_heston_normals(rng, N, M)
    S32 = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N,
                                      Z1, Z2).astype(np.float32)
    ref = orc.lsm_sweep(S32.astype(np.float64), 100.0, 0.05, 1.0, "put")
    res = eng.lsm(_slab(eng, S32, torch.float32), 100.0, 0.05, 1.0, "put", impl=impl)
    assert res.price == pytest.approx(ref.price, rel=1e-4)
    assert res.stderr == pytest.approx(ref.stderr, rel=1e-4)
    assert np.abs(res.ex_count - ref.ex_count).sum() <= 1e-4 * M  # a handful of borderline decisions may flip
    np.testing.assert_array_equal(res.n_itm[N - 1], ref.n_itm[N - 1])


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 1e-4)])
def test_price_american_philox_vs_oracle_same_draws(eng, mods, dtype, tol):
    """Whole fused call (Philox paths + sweep) vs the oracle fed the kernel's own normals."""
    L, E, orc = mods
    M, N, K = 32768, 24, 100.0
    rng = E.RngSpec(seed=99, stream=3)
    res = eng.price_american(E.heston(100.0, 0.05, 1.0, **HP), M, N, K, "put", dtype, rng, arrays=True)
    z1 = eng.philox_normals(L.MODEL_HESTON, M, N, 0, dtype, rng).double().cpu().numpy()
    z2 = eng.philox_normals(L.MODEL_HESTON, M, N, 1, dtype, rng).double().cpu().numpy()
    S = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, z1, z2)
    ref = orc.lsm_sweep(S, K, 0.05, 1.0, "put")
    assert res.price == pytest.approx(ref.price, rel=tol)
    if dtype == "f64":
        _check_sweep(res, ref, tol, boundary_rtol=1e-12)


def test_independent_rng_within_3_se(eng, mods):
    """Config 1 with in-kernel Philox vs the oracle on numpy PCG64 draws: |diff| < 3 combined standard errors."""
    L, E, orc = mods
    ref = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 100_000, 50, orc.RNGManager(42), sigma=0.2)
    res = eng.price_american(E.gbm(100.0, 0.05, 1.0, 0.2), 100_000, 50, 100.0, "put", "f32", E.RngSpec(seed=2024))
    assert abs(res.price - ref.price) < 3 * np.hypot(res.stderr, ref.stderr)
    ref_h = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 100_000, 50, orc.RNGManager(42), heston_params=HP)
    res_h = eng.price_american(E.heston(100.0, 0.05, 1.0, **HP), 100_000, 50, 100.0, "put", "f32", E.RngSpec(seed=2025))
    assert abs(res_h.price - ref_h.price) < 3 * np.hypot(res_h.stderr, ref_h.stderr)


@pytest.mark.parametrize("impl", ["resident", "split"])
def test_edge_cases(eng, mods, impl):
    L, E, orc = mods
    rng = np.random.default_rng(2)
    # (a) never in the money: deep OTM put -> no regression at any date, price = discounted terminal payoff = 0
    M, N = 1000, 6
    S = orc.gbm_paths_antithetic(100.0, 0.05, 0.1, 0.5, M, N, orc.draw_gbm_normals(rng, N, M))
    ref = orc.lsm_sweep(S, 1.0, 0.05, 0.5, "put")
    res = eng.lsm(_slab(eng, S, torch.float64), 1.0, 0.05, 0.5, "put", impl=impl)
    assert res.price == 0.0 == ref.price and np.all(np.isnan(res.betas)) and res.n_itm.sum() == 0
    # (b) N = 1: no exercise date at all (range(N-1, 0, -1) is empty) -> mean terminal payoff, zero discounts
    S = orc.gbm_paths_antithetic(100.0, 0.05, 0.2, 1.0, M, 1, orc.draw_gbm_normals(rng, 1, M))
    ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put")
    res = eng.lsm(_slab(eng, S, torch.float64), 100.0, 0.05, 1.0, "put", impl=impl)
    assert res.price == pytest.approx(ref.price, rel=1e-13)
    # (c) ragged sizes: M = 2, M % 4 == 2, N = 2; fewer ITM rows than basis columns -> "no exercise" (8(c))
    for M2, N2 in ((2, 2), (1002, 3), (514, 7)):
        S = orc.gbm_paths_antithetic(100.0, 0.05, 0.2, 1.0, M2, N2, orc.draw_gbm_normals(rng, N2, M2))
        ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put")
        res = eng.lsm(_slab(eng, S, torch.float64), 100.0, 0.05, 1.0, "put", impl=impl)
        _check_sweep(res, ref, 1e-12)
    # (d) degenerate regression: every path identical -> pivot guard -> no exercise, price = discounted payoff
    S = np.tile(np.linspace(100.0, 90.0, 5)[:, None], (1, 64))
    ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put")
    res = eng.lsm(_slab(eng, S, torch.float64), 100.0, 0.05, 1.0, "put", impl=impl)
    _check_sweep(res, ref, 1e-13)
    assert res.ex_count.sum() == 0


def test_unaligned_slab_falls_back_to_split(eng, mods):
    """fp32 rows of 1002 elements are 4008 bytes: not a multiple of 16, so the bulk copies cannot be used.
    AUTO must pick SPLIT, RESIDENT must refuse (NotImplementedError), results must still match the oracle."""
    L, E, orc = mods
    rng = np.random.default_rng(8)
    M, N = 1002, 5
    S32 = orc.gbm_paths_antithetic(100.0, 0.05, 0.2, 1.0, M, N, orc.draw_gbm_normals(rng, N, M)).astype(np.float32)
    Sd = _dev(S32)  # contiguous: ld == M
    ref = orc.lsm_sweep(S32.astype(np.float64), 100.0, 0.05, 1.0, "put")
    res = eng.lsm(Sd, 100.0, 0.05, 1.0, "put", impl="auto")
    assert res.impl_used == L.SWEEP_SPLIT
    assert res.price == pytest.approx(ref.price, rel=1e-5)
    np.testing.assert_array_equal(res.n_itm, ref.n_itm)
    with pytest.raises(NotImplementedError):
        eng.lsm(Sd, 100.0, 0.05, 1.0, "put", impl="resident")


def test_path_sharded_sweep_equals_single(mods):
    from options_model_b200 import engine as E2
    from options_model_b200.sharded import shard_pairs

    L, E, orc = mods
    M, N, K = 8192, 12, 100.0
    model = E.heston(100.0, 0.05, 1.0, **HP)
    e0, e1 = E2.Engine(0), E2.Engine(0)
    try:
        full = e0.paths(model, M, N, "f64", E.RngSpec(seed=77)).clone()
        single = e0.lsm(full, K, 0.05, 1.0, "put", impl="split")
        shards = []
        for rank, e in enumerate((e0, e1)):
            off, m_loc = shard_pairs(M, rank, 2)
            shards.append(e.paths(model, m_loc, N, "f64", E.RngSpec(seed=77, pair_offset=off)).clone())
        q = e0.gram_len("poly2")
        g = [torch.zeros(q, dtype=torch.float64, device="cuda") for _ in range(2)]
        for e, S in zip((e0, e1), shards):
            e.lsm_begin(S, K, 0.05, 1.0, "put")
        for t in range(N - 1, 0, -1):
            for e, gi in zip((e0, e1), g):
                e.lsm_gram_date(t, gi)
            tot = g[0] + g[1]  # stands in for the NCCL all-reduce
            for e in (e0, e1):
                e.lsm_update_date(t, tot)
        s = [torch.zeros(3, dtype=torch.float64, device="cuda") for _ in range(2)]
        for e, si in zip((e0, e1), s):
            e.lsm_finish(si)
        tot = (s[0] + s[1]).cpu().numpy()
        assert tot[2] == M
        assert tot[0] / tot[2] == pytest.approx(single.price, rel=1e-12)
    finally:
        e0.close(); e1.close()


def test_config2_full_size_properties(eng, mods):
    """Config 2 (Heston put, 1M x 252, fp32): resident == split, American >= European, terminal-row checksum,
    and the documented reference-semantics value (5.83 at N=252) within Monte-Carlo error."""
    L, E, orc = mods
    M, N, K = 1_000_000, 252, 100.0
    model = E.heston(100.0, 0.05, 1.0, **HP)
    S = eng.paths(model, M, N, "f32", E.RngSpec(seed=42))
    assert torch.all(S[0] == 100.0) and torch.isfinite(S[N]).all() and (S[N] > 0).all()
    a = eng.lsm(S, K, 0.05, 1.0, "put", impl="resident")
    b = eng.lsm(S, K, 0.05, 1.0, "put", impl="split")
    assert a.impl_used == L.SWEEP_RESIDENT and b.impl_used == L.SWEEP_SPLIT
    assert a.price == pytest.approx(b.price, rel=1e-6)
    assert np.abs(a.ex_count - b.ex_count).sum() <= 1e-5 * M * N
    np.testing.assert_array_equal(a.n_itm[N - 1], b.n_itm[N - 1])
    eu, eu_se = eng.european_from_slab(S[N].contiguous(), K, 0.05, 1.0, "put")
    assert a.price > eu
    assert abs(eu - 5.3234) < 0.05 and abs(a.price - 5.83) < 0.06
    tb = eng.lsm(S, K, 0.05, 1.0, "put", semantics="textbook")
    assert eu < tb.price < a.price  # look-ahead bias of the sticky mask (App. A, Q1)


@pytest.mark.parametrize("M", [4096, 40_000])
def test_batch_matches_oracle_and_single(eng, mods, M):
    """optmc_price_american_batch: heterogeneous S0 / K / T / N / put-call options, each on its own Philox
    stream.  fp64: every price equals the oracle fed the same normals; it also equals the single-option call.
    M = 4096 -> one CTA per option (no exchange), M = 40 000 -> several CTAs per option and several waves."""
    L, E, orc = mods
    S0 = np.array([100.0, 95.0, 105.0, 100.0, 110.0, 90.0, 100.0])
    K = np.array([100.0, 100.0, 100.0, 105.0, 100.0, 100.0, 98.0])
    T = np.array([1.0, 0.5, 0.25, 1.0, 0.75, 0.1, 1.0])
    N = np.array([20, 13, 10, 24, 16, 10, 11])
    put = np.array([1, 1, 1, 1, 0, 0, 1])
    streams = np.array([3, 5, 8, 13, 21, 34, 55])
    model = E.heston(100.0, 0.05, 1.0, **HP)
    for sem in ("reference", "textbook"):
        price, se = eng.price_american_batch(model, M, S0, K, T, N, put, "f64", E.RngSpec(seed=77), semantics=sem,
                                             streams=streams)
        for i in range(len(S0)):
            ot = "put" if put[i] else "call"
            ref = _oracle_price_philox(eng, mods, None, S0[i], K[i], T[i], int(N[i]), M, ot, 77, int(streams[i]),
                                       semantics=sem)
            assert price[i] == pytest.approx(ref.price, rel=1e-10), (sem, i)
            assert se[i] == pytest.approx(ref.stderr, rel=1e-8, abs=1e-12)
            single = eng.price_american(E.heston(S0[i], 0.05, T[i], **HP), M, int(N[i]), K[i], ot, "f64",
                                        E.RngSpec(seed=77, stream=int(streams[i])), semantics=sem)
            assert price[i] == pytest.approx(single.price, rel=1e-12)


def test_batch_many_small_options_fp32(eng, mods):
    """The curve-driver shape (om3:697-713): hundreds of small pricings, N = max(10, min(130, ceil(days)))."""
    L, E, orc = mods
    days = np.arange(200, 0, -1, dtype=np.float64)
    N = np.maximum(10, np.minimum(130, np.ceil(days))).astype(np.int64)
    T = days / 365
    model = E.gbm(100.0, 0.05, 1.0, 0.2)
    price, se = eng.price_american_batch(model, 10_000, 100.0, 100.0, T, N, 1, "f32", E.RngSpec(seed=5))
    assert np.isfinite(price).all() and (se > 0).all()
    # American put >= European put (Black-Scholes) within Monte-Carlo error, and values grow with maturity overall
    from options_model_b200 import compat

    bs = np.array([compat.BlackScholesGreeks.black_scholes_price(100.0, 100.0, t, 0.05, 0.2, "put") for t in T])
    assert np.all(price > bs - 4 * se)
    assert price[0] > price[-1]
    # spot-check three grid points against the single-option entry on the same stream
    for i in (0, 77, 199):
        single = eng.price_american(E.gbm(100.0, 0.05, T[i], 0.2), 10_000, int(N[i]), 100.0, "put", "f32",
                                    E.RngSpec(seed=5, stream=i))
        assert price[i] == pytest.approx(single.price, rel=1e-5)


@pytest.mark.parametrize("semantics", ["textbook", "reference"])
@pytest.mark.parametrize("basis,dtype", [("poly2", "f64"), ("poly3", "f64"), ("poly2", "f32")])
def test_apply_policy_out_of_sample_vs_oracle(eng, mods, semantics, basis, dtype):
    L, E, orc = mods
    M, N, K = 20_000, 20, 100.0
    model = E.heston(100.0, 0.05, 1.0, **HP)
    S_fit = eng.paths(model, M, N, "f64", E.RngSpec(seed=101))
    fit = eng.lsm(S_fit, K, 0.05, 1.0, "put", basis, semantics)
    S_new = eng.paths(model, M, N, dtype, E.RngSpec(seed=202))
    got = eng.lsm_apply_policy(S_new, fit.betas, K, 0.05, 1.0, "put", basis, semantics)
    ref = orc.lsm_sweep(S_new.cpu().numpy().astype(np.float64), K, 0.05, 1.0, "put", orc.FixedPolicyRegressor(K, fit.betas),
                        semantics=semantics)
    assert got.price == pytest.approx(ref.price, rel=1e-12 if dtype == "f64" else 1e-6)
    assert got.stderr == pytest.approx(ref.stderr, rel=1e-9 if dtype == "f64" else 1e-5)
    np.testing.assert_array_equal(got.ex_count, ref.ex_count)
    np.testing.assert_array_equal(np.isnan(got.boundary), np.isnan(ref.boundary))
    np.testing.assert_allclose(got.boundary[~np.isnan(ref.boundary)], ref.boundary[~np.isnan(ref.boundary)], rtol=0)
    if semantics == "textbook":  # a fixed policy on fresh paths is a lower bound in expectation: below the in-sample value
        assert got.price < fit.price + 3 * fit.stderr


def test_ref7_per_date_spans_poly3(eng, mods):
    """SURVEY 8(d) C2 "poly2 and ref7 bases": within one date the seven reference features span [1, x, x^2, x^3], so a
    per-date least-squares fit on them has the fitted values of POLY3 -- checked against numpy's minimum-norm lstsq on
    the real seven-column design matrix (oracle.features_ref7)."""
    L, E, orc = mods
    M, N, K = 4096, 10, 100.0
    S = eng.paths(E.heston(100.0, 0.05, 1.0, **HP), M, N, "f64", E.RngSpec(seed=12))
    got = eng.lsm(S, K, 0.05, 1.0, "put", "ref7", "textbook")
    p3 = eng.lsm(S, K, 0.05, 1.0, "put", "poly3", "textbook")
    assert got.price == p3.price and got.betas.shape == (N + 1, 4)

    class Ref7Lstsq:  # per-date regression on all seven features, minimum-norm solution
        p = 7

        def __call__(self, t, t_current, S_itm, Y):
            F = orc.features_ref7(S_itm, K, 0.05, 1.0, t_current)
            if len(S_itm) < 4:
                return None, None
            w, *_ = np.linalg.lstsq(F, np.asarray(Y, dtype=np.float64), rcond=1e-12)
            return F @ w, w

    ref = orc.lsm_sweep(S.cpu().numpy(), K, 0.05, 1.0, "put", Ref7Lstsq(), semantics="textbook")
    assert got.price == pytest.approx(ref.price, rel=1e-7)
    np.testing.assert_array_equal(got.ex_count, ref.ex_count)


def test_error_paths_of_the_newer_entry_points(eng, mods):
    """Loud failures instead of silent fallbacks: argument errors map to ValueError with a message, unsupported shapes to
    NotImplementedError / RuntimeError."""
    L, E, orc = mods
    gbm = E.gbm(100.0, 0.05, 1.0, 0.2)
    S = eng.paths(gbm, 4096, 8, "f32", E.RngSpec(seed=1))
    with pytest.raises(ValueError, match="optmc_comm_init must be called first"):
        eng.lsm_sharded(S, 8192, 100.0, 0.05, 1.0)
    with pytest.raises(NotImplementedError, match="SingleLSMNet"):
        eng.lsm_gnet(S, 100.0, 0.05, 1.0, hidden=64)
    with pytest.raises(ValueError, match="bad training parameters"):
        eng.lsm_gnet(S, 100.0, 0.05, 1.0, lr=-1.0)
    net = dict(hidden=64, layers=4, weights=np.zeros(100, dtype=np.float32), m_scale=0.1, tau_scale=0.3, epsilon=1e-4)
    with pytest.raises(ValueError, match="weight count"):
        eng.paths_localvol(100.0, 0.05, 1.0, net, 100.0, 1024, 4)
    net48 = dict(net, hidden=48, weights=np.zeros(3 * 48 + 4 * (48 * 48 + 144) + 49, dtype=np.float32))
    with pytest.raises(NotImplementedError, match="hidden_dim"):
        eng.paths_localvol(100.0, 0.05, 1.0, net48, 100.0, 1024, 4)
    with pytest.raises(ValueError, match="antithetic"):
        eng.paths_localvol(100.0, 0.05, 1.0, dict(net, weights=np.zeros(17409, dtype=np.float32)), 100.0, 1023, 4)
    res = eng.lsm(S, 100.0, 0.05, 1.0, "put", impl="resident", arrays=False)
    assert np.isfinite(res.price)
    with pytest.raises(ValueError, match="cash-flows in device memory"):
        eng.lsm_zero_cashflows()  # the persistent sweep keeps them in registers
    eng.lsm(S, 100.0, 0.05, 1.0, "put", impl="split", arrays=False)
    assert 0 <= eng.lsm_zero_cashflows() <= 4096
    with pytest.raises(AssertionError):
        eng.lsm_apply_policy(S, np.zeros((3, 3)), 100.0, 0.05, 1.0)  # betas must be [(N+1), p]


def test_bench_shape_4x1M_fp32_vs_oracle_same_draws(eng, mods):
    """The launch bench.py times -- optmc_price_american_batch, 4 options x 1 M paths x 252 dates, fp32 slabs, one
    grouped persistent sweep (4 groups of 37 CTAs, 768 threads x 36 paths) -- against the oracle fed the kernel's own
    Philox normals, both semantics: prices within 1e-4 (north star, fp32), exercise counts within 2e-4 of the paths
    (fp32 paths vs fp64 oracle paths: a handful of borderline decisions may flip), regression-row counts of the
    first regression equal up to the same flips."""
    L, E, orc = mods
    M, N, K, B = 1_000_000, 252, 100.0, 4
    model = E.heston(100.0, 0.05, 1.0, **HP)
    streams = np.array([11, 12, 13, 14])
    out = {}
    for sem in ("reference", "textbook"):
        price, se, ex = eng.price_american_batch(model, M, 100.0, K, 1.0, np.full(B, N), 1, "f32", E.RngSpec(seed=42),
                                                 semantics=sem, streams=streams, details=True, european=True)
        out[sem] = (price, se, ex)
        nt, ppt, cpg, G = ex["shape"]
        assert (cpg, G) == (eng.sm_count // B, B) and nt * ppt * cpg >= M, ex["shape"]
        if sem == "reference" and eng.sm_count == 148:
            # the wide shape of the headline bench: single-role 768 x 36 or speculative (736 + 32) x 37, whichever the
            # batch entry point timed faster on this box (both kernels give bit-identical results)
            assert (nt, ppt) in ((768, 36), (736, 37)), ex["shape"]
    for i in range(B):
        rng = E.RngSpec(seed=42, stream=int(streams[i]))
        z1 = eng.philox_normals(L.MODEL_HESTON, M, N, 0, "f32", rng).double().cpu().numpy()
        z2 = eng.philox_normals(L.MODEL_HESTON, M, N, 1, "f32", rng).double().cpu().numpy()
        S = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, z1, z2)
        del z1, z2
        eu_ref = orc.european_from_paths(S[N], K, 0.05, 1.0, "put")
        for sem in ("reference", "textbook"):
            ref = orc.lsm_sweep(S, K, 0.05, 1.0, "put", semantics=sem)
            price, se, ex = out[sem]
            assert price[i] == pytest.approx(ref.price, rel=1e-4), (sem, i)
            assert se[i] == pytest.approx(ref.stderr, rel=1e-3), (sem, i)
            # borderline decisions flip between the fp32 paths and the oracle's fp64 paths: ~1e-4 M under the sticky mask
            # (every path decides once); textbook semantics count every date's decisions (~1.4e5 per date)
            flips = np.abs(ex["ex_count"][i] - ref.ex_count).sum()
            assert flips <= (2e-4 * M if sem == "reference" else 1e-3 * ref.ex_count.sum()), (sem, i, flips)
            assert abs(int(ex["n_itm"][i, N - 1]) - int(ref.n_itm[N - 1])) <= 1e-5 * M
            assert ex["european"][i, 0] == pytest.approx(eu_ref[0], rel=1e-5)
            assert ex["european"][i, 1] == pytest.approx(eu_ref[1], rel=1e-4)
        del S


def test_batch_extras_equal_single_option_outputs(eng, mods):
    """optmc_price_american_batch_ex: per-date outputs and the European leg of every option equal the single-option
    calls on the same Philox stream (fp64, bit-level agreement of integers, 1e-12 of floats)."""
    L, E, orc = mods
    M = 20_000
    S0 = np.array([100.0, 96.0, 104.0]); K = np.array([100.0, 100.0, 98.0]); T = np.array([1.0, 0.5, 0.75])
    N = np.array([30, 17, 24]); put = np.array([1, 1, 0]); streams = np.array([2, 4, 6])
    model = E.heston(100.0, 0.05, 1.0, **HP)
    for sem in ("reference", "textbook"):
        price, se, ex = eng.price_american_batch(model, M, S0, K, T, N, put, "f64", E.RngSpec(seed=31), semantics=sem,
                                                 streams=streams, details=True, european=True)
        for i in range(3):
            ot = "put" if put[i] else "call"
            m = E.heston(S0[i], 0.05, T[i], **HP)
            rs = E.RngSpec(seed=31, stream=int(streams[i]))
            single = eng.price_american(m, M, int(N[i]), K[i], ot, "f64", rs, semantics=sem, arrays=True)
            n1 = int(N[i]) + 1
            assert price[i] == pytest.approx(single.price, rel=1e-12)
            np.testing.assert_array_equal(ex["ex_count"][i, :n1], single.ex_count)
            np.testing.assert_array_equal(ex["n_itm"][i, :n1], single.n_itm)
            np.testing.assert_array_equal(np.isnan(ex["boundary"][i, :n1]), np.isnan(single.boundary))
            np.testing.assert_allclose(np.nan_to_num(ex["boundary"][i, :n1]), np.nan_to_num(single.boundary), rtol=0)
            # different launch geometry (the batch entry point times both sweep kernels): the fixed-point grid sums truncate per
            # CTA, which moves the ill-conditioned individual betas in the 9th digit
            np.testing.assert_allclose(np.nan_to_num(ex["betas"][i, :n1, :3]), np.nan_to_num(single.betas), rtol=1e-7, atol=1e-10)
            Ssl = eng.paths(m, M, int(N[i]), "f64", rs)
            eu, eu_se = eng.european_from_slab(Ssl[int(N[i])].contiguous(), K[i], 0.05, T[i], ot)
            assert ex["european"][i, 0] == pytest.approx(eu, rel=1e-12)
            assert ex["european"][i, 1] == pytest.approx(eu_se, rel=1e-9)


def test_path_sharded_batch_single_rank_group(eng, mods):
    """optmc_price_american_batch_ex with M_total on a one-rank group (comm_export / comm_init with itself): the
    in-kernel cross-rank exchange path runs (tagged slot words, re-centred payloads, one slot block per option) and
    must reproduce the plain batch bit for bit; with two emulated halves of the paths priced separately the Philox
    counters line up through pair_offset.  (Two real GPUs: tests/test_multi_gpu.py.)"""
    L, E, orc = mods
    model = E.heston(100.0, 0.05, 1.0, **HP)
    e = E.Engine(0)
    try:
        e.comm_init(0, 1, [e.comm_export()])
        for M, N, B in ((40_000, 30, 3), (600_000, 25, 4)):
            streams = np.arange(B) + 3
            plain, se = eng.price_american_batch(model, M, 100.0, 100.0, 1.0, np.full(B, N), 1, "f32", E.RngSpec(seed=8),
                                                 streams=streams)
            for _ in range(2):  # twice: the exchange counter keeps running across launches
                sh, se2, ex = e.price_american_batch(model, M, 100.0, 100.0, 1.0, np.full(B, N), 1, "f32", E.RngSpec(seed=8),
                                                     streams=streams, M_total=M)
                np.testing.assert_array_equal(sh, plain)
                np.testing.assert_array_equal(se2, se)
        with pytest.raises(Exception):  # one N per sharded batch
            e.price_american_batch(model, 40_000, 100.0, 100.0, 1.0, np.array([20, 30]), 1, "f32", E.RngSpec(seed=8), M_total=40_000)
        e.comm_finalize()
    finally:
        e.close()


def test_price_only_call_skips_per_date_outputs(eng, mods):
    """optmc_price_american with a result block that carries no per-date arrays (what the reference's
    price_american_enhanced_lsm needs: it returns a float) tells the persistent sweep to skip them: same price and
    standard error bit for bit; a later optmc_lsm_fetch then reads NaN / none / 0 for every date, whereas a call that
    asks for the arrays (or an asynchronous call, whose caller fetches later) gets them."""
    L, E, orc = mods
    model = E.heston(100.0, 0.05, 1.0, **HP)
    for sem in ("reference", "textbook"):
        full = eng.price_american(model, 200_000, 40, 100.0, "put", "f32", E.RngSpec(seed=5), semantics=sem, arrays=True)
        lean = eng.price_american(model, 200_000, 40, 100.0, "put", "f32", E.RngSpec(seed=5), semantics=sem, arrays=False)
        assert lean.price == full.price and lean.stderr == full.stderr
        after = eng.lsm_fetch(40)
        assert np.isnan(after.betas).all() and int(after.ex_count.sum()) == 0 and np.isnan(after.boundary).all()
        assert full.ex_count[1:40].sum() > 0 and np.isfinite(full.betas[1:39]).any()
        eng.price_american(model, 200_000, 40, 100.0, "put", "f32", E.RngSpec(seed=5), semantics=sem, asynchronous=True)
        late = eng.lsm_fetch(40)
        np.testing.assert_array_equal(late.ex_count, full.ex_count)
        np.testing.assert_array_equal(late.betas, full.betas)
        assert late.price == full.price
