"""GPU tests of the drop-in layer (options_model_b200.compat): the reference's call signatures, error behaviour and European / calibration paths.

All calls go through the C ABI (ctypes).  Tolerances (north star): fed identical draws, prices / betas / boundary within
1e-5 relative in fp64 and 1e-4 in fp32 -- the fp64 assertions are far tighter; integer outputs are compared exactly.
"""
import os  # noqa: F401

import numpy as np
import pytest

from gpu_common import HP  # noqa: F401

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("tag", ["even", "odd"])
def test_compat_simulate_heston_is_drop_in(golden_dir, tag):
    """Same call as om3.simulate_heston_paths_antithetic(..., rng): identical draws, identical paths."""
    from options_model_b200 import compat

    g = np.load(os.path.join(golden_dir, f"ref_heston_paths_{tag}.npz"))
    S0, r, T, v0, kappa, theta, xi, rho = g["args"]
    S = compat.simulate_heston_paths_antithetic(S0, r, T, v0, kappa, theta, xi, rho, int(g["M"]), int(g["N"]),
                                                np.random.default_rng(int(g["seed"])))
    assert S.shape == g["S"].shape
    np.testing.assert_allclose(S, g["S"], rtol=1e-12)


def test_error_behaviour_matches_reference():
    """om3:447-452, om3:471-472: ValueError with the reference's messages; no silent CPU path."""
    from options_model_b200 import compat

    p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", use_control_variate=False)
    with pytest.raises(ValueError, match="S0, K, T must be positive"):
        p.price_american_enhanced_lsm(-1.0, 1.0, 100, 10)
    with pytest.raises(ValueError, match="positive integers"):
        p.price_american_enhanced_lsm(100.0, 1.0, 0, 10)
    with pytest.raises(ValueError, match="r must be non-negative"):
        compat.AdvancedOptionPricer(K=100.0, r=-0.01, sigma=0.2).price_american_enhanced_lsm(100.0, 1.0, 100, 10)
    with pytest.raises(ValueError, match="sigma is None"):
        compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=None).price_american_enhanced_lsm(100.0, 1.0, 100, 10)
    assert compat.compute_curve_worker_enhanced(-5.0, 100.0, 0.05, 0.2, "put", 1, 1, 2, 100, False, False, None) == []


def test_european_fused_equals_slab_reduction(eng, mods):
    """The no-store kernel consumes the same Philox counters as the path kernel: identical terminal values."""
    L, E, orc = mods
    M, N = 20000, 30
    model = E.heston(100.0, 0.05, 0.7, **HP)
    rng = E.RngSpec(seed=5, stream=11)
    mean, se = eng.price_european_batch(model, M, N, [100.0], [0.7], [1], "f64", rng, stream_id=[0])
    S = eng.paths(model, M, N, "f64", rng)
    m2, s2 = eng.european_from_slab(S[N].contiguous(), 100.0, 0.05, 0.7, "put")
    assert mean[0] == pytest.approx(m2, rel=1e-12) and se[0] == pytest.approx(s2, rel=1e-10)
    ref = orc.european_from_paths(S[N].cpu().numpy(), 100.0, 0.05, 0.7, "put")
    assert m2 == pytest.approx(ref[0], rel=1e-12) and s2 == pytest.approx(ref[1], rel=1e-10)


def test_european_batch_calibration_grid(eng, mods):
    """Config-5 shape at reduced size: per-option K, T; options with equal stream ids share their paths."""
    L, E, orc = mods
    K = np.array([90.0, 100.0, 110.0, 100.0])
    T = np.array([0.5, 0.5, 0.5, 1.0])
    model = E.heston(100.0, 0.05, 1.0, **HP, scheme=L.SCHEME_HESTON_REF_CALIB)
    mean, se = eng.price_european_batch(model, 50_000, 50, K, T, [0, 0, 0, 0], "f32", E.RngSpec(seed=1),
                                        stream_id=[0, 0, 0, 1])
    assert mean[0] > mean[1] > mean[2] > 0  # same paths, decreasing in strike: strictly monotone
    # against the calibrator scheme on numpy draws (independent RNG): 4 standard errors
    Z1, Z2i = orc.hc_draw_normals(np.random.default_rng(0), 50_000, 50)
    Sh, _ = orc.hc_simulate_paths(HP["kappa"], HP["theta"], HP["xi"], HP["rho"], HP["v0"], 100.0, 0.5, 0.05, 50_000, 50, Z1, Z2i)
    for i in range(3):
        ref = orc.hc_price_european(Sh[:, -1], K[i], 0.5, 0.05, "call")
        assert abs(mean[i] - ref) < 4 * np.hypot(se[i], se[i])


def test_compat_pricer_end_to_end(mods):
    """AdvancedOptionPricer through the compat layer: control-variate route, European route, curve driver."""
    from options_model_b200 import compat

    L, E, orc = mods
    p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(42),
                                    use_control_variate=False)
    v = p.price_american_enhanced_lsm(100.0, 1.0, 100_000, 50)
    assert abs(v - 6.5424) < 4 * 0.0255 * np.sqrt(2)  # reference-semantics value of config 1
    bs = compat.BlackScholesGreeks.black_scholes_price(100.0, 100.0, 1.0, 0.05, 0.2, "put")
    assert bs == pytest.approx(5.573526, abs=1e-6)
    eu = p.price_european_streaming(100.0, 1.0, 200_000, 50)
    assert abs(eu - bs) < 0.06
    pc = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(1))
    rec = pc.compute_curve_for_S0(100.0, 1, 3, 20000, False)
    assert [r["Days to Expiry"] for r in rec] == [3.0, 2.0, 1.0] and all(r["Option Value"] > 0 for r in rec)


def test_compat_curve_uses_batch(mods):
    from options_model_b200 import compat

    kw = dict(K=100.0, r=0.05, sigma=None, option_type="put", use_heston=True, heston_params=HP,
              use_control_variate=False)
    a = compat.AdvancedOptionPricer(rng_manager=compat.RNGManager(9), **kw).compute_curve_for_S0(100.0, 1, 12, 20_000, False)
    b = compat.AdvancedOptionPricer(rng_manager=compat.RNGManager(9), batched=False, **kw).compute_curve_for_S0(
        100.0, 1, 12, 20_000, False)
    assert [r["Days to Expiry"] for r in a] == [r["Days to Expiry"] for r in b]
    va = np.array([r["Option Value"] for r in a]); vb = np.array([r["Option Value"] for r in b])
    assert np.all(np.abs(va - vb) < 0.25)  # different Philox streams, same distribution
    assert np.all(va > 0)


def test_compat_om2_option_pricer_nn_and_poly(mods):
    """om2.OptionPricer.price_american_option: the reference's per-date network by default, polynomial on request."""
    from options_model_b200 import compat

    L, E, orc = mods
    nn = compat.OptionPricer(100.0, 0.05, 0.2, "put", seed=3, nn_hidden=32, nn_epochs=10).price_american_option(100.0, 1.0, 20_000, 20)
    nn128 = compat.OptionPricer(100.0, 0.05, 0.2, "put", seed=3, nn_hidden=128, nn_epochs=10).price_american_option(100.0, 1.0, 20_000, 20)
    poly = compat.OptionPricer(100.0, 0.05, 0.2, "put", seed=3, regressor="poly").price_american_option(100.0, 1.0, 20_000, 20)
    bs = compat.BlackScholesGreeks.black_scholes_price(100.0, 100.0, 1.0, 0.05, 0.2, "put")
    for v in (nn, nn128, poly):
        assert np.isfinite(v) and bs - 0.5 < v < bs + 2.5
    with pytest.raises(NotImplementedError):
        compat.OptionPricer(100.0, 0.05, 0.2, "put", nn_hidden=64).price_american_option(100.0, 1.0, 1000, 5)
    assert compat.compute_curve_worker(-1.0, 100.0, 0.05, 0.2, "put", 2, 1, 1, 2, 100, False, False, None) == []


def test_compat_pricer_nn_regressor(mods):
    from options_model_b200 import compat

    p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(42),
                                    lsm_regressor="nn", nn_epochs=3)
    v = p.price_american_enhanced_lsm_gpu(100.0, 1.0, num_simulations=50_000, num_time_steps=30)
    assert 5.5 < v < 8.5 and p.last_result["epochs_run"] >= 1
    with pytest.raises(ValueError):
        compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, lsm_regressor="nn", nn_hidden=64).price_american_enhanced_lsm(100.0, 1.0, 1000, 10)
    # a fresh SingleLSMNet per exercise date behind the same signature (BASELINE config 3)
    q = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(42),
                                    lsm_regressor="nn_per_date", nn_epochs=3)
    w = q.price_american_enhanced_lsm_gpu(100.0, 1.0, num_simulations=50_000, num_time_steps=10)
    assert 5.5 < w < 8.5 and q.last_result["n_rows"] > 20_000 and q.last_result["epochs_run"] >= 9


def test_compat_out_of_sample_flag(mods):
    from options_model_b200 import compat

    kw = dict(K=100.0, r=0.05, sigma=0.2, option_type="put", semantics="textbook")
    a = compat.AdvancedOptionPricer(rng_manager=compat.RNGManager(7), **kw).price_american_enhanced_lsm(100.0, 1.0, 200_000, 50)
    b = compat.AdvancedOptionPricer(rng_manager=compat.RNGManager(7), out_of_sample=True, **kw).price_american_enhanced_lsm(100.0, 1.0, 200_000, 50)
    # American put, GBM: binomial value 6.09; in-sample slightly above the out-of-sample (low-biased) estimate, both close
    assert 5.95 < b < 6.2 and 5.95 < a < 6.25 and abs(a - b) < 0.12


def test_calibration_objective_is_one_launch_and_matches_row_by_row(mods):
    """hc:404-472 on the reference's synthetic smile (hc:751-756): the batched objective equals the row-by-row
    restatement on the same prices, common random numbers make it deterministic, invalid parameters give 1e6."""
    from options_model_b200 import compat

    S0, r = 100.0, 0.05
    Kg, Tg = np.meshgrid(np.linspace(80, 120, 5), np.linspace(0.25, 1.0, 3))
    K, T = Kg.ravel(), Tg.ravel()
    iv = 0.2 + 0.1 * np.abs(np.log(K / S0)) + 0.02 * np.sqrt(T)
    cfg = compat.CalibrationConfig(n_mc_paths=40_000, n_time_steps=50, verbose=False)
    f = compat.HestonObjective(compat.HestonPricer(cfg), S0, r, K, T, iv, common_random_numbers=True)
    x = np.array([2.0, 0.04, 0.5, -0.7, 0.04])
    eng = compat._engine()
    n0 = eng.launch_count()
    v = f(x)
    assert eng.launch_count() - n0 == 1  # the whole surface in one kernel launch
    assert 0.0 < v < 1e5 and v == f(x)   # common random numbers: bit-identical re-evaluation
    # row-by-row restatement of the reference loop on the same model prices
    tot = wsum = 0.0
    for p, k, t, s in zip(f.last_prices, K, T, iv):
        bs = compat.bs_price(S0, k, t, r, s, "call")
        w = max(compat.bs_vega(S0, k, t, r, s) / 100.0, cfg.min_vega_weight)
        tot += w * np.log(p / bs) ** 2
        wsum += w
    feller = 100.0 * abs(2 * 2.0 * 0.04 - 0.5**2)  # violated for these parameters
    assert v == pytest.approx(np.sqrt(tot / wsum) + feller, rel=1e-12)
    # each price within 4 SE of the calibrator scheme priced option by option
    single = compat.HestonPricer(cfg).price_european_option(compat.HestonParams.from_array(x), S0, K[7], T[7], r, "call")
    assert abs(single - f.last_prices[7]) < 0.25
    assert f(np.array([-1.0, 0.04, 0.5, -0.7, 0.04])) == 1e6
    # a fresh-noise objective (the reference's behaviour) differs from call to call
    g = compat.HestonObjective(compat.HestonPricer(cfg), S0, r, K, T, iv)
    assert g(x) != g(x)


def test_compat_control_variate_on_the_same_paths(mods):
    """the reference's control variate uses an independent European simulation (om3:653-677); the flag
    puts the European leg on the American paths (one slab, one sweep, one reduction of the terminal row).  Both are
    unbiased estimates of the same quantity."""
    from options_model_b200 import compat

    kw = dict(K=100.0, r=0.05, sigma=0.2, option_type="put", semantics="textbook")
    ref, same = [], []
    for sd in range(8):
        ref.append(compat.AdvancedOptionPricer(rng_manager=compat.RNGManager(sd), **kw)
                   .price_american_with_control_variate(100.0, 1.0, 20_000, 25))
        same.append(compat.AdvancedOptionPricer(rng_manager=compat.RNGManager(sd), control_variate_same_paths=True, **kw)
                    .price_american_with_control_variate(100.0, 1.0, 20_000, 25))
    assert abs(np.mean(ref) - np.mean(same)) < 0.08 and 5.9 < np.mean(same) < 6.25
    assert np.std(same) < 1.5 * np.std(ref)


def test_compat_om1_api_mean_std_zero_prob(mods):
    """Options_model.py (om1:44-211): (mean, std, P(worthless)) on the reference's own numpy draws.  The paths are the
    reference's bit for bit (np.random.seed(seed) stream), so the European-style statistics that do not depend on the
    network are checked exactly against numpy; the price within the usual network tolerance."""
    from options_model_b200 import compat

    L, E, orc = mods
    S0, K, T, r, sigma, M, N, seed = 100.0, 100.0, 0.5, 0.05, 0.25, 20_000, 20, 42
    mean, std, zp = compat.om1.price_american_option(S0, K, T, r, sigma, M, N, "put", 2, False, seed)
    Z = np.random.RandomState(seed).standard_normal((N, M // 2))
    S = orc.gbm_paths_antithetic(S0, r, sigma, T, M, N, Z)
    # a path expires worthless iff it is never exercised and ends out of the money: at least every path that is never in
    # the money does; at most every path that ends out of the money
    never_itm = np.mean((S[1:] >= K).all(axis=0))
    ends_otm = np.mean(S[-1] >= K)
    assert never_itm <= zp <= ends_otm
    bs = compat.BlackScholesGreeks.black_scholes_price(S0, K, T, r, sigma, "put")
    assert bs - 0.3 < mean < bs + 1.5 and 0.5 * mean < std < 3.0 * mean
    rec = compat.om1.compute_curve_for_S0(S0, K, r, sigma, 4000, 1, 2, "put", 2, False, seed)
    assert [x["Days to Expiry"] for x in rec] == [2.0, 1.0] and set(rec[0]) == {"S0", "Days to Expiry", "Option Value", "Std Dev", "Zero Prob"}
    with pytest.raises(ValueError):
        compat.om1.price_american_option(S0, K, T, r, -1.0)


def test_compat_batched_curve_equals_per_point_loop(mods):
    """compute_curve_for_S0(batched=True) and the per-point loop draw the same Philox streams (key = master seed,
    stream = child seed): identical prices point by point, without and with the control variate."""
    from options_model_b200 import compat

    for cv, same in ((False, False), (True, False), (True, True)):
        out = []
        for batched in (True, False):
            p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(7),
                                            use_control_variate=cv, control_variate_same_paths=same, batched=batched)
            rec = p.compute_curve_for_S0(100.0, 1, 12, 4000, False)
            out.append(np.array([r["Option Value"] for r in rec]))
            state = p.rng_manager.master_rng.integers(0, 2**31 - 1)  # both routes advance the generators alike
            out.append(state)
        np.testing.assert_allclose(out[0], out[2], rtol=1e-6, atol=1e-9)  # fp32 slabs: grouped vs single launch geometry
        assert out[1] == out[3]


def test_compat_curve_control_variate_is_batched_and_reduces_variance(mods):
    """The reference's default route (om3:692-693) for a whole curve: <= 3 launches per grid wave, and the same-path
    European leg removes variance the independent leg adds (spread of the CV-adjusted price over seeds)."""
    from options_model_b200 import compat

    eng = compat._engine(0)
    spread = {}
    for same in (False, True):
        vals = []
        for seed in range(12):
            p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(seed),
                                            use_control_variate=True, control_variate_same_paths=same, semantics="textbook")
            l0 = eng.launch_count()
            rec = p.compute_curve_for_S0(100.0, 1, 30, 6000, False)
            launches = eng.launch_count() - l0
            assert launches <= 3 * 2, launches  # reset + paths + grouped sweep (+ one fused European launch), per wave
            vals.append([r["Option Value"] for r in rec])
        spread[same] = np.std(np.array(vals), axis=0).mean()  # spread over seeds, averaged over the 30 curve points
    assert spread[True] < 0.8 * spread[False], spread


def test_compat_iv_model_with_control_variate(mods, golden_dir):
    """an iv_model pricer with the constructor defaults (use_control_variate=True, sigma given) prices the
    American leg on local-volatility paths and the European control leg through the same network; the curve driver
    falls back to the per-point loop."""
    import os

    from options_model_b200 import compat

    g = np.load(os.path.join(golden_dir, "ref_localvol.npz"))

    # the flattened ImprovedIVNetwork of the fixture, as compat.IVModel stores it
    H, Lh, ms, ts, eps = g["h32_meta"]
    ivm = compat.IVModel.__new__(compat.IVModel)
    ivm.net = dict(hidden=int(H), layers=int(Lh), weights=g["h32_weights"].astype(np.float32), m_scale=float(ms),
                   tau_scale=float(ts), epsilon=float(eps))
    ivm.m_scale, ivm.tau_scale = float(ms), float(ts)
    p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(3),
                                    iv_model=ivm)
    v = p.price_american_option(100.0, 0.5, 4000, 20)
    assert np.isfinite(v) and v > 0
    e = p.price_european_streaming(100.0, 0.5, 4000, 20)
    assert np.isfinite(e) and e > 0
    rec = p.compute_curve_for_S0(100.0, 1, 3, 2000, False)
    assert len(rec) == 3 and all(np.isfinite(r["Option Value"]) for r in rec)


def test_compat_qmc_flag(mods):
    """AdvancedOptionPricer(qmc=True): Sobol' + Brownian-bridge draws behind the reference's call signature (GBM and
    Heston); same estimator, tighter: the spread over seeds is smaller than with pseudo-random draws."""
    from options_model_b200 import compat

    vals = {False: [], True: []}
    for q in (False, True):
        for sd in range(8):
            p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(sd),
                                            use_control_variate=False, semantics="textbook", qmc=q)
            vals[q].append(p.price_american_option(100.0, 1.0, 50_000, 50))
    assert abs(np.mean(vals[True]) - np.mean(vals[False])) < 0.05 and np.std(vals[True]) < np.std(vals[False])
    ph = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=None, option_type="put", rng_manager=compat.RNGManager(1),
                                     use_heston=True, heston_params=HP, use_control_variate=False, qmc=True)
    v = ph.price_american_option(100.0, 1.0, 50_000, 50)
    assert 5.0 < v < 7.0
