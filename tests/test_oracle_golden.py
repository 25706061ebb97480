"""CPU: the numpy oracle against fixtures produced by the REAL reference (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import lsm_oracle as orc

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_rng_seed_tree(golden_meta):
    g = golden_meta["rng"]
    m = orc.RNGManager(g["master_seed"])
    assert [int(m.get_child_seed()) for _ in range(4)] == g["child_seeds"]
    m = orc.RNGManager(g["master_seed"])
    assert m.get_child_rng().standard_normal(4).tolist() == g["child0_normals"]


@pytest.mark.parametrize("tag", ["even", "odd"])
def test_heston_paths_bitwise(golden_dir, tag):
    g = _load(golden_dir, f"ref_heston_paths_{tag}.npz")
    S0, r, T, v0, kappa, theta, xi, rho = g["args"]
    M, N = int(g["M"]), int(g["N"])
    Me = M // 2 * 2
    S = orc.heston_paths_antithetic(S0, r, T, v0, kappa, theta, xi, rho, Me, N, g["Z1"], g["Z2"])
    assert np.array_equal(S, g["S"][:, :Me])  # same numpy ops in the same order: bit-exact


def test_features_bitwise(golden_dir):
    g = _load(golden_dir, "ref_features.npz")
    assert np.array_equal(orc.features_ref7(g["S"], 100.0, 0.05, 1.0, 0.3), g["F"])
    assert np.array_equal(orc.features_ref7(g["S"], 100.0, 0.05, 1.0, 1.0), g["F_end"])


def test_welford(golden_dir):
    g = _load(golden_dir, "ref_welford.npz")
    flat, sizes = g["flat"], g["sizes"]
    st = (0.0, 0.0, 0)
    off = 0
    for i, n in enumerate(sizes):
        st = orc.welford_batch_update(*st, flat[off:off + n])
        off += n
        assert [float(st[0]), float(st[1]), int(st[2])] == g["states"][i].tolist()
    # closed form cross-check
    assert st[0] == pytest.approx(flat.mean(), rel=1e-14)
    assert st[1] == pytest.approx(((flat - flat.mean()) ** 2).sum(), rel=1e-13)


def test_european_streaming(golden_meta):
    g = golden_meta["european_streaming"]
    cases = {"gbm_put": dict(sigma=0.2, option_type="put"), "gbm_call": dict(sigma=0.25, option_type="call"),
             "heston_put": dict(sigma=None, option_type="put", heston_params=HP)}
    for name, kw in cases.items():
        mean, se, n = orc.price_european_streaming(g["K"], g["r"], kw["sigma"], kw["option_type"],
                                                   orc.RNGManager(g["master_seed"]), g["S0"], g["T"], g["n"],
                                                   g["steps"], g["chunk"], kw.get("heston_params"))
        assert n == g["n"]
        assert mean == pytest.approx(g["prices"][name], rel=1e-14), name


def _global_price(g, name):
    W = np.array(g["W"])
    B = g["B"]

    def fit(Xn, Ys):
        return lambda fn: fn.astype(np.float32).astype(np.float64) @ W + B

    mgr = orc.RNGManager(g["master_seed"])
    rng = mgr.get_child_rng()
    mgr.get_child_seed()
    M, N = g["M"], g["N"]
    if name.startswith("heston"):
        Z1, Z2 = orc.draw_heston_normals(rng, N, M)
        S = orc.heston_paths_antithetic(g["S0"], g["r"], g["T"], HP["v0"], HP["kappa"], HP["theta"], HP["xi"],
                                        HP["rho"], M, N, Z1, Z2)
    else:
        S = orc.gbm_paths_antithetic(g["S0"], g["r"], g["sigma"][name], g["T"], M, N, orc.draw_gbm_normals(rng, N, M))
    return orc.lsm_global(S, g["K"], g["r"], g["T"], "call" if name.endswith("call") else "put", fit)[0]


@pytest.mark.parametrize("name", ["gbm_put", "gbm_call", "heston_put"])
def test_lsm_skeleton_global_matches_real_reference(golden_meta, name):
    """om3:439-651 run for real with a deterministic stand-in network vs the restated loop."""
    g = golden_meta["lsm_skeleton_global"]
    assert _global_price(g, name) == pytest.approx(g["prices"][name], rel=1e-12)


@pytest.mark.parametrize("name", ["gbm_put", "heston_put"])
def test_lsm_skeleton_perdate_matches_real_reference(golden_meta, name):
    """om2:216-334 run for real with a deterministic stand-in ContNet vs lsm_sweep + custom regressor."""
    g = golden_meta["lsm_skeleton_perdate"]
    A0, A1, A2 = g["A"]
    M, N = g["M"], g["N"]
    np.random.seed(g["seed"])
    if name.startswith("heston"):
        Z1 = np.empty((N, M)); Z2 = np.empty((N, M))   # om2:150-170 is NOT antithetic
        for t in range(N):
            Z1[t] = np.random.standard_normal(M)
            Z2[t] = np.random.standard_normal(M)
        dt = g["T"] / N
        S = np.zeros((N + 1, M)); v = np.zeros((N + 1, M)); S[0] = g["S0"]; v[0] = HP["v0"]
        for t in range(1, N + 1):
            w2 = HP["rho"] * Z1[t - 1] + np.sqrt(1 - HP["rho"] ** 2) * Z2[t - 1]
            vp = np.maximum(v[t - 1], 0)
            v[t] = np.maximum(vp + HP["kappa"] * (HP["theta"] - vp) * dt + HP["xi"] * np.sqrt(vp * dt) * w2, 0)
            S[t] = S[t - 1] * np.exp((g["r"] - 0.5 * vp) * dt + np.sqrt(vp * dt) * Z1[t - 1])
    else:
        S = orc.gbm_paths_antithetic(g["S0"], g["r"], 0.2, g["T"], M, N, np.random.standard_normal((N, M // 2)))

    def regressor(t, t_cur, X, Y):
        Xs = (X - X.mean()) / X.std() if X.std() > 0 else X - X.mean()   # om2:289
        xd = Xs.astype(np.float32).astype(np.float64)                      # om2:295 .float()
        return A0 + A1 * xd + A2 * xd * xd, None

    res = orc.lsm_sweep(S, g["K"], g["r"], g["T"], "put", regressor=regressor)
    assert res.price == pytest.approx(g["prices"][name], rel=1e-12)


def test_hc_scheme(golden_dir, golden_meta):
    g = _load(golden_dir, "ref_hc_paths.npz")
    kappa, theta, sigma, rho, v0 = g["params"]
    S, V = orc.hc_simulate_paths(kappa, theta, sigma, rho, v0, float(g["S0"]), float(g["T"]), float(g["r"]), 64, 10,
                                 g["Z1"], g["Z2i"])
    assert np.array_equal(S, g["S"]) and np.array_equal(V, g["V"])
    # persistent generator: first call prices the call, the second continues the stream (hc:202)
    rng = np.random.default_rng(42)
    Z1, Z2i = orc.hc_draw_normals(rng, 64, 10)
    S1, _ = orc.hc_simulate_paths(kappa, theta, sigma, rho, v0, 100.0, 0.75, 0.03, 64, 10, Z1, Z2i)
    assert orc.hc_price_european(S1[:, -1], 95.0, 0.75, 0.03, "call") == pytest.approx(float(g["call_95"]), rel=1e-14)
    Z1, Z2i = orc.hc_draw_normals(rng, 64, 10)
    S2, _ = orc.hc_simulate_paths(kappa, theta, sigma, rho, v0, 100.0, 0.75, 0.03, 64, 10, Z1, Z2i)
    assert orc.hc_price_european(S2[:, -1], 105.0, 0.75, 0.03, "put") == pytest.approx(float(g["put_105"]), rel=1e-14)


def test_hc_full_size_matches_real_reference(golden_meta):
    """hc.HestonPricer(seed=42).price_european_option at 50k x 100 (config-5 unit of work), run for real."""
    Z1, Z2i = orc.hc_draw_normals(np.random.default_rng(42), 50_000, 100)
    S, _ = orc.hc_simulate_paths(2.0, 0.04, 0.5, -0.7, 0.04, 100.0, 1.0, 0.05, 50_000, 100, Z1, Z2i)
    assert orc.hc_price_european(S[:, -1], 100.0, 1.0, 0.05, "call") == pytest.approx(
        golden_meta["hc_call_50k_x100"], rel=1e-13)


def test_torch_fp32_paths(golden_dir):
    g = _load(golden_dir, "ref_torch_paths.npz")
    M, N = int(g["M"]), int(g["N"])
    np.testing.assert_allclose(orc.bs_paths_fp32(100.0, 0.05, 1.0, 0.2, M, N, g["Zh"]), g["S_bs"], rtol=3e-6)
    np.testing.assert_allclose(orc.bs_paths_logspace_fp32(100.0, 0.05, 1.0, 0.2, M, N, g["Zbw"]), g["S_bw"], rtol=3e-6)
    np.testing.assert_allclose(orc.heston_paths_fp32(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"],
                                                     HP["rho"], M, N, g["Z1"], g["Z2"]), g["S_h"], rtol=5e-6)
    f = _load(golden_dir, "ref_features_torch.npz")
    np.testing.assert_allclose(orc.features_ref7(f["S"].astype(np.float64), 100.0, 0.05, 1.0, 0.3), f["F"], rtol=1e-6)


def test_poly_restatement_pins(golden_meta):
    """SURVEY.md 8(c) survey-time pins, regenerated by the oracle (C1, 100k x 50)."""
    p = golden_meta["poly_pins"]["c1_gbm_put_100k_50"]
    assert p["S_1_0"] == pytest.approx(98.25145895, abs=1e-8)
    assert p["S_50_0"] == pytest.approx(86.3392302179, abs=1e-9)
    assert p["reference"]["price"] == pytest.approx(6.542437, abs=1e-6)
    assert p["reference"]["stderr"] == pytest.approx(0.02548, abs=1e-5)
    assert p["reference"]["boundary_25"] == pytest.approx(99.8411, abs=1e-4)
    assert p["textbook"]["price"] == pytest.approx(6.042107, abs=1e-6)
    assert p["textbook"]["boundary_45"] == pytest.approx(88.0067, abs=1e-4)
    h = golden_meta["poly_pins"]["heston_put_100k_50"]
    assert h["S_50_0"] == pytest.approx(104.4809505622, abs=1e-9)
    assert h["european"] == pytest.approx(5.36626, abs=1e-5)
    assert h["reference"] == pytest.approx(6.470907, abs=1e-6)
    assert h["textbook"] == pytest.approx(5.765446, abs=1e-6)


def test_poly_small_case_regenerates(golden_dir):
    g = _load(golden_dir, "oracle_heston_poly2_small.npz")
    res, S, _ = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 4096, 20, orc.RNGManager(1),
                                       heston_params=HP, return_paths=True)
    assert res.price == pytest.approx(float(g["price"]), rel=1e-13)
    np.testing.assert_allclose(res.betas, g["betas"], rtol=1e-9, equal_nan=True)
    np.testing.assert_array_equal(res.ex_count, g["ex_count"])
    np.testing.assert_allclose(res.boundary, g["boundary"], rtol=0, atol=0, equal_nan=True)


def test_cholesky_guard_vs_lstsq():
    rng = np.random.default_rng(0)
    x = rng.uniform(0.5, 1.0, 5000)
    Phi = np.column_stack([np.ones_like(x), x, x * x])
    y = rng.standard_normal(5000)
    beta = orc.cholesky_solve_guarded(Phi.T @ Phi, Phi.T @ y)
    ref = np.linalg.lstsq(Phi, y, rcond=None)[0]
    np.testing.assert_allclose(Phi @ beta, Phi @ ref, atol=1e-9)
    # degenerate: all x identical -> rank 1 -> None
    Phi = np.column_stack([np.ones(10), np.full(10, 0.9), np.full(10, 0.81)])
    assert orc.cholesky_solve_guarded(Phi.T @ Phi, Phi.T @ np.ones(10)) is None


def test_textbook_vs_reference_semantics_differ():
    res, S, _ = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 20000, 25, orc.RNGManager(3), sigma=0.2,
                                       return_paths=True)
    tb = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", semantics="textbook")
    assert res.price > tb.price + 0.2  # look-ahead bias of the sticky mask (App. A, Q1)
    assert 5.9 < tb.price < 6.25       # binomial value ~6.09


def test_single_lsm_net_fit_runs_the_reference_training_loop():
    """The torch restatement of om3:565-613 / om3gpu:740-798 used as the checker of optmc_lsm_gnet: tiny problem,
    both variants; the fit must reduce the loss and produce a price between the European value and the strike."""
    torch = pytest.importorskip("torch")  # noqa: F841
    rng = np.random.default_rng(3)
    M, N = 2000, 8
    Z = orc.draw_gbm_normals(rng, N, M)
    S = orc.gbm_paths_antithetic(100.0, 0.05, 0.2, 1.0, M, N, Z)
    for variant in ("cpu", "gpu"):
        log = []
        price, st = orc.lsm_global(S, 100.0, 0.05, 1.0, "put", orc.single_lsm_net_fit(variant, epochs=4, seed=1, log=log),
                                   target_ddof=0 if variant == "cpu" else 1)
        assert st["n_rows"] > 0 and len(log) >= 1 and log[-1] <= log[0] + 1e-9
        assert 4.0 < price < 12.0


@pytest.mark.parametrize("tag", ["h64", "h32"])
def test_localvol_oracle_vs_reference_golden(golden_dir, tag):
    """The numpy restatement of IVModel.get_volatility_batch + simulate_local_vol_paths_antithetic (om3:263-333) against
    vectors produced by the real reference (oracle/gen_golden_localvol.py)."""
    g = np.load(os.path.join(golden_dir, "ref_localvol.npz"))
    H, Lh, ms, ts, eps = g[f"{tag}_meta"]
    net = dict(hidden=int(H), layers=int(Lh), weights=g[f"{tag}_weights"], m_scale=float(ms), tau_scale=float(ts), epsilon=float(eps))
    S0, r, T, K, M, N = g[f"{tag}_args"]
    for i, tau in enumerate((1.0, 0.3, 1e-9)):
        np.testing.assert_allclose(orc.ivnet_sigma(net, K, g[f"{tag}_spots"], tau), g[f"{tag}_sigma"][i], rtol=1e-5)
    S = orc.localvol_paths_antithetic(S0, r, T, int(M), int(N), net, K, g[f"{tag}_Zh"])
    np.testing.assert_allclose(S, g[f"{tag}_S"], rtol=1e-5)


def test_global_network_lsm_oracle_reproduces_the_real_reference(golden_dir):
    """oracle.price_american_enhanced_lsm_nn against prices produced by the REAL AdvancedOptionPricer (om3:439-651, network
    and all; oracle/gen_golden_gnet.py).  Same numpy draws, torch's global generator consumed in the same order: equal to
    the last bit on the torch version that made the fixture, to training-noise level otherwise."""
    import json

    torch = pytest.importorskip("torch")
    g = json.load(open(os.path.join(golden_dir, "ref_gnet_prices.json")))
    same = torch.__version__ == g["torch"] and np.__version__ == g["numpy"]
    hp = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
    for c in g["cases"]:
        price, stats = orc.price_american_enhanced_lsm_nn(c["S0"], c["K"], c["r"], c["T"], c["option_type"], c["M"], c["N"],
                                                          c["master_seed"], sigma=c["sigma"], heston_params=hp if c["heston"] else None,
                                                          nn_hidden=c["nn_hidden"], nn_epochs=c["nn_epochs"], nn_lr=c["nn_lr"])
        assert price == pytest.approx(c["reference_price"], rel=1e-12 if same else 5e-2), c["name"]
        assert stats["n_rows"] > 0


def test_per_date_network_lsm_oracle_reproduces_the_real_reference(golden_dir):
    """oracle.price_american_om2_nn against prices produced by the REAL om2.OptionPricer.price_american_option (om2:216-330:
    a default-initialised ContNet per date, 10 full-batch Adam steps)."""
    import json

    torch = pytest.importorskip("torch")
    g = json.load(open(os.path.join(golden_dir, "ref_gnet_prices.json")))
    same = torch.__version__ == g["torch"] and np.__version__ == g["numpy"]
    for c in g["om2_cases"]:
        mean, std, zp = orc.price_american_om2_nn(c["S0"], c["K"], c["r"], c["sigma"], c["T"], c["option_type"], c["M"], c["N"],
                                                  c["seed"], c["nn_hidden"], c["nn_epochs"], 1e-3)
        assert mean == pytest.approx(c["reference_price"], rel=1e-12 if same else 5e-2)
        assert std > 0 and 0 <= zp <= 1


def test_calibration_objective_vs_the_real_reference(golden_dir):
    """hc:404-472: the oracle's row-by-row restatement and the product's host-side objective (compat.objective_from_prices,
    fed the reference's own per-row prices) against values produced by the REAL HestonCalibrator._objective_function."""
    import json

    import options_model_b200  # noqa: F401
    from options_model_b200 import compat

    g = json.load(open(os.path.join(golden_dir, "ref_gnet_prices.json")))
    for c in g["hc_objective"]:
        K, T, iv = np.array(c["K"]), np.array(c["T"]), np.array(c["sigma_iv"])
        val, prices = orc.hc_objective(c["x"], c["S0"], c["r"], K, T, iv, c["n_mc_paths"], c["n_time_steps"], c["seed"])
        assert val == pytest.approx(c["reference_value"], rel=1e-12)
        np.testing.assert_allclose(prices, c["prices"], rtol=1e-12)
        mine = compat.objective_from_prices(np.array(c["x"]), np.array(c["prices"]), c["S0"], c["r"], K, T, iv)
        assert mine == pytest.approx(c["reference_value"], rel=1e-10)
