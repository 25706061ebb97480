"""GPU parity, path generation: Philox, GBM / Heston schemes against fixtures made by the real reference, QE, local volatility.

All calls go through the C ABI (ctypes).  Tolerances (north star): fed identical draws, prices / betas / boundary within
1e-5 relative in fp64 and 1e-4 in fp32 -- the fp64 assertions are far tighter; integer outputs are compared exactly.
"""
import os  # noqa: F401

import numpy as np
import pytest

from gpu_common import _dev, HP  # noqa: F401

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_philox_kat_on_device(eng):
    """Random123 known-answer vectors."""
    ctr = [[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]]
    key = [[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0xA4093822, 0x299F31D0]]
    out = eng.philox_kat(ctr, key)
    exp = np.array([[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD],
                    [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]], dtype=np.uint32)
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_philox_normal_moments(eng, mods, dtype):
    L, E, _ = mods
    z = eng.philox_normals(L.MODEL_HESTON, 1 << 20, 8, 0, dtype, E.RngSpec(seed=7)).double().cpu().numpy().ravel()
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs((z**3).mean()) < 5 * np.sqrt(15 / n)
    assert abs((z**4).mean() - 3) < 5 * np.sqrt(96 / n)
    z2 = eng.philox_normals(L.MODEL_HESTON, 1 << 20, 8, 1, dtype, E.RngSpec(seed=7)).double().cpu().numpy().ravel()
    assert abs(np.corrcoef(z, z2)[0, 1]) < 5 / np.sqrt(n)


def test_philox_paths_invariant_to_sharding(eng, mods):
    """Counters are global pair indices: generating [0,M) at once or as two shards gives identical bits."""
    L, E, _ = mods
    M, N = 4096, 9
    model = E.heston(100, 0.05, 1.0, **HP)
    full = eng.paths(model, M, N, "f32", E.RngSpec(seed=3)).clone()
    h = M // 4  # pairs per shard (M/2 pairs split in two)
    a = eng.paths(model, M // 2, N, "f32", E.RngSpec(seed=3, pair_offset=0)).clone()
    b = eng.paths(model, M // 2, N, "f32", E.RngSpec(seed=3, pair_offset=h)).clone()
    # shard columns: [+z block | -z block] locally
    assert torch.equal(a[:, :h], full[:, :h]) and torch.equal(b[:, :h], full[:, h:2 * h])
    assert torch.equal(a[:, h:], full[:, M // 2:M // 2 + h]) and torch.equal(b[:, h:], full[:, M // 2 + h:])


def test_heston_paths_vs_reference_golden(eng, mods, golden_dir):
    L, E, _ = mods
    g = np.load(os.path.join(golden_dir, "ref_heston_paths_even.npz"))
    S0, r, T, v0, kappa, theta, xi, rho = g["args"]
    M, N = int(g["M"]), int(g["N"])
    S = eng.paths(E.heston(S0, r, T, v0, kappa, theta, xi, rho), M, N, "f64",
                  E.RngSpec(z1=_dev(g["Z1"]), z2=_dev(g["Z2"])))
    np.testing.assert_allclose(S.cpu().numpy(), g["S"], rtol=1e-12)


def test_torch_fp32_variants_vs_reference_golden(eng, mods, golden_dir):
    """om3gpu:117-248 run for real on CPU torch (fixtures) vs the fp32 kernels on the same torch.randn draws."""
    L, E, _ = mods
    g = np.load(os.path.join(golden_dir, "ref_torch_paths.npz"))
    M, N = int(g["M"]), int(g["N"])
    S = eng.paths(E.gbm(100, 0.05, 1.0, 0.2), M, N, "f32", E.RngSpec(z1=_dev(g["Zh"])))
    np.testing.assert_allclose(S.cpu().numpy(), g["S_bs"], rtol=2e-5)
    S = eng.paths(E.gbm(100, 0.05, 1.0, 0.2, scheme=L.SCHEME_GBM_LOGSPACE), M, N, "f32",
                  E.RngSpec(z1=_dev(g["Zbw"]), antithetic=False))
    np.testing.assert_allclose(S.cpu().numpy(), g["S_bw"], rtol=2e-5)
    S = eng.paths(E.heston(100, 0.05, 1.0, **HP), M, N, "f32", E.RngSpec(z1=_dev(g["Z1"]), z2=_dev(g["Z2"])))
    np.testing.assert_allclose(S.cpu().numpy(), g["S_h"], rtol=5e-5)


def test_calibrator_scheme_vs_reference_golden(golden_dir):
    """hc.HestonPricer(seed=42): simulate_paths and two consecutive price_european_option calls."""
    from options_model_b200 import compat

    g = np.load(os.path.join(golden_dir, "ref_hc_paths.npz"))
    kappa, theta, sigma, rho, v0 = g["params"]
    cfg = compat.CalibrationConfig(n_mc_paths=64, n_time_steps=10, seed=42, verbose=False, plot_results=False)
    params = compat.HestonParams(kappa=kappa, theta=theta, sigma=sigma, rho=rho, v0=v0)
    S, V = compat.HestonPricer(cfg).simulate_paths(params, 100.0, 0.75, 0.03)
    np.testing.assert_allclose(S, g["S"], rtol=1e-12)
    np.testing.assert_allclose(V, g["V"], rtol=1e-10, atol=1e-18)
    pr = compat.HestonPricer(cfg, reference_draws=True)
    assert pr.price_european_option(params, 100.0, 95.0, 0.75, 0.03, "call") == pytest.approx(float(g["call_95"]), rel=1e-12)
    assert pr.price_european_option(params, 100.0, 105.0, 0.75, 0.03, "put") == pytest.approx(float(g["put_105"]), rel=1e-12)


def test_features_vs_reference_golden(golden_dir):
    from options_model_b200 import compat

    g = np.load(os.path.join(golden_dir, "ref_features.npz"))
    np.testing.assert_allclose(compat.create_regression_features(g["S"], 100.0, 0.05, 1.0, 0.3), g["F"], rtol=1e-15)
    np.testing.assert_allclose(compat.create_regression_features(g["S"], 100.0, 0.05, 1.0, 1.0), g["F_end"], rtol=1e-15)
    f = np.load(os.path.join(golden_dir, "ref_features_torch.npz"))
    F = compat.create_regression_features_torch(torch.as_tensor(f["S"]).cuda(), 100.0, 0.05, 1.0, 0.3)
    np.testing.assert_allclose(F.cpu().numpy(), f["F"], rtol=1e-6)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 2e-3)])
def test_qe_paths_vs_oracle_same_draws(eng, mods, dtype, tol):
    L, E, orc = mods
    M, N = 4096, 24
    Z1, Z2 = orc.draw_heston_normals(np.random.default_rng(21), N, M)
    model = E.heston(100.0, 0.05, 1.0, **HP, scheme=L.SCHEME_HESTON_QE)
    S, V = eng.paths(model, M, N, dtype, E.RngSpec(z1=_dev(Z1), z2=_dev(Z2)), return_v=True)
    ref, Vref = orc.heston_paths_qe(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2,
                                    return_v=True)
    Sn, Vn = S.cpu().numpy().astype(np.float64), V.cpu().numpy().astype(np.float64)
    if dtype == "f64":
        np.testing.assert_allclose(Sn, ref, rtol=tol)
        np.testing.assert_allclose(Vn, Vref, rtol=tol, atol=1e-14)
        assert (Vn >= 0).all() and (Vn == 0).any()  # the mass at zero of the exponential branch is exercised
    else:  # fp32 state: a path whose psi sits at the 1.5 switch may take the other branch; compare in bulk
        rel = np.abs(Sn[-1] - ref[-1]) / ref[-1]
        assert np.median(rel) < 1e-5 and np.mean(rel < tol) > 0.99


def test_qe_european_unbiased_at_coarse_steps(eng, mods):
    """8 steps per year: QE reproduces the semi-analytic Heston put within 3 standard errors, the reference's
    absorption Euler is off by far more (the reason the scheme exists)."""
    L, E, orc = mods
    M, N = 4_000_000, 8
    exact = orc.heston_european_analytic(100.0, 100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], "put")
    out = {}
    for name, sch in (("qe", L.SCHEME_HESTON_QE), ("absorb", L.SCHEME_HESTON_REF_ABSORB)):
        model = E.heston(100.0, 0.05, 1.0, **HP, scheme=sch)
        mean, se = eng.price_european_batch(model, M, N, [100.0], [1.0], [1], "f32", E.RngSpec(seed=3))
        out[name] = (mean[0], se[0])
    assert abs(out["qe"][0] - exact) < 3 * out["qe"][1]
    assert abs(out["absorb"][0] - exact) > 10 * out["absorb"][1]


def test_qe_american_batch_runs_through_the_sweep(eng, mods):
    L, E, orc = mods
    model = E.heston(100.0, 0.05, 1.0, **HP, scheme=L.SCHEME_HESTON_QE)
    res = eng.price_american(model, 200_000, 50, 100.0, "put", "f32", E.RngSpec(seed=9), semantics="textbook")
    eu = orc.heston_european_analytic(100.0, 100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], "put")
    assert eu - 3 * res.stderr < res.price < eu + 1.0  # early-exercise premium of an ATM 1y put: a few tenths
    pb, _ = eng.price_american_batch(model, 200_000, 100.0, [100.0], [1.0], [50], 1, "f32", E.RngSpec(seed=9), semantics="textbook")
    assert pb[0] == pytest.approx(res.price, rel=1e-6)


def _lv_net(g, tag):
    H, Lh, ms, ts, eps = g[f"{tag}_meta"]
    return dict(hidden=int(H), layers=int(Lh), weights=g[f"{tag}_weights"], m_scale=float(ms), tau_scale=float(ts), epsilon=float(eps))


@pytest.mark.parametrize("tag", ["h64", "h32"])
def test_localvol_sigma_and_paths_vs_reference_golden(eng, mods, golden_dir, tag):
    L, E, orc = mods
    g = np.load(os.path.join(golden_dir, "ref_localvol.npz"))
    net = _lv_net(g, tag)
    S0, r, T, K, M, N = g[f"{tag}_args"]
    M, N = int(M), int(N)
    # IVModel.get_volatility_batch: fp32 network, identical inputs -> fp32 rounding of a 5-layer network
    for i, tau in enumerate((1.0, 0.3, 1e-9)):
        sig = eng.ivnet_sigma(net, K, g[f"{tag}_spots"], tau).cpu().numpy()
        np.testing.assert_allclose(sig, g[f"{tag}_sigma"][i], rtol=2e-5)
    # simulate_local_vol_paths_antithetic on the reference's own draws: fp64 state, fp32 network
    S = eng.paths_localvol(S0, r, T, net, K, M, N, "f64", E.RngSpec(z1=_dev(g[f"{tag}_Zh"])))
    np.testing.assert_allclose(S.cpu().numpy(), g[f"{tag}_S"], rtol=2e-5)
    # fp32 storage (the production layout): 1e-4, the north star's fp32 tolerance
    S32 = eng.paths_localvol(S0, r, T, net, K, M, N, "f32", E.RngSpec(z1=_dev(g[f"{tag}_Zh"])))
    np.testing.assert_allclose(S32.cpu().numpy(), g[f"{tag}_S"], rtol=1e-4)


def test_localvol_compat_drop_in_and_pricing(mods, golden_dir):
    """compat.IVModel + simulate_local_vol_paths_antithetic take what the reference takes (a torch module with the
    ImprovedIVNetwork state_dict and a fitted scaler, a numpy Generator) and return the reference's paths."""
    from options_model_b200 import compat

    g = np.load(os.path.join(golden_dir, "ref_localvol.npz"))
    net = _lv_net(g, "h64")
    H, Lh = net["hidden"], net["layers"]

    class Net(torch.nn.Module):  # same parameter names / order as nniv.ImprovedIVNetwork
        def __init__(self):
            super().__init__()
            self.input_proj = torch.nn.Linear(2, H)
            self.layers = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.LayerNorm(H), torch.nn.GELU(),
                                                                   torch.nn.Identity()) for _ in range(Lh)])
            self.output = torch.nn.Linear(H, 1)

    m = Net()
    o = 0
    with torch.no_grad():
        for p in m.state_dict().values():
            n = p.numel()
            p.copy_(torch.from_numpy(net["weights"][o:o + n].reshape(tuple(p.shape))))
            o += n
    m.scaler = type("S", (), dict(m_scale=net["m_scale"], tau_scale=net["tau_scale"]))()
    m.config = type("C", (), dict(epsilon=net["epsilon"]))()
    ivm = compat.IVModel(m)
    S0, r, T, K, M, N = g["h64_args"]
    S = compat.simulate_local_vol_paths_antithetic(S0, r, T, int(M), int(N), ivm, K, np.random.default_rng(11))
    np.testing.assert_allclose(S, g["h64_S"], rtol=2e-5)
    S_odd = compat.simulate_local_vol_paths_antithetic(S0, r, T, 5, 6, ivm, K, np.random.default_rng(1))
    assert S_odd.shape == (7, 5) and np.isfinite(S_odd).all()
    with pytest.raises(ValueError):
        ivm.get_volatility_batch(K, np.array([1.0, -1.0]), 0.5)
    # torch-file variants (om3gpu:250-298, 498-519): torch.randn draws on the device, fp32 tensors out
    torch.manual_seed(3)
    St = compat.simulate_local_vol_paths_torch(S0, r, T, 1001, 6, ivm, K, torch.device("cuda"))
    torch.manual_seed(3)
    Zh = torch.randn(6, 500, device="cuda")
    ref32 = mods[2].localvol_paths_antithetic(S0, r, T, 1000, 6, net, K, Zh.double().cpu().numpy())
    assert St.shape == (7, 1001) and St.dtype == torch.float32
    np.testing.assert_allclose(St[:, :1000].cpu().numpy(), ref32, rtol=1e-4)
    sig_t = ivm.get_volatility_batch_torch(K, torch.tensor([90.0, 100.0, 110.0], device="cuda"), 0.5)
    np.testing.assert_allclose(sig_t.cpu().numpy(), ivm.get_volatility_batch(K, np.array([90.0, 100.0, 110.0]), 0.5), rtol=1e-6)
    # the pricer routes iv_model through the local-volatility paths and the same sweep (om3:461-462)
    p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=None, option_type="put", rng_manager=compat.RNGManager(42), iv_model=ivm)
    v = p.price_american_enhanced_lsm(100.0, 1.0, num_simulations=50_000, num_time_steps=25)
    assert np.isfinite(v) and 0.5 < v < 40.0


def test_localvol_philox_counters_and_throughput(eng, mods, golden_dir):
    L, E, orc = mods
    g = np.load(os.path.join(golden_dir, "ref_localvol.npz"))
    net = _lv_net(g, "h64")
    M, N = 8192, 12
    rng = E.RngSpec(seed=31)
    S = eng.paths_localvol(100.0, 0.05, 1.0, net, 105.0, M, N, "f64", rng)
    Z = eng.philox_normals(L.MODEL_GBM, M, N, 0, "f64", rng)  # the GBM counter layout: 4 steps per Philox block
    ref = orc.localvol_paths_antithetic(100.0, 0.05, 1.0, M, N, net, 105.0, Z.cpu().numpy())
    np.testing.assert_allclose(S.cpu().numpy(), ref, rtol=5e-5)


@pytest.mark.parametrize("bridge", [False, True])
def test_qmc_normals_vs_oracle(eng, mods, bridge):
    """Sobol' (Joe-Kuo) + Brownian-bridge normals against the oracle's restatement (scipy's Sobol' engine as
    the independent point set, Jaeckel's bridge): plain and digitally shifted, one and two factors, and sharded blocks
    (pair_offset) equal to the corresponding columns of the whole set."""
    L, E, orc = mods
    M, N = 4096, 50
    for factors, seed in ((1, None), (2, 5)):
        got = eng.qmc_normals(M, N, factors=factors, bridge=bridge, dtype="f64", shift_seed=seed)
        got = [got] if factors == 1 else list(got)
        shift = None
        if seed is not None:
            shift = np.random.default_rng(seed).integers(0, 2**32, size=factors * N, dtype=np.uint64).astype(np.uint32)
        ref = orc.sobol_bridge_normals(M, N, factors, bridge, shift)
        for f in range(factors):
            np.testing.assert_allclose(got[f].cpu().numpy(), ref[f], rtol=1e-9, atol=1e-9)
    whole = eng.qmc_normals(M, N, factors=1, bridge=bridge, dtype="f64", shift_seed=9).cpu().numpy()
    part = eng.qmc_normals(M // 2, N, factors=1, bridge=bridge, dtype="f64", shift_seed=9, pair_offset=M // 4).cpu().numpy()
    np.testing.assert_array_equal(part, whole[:, M // 4:])
    z = eng.qmc_normals(1 << 16, N, factors=1, bridge=bridge, dtype="f32", shift_seed=1).double().cpu().numpy()
    assert abs(z.mean()) < 1e-3 and abs(z.var() - 1) < 2e-3  # far tighter than 65536 x 50 pseudo-random draws would be


def test_qmc_brownian_bridge_reduces_the_error_of_config1(eng, mods):
    """Randomised QMC (16 digital shifts) vs Philox (16 seeds) on BASELINE config 1 (GBM put 100 k x 50, textbook
    semantics, judged against the textbook value, not the reference): European leg and American price agree within
    the spreads, and the spread over shifts is several times smaller than over pseudo-random seeds."""
    L, E, orc = mods
    M, N, K = 100_000, 50, 100.0
    gbm = E.gbm(100.0, 0.05, 1.0, 0.2)
    from options_model_b200 import compat

    bs = compat.BlackScholesGreeks.black_scholes_price(100.0, K, 1.0, 0.05, 0.2, "put")
    res = {"qmc": [], "philox": []}
    for s in range(16):
        z = eng.qmc_normals(M, N, factors=1, bridge=True, dtype="f32", shift_seed=100 + s)
        Sq = eng.paths(gbm, M, N, "f32", E.RngSpec(z1=z))
        Sp = eng.paths(gbm, M, N, "f32", E.RngSpec(seed=100 + s))
        for name, S in (("qmc", Sq), ("philox", Sp)):
            eu, _ = eng.european_from_slab(S[N].contiguous(), K, 0.05, 1.0, "put")
            am = eng.lsm(S, K, 0.05, 1.0, "put", semantics="textbook", arrays=False).price
            res[name].append((eu, am))
    q, p = np.array(res["qmc"]), np.array(res["philox"])
    assert abs(q[:, 0].mean() - bs) < 4 * q[:, 0].std() / 4 + 1e-4      # unbiased: mean over shifts vs Black-Scholes
    assert q[:, 0].std() < 0.25 * p[:, 0].std()                            # European leg: > 4x smaller spread
    assert q[:, 1].std() < 0.6 * p[:, 1].std()                             # American price: the regression noise remains
    assert abs(q[:, 1].mean() - p[:, 1].mean()) < 3 * np.hypot(q[:, 1].std(), p[:, 1].std()) / 4 + 5e-3
    assert 6.0 < q[:, 1].mean() < 6.15                                     # textbook value of config 1 (binomial 6.09)
