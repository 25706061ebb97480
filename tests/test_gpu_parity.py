"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against (a) fixtures produced by the
real reference (tests/golden, see oracle/gen_golden.py) and (b) the numpy oracle on identical Gaussian draws.

Tolerances (north star): fed identical draws, prices / betas / boundary within 1e-5 relative in fp64 and
1e-4 in fp32 -- the fp64 assertions below are far tighter (1e-9 or better); integer outputs (exercise
counts, ITM counts) and the boundary (a selected input value) are compared exactly.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


@pytest.fixture(scope="module")
def eng():
    from options_model_b200 import engine as E

    e = E.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def mods():
    from options_model_b200 import _lib as L
    from options_model_b200 import engine as E
    from oracle import lsm_oracle as orc

    return L, E, orc


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _slab(eng, S_np, dtype):
    """Upload a host [(N+1), M] array into a padded step-major slab; returns the [:, :M] view."""
    N1, M = S_np.shape
    slab = eng.alloc_slab(M, N1 - 1, "f64" if dtype == torch.float64 else "f32")
    slab[:, :M] = _dev(S_np, dtype)
    return slab[:, :M]


# ------------------------------------------------------------------------------------------------------
# RNG
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# path schemes vs the REAL reference (golden fixtures)
# ------------------------------------------------------------------------------------------------------
def test_heston_paths_vs_reference_golden(eng, mods, golden_dir):
    L, E, _ = mods
    g = np.load(os.path.join(golden_dir, "ref_heston_paths_even.npz"))
    S0, r, T, v0, kappa, theta, xi, rho = g["args"]
    M, N = int(g["M"]), int(g["N"])
    S = eng.paths(E.heston(S0, r, T, v0, kappa, theta, xi, rho), M, N, "f64",
                  E.RngSpec(z1=_dev(g["Z1"]), z2=_dev(g["Z2"])))
    np.testing.assert_allclose(S.cpu().numpy(), g["S"], rtol=1e-12)


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


# ------------------------------------------------------------------------------------------------------
# LSM sweep vs the oracle on identical paths
# ------------------------------------------------------------------------------------------------------
def _check_sweep(res, ref, price_rtol, beta_rtol=1e-6, boundary_rtol=0.0):
    assert res.price == pytest.approx(ref.price, rel=price_rtol)
    # one-pass variance: absolute floor for the degenerate "all cash-flows equal" case
    assert res.stderr == pytest.approx(ref.stderr, rel=max(price_rtol, 1e-9), abs=1e-7 * max(1.0, abs(ref.price)))
    np.testing.assert_array_equal(res.n_itm, ref.n_itm)
    np.testing.assert_array_equal(res.ex_count, ref.ex_count)
    np.testing.assert_array_equal(np.isnan(res.boundary), np.isnan(ref.boundary))
    # the boundary is one of the input path values: exact when kernel and oracle saw the same array
    np.testing.assert_allclose(np.nan_to_num(res.boundary), np.nan_to_num(ref.boundary), rtol=boundary_rtol, atol=0)
    np.testing.assert_array_equal(np.isnan(res.betas), np.isnan(ref.betas))
    # betas are ill-conditioned individually; compare the fitted continuation over the ITM range instead
    x = np.linspace(0.8, 1.0, 9)
    for t in range(res.betas.shape[0]):
        if not np.isnan(ref.betas[t, 0]):
            p = ref.betas.shape[1]
            fit_ref = sum(ref.betas[t, i] * x**i for i in range(p))
            fit_gpu = sum(res.betas[t, i] * x**i for i in range(p))
            np.testing.assert_allclose(fit_gpu, fit_ref, rtol=beta_rtol, atol=beta_rtol)


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


# ------------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# European reductions
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# multi-GPU building blocks on one device: two contexts each own half the pairs
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# full BASELINE size: size-independent properties
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# batched pricing (grouped persistent sweep)
# ------------------------------------------------------------------------------------------------------
def _oracle_price_philox(eng, mods, model_fn, S0, K, T, N, M, ot, seed, stream, dtype="f64", semantics="reference"):
    """Oracle sweep on the kernel's own Philox normals for one option."""
    L, E, orc = mods
    rng = E.RngSpec(seed=seed, stream=stream)
    z1 = eng.philox_normals(L.MODEL_HESTON, M, N, 0, dtype, rng).double().cpu().numpy()
    z2 = eng.philox_normals(L.MODEL_HESTON, M, N, 1, dtype, rng).double().cpu().numpy()
    S = orc.heston_paths_antithetic(S0, 0.05, T, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, z1, z2)
    return orc.lsm_sweep(S, K, 0.05, T, ot, semantics=semantics)


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


# ------------------------------------------------------------------------------------------------------
# global-regression LSM (the reference's v3 structure with a linear model on the seven reference features)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["heston_put", "gbm_call", "heston_put_f32"])
def test_global_lsm_vs_oracle(eng, mods, case):
    """optmc_lsm_global vs oracle.lsm_global (pinned against the real om3 loop by tests/test_oracle_golden.py)
    with a least-squares fit on the z-scored reference features, same paths."""
    L, E, orc = mods
    rng = np.random.default_rng(23)
    M, N = 20_000, 25
    if case.startswith("heston"):
        Z1, Z2 = orc.draw_heston_normals(rng, N, M)
        S = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
        K, ot = 100.0, "put"
    else:
        S = orc.gbm_paths_antithetic(100.0, 0.05, 0.25, 1.0, M, N, orc.draw_gbm_normals(rng, N, M))
        K, ot = 97.5, "call"
    f32 = case.endswith("f32")
    if f32:
        S = S.astype(np.float32).astype(np.float64)
    price, st = orc.lsm_global(S, K, 0.05, 1.0, ot, orc.linear_fit)
    res = eng.lsm_global(_slab(eng, S, torch.float32 if f32 else torch.float64), K, 0.05, 1.0, ot)
    assert res["n_rows"] == st["n_rows"]
    assert res["price"] == pytest.approx(price, rel=1e-4 if f32 else 1e-7)
    assert res["stderr"] == pytest.approx(st["stderr"], rel=1e-4 if f32 else 1e-6)
    assert np.abs(res["ex_count"] - st["ex_count"]).sum() <= (20 if f32 else 2)
    same = res["ex_count"] == st["ex_count"]
    both = same & ~np.isnan(st["boundary"])
    np.testing.assert_allclose(res["boundary"][both], st["boundary"][both], rtol=1e-6 if f32 else 1e-9)
    # the model itself: evaluate both on a few (x, tau) points through the reference features
    assert res["rank"] == 6 and res["beta"][4] == 0.0
    xs = np.linspace(0.7, 0.99, 5) if ot == "put" else np.linspace(1.01, 1.3, 5)
    for t_cur in (0.2, 0.6):
        F = orc.features_ref7(xs * K, K, 0.05, 1.0, t_cur)
        got = F @ res["beta"]
        # oracle predictor in its normalised space
        import numpy.linalg as la  # noqa: F401

        feats, targs = [], []
        dt = 1.0 / N
        cf = orc.payoff(S[-1], K, ot).astype(np.float64)
        for t in range(N - 1, 0, -1):
            cf *= np.exp(-0.05 * dt)
            itm = orc.payoff(S[t], K, ot) > 0
            feats.append(orc.features_ref7(S[t, itm], K, 0.05, 1.0, t * dt)); targs.append(cf[itm])
        X_all = np.vstack(feats); Y_all = np.concatenate(targs)
        A = np.column_stack([X_all[:, 1:4], X_all[:, 5:7]])  # [x, x^2, x^3, s, x s] + intercept
        A = np.column_stack([np.ones(len(A)), A])
        w, *_ = np.linalg.lstsq(A, Y_all, rcond=None)
        want = np.column_stack([np.ones(len(F)), F[:, 1:4], F[:, 5:7]]) @ w
        np.testing.assert_allclose(got, want, rtol=2e-4 if f32 else 1e-6, atol=1e-6)


def test_global_lsm_full_size_streams_at_hbm_rate(eng, mods):
    """Config-2-sized slab through both streaming passes: all six informative columns are kept, the regression
    sees every ITM (date, path) row, the sticky-mask price exceeds the textbook one (look-ahead bias, App. A Q1),
    and both stay near the European value -- the regression target is the European payoff (App. A Q4), so the
    exercise rule is far from optimal and the textbook value may even fall below the European one."""
    L, E, orc = mods
    M, N, K = 1_000_000, 252, 100.0
    S = eng.paths(E.heston(100.0, 0.05, 1.0, **HP), M, N, "f32", E.RngSpec(seed=11))
    g = eng.lsm_global(S, K, 0.05, 1.0, "put", arrays=False)
    gt = eng.lsm_global(S, K, 0.05, 1.0, "put", semantics="textbook", arrays=False)
    eu, _ = eng.european_from_slab(S[N].contiguous(), K, 0.05, 1.0, "put")
    assert g["rank"] == 6 and g["n_rows"] > 0.3 * M * (N - 1)
    assert gt["price"] < g["price"] and np.isfinite(g["stderr"])
    assert abs(gt["price"] - eu) < 0.3


# ------------------------------------------------------------------------------------------------------
# per-date neural-network LSM (om2:277-310)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("semantics,epochs", [("reference", 10), ("textbook", 10), ("reference", 40)])
def test_mlp_lsm_vs_torch_oracle_same_init(eng, mods, semantics, epochs):
    """optmc_lsm_mlp vs the oracle loop with the reference's torch ContNet fit, both started from the same
    per-date initial weights (optmc_mlp_init_params).  fp32 network arithmetic on both sides: the fits agree to
    rounding, so prices match closely and only borderline paths may decide differently."""
    L, E, orc = mods
    rng = np.random.default_rng(31)
    M, N = 8192, 12
    Z1, Z2 = orc.draw_heston_normals(rng, N, M)
    S = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
    reg = orc.ContNetRegressor(lambda t: eng.mlp_init_params(1234, t), hidden=32, epochs=epochs, lr=1e-3)
    ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", regressor=reg, semantics=semantics)
    res = eng.lsm_mlp(_slab(eng, S, torch.float64), 100.0, 0.05, 1.0, "put", semantics, hidden=32, epochs=epochs, lr=1e-3,
                      seed=1234)
    np.testing.assert_array_equal(res.n_itm[N - 1], ref.n_itm[N - 1])
    assert np.abs(res.ex_count - ref.ex_count).sum() <= 0.002 * M
    assert res.price == pytest.approx(ref.price, rel=2e-3)
    assert res.stderr == pytest.approx(ref.stderr, rel=1e-2)


def test_mlp_init_matches_torch_default_range(eng):
    """Initial weights are uniform in torch's default nn.Linear range: (-1, 1) for fan-in 1, +-1/sqrt(32) else."""
    p = np.concatenate([eng.mlp_init_params(7, t) for t in range(1, 40)]).reshape(39, -1)
    H = 32
    first, rest = p[:, :2 * H], p[:, 2 * H:]
    assert np.abs(first).max() < 1.0 and np.abs(first).max() > 0.95
    b = 1 / np.sqrt(H)
    assert np.abs(rest).max() < b and np.abs(rest).max() > 0.95 * b
    assert abs(rest.mean()) < 0.01 * b and abs(rest.std() - b / np.sqrt(3)) < 0.01 * b
    assert not np.array_equal(p[0], p[1])  # a fresh network per date


def _torch_contnet_grads(H, xs, ys, p0):
    from torch import nn

    net = nn.Sequential(nn.Linear(1, H), nn.ReLU(), nn.Linear(H, H), nn.ReLU(), nn.Linear(H, 1))
    with torch.no_grad():
        net[0].weight.copy_(torch.from_numpy(p0[0:H].reshape(H, 1)))
        net[0].bias.copy_(torch.from_numpy(p0[H:2 * H]))
        net[2].weight.copy_(torch.from_numpy(p0[2 * H:2 * H + H * H].reshape(H, H)))
        net[2].bias.copy_(torch.from_numpy(p0[2 * H + H * H:3 * H + H * H]))
        net[4].weight.copy_(torch.from_numpy(p0[3 * H + H * H:4 * H + H * H].reshape(1, H)))
        net[4].bias.copy_(torch.from_numpy(p0[4 * H + H * H:4 * H + H * H + 1]))
    X = torch.from_numpy(xs.reshape(-1, 1)); Y = torch.from_numpy(ys.reshape(-1, 1))
    out = net(X)
    loss = nn.MSELoss()(out, Y)
    loss.backward()
    g = np.concatenate([net[0].weight.grad.numpy().ravel(), net[0].bias.grad.numpy().ravel(),
                        net[2].weight.grad.numpy().ravel(), net[2].bias.grad.numpy().ravel(),
                        net[4].weight.grad.numpy().ravel(), net[4].bias.grad.numpy().ravel()])
    return g, out.detach().numpy().ravel()


@pytest.mark.parametrize("H,n", [(32, 1000), (32, 70_000), (128, 1000), (128, 70_000)])
def test_mlp_gradients_vs_torch(eng, H, n):
    """One full-batch ContNet gradient: hidden 32 = fp32 CUDA cores (tight), hidden 128 = bf16 tcgen05 MMAs with
    fp32 accumulation in tensor memory (bf16 tolerance).  n is not a multiple of the tile size on purpose."""
    rng = np.random.default_rng(H + n)
    xs = rng.standard_normal(n).astype(np.float32)
    ys = (np.maximum(0.0, 3.0 - 2.0 * xs) + 0.3 * rng.standard_normal(n)).astype(np.float32)
    p0 = eng.mlp_init_params(99, 5, H)
    g_ref, out_ref = _torch_contnet_grads(H, xs, ys, p0)
    g, out = eng.mlp_grad_debug(H, xs, ys, p0)
    tol_out, tol_g = (2e-6, 2e-5) if H == 32 else (2e-2, 3e-2)
    assert np.abs(out - out_ref).max() <= tol_out * max(1.0, np.abs(out_ref).max())
    # per-block relative L2 error (weights and biases of each layer have very different scales)
    b = [0, H, 2 * H, 2 * H + H * H, 3 * H + H * H, 4 * H + H * H, 4 * H + H * H + 1]
    for lo, hi in zip(b[:-1], b[1:]):
        den = np.linalg.norm(g_ref[lo:hi]) + 1e-12
        assert np.linalg.norm(g[lo:hi] - g_ref[lo:hi]) / den <= tol_g, (lo, hi)


def test_mlp_lsm_tensor_core_hidden128(eng, mods):
    """Per-date NN-LSM with hidden = 128 on tcgen05 vs the torch fp32 oracle from the same initial weights: bf16
    operands perturb the fit slightly, so the comparison is statistical (price within 1%, few decisions differ)."""
    L, E, orc = mods
    rng = np.random.default_rng(37)
    M, N = 8192, 10
    Z1, Z2 = orc.draw_heston_normals(rng, N, M)
    S = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
    reg = orc.ContNetRegressor(lambda t: eng.mlp_init_params(4321, t, 128), hidden=128, epochs=10, lr=1e-3)
    ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", regressor=reg, semantics="textbook")
    res = eng.lsm_mlp(_slab(eng, S, torch.float64), 100.0, 0.05, 1.0, "put", "textbook", hidden=128, epochs=10, lr=1e-3,
                      seed=4321)
    np.testing.assert_array_equal(res.n_itm[N - 1], ref.n_itm[N - 1])
    assert res.price == pytest.approx(ref.price, rel=1e-2)
    assert np.abs(res.ex_count - ref.ex_count).sum() <= 0.03 * M


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


# ------------------------------------------------------------------------------------------------------
# Andersen QE scheme (north-star scheme, not in the reference: oracle = the published algorithm restated in
# oracle.lsm_oracle.heston_paths_qe; accuracy pinned by the semi-analytic Heston price)
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# Global network LSM: SingleLSMNet(7,128,3) on the tensor cores
# ------------------------------------------------------------------------------------------------------
def _single_lsm_net():
    layers = [torch.nn.Linear(7, 128), torch.nn.ReLU()]
    for _ in range(2):
        layers += [torch.nn.Linear(128, 128), torch.nn.ReLU()]
    return torch.nn.Sequential(*layers, torch.nn.Linear(128, 1))


@pytest.mark.parametrize("n", [100, 128, 1000, 5000])
def test_gnet_gradients_vs_torch(eng, n):
    """Loss and gradient of one batch against torch autograd (fp32) -- the 128x128 contractions run in bf16 with fp32
    accumulation, so the tolerance is bf16's (relative L2 per parameter block); targets are shifted so that the
    gradient is a coherent sum (with zero-mean errors it is cancellation noise that amplifies ReLU sign flips)."""
    torch.manual_seed(n)
    net = _single_lsm_net()
    X, y = torch.randn(n, 7), torch.randn(n) - 5.0
    loss = torch.nn.functional.mse_loss(net(X).squeeze(1), y)
    loss.backward()
    flat = lambda g: np.concatenate([(p.grad if g else p.data).detach().reshape(-1).numpy() for p in net.parameters()])  # noqa: E731
    g, l = eng.gnet_grad_debug(X.numpy(), y.numpy(), flat(False))
    g_ref = flat(True)
    assert l == pytest.approx(float(loss.detach()), rel=2e-4)
    seg = {"W1": (0, 896), "b1": (896, 1024), "W2": (1024, 17408), "b2": (17408, 17536), "W3": (17536, 33920),
           "b3": (33920, 34048), "w4": (34048, 34176), "b4": (34176, 34177)}
    for k, (a, b) in seg.items():
        err = np.linalg.norm(g[a:b] - g_ref[a:b]) / np.linalg.norm(g_ref[a:b])
        assert err < 4e-2, (k, err)


def test_gnet_training_and_prices_vs_torch_restatement(eng, mods):
    """The whole v3 algorithm against its torch restatement (oracle.single_lsm_net_fit) on the same paths.  The
    reference's initialisation / shuffle / dropout streams are torch's global RNG, so agreement is statistical: the
    training loss (a smooth functional of the fit) within 1 %, the price within the seed-to-seed spread of the
    estimator itself."""
    L, E, orc = mods
    model = E.heston(100.0, 0.05, 1.0, **HP)
    S = eng.paths(model, 40_000, 25, "f64", E.RngSpec(seed=8))
    Sn = S.cpu().numpy()
    got = [eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", variant="gpu", epochs=12, seed=sd, stop_patience=0) for sd in (1, 2, 3, 4)]
    ref, ref_loss, ref_rows = [], [], 0
    for sd in (1, 2):
        log = []
        p, st = orc.lsm_global(Sn, 100.0, 0.05, 1.0, "put", orc.single_lsm_net_fit("gpu", epochs=12, seed=sd, log=log), target_ddof=1)
        ref.append(p); ref_loss.append(min(log)); ref_rows = st["n_rows"]
    assert got[0]["n_rows"] == ref_rows
    assert np.mean([r["best_loss"] for r in got]) == pytest.approx(np.mean(ref_loss), rel=1e-2)
    assert abs(np.mean([r["price"] for r in got]) - np.mean(ref)) < 0.35
    # reproducible: same seed, same price, bit for bit
    again = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", variant="gpu", epochs=12, seed=1, stop_patience=0)
    assert again["price"] == got[0]["price"] and again["best_loss"] == got[0]["best_loss"]
    # exercise statistics are consistent with the price pass
    assert got[0]["ex_count"][1:25].sum() > 0 and np.isnan(got[0]["boundary"][0])


def test_gnet_textbook_policy_is_sane_and_edge_cases(eng, mods):
    L, E, orc = mods
    gbm = E.gbm(100.0, 0.05, 1.0, 0.2)
    S = eng.paths(gbm, 100_000, 50, "f32", E.RngSpec(seed=4))
    r = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "textbook", variant="gpu", epochs=15, seed=3)
    # American put, GBM: binomial value 6.09; an in-sample network policy lands near it
    assert 5.85 < r["price"] < 6.35
    assert r["epochs_run"] >= 3 and r["n_rows"] > 1_000_000
    # no in-the-money row at all (deep OTM call): the price is the discounted terminal payoff mean, no training
    r0 = eng.lsm_gnet(S, 1e6, 0.05, 1.0, "call", "reference", variant="gpu", epochs=3)
    assert r0["n_rows"] == 0 and r0["epochs_run"] == 0 and r0["price"] == 0.0
    # N = 1: no exercise date before maturity
    S1 = eng.paths(gbm, 4096, 1, "f64", E.RngSpec(seed=4))
    r1 = eng.lsm_gnet(S1, 100.0, 0.05, 1.0, "put", "reference", variant="cpu", epochs=2)
    pay = np.maximum(100.0 - S1[1].cpu().numpy(), 0.0)
    assert r1["n_rows"] == 0 and r1["price"] == pytest.approx(pay.mean(), rel=1e-12)


def test_compat_pricer_nn_regressor(mods):
    from options_model_b200 import compat

    p = compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, option_type="put", rng_manager=compat.RNGManager(42),
                                    lsm_regressor="nn", nn_epochs=3)
    v = p.price_american_enhanced_lsm_gpu(100.0, 1.0, num_simulations=50_000, num_time_steps=30)
    assert 5.5 < v < 8.5 and p.last_result["epochs_run"] >= 1
    with pytest.raises(ValueError):
        compat.AdvancedOptionPricer(K=100.0, r=0.05, sigma=0.2, lsm_regressor="nn", nn_hidden=64).price_american_enhanced_lsm(100.0, 1.0, 1000, 10)


# ------------------------------------------------------------------------------------------------------
# Local volatility: the IV network inside the path step against the REAL reference
# (tests/golden/ref_localvol.npz, oracle/gen_golden_localvol.py)
# ------------------------------------------------------------------------------------------------------
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


# ------------------------------------------------------------------------------------------------------
# Out-of-sample exercise: coefficients fitted on one path set applied to another
# ------------------------------------------------------------------------------------------------------
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


def test_gnet_warm_start_roundtrip(eng, mods):
    """init_params / final_params of optmc_gnet_params: 0 epochs with given weights reproduces the decision pass of the run
    that produced them (the torch-GPU file's cached network, om3gpu:741-748); a warm start continues from them."""
    L, E, orc = mods
    S = eng.paths(E.gbm(100.0, 0.05, 1.0, 0.2), 20_000, 12, "f32", E.RngSpec(seed=6))
    a = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "textbook", variant="gpu", epochs=4, seed=5, return_params=True, stop_patience=0)
    assert a["params"].shape == (eng.GNET_PARAMS,) and np.isfinite(a["params"]).all()
    b = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "textbook", variant="gpu", epochs=0, seed=99, init_params=a["params"])
    assert b["price"] == a["price"] and b["epochs_run"] == 0
    c = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "textbook", variant="gpu", epochs=2, seed=5, init_params=a["params"], stop_patience=0)
    assert c["best_loss"] < a["best_loss"] + 5e-3


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
