"""GPU parity, regressions beyond the per-date polynomial: global linear model, per-date ContNet (CUDA cores / tcgen05), global SingleLSMNet on tcgen05.

All calls go through the C ABI (ctypes).  Tolerances (north star): fed identical draws, prices / betas / boundary within
1e-5 relative in fp64 and 1e-4 in fp32 -- the fp64 assertions are far tighter; integer outputs are compared exactly.
"""
import os  # noqa: F401

import numpy as np
import pytest

from gpu_common import _slab, HP  # noqa: F401

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


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


def test_gnet_streams_device_equals_numpy_restatement(eng, mods):
    """The shuffle and the dropout masks the training / decision kernels evaluate (optmc_gnet_streams_debug runs the same
    device functions) against oracle/engine_streams.py, bit for bit -- what makes the paired test below a paired one."""
    from oracle import engine_streams as es

    rows = np.concatenate([np.arange(300, dtype=np.uint32), es.walk_row_id(np.arange(39_000, 39_300), 17)])
    for seed, ep, step, n in ((3, 0, 1, 1), (3, 1, 77, 1000), (2**40 + 5, 4, 301, 481_233), (42, 0, 0, 4097)):
        perm, keep = eng.gnet_streams_debug(seed, ep, step, 0.1, n, rows)
        assert np.array_equal(perm, es.feistel_perm(n, es.perm_key(seed, ep)))
        key = es.train_drop_key(seed, step) if step else es.walk_drop_key(seed)
        for layer in range(3):
            assert np.array_equal(keep[:, layer], es.keep_mask(key, rows, layer, 0.1))
    _, keep = eng.gnet_streams_debug(3, 0, 5, 0.0, 0, rows)
    assert keep.all()


def test_gnet_paired_training_vs_torch_same_init_shuffle_and_masks(eng, mods):
    """the engine and the torch fp32 restatement of om3gpu:700-830 trained from the SAME initial weights
    on the SAME mini-batches with the SAME dropout masks (oracle/engine_streams.py).  What differs is arithmetic: bf16
    tensor-core operands in the two hidden layers.  Tolerances: per-epoch loss 0.2 %, price 1 %, exercise counts
    within 2 % of the paths in total (replaces the 0.35-absolute statistical bound of the unpaired test above)."""
    from oracle import engine_streams as es

    L, E, orc = mods
    model = E.heston(100.0, 0.05, 1.0, **HP)
    M, N, epochs, seed = 40_000, 25, 12, 11
    S = eng.paths(model, M, N, "f64", E.RngSpec(seed=8))
    init = es.torch_default_init(5)
    got = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", variant="gpu", epochs=epochs, seed=seed, stop_patience=0,
                       init_params=init, return_params=True)
    log = []
    price, st = es.paired_gnet(S.cpu().numpy(), 100.0, 0.05, 1.0, "put", init, seed, epochs=epochs, log=log)
    assert got["n_rows"] == st["n_rows"] and got["epochs_run"] == epochs
    assert got["best_loss"] == pytest.approx(st["best_loss"], rel=2e-3)
    assert got["price"] == pytest.approx(price, rel=1e-2)
    assert np.abs(got["ex_count"] - st["ex_count"]).sum() <= 0.02 * M
    # the trained weights themselves stay close (300 optimiser steps of bf16-vs-fp32 drift): relative L2 distance
    d = np.linalg.norm(got["params"] - st["params"]) / np.linalg.norm(st["params"])
    print(f"paired gnet: price {got['price']:.5f} vs {price:.5f}, best loss {got['best_loss']:.6f} vs {st['best_loss']:.6f}, "
          f"|d ex_count| {np.abs(got['ex_count'] - st['ex_count']).sum()} of {M}, param distance {d:.4f}")
    assert d < 0.15, d


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


def test_gnet_sharded_single_rank_group_is_bit_identical(eng, mods):
    """optmc_lsm_gnet_sharded on a one-rank group (comm_export / comm_init with itself): the peer-memory path runs --
    row counts and moments gathered through tagged words, per-step gradients pushed as {tag, fp32} words and summed by
    the optimiser kernel from the slots, the final value sums gathered -- and must reproduce optmc_lsm_gnet bit for bit
    (same batches by construction of optmc_gnet_shard_plan, rank 0 keys unchanged).  Two real GPUs: tests/test_multi_gpu.py."""
    L, E, orc = mods
    e = E.Engine(0)
    try:
        e.comm_init(0, 1, [e.comm_export()])
        for mdl, M, N, batch in ((E.gbm(100.0, 0.05, 1.0, 0.2), 20_000, 20, 256), (E.heston(100.0, 0.05, 1.0, **HP), 60_000, 30, 8192)):
            S = eng.paths(mdl, M, N, "f32", E.RngSpec(seed=21))
            kw = dict(variant="cpu" if batch == 256 else "gpu", epochs=3, batch=batch, seed=9, return_params=True)
            plain = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", **kw)
            for _ in range(2):  # twice: the exchange tags keep running across calls
                S2 = e.paths(mdl, M, N, "f32", E.RngSpec(seed=21))
                sh = e.lsm_gnet(S2, 100.0, 0.05, 1.0, "put", "reference", M_total=M, **kw)
                assert sh["price"] == plain["price"] and sh["stderr"] == plain["stderr"]
                assert sh["best_loss"] == plain["best_loss"] and sh["n_rows"] == plain["n_rows"] and sh["epochs_run"] == plain["epochs_run"]
                np.testing.assert_array_equal(sh["params"], plain["params"])
                np.testing.assert_array_equal(sh["ex_count"], plain["ex_count"])
        # a slab without in-the-money rows: nothing to fit, the gathers still run
        S = eng.paths(E.gbm(100.0, 0.05, 0.01, 0.01), 4096, 4, "f32", E.RngSpec(seed=3))
        a = eng.lsm_gnet(S, 50.0, 0.05, 0.01, "put", "reference", epochs=2)
        b = e.lsm_gnet(e.paths(E.gbm(100.0, 0.05, 0.01, 0.01), 4096, 4, "f32", E.RngSpec(seed=3)), 50.0, 0.05, 0.01, "put", "reference",
                       epochs=2, M_total=4096)
        assert a["n_rows"] == b["n_rows"] == 0 and a["price"] == b["price"] == 0.0
        e.comm_finalize()
    finally:
        e.close()
    with pytest.raises(Exception):  # no group wired
        eng.lsm_gnet(eng.paths(E.gbm(100.0, 0.05, 1.0, 0.2), 4096, 4, "f32", E.RngSpec(seed=3)), 100.0, 0.05, 1.0, "put", "reference",
                     epochs=1, M_total=4096)


def test_gnet_per_date_one_exercise_date_equals_global_fit(eng, mods):
    """optmc_gnet_params.per_date on a slab with ONE exercise date (N = 2): the per-date loop collects the same rows
    (live at t = 1 = in the money), the same targets (cash-flow at t = 1 = discounted terminal payoff), z-scores with the
    same moments, starts from the same weights and uses the same shuffle / dropout streams as the global fit, so training
    is identical; the decision uses the decision pass's stream.  Weights and exercise decisions bit for bit, price to
    rounding (the per-date loop keeps its cash-flows in date-N money in storage precision)."""
    L, E, orc = mods
    for mdl, dtype in ((E.heston(100.0, 0.05, 0.5, **HP), "f64"), (E.gbm(100.0, 0.05, 0.5, 0.25), "f32")):
        S = eng.paths(mdl, 30_000, 2, dtype, E.RngSpec(seed=4))
        for sem in ("reference", "textbook"):
            kw = dict(variant="gpu", epochs=4, batch=2048, seed=6, return_params=True)
            g = eng.lsm_gnet(S, 100.0, 0.05, 0.5, "put", sem, **kw)
            d = eng.lsm_gnet(S, 100.0, 0.05, 0.5, "put", sem, per_date=1, **kw)
            assert d["n_rows"] == g["n_rows"] > 5000 and d["epochs_run"] == g["epochs_run"]
            np.testing.assert_array_equal(d["params"], g["params"])
            assert d["best_loss"] == g["best_loss"]
            np.testing.assert_array_equal(d["ex_count"], g["ex_count"])
            np.testing.assert_array_equal(d["boundary"], g["boundary"])
            assert d["price"] == pytest.approx(g["price"], rel=1e-12 if dtype == "f64" else 2e-7)
            assert d["stderr"] == pytest.approx(g["stderr"], rel=1e-9 if dtype == "f64" else 2e-6)


def test_gnet_per_date_vs_torch_restatement(eng, mods):
    """A fresh SingleLSMNet(7,128,3) per exercise date (the om2:277-310 loop with om3's regressor) against its torch
    restatement (oracle.SingleLSMNetDateRegressor inside oracle.lsm_sweep) on the same paths: the streams differ (Philox /
    Feistel vs torch's generator), so agreement is statistical -- mean training loss over the dates within 3 %, prices
    within 1 % (textbook semantics, no inference dropout, early stopping off on both sides -- with the torch-GPU file's
    patience of 3 BOTH implementations stop under-trained and scatter by +-3 % from seed to seed: engine 5.26-5.59, torch
    5.49-5.63, losses 0.529-0.547 vs 0.534-0.539 on this problem), both within 1.5 % of the polynomial LSM on the same
    paths; row counts equal."""
    L, E, orc = mods
    M, N = 20_000, 10
    S = eng.paths(E.heston(100.0, 0.05, 1.0, **HP), M, N, "f64", E.RngSpec(seed=12))
    Sn = S.cpu().numpy()
    kw = dict(variant="gpu", epochs=30, batch=4096, dropout=0.0, inference_dropout=0, stop_patience=0)
    got = [eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "textbook", per_date=1, seed=sd, **kw) for sd in (1, 2)]
    regs = [orc.SingleLSMNetDateRegressor(100.0, 0.05, 1.0, variant="gpu", epochs=30, batch=4096, dropout=0.0, seed=sd, patience=10**9) for sd in (1, 2)]
    ref = [orc.lsm_sweep(Sn, 100.0, 0.05, 1.0, "put", regressor=rg, semantics="textbook") for rg in regs]
    poly = orc.lsm_sweep(Sn, 100.0, 0.05, 1.0, "put", semantics="textbook")
    assert got[0]["n_rows"] == int(ref[0].n_itm.sum()) and got[0]["epochs_run"] == 30 * (N - 1)
    p_eng, p_ref = np.mean([g["price"] for g in got]), np.mean([r.price for r in ref])
    assert p_eng == pytest.approx(p_ref, rel=1e-2)
    assert p_eng == pytest.approx(poly.price, rel=1.5e-2) and p_ref == pytest.approx(poly.price, rel=1.5e-2)
    l_ref = np.mean([np.mean(list(rg.loss.values())) for rg in regs])
    assert np.mean([g["best_loss"] for g in got]) == pytest.approx(l_ref, rel=3e-2)
    # reproducible, and the reference semantics run (sticky mask, dropout left on at inference)
    again = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "textbook", per_date=1, seed=1, **kw)
    assert again["price"] == got[0]["price"]
    r = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", per_date=1, seed=1, variant="gpu", epochs=10, batch=4096)
    assert r["price"] > poly.price * 0.95 and r["ex_count"][1:N].sum() > 0


def test_gnet_device_side_epoch_end_equals_host_loop(eng, mods):
    """Without a scheduler and without early stopping the best-weights snapshot (om3:599-603) is taken on the device and
    the host never waits for an epoch; a patience that cannot trigger forces the host loop -- same weights, loss, price."""
    L, E, orc = mods
    S = eng.paths(E.gbm(100.0, 0.05, 1.0, 0.2), 30_000, 12, "f32", E.RngSpec(seed=2))
    kw = dict(variant="gpu", epochs=7, batch=4096, seed=3, return_params=True)
    dev = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", stop_patience=0, **kw)
    host = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", stop_patience=1_000_000, **kw)
    assert dev["epochs_run"] == host["epochs_run"] == 7 and dev["n_launches"] != host["n_launches"]
    assert dev["best_loss"] == host["best_loss"] and dev["price"] == host["price"]
    np.testing.assert_array_equal(dev["params"], host["params"])
