#!/usr/bin/env python
"""Generate tests/golden/* by running the REAL reference (read-only at /root/reference).

Run in the build container only:   python oracle/gen_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.
yfinance / plotly / matplotlib are import-only dependencies of the reference (I/O and plots) and are
replaced by empty stub modules; option_model_3_gpu.py is exec'd with its three mis-indented lines
(594-596) fixed in memory.  Nothing from the reference is copied into the repo.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("OPTMC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))


def _stub_modules():
    for name in ["yfinance", "plotly", "plotly.graph_objects", "plotly.io", "matplotlib", "matplotlib.pyplot",
                 "streamlit"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plotly.io"].renderers = types.SimpleNamespace(default=None)
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    sys.modules["plotly"].io = sys.modules["plotly.io"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load_reference():
    _stub_modules()
    sys.path.insert(0, os.path.join(REF, "options_model_3"))
    sys.path.insert(0, REF)
    import importlib

    om3 = importlib.import_module("options_model_3")
    hc = importlib.import_module("heston_calibration")
    om2 = importlib.import_module("options_model_2")
    # om3gpu: fix the class-body-level "self." lines in memory (App. B-1)
    src = open(os.path.join(REF, "options_model_3", "option_model_3_gpu.py")).read().split("\n")
    for i in range(len(src)):
        s = src[i]
        if s.startswith("    self.") or s.startswith("    # Cache a reusable") or s.startswith("    # Device and"):
            src[i] = "    " + s
    om3gpu = types.ModuleType("option_model_3_gpu")
    om3gpu.__file__ = "option_model_3_gpu.py(patched)"
    sys.modules["option_model_3_gpu"] = om3gpu
    exec(compile("\n".join(src), om3gpu.__file__, "exec"), om3gpu.__dict__)
    return om3, om3gpu, hc, om2


HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


def main():
    import torch
    import torch.nn as nn

    os.makedirs(OUT, exist_ok=True)
    om3, om3gpu, hc, om2 = load_reference()
    from oracle import lsm_oracle as orc

    meta = {"numpy": np.__version__, "torch": torch.__version__, "reference": REF}

    # ---- 1. RNG seed tree (om3:69-79) -------------------------------------------------------
    m = om3.RNGManager(42)
    seeds = [int(m.get_child_seed()) for _ in range(4)]
    m = om3.RNGManager(42)
    z = m.get_child_rng().standard_normal(4)
    meta["rng"] = {"master_seed": 42, "child_seeds": seeds, "child0_normals": z.tolist()}

    # ---- 2. Heston absorption-Euler paths (om3:211-251) -------------------------------------
    for tag, M, N, seed in [("even", 64, 16, 7), ("odd", 7, 5, 11)]:
        S = om3.simulate_heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"],
                                                 HP["rho"], M, N, np.random.default_rng(seed))
        Me = M // 2 * 2
        Z1, Z2 = orc.draw_heston_normals(np.random.default_rng(seed), N, Me)
        np.savez_compressed(os.path.join(OUT, f"ref_heston_paths_{tag}.npz"), S=S, Z1=Z1, Z2=Z2,
                            args=np.array([100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"]]),
                            M=M, N=N, seed=seed)

    # ---- 3. features (om3:105-121) ------------------------------------------------------------
    Sx = np.array([55.0, 80.0, 99.5, 100.0, 100.5, 130.0, 250.0])
    F = om3.create_regression_features(Sx, 100.0, 0.05, 1.0, 0.3)
    F_end = om3.create_regression_features(Sx, 100.0, 0.05, 1.0, 1.0)  # tau clamp 1e-6
    np.savez_compressed(os.path.join(OUT, "ref_features.npz"), S=Sx, F=F, F_end=F_end)

    # ---- 4. Welford (om3:33-63) ---------------------------------------------------------------
    rng = np.random.default_rng(3)
    batches = [rng.standard_normal(n) * 3 + 1 for n in (5, 1, 0, 17, 2)]
    st = (0.0, 0.0, 0)
    states = []
    for b in batches:
        st = om3.welford_batch_update(*st, b)
        states.append([float(st[0]), float(st[1]), int(st[2])])
    it = iter(batches)
    mc = om3.monte_carlo_price_streaming(lambda n: next(it)[:n], 25, 100)
    np.savez_compressed(os.path.join(OUT, "ref_welford.npz"), flat=np.concatenate(batches),
                        sizes=np.array([b.size for b in batches]), states=np.array(states),
                        mc_first=np.array([float(mc[0]), float(mc[1]), float(mc[2])]))

    # ---- 5. European streaming (om3:382-437) ----------------------------------------------------
    eu = {}
    for name, kw in [("gbm_put", dict(sigma=0.2, option_type="put")),
                     ("gbm_call", dict(sigma=0.25, option_type="call")),
                     ("heston_put", dict(sigma=None, option_type="put", use_heston=True, heston_params=HP))]:
        pr = om3.AdvancedOptionPricer(K=100.0, r=0.05, rng_manager=om3.RNGManager(5), chunk_size=500, **kw)
        eu[name] = float(pr.price_european_streaming(100.0, 0.5, 1001, 10))
    meta["european_streaming"] = {"S0": 100.0, "K": 100.0, "r": 0.05, "T": 0.5, "n": 1001, "steps": 10,
                                  "chunk": 500, "master_seed": 5, "prices": eu}

    # ---- 6. LSM skeleton of om3:439-651 with a deterministic stand-in network -------------------
    W = np.array([0.0, -0.8, 0.1, 0.05, 0.0, 0.02, 0.3])
    B = 0.1

    class FakeNet(nn.Module):
        def __init__(self, input_dim=7, hidden_dim=128, num_layers=3, dropout=0.1):
            super().__init__()
            self.dummy = nn.Parameter(torch.zeros(1))
            self.w = torch.tensor(W, dtype=torch.float64)

        def forward(self, x):
            out = (x.double() @ self.w + B).unsqueeze(1)
            if torch.is_grad_enabled():  # training step: MSELoss needs the target's dtype
                return out.float() + 0.0 * self.dummy
            return out  # decision time (torch.no_grad): fp64, no dropout

    real_net = om3.SingleLSMNet
    om3.SingleLSMNet = FakeNet
    sk = {}
    try:
        for name, kw in [("gbm_put", dict(sigma=0.2, option_type="put")),
                         ("gbm_call", dict(sigma=0.3, option_type="call")),
                         ("heston_put", dict(sigma=None, option_type="put", use_heston=True, heston_params=HP))]:
            pr = om3.AdvancedOptionPricer(K=100.0, r=0.05, rng_manager=om3.RNGManager(42), nn_epochs=1,
                                          use_control_variate=False, **kw)
            sk[name] = float(pr.price_american_enhanced_lsm(100.0, 1.0, 2000, 12))
    finally:
        om3.SingleLSMNet = real_net
    meta["lsm_skeleton_global"] = {"W": W.tolist(), "B": B, "S0": 100.0, "K": 100.0, "r": 0.05, "T": 1.0,
                                   "M": 2000, "N": 12, "master_seed": 42, "prices": sk,
                                   "sigma": {"gbm_put": 0.2, "gbm_call": 0.3}}

    # ---- 7. per-date loop of om2:277-310 with a deterministic stand-in ContNet ------------------
    A0, A1, A2 = 4.0, -3.0, 0.5

    class FakeCont(nn.Module):
        def __init__(self, hidden=32):
            super().__init__()
            self.dummy = nn.Parameter(torch.zeros(1))

        def forward(self, x):
            xd = x.double()
            out = A0 + A1 * xd + A2 * xd * xd
            if torch.is_grad_enabled():
                return out.float() + 0.0 * self.dummy
            return out

    real_cont = om2.ContNet
    om2.ContNet = FakeCont
    pd_prices = {}
    try:
        for name, kw in [("gbm_put", dict(sigma=0.2, option_type="put")),
                         ("heston_put", dict(sigma=None, option_type="put", use_heston=True, heston_params=HP))]:
            pr = om2.OptionPricer(K=100.0, r=0.05, seed=42, nn_epochs=1, **kw)
            pd_prices[name] = float(pr.price_american_option(100.0, 1.0, 2000, 12))
    finally:
        om2.ContNet = real_cont
    meta["lsm_skeleton_perdate"] = {"A": [A0, A1, A2], "S0": 100.0, "K": 100.0, "r": 0.05, "T": 1.0, "M": 2000,
                                    "N": 12, "seed": 42, "prices": pd_prices}

    # ---- 8. calibrator scheme (hc:197-281) -------------------------------------------------------
    cfg = hc.CalibrationConfig(n_mc_paths=64, n_time_steps=10, seed=42, verbose=False, plot_results=False)
    params = hc.HestonParams(kappa=2.0, theta=0.04, sigma=0.5, rho=-0.7, v0=0.04)
    S, V = hc.HestonPricer(cfg).simulate_paths(params, 100.0, 0.75, 0.03)
    Z1, Z2i = orc.hc_draw_normals(np.random.default_rng(42), 64, 10)
    pricer = hc.HestonPricer(cfg)
    c1 = pricer.price_european_option(params, 100.0, 95.0, 0.75, 0.03, "call")
    p2 = pricer.price_european_option(params, 100.0, 105.0, 0.75, 0.03, "put")  # rng stream continues
    np.savez_compressed(os.path.join(OUT, "ref_hc_paths.npz"), S=S, V=V, Z1=Z1, Z2i=Z2i,
                        params=params.to_array(), S0=100.0, T=0.75, r=0.03, call_95=c1, put_105=p2)
    cfg_big = hc.CalibrationConfig(n_mc_paths=50_000, n_time_steps=100, seed=42, verbose=False, plot_results=False)
    meta["hc_call_50k_x100"] = float(hc.HestonPricer(cfg_big).price_european_option(params, 100.0, 100.0, 1.0, 0.05, "call"))

    # ---- 9. torch fp32 path functions (om3gpu:117-248) on CPU ------------------------------------
    dev = torch.device("cpu")
    M, N = 32, 8
    torch.manual_seed(123)
    S_bs = om3gpu.simulate_bs_paths_torch(100.0, 0.05, 1.0, 0.2, M, N, dev).numpy()
    torch.manual_seed(123)
    Zh = torch.randn(N, M // 2).numpy()
    torch.manual_seed(124)
    S_bw = om3gpu.simulate_bs_paths_torch_bandwidth_optimized(100.0, 0.05, 1.0, 0.2, M, N, dev).numpy()
    torch.manual_seed(124)
    Zbw = torch.randn(N, M).numpy()
    torch.manual_seed(125)
    S_h = om3gpu.simulate_heston_paths_torch(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"],
                                             HP["rho"], M, N, dev).numpy()
    torch.manual_seed(125)
    Z1t, Z2t = [], []
    for _ in range(N):
        Z1t.append(torch.randn(M // 2).numpy())
        Z2t.append(torch.randn(M // 2).numpy())
    np.savez_compressed(os.path.join(OUT, "ref_torch_paths.npz"), S_bs=S_bs, Zh=Zh, S_bw=S_bw, Zbw=Zbw, S_h=S_h,
                        Z1=np.stack(Z1t), Z2=np.stack(Z2t), M=M, N=N)
    Ft = om3gpu.create_regression_features_torch(torch.tensor(Sx, dtype=torch.float32), 100.0, 0.05, 1.0, 0.3).numpy()
    np.savez_compressed(os.path.join(OUT, "ref_features_torch.npz"), S=Sx.astype(np.float32), F=Ft)

    # ---- 10. restatement pins: polynomial LSM; oracle's own values ---------------
    pins = {}
    # C1: GBM put 100k x 50
    res, S, _ = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 100_000, 50, orc.RNGManager(42), sigma=0.2,
                                       return_paths=True)
    tb = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", semantics="textbook")
    pins["c1_gbm_put_100k_50"] = {
        "S_1_0": float(S[1, 0]), "S_1_1": float(S[1, 1]), "S_50_0": float(S[50, 0]),
        "reference": {"price": res.price, "stderr": res.stderr, "boundary_25": float(res.boundary[25]),
                      "boundary_45": float(res.boundary[45]), "beta_25": res.betas[25].tolist()},
        "textbook": {"price": tb.price, "stderr": tb.stderr, "boundary_25": float(tb.boundary[25]),
                     "boundary_45": float(tb.boundary[45])},
    }
    # Heston put 100k x 50
    res, S, _ = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 100_000, 50, orc.RNGManager(42),
                                       heston_params=HP, return_paths=True)
    tb = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", semantics="textbook")
    eu = orc.european_from_paths(S[-1], 100.0, 0.05, 1.0, "put")
    pins["heston_put_100k_50"] = {"S_50_0": float(S[50, 0]), "european": eu[0], "reference": res.price,
                                  "reference_se": res.stderr, "textbook": tb.price}
    # small full-vector case used by CPU and GPU tests
    res, S, _ = orc.price_american_lsm(100.0, 100.0, 0.05, 1.0, "put", 4096, 20, orc.RNGManager(1),
                                       heston_params=HP, return_paths=True)
    np.savez_compressed(os.path.join(OUT, "oracle_heston_poly2_small.npz"), price=res.price, stderr=res.stderr,
                        betas=res.betas, boundary=res.boundary, ex_count=res.ex_count, n_itm=res.n_itm,
                        S_last=S[-1])
    meta["poly_pins"] = pins

    with open(os.path.join(OUT, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True)[:3000])


if __name__ == "__main__":
    main()
