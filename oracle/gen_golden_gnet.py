"""Golden prices of the reference's v3 algorithm with its own network (om3:439-651), produced by the REAL
`AdvancedOptionPricer.price_american_enhanced_lsm` in this container (torch CPU), together with the oracle's restatement
run on the same streams.  Writes tests/golden/ref_gnet_prices.json.  Run: python oracle/gen_golden_gnet.py
Test infrastructure only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import gen_golden as G  # noqa: E402
from oracle import lsm_oracle as orc  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


def oracle_price(case):
    price, _ = orc.price_american_enhanced_lsm_nn(case["S0"], case["K"], case["r"], case["T"], case["option_type"], case["M"],
                                                  case["N"], case["master_seed"], sigma=case["sigma"],
                                                  heston_params=HP if case["heston"] else None, nn_hidden=case["nn_hidden"],
                                                  nn_epochs=case["nn_epochs"], nn_lr=case["nn_lr"])
    return price


def main():
    import torch

    om3, _, _, _ = G.load_reference()
    cases = [
        dict(name="gbm_put", heston=False, S0=100.0, K=100.0, r=0.05, sigma=0.2, T=1.0, option_type="put", M=2000, N=10,
             nn_hidden=128, nn_epochs=3, nn_lr=1e-3, master_seed=42),
        dict(name="heston_put", heston=True, S0=100.0, K=100.0, r=0.05, sigma=0.2, T=1.0, option_type="put", M=3000, N=12,
             nn_hidden=128, nn_epochs=4, nn_lr=1e-3, master_seed=7),
        dict(name="gbm_call", heston=False, S0=100.0, K=95.0, r=0.03, sigma=0.3, T=0.5, option_type="call", M=1500, N=8,
             nn_hidden=128, nn_epochs=2, nn_lr=2e-3, master_seed=3),
    ]
    out = {"torch": torch.__version__, "numpy": np.__version__, "cases": []}
    for c in cases:
        pr = om3.AdvancedOptionPricer(K=c["K"], r=c["r"], sigma=c["sigma"], option_type=c["option_type"],
                                      rng_manager=om3.RNGManager(c["master_seed"]), use_heston=c["heston"],
                                      heston_params=HP if c["heston"] else None, nn_hidden=c["nn_hidden"], nn_epochs=c["nn_epochs"],
                                      nn_lr=c["nn_lr"])
        ref = float(pr.price_american_enhanced_lsm(c["S0"], c["T"], c["M"], c["N"]))
        mine = float(oracle_price(c))
        print(c["name"], "reference", ref, "oracle", mine, "rel diff", abs(ref - mine) / ref)
        out["cases"].append(dict(c, reference_price=ref, oracle_price_at_generation=mine))
    # the per-date network pricer the Streamlit UI imports (om2:216-330) and its om1 twin (om1:44-170)
    _, _, _, om2 = G.load_reference()
    out["om2_cases"] = []
    for c in (dict(S0=100.0, K=100.0, r=0.05, sigma=0.2, T=1.0, option_type="put", M=2000, N=10, seed=42, nn_hidden=32, nn_epochs=10),
              dict(S0=100.0, K=105.0, r=0.03, sigma=0.3, T=0.5, option_type="call", M=1200, N=8, seed=7, nn_hidden=32, nn_epochs=5)):
        pr = om2.OptionPricer(c["K"], c["r"], c["sigma"], c["option_type"], 2, c["seed"], False, None, c["nn_hidden"], c["nn_epochs"],
                              1e-3, False)
        ref = float(pr.price_american_option(c["S0"], c["T"], c["M"], c["N"]))
        mine = orc.price_american_om2_nn(c["S0"], c["K"], c["r"], c["sigma"], c["T"], c["option_type"], c["M"], c["N"], c["seed"],
                                         c["nn_hidden"], c["nn_epochs"], 1e-3)
        print("om2", c["option_type"], "reference", ref, "oracle", mine, "rel diff", abs(ref - mine[0]) / ref)
        out["om2_cases"].append(dict(c, reference_price=ref, oracle_at_generation=list(mine)))
    # the calibration objective (hc:404-472) of a fresh HestonCalibrator on a small synthetic smile
    import pandas as pd

    _, _, hc, _ = G.load_reference()
    S0, r = 100.0, 0.05
    Kg, Tg = np.meshgrid(np.linspace(85, 115, 4), np.array([0.25, 0.75]))
    K, T = Kg.ravel(), Tg.ravel()
    iv = 0.2 + 0.1 * np.abs(np.log(K / S0)) + 0.02 * np.sqrt(T)
    cfg = hc.CalibrationConfig(n_mc_paths=4000, n_time_steps=20, verbose=False, plot_results=False)
    cal = hc.HestonCalibrator(cfg)
    md = object.__new__(hc.MarketData)  # the constructor reads self.S0 before setting it
    md.S0, md.r = S0, r
    md.df = md._validate_data(pd.DataFrame({"K": K, "T": T, "sigma_IV": iv}))
    md.regime = md._detect_regime()
    cal.market_data = md
    out["hc_objective"] = []
    for x in ([2.0, 0.04, 0.5, -0.7, 0.04], [3.0, 0.06, 0.3, -0.4, 0.05]):
        cal.pricer = hc.HestonPricer(cfg)  # fresh generator per evaluation, as a new calibrator would have
        ref = float(cal._objective_function(np.array(x)))
        mine, prices = orc.hc_objective(x, S0, r, K, T, iv, cfg.n_mc_paths, cfg.n_time_steps, cfg.seed)
        print("hc objective", x, "reference", ref, "oracle", mine, "rel diff", abs(ref - mine) / ref)
        out["hc_objective"].append(dict(x=x, S0=S0, r=r, K=K.tolist(), T=T.tolist(), sigma_iv=iv.tolist(), n_mc_paths=cfg.n_mc_paths,
                                        n_time_steps=cfg.n_time_steps, seed=cfg.seed, reference_value=ref, prices=[float(p) for p in prices]))
    with open(os.path.join(G.OUT, "ref_gnet_prices.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
