"""Development check: training loss / price of the global network LSM against the torch restatement, with and without dropout."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa
from options_model_b200 import engine as E
from oracle import lsm_oracle as orc
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, **HP)
S = eng.paths(model, 40_000, 25, "f64", E.RngSpec(seed=8))
Sn = S.cpu().numpy()
for drop in (0.0, 0.1):
    for variant, ep in (("gpu", 30),):
        got = []
        for sd in (1, 2, 3, 4):
            r = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", variant=variant, epochs=ep, seed=sd, dropout=drop, stop_patience=0)
            got.append((round(r["price"], 4), round(r["best_loss"], 5), r["epochs_run"], int(r["ex_count"].sum())))
        ref = []
        for sd in (1, 2, 3):
            log = []
            p, st = orc.lsm_global(Sn, 100.0, 0.05, 1.0, "put", orc.single_lsm_net_fit(variant, epochs=ep, seed=sd, dropout=drop,
                                   inference_dropout=drop > 0, log=log), target_ddof=1)
            ref.append((round(p, 4), round(min(log), 5), len(log), int(st["ex_count"].sum())))
        print("dropout", drop, variant, "engine", got, "\n   torch", ref, flush=True)
