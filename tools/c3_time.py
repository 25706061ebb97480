"""Timing of BASELINE config 3 (4 M Heston paths x 252 dates, per-date network fits on tcgen05), sweep only:
ContNet(1,128,128,1) (optmc_lsm_mlp) under reference and textbook semantics, SingleLSMNet(7,128,3) per date
(optmc_gnet_params.per_date).  argv[1] = dates (default 252)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa
from options_model_b200 import engine as E
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 252
S = eng.paths(E.heston(100.0, 0.05, 1.0, **HP), 4_000_000, N, "f32", E.RngSpec(seed=11))
def timed(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
for sem in ("reference", "textbook"):
    ms, r = timed(lambda: eng.lsm_mlp(S, 100.0, 0.05, 1.0, "put", sem, hidden=128, epochs=10, lr=1e-3, seed=1, arrays=True))
    rows = float(r.n_itm[1:N].sum())
    print(f"contnet128 {sem}: {ms:.1f} ms price {r.price:.4f} rows/date {rows / (N - 1):.0f} bf16-equivalent TFLOP/s {rows * 10 * 98304 * 1e-12 / (ms * 1e-3):.1f}")
if "--gnet" in sys.argv:
    ms, r = timed(lambda: eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", "reference", variant="gpu", per_date=1, epochs=10, batch=131072, stop_patience=0, seed=1, arrays=False))
    print(f"single_lsm_net per date reference: {ms:.1f} ms price {r['price']:.4f} rows {r['n_rows']}")
