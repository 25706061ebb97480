import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import options_model_b200
from options_model_b200 import engine as E
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0); model = E.heston(100.0, 0.05, 1.0, **HP)
def timed(fn, reps=5):
    for _ in range(2): out = fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps, out
for B in (1, 4):
    ref = None
    for spec in ("1", "0"):
        os.environ["OPTMC_RES_SPEC"] = spec
        ms, (p, se) = timed(lambda: eng.price_american_batch(model, 1_000_000, 100.0, 100.0, 1.0, np.full(B, 252), 1, "f32", E.RngSpec(seed=3)))
        kp, ks = eng.kernel_times()
        same = ref is None or tuple(p) == ref; ref = ref or tuple(p)
        print(f"batch {B} spec={spec}: total {ms:.3f} paths {kp:.3f} sweep {ks:.3f} identical={same}", flush=True)
