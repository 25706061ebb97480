"""Persistent-sweep A/B (development tool): the speculative warp-specialised kernel against the single-role sparse
kernel (OPTMC_RES_SPEC=0) and the split sweep, single option and grouped batches; prices / exercise counts must agree
exactly between the resident variants (same arithmetic per path, order-independent grid sums)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import options_model_b200  # noqa: E402,F401
from options_model_b200 import engine as E  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, **HP)
N = 252


def timed(fn, reps=10):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


for M in (100_000, 1_000_000, 4_000_000):
    S = eng.paths(model, M, N, "f32", E.RngSpec(seed=1))
    ref = None
    for label, env in (("default", dict(OPTMC_RES_SPEC=None)), ("single-role", dict(OPTMC_RES_SPEC=0))):
        setenv(**env)
        for sem in ("reference",):
            ms, r = timed(lambda: eng.lsm(S, 100.0, 0.05, 1.0, "put", semantics=sem, impl="resident"))
            key = (r.price, tuple(r.ex_count.tolist()))
            same = ref is None or key == ref
            ref = ref or key
            print(f"single option M={M:8d} {sem:9s} {label:12s} sweep {ms:7.3f} ms  {M * N / ms / 1e6:7.1f} G path-steps/s  "
                  f"price {r.price:.6f}  identical={same}", flush=True)
    setenv(OPTMC_RES_SPEC=None)
    ms, r = timed(lambda: eng.lsm(S, 100.0, 0.05, 1.0, "put", semantics="textbook", impl="resident"))
    print(f"single option M={M:8d} textbook  dense        sweep {ms:7.3f} ms  price {r.price:.6f}", flush=True)
    del S

for B in (1, 2, 4, 8):
    ref = None
    for label, env in (("default", dict(OPTMC_RES_SPEC=None)), ("single-role", dict(OPTMC_RES_SPEC=0))):
        setenv(**env)
        ms, (p, se) = timed(lambda: eng.price_american_batch(model, 1_000_000, 100.0, 100.0, 1.0, np.full(B, N), 1, "f32",
                                                               E.RngSpec(seed=3)), reps=5)
        kp, ks = eng.kernel_times()
        key = tuple(p.tolist())
        same = ref is None or key == ref
        ref = ref or key
        print(f"batch of {B} x 1M {label:12s} total {ms:7.3f} ms  paths {kp:.3f} sweep {ks:.3f} ms  "
              f"{B * 1e6 * N / ms / 1e6:7.1f} G path-steps/s  price[0] {p[0]:.6f} identical={same}", flush=True)
setenv(OPTMC_RES_SPEC=1)
for cap in (74, 49):
    setenv(OPTMC_BATCH_CPG=cap)
    ms, (p, se) = timed(lambda: eng.price_american_batch(model, 1_000_000, 100.0, 100.0, 1.0, np.full(4, N), 1, "f32",
                                                           E.RngSpec(seed=3)), reps=5)
    kp, ks = eng.kernel_times()
    print(f"batch of 4 x 1M spec, {cap} CTAs per option: total {ms:7.3f} ms  paths {kp:.3f} sweep {ks:.3f} ms  price[0] {p[0]:.6f}",
          flush=True)
setenv(OPTMC_BATCH_CPG=None)
