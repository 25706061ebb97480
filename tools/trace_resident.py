"""Dump per-date phase clocks of the resident sweep (OPTMC_TRACE) and print a per-phase summary (cycles)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.txt"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 252
SEM = sys.argv[4] if len(sys.argv) > 4 else "reference"
from options_model_b200 import engine as E  # noqa: E402

eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
S = eng.paths(model, M, N, "f32", E.RngSpec(seed=1))
eng.lsm(S, 100.0, 0.05, 1.0, "put", semantics=SEM, impl="resident")  # warm
os.environ["OPTMC_TRACE"] = out
r = eng.lsm(S, 100.0, 0.05, 1.0, "put", semantics=SEM, impl="resident")
del os.environ["OPTMC_TRACE"]
print("price", r.price)
rows = np.loadtxt(out, comments="#")
# iteration t: stamps 0 pass start (beta visible, stage ready), 1 fused decide+Gram pass done, 2 CTA totals in
# warp 0 (sync A), 4 grid sum complete, 5 solved; the next iteration's stamp 0 closes sync B + stage wait.
cols = [0, 1, 2, 4, 5]
names = ["fused decide+gram pass", "block reduce (sync A)", "publish + grid sum", "solve", "sync B + stage wait"]
for c in (0, 1):
    a = rows[rows[:, 0] == c][:, 2:]
    a = a[(a[:, 0] > 0) & (a[:, 5] > 0)]
    per_date = a[1:, 0] - a[:-1, 0]
    print(f"cta {'first' if c == 0 else 'last'}: per-date total {np.median(per_date):.0f} cycles; spins median "
          f"{np.median(a[:, 7]):.0f} p90 {np.percentile(a[:, 7], 90):.0f}")
    d = np.diff(a[:, cols], axis=1)
    for k in range(4):
        print(f"   {names[k]:24s} median {np.median(d[:, k]):8.0f}  p90 {np.percentile(d[:, k], 90):8.0f}")
    nxt = a[1:, 0] - a[:-1, 5]
    print(f"   {names[4]:24s} median {np.median(nxt):8.0f}  p90 {np.percentile(nxt, 90):8.0f}")
