"""Dump per-date phase clocks of the resident sweep (OPTMC_TRACE) and print a per-phase summary (SM cycles).

  python tools/trace_resident.py [out] [M] [N] [semantics] [max_ctas]

Speculative kernel (sticky semantics, default): stamps of thread 96 (compute) 6 = S(t) starts, 1 = S(t) done, 2 = beta_t seen (BETA
barrier passed), 3 = C(t) + warp reduce done; of the communication warp 0 = TOT barrier passed (all warps done with date
t), 4 = grid-wide sum complete, 5 = beta_(t-1) solved.  Single-role kernel (OPTMC_RES_SPEC=0 or textbook semantics):
0 pass start, 1 pass done, 2 CTA totals in warp 0, 4 grid sum complete, 5 solved.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.txt"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 252
SEM = sys.argv[4] if len(sys.argv) > 4 else "reference"
if len(sys.argv) > 5:
    os.environ["OPTMC_RES_MAXCTAS"] = sys.argv[5]
from options_model_b200 import engine as E  # noqa: E402

eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
S = eng.paths(model, M, N, "f32", E.RngSpec(seed=1))
eng.lsm(S, 100.0, 0.05, 1.0, "put", semantics=SEM, impl="resident")  # warm
os.environ["OPTMC_TRACE"] = out
r = eng.lsm(S, 100.0, 0.05, 1.0, "put", semantics=SEM, impl="resident")
del os.environ["OPTMC_TRACE"]
print("price", r.price, open(out).readline().strip())
rows = np.loadtxt(out, comments="#")
spec = SEM == "reference" and os.environ.get("OPTMC_RES_SPEC", "1") != "0"


def med(x):
    return f"median {np.median(x):8.0f}  p90 {np.percentile(x, 90):8.0f}"


for c in (0, 1):
    a = rows[rows[:, 0] == c][:, 2:]  # rows ordered t = N .. 1
    a = a[(a[:, 5] > 0)]
    if spec:
        a = a[(a[:, 1] > 0) & (a[:, 3] > 0)]
        per_date = a[1:, 5] - a[:-1, 5]
        print(f"cta {'first' if c == 0 else 'last'}: per-date total {med(per_date)} cycles; spins {med(a[:, 7])}")
        print(f"   S(t) pass                  (1 - 6)        {med(a[:, 1] - a[:, 6])}")
        print(f"   arrive TOT + stats         (12 - 3)       {med(a[1:, 12] - a[1:, 3])}")
        print(f"   cand_apply                 (11 - 12)      {med(a[1:, 11] - a[1:, 12])}")
        print(f"   loop top + stage wait      (6[t-1] - 11[t]) {med(a[2:, 6] - a[1:-1, 11])}")
        print(f"   BETA arrive -> beta seen   (2 - 10)       {med(a[:, 2] - a[:, 10])}")
        print(f"   C(t) candidates            (8 - 2)        {med(a[1:, 8] - a[1:, 2])}")
        print(f"   stats + warp reduce        (3 - 8)        {med(a[1:, 3] - a[1:, 8])}")
        print(f"   CTA totals                 (9 - 0)        {med(a[:, 9] - a[:, 0])}")
        print(f"   grid sum                   (4 - 9)        {med(a[:, 4] - a[:, 9])}")
        print(f"   C(t) + warp reduce        (3 - 2)        {med(a[:, 3] - a[:, 2])}")
        print(f"   TOT barrier (slowest warp) (0 - 3)        {med(a[:, 0] - a[:, 3])}")
        print(f"   CTA totals + grid sum      (4 - 0)        {med(a[:, 4] - a[:, 0])}")
        print(f"   solve                      (5 - 4)        {med(a[:, 5] - a[:, 4])}")
        print(f"   beta_t seen after solved   (2[t-1] - 5[t]) {med(a[1:, 2] - a[:-1, 5])}")
        print(f"   S(t-1) done after TOT(t)   (1[t-1] - 0[t]) {med(a[1:, 1] - a[:-1, 0])}   (negative: S finished before)")
        print(f"   S(t-1) done vs beta solved (1[t-1] - 5[t]) {med(a[1:, 1] - a[:-1, 5])}   (positive: the pass, not the exchange, is critical)")
    else:
        a = a[(a[:, 0] > 0)]
        per_date = a[1:, 0] - a[:-1, 0]
        print(f"cta {'first' if c == 0 else 'last'}: per-date total {med(per_date)} cycles; spins {med(a[:, 7])}")
        cols = [0, 1, 2, 4, 5]
        names = ["fused decide+gram pass", "block reduce (sync A)", "publish + grid sum", "solve"]
        d = np.diff(a[:, cols], axis=1)
        for k in range(4):
            print(f"   {names[k]:24s} {med(d[:, k])}")
        print(f"   {'sync B + stage wait':24s} {med(a[1:, 0] - a[:-1, 5])}")
