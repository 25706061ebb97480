"""ncu target: BASELINE config 3 (per-date tcgen05 network LSM, 4 M paths) with a reduced number of dates.  argv[1] = dates."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa
from options_model_b200 import engine as E
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
S = eng.paths(E.heston(100.0, 0.05, 1.0, **HP), 4_000_000, N, "f32", E.RngSpec(seed=11))
for _ in range(2):
    r = eng.lsm_mlp(S, 100.0, 0.05, 1.0, "put", "reference", hidden=128, epochs=10, lr=1e-3, seed=1, arrays=True)
    torch.cuda.synchronize()
print(r.price, r.n_itm[1:N].mean(), eng.kernel_times())
