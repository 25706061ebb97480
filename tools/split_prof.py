"""ncu target: the streaming (split) sweep on a 16 M-path option with few dates."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa
from options_model_b200 import engine as E
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
M, N = 16_000_000, 12
model = E.heston(100.0, 0.05, 1.0, **HP)
for _ in range(2):
    r = eng.price_american(model, M, N, 100.0, "put", "f32", E.RngSpec(seed=17))
    torch.cuda.synchronize()
print(r.price, r.impl_used, eng.kernel_times())
