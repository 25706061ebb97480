"""ncu target: per-date tcgen05 network LSM under textbook semantics (many tiles per CTA), few dates."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import options_model_b200
from options_model_b200 import engine as E
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
S = eng.paths(E.heston(100.0, 0.05, 1.0, **HP), 2_000_000, 6, "f32", E.RngSpec(seed=3))
for _ in range(2):
    r = eng.lsm_mlp(S, 100.0, 0.05, 1.0, "put", "textbook", hidden=128, epochs=5, arrays=False)
    torch.cuda.synchronize()
print(r.price)
