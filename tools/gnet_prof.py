"""ncu target: one epoch of the global network LSM at BASELINE config-1 size.  argv: variant (gpu|cpu) [batch]."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa
from options_model_b200 import engine as E
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, **HP)
variant = sys.argv[1] if len(sys.argv) > 1 else "gpu"
S = eng.paths(model, 100_000, 50, "f32", E.RngSpec(seed=3))
r = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", variant=variant, epochs=1, seed=1, batch=int(sys.argv[2]) if len(sys.argv) > 2 else None)
print({k: v for k, v in r.items() if k not in ("boundary", "ex_count")}, eng.kernel_times())
