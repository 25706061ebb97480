"""Second ncu target: the bench step (batched paths + grouped sweep) and the kernels added after the first capture --
global network LSM (tcgen05), local-volatility paths, QE paths.  Small problem sizes where only per-launch metrics matter."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import options_model_b200  # noqa: E402,F401
from options_model_b200 import _lib as L  # noqa: E402
from options_model_b200 import engine as E  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, **HP)
for rep in range(2):  # rep 0 = warm-up, rep 1 = captured
    eng.price_american_batch(model, 1_000_000, 100.0, 100.0, 1.0, np.full(4, 252), 1, "f32", E.RngSpec(seed=1 + rep))
    Sg = eng.paths(E.gbm(100.0, 0.05, 1.0, 0.2), 40_000, 6, "f32", E.RngSpec(seed=7))
    eng.lsm_gnet(Sg, 100.0, 0.05, 1.0, "put", variant="gpu", epochs=1, seed=1, arrays=False, batch=8192)
    rs = np.random.default_rng(0)
    w = (0.1 * rs.standard_normal(3 * 64 + 4 * (64 * 64 + 3 * 64) + 64 + 1)).astype(np.float32)
    w[-1] = 0.2
    eng.paths_localvol(100.0, 0.05, 1.0, dict(hidden=64, layers=4, weights=w, m_scale=0.15, tau_scale=0.4, epsilon=1e-4),
                       100.0, 100_000, 50, "f32", E.RngSpec(seed=3))
    qe = E.heston(100.0, 0.05, 1.0, scheme=L.SCHEME_HESTON_QE, **HP)
    eng.paths(qe, 1_000_000, 50, "f32", E.RngSpec(seed=5))
    torch.cuda.synchronize()
print("profile_new done")
