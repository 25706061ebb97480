"""Thread-count sweep of the local-volatility path kernel (OPTMC_LV_NT) at three path counts."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa
from options_model_b200 import engine as E
eng = E.Engine(0)
rs = np.random.default_rng(0)
H, Ln = 64, 4
w = (0.1 * rs.standard_normal(3 * H + Ln * (H * H + 3 * H) + H + 1)).astype(np.float32); w[-1] = 0.2
net = dict(hidden=H, layers=Ln, weights=w, m_scale=0.15, tau_scale=0.4, epsilon=1e-4)
for M in (75776, 100_000, 1_000_000):
    for nt in (256, 384, 512):
        os.environ["OPTMC_LV_NT"] = str(nt)
        S = eng.alloc_slab(M, 50, "f32")
        eng.paths_localvol(100.0, 0.05, 1.0, net, 100.0, M, 50, "f32", E.RngSpec(seed=3), out=S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.paths_localvol(100.0, 0.05, 1.0, net, 100.0, M, 50, "f32", E.RngSpec(seed=3), out=S)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(M, nt, f"{ms:.2f} ms", f"{M*50/ms/1e6:.3f} G path-steps/s", f"{2*(2*H+Ln*H*H+H)*M*50/ms/1e9:.1f} TFLOP/s", flush=True)
        del S
