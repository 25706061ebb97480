"""Size / shape stress of every entry point (development tool): checks that nothing crashes, AUTO picks a working
implementation and resident == split where both apply."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import options_model_b200  # noqa: E402,F401
from options_model_b200 import _lib as L  # noqa: E402
from options_model_b200 import engine as E  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)


def t(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label:70s} {1e3 * (time.perf_counter() - t0):9.2f} ms  {out}", flush=True)


h = E.heston(100.0, 0.05, 1.0, **HP)
g = E.gbm(100.0, 0.05, 1.0, 0.2)
for M, N, dt in ((8_000_000, 50, "f32"), (4_000_000, 20, "f32"), (4_090_000, 20, "f32"), (1_000_000, 50, "f64"),
                 (100_000, 1000, "f32"), (100_002, 7, "f32"), (3_000_000, 30, "f64")):
    def run(M=M, N=N, dt=dt):
        r = eng.price_american(h, M, N, 100.0, "put", dt, E.RngSpec(seed=M % 97))
        return f"price {r.price:.4f} impl {r.impl_used} launches {r.n_launches}"
    t(f"price_american Heston M={M} N={N} {dt}", run)
for M, n in ((300_000, 30), (27_000, 200), (1_500_000, 5), (5_000_000, 3), (64, 300)):
    def run(M=M, n=n):
        N = np.random.default_rng(1).integers(5, 60, n)
        p, se = eng.price_american_batch(h, M, 100.0, np.linspace(90, 110, n), np.linspace(0.1, 1.5, n), N, 1, "f32")
        return f"prices {p.min():.3f}..{p.max():.3f} finite {np.isfinite(p).all()}"
    t(f"price_american_batch M={M} n={n}", run)
S = eng.paths(g, 1_000_000, 600, "f64", E.RngSpec(seed=2))
t("lsm_global f64 1M x 600", lambda: {k: v for k, v in eng.lsm_global(S, 100.0, 0.05, 1.0, "put", arrays=True).items() if k in ("price", "rank", "n_rows")})
a = eng.lsm(S, 100.0, 0.05, 1.0, "put", impl="resident", arrays=False)
b = eng.lsm(S, 100.0, 0.05, 1.0, "put", impl="split", arrays=False)
print("resident vs split f64 1M x 600:", a.price, b.price, abs(a.price - b.price) / b.price)
del S
S = eng.paths(h, 2_000_000, 40, "f32", E.RngSpec(seed=3))
for H in (32, 128):
    t(f"lsm_mlp 2M x 40 hidden {H} textbook", lambda H=H: eng.lsm_mlp(S, 100.0, 0.05, 1.0, "put", "textbook", hidden=H, epochs=5, arrays=False).price)
K, T = np.meshgrid(np.linspace(60, 140, 64), np.linspace(0.05, 2, 32))
t("price_european_batch 2048 options x 100k x 64 (GBM)", lambda: eng.price_european_batch(g, 100_000, 64, K.ravel(), T.ravel(), np.ones(2048, dtype=np.int32))[0][:3])
# global network LSM: odd sizes, both dtypes, large row tables, partial tiles / partial batches
for M, N, dt, variant, bs in ((100_003 * 2, 37, "f32", "gpu", None), (50_000, 50, "f64", "cpu", 1000), (2_000_000, 100, "f32", "gpu", 131072),
                              (130, 3, "f64", "gpu", None), (1_000_000, 252, "f32", "gpu", 65536)):
    def run(M=M, N=N, dt=dt, variant=variant, bs=bs):
        Sx = eng.paths(h, M, N, dt, E.RngSpec(seed=5))
        r = eng.lsm_gnet(Sx, 100.0, 0.05, 1.0, "put", variant=variant, epochs=1, batch=bs, seed=2, arrays=True)
        return f"price {r['price']:.4f} rows {r['n_rows']} loss {r['best_loss']:.4f} launches {r['n_launches']}"
    t(f"lsm_gnet M={M} N={N} {dt} {variant} batch={bs}", run)
# local volatility: sizes around the one-wave partition, both widths and dtypes
rs = np.random.default_rng(0)
for Hn, Ln in ((64, 4), (32, 1), (64, 0), (32, 8)):
    w = (0.1 * rs.standard_normal(3 * Hn + Ln * (Hn * Hn + 3 * Hn) + Hn + 1)).astype(np.float32)
    w[-1] = 0.2
    net = dict(hidden=Hn, layers=Ln, weights=w, m_scale=0.15, tau_scale=0.4, epsilon=1e-4)
    for M, N, dt in ((2, 3, "f64"), (146, 5, "f32"), (148 * 512 + 2, 7, "f32"), (300_000, 20, "f64")):
        def run(M=M, N=N, dt=dt, net=net):
            Sx = eng.paths_localvol(100.0, 0.05, 1.0, net, 100.0, M, N, dt, E.RngSpec(seed=5))
            return f"S_T mean {float(Sx[N].double().mean()):.3f} finite {bool(torch.isfinite(Sx).all())}"
        t(f"paths_localvol H={Hn} L={Ln} M={M} N={N} {dt}", run)
print("stress done")
