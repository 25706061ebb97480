"""Throughput of optmc_price_american_batch on the BASELINE batch shapes (development tool)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import options_model_b200  # noqa: E402,F401
from options_model_b200 import engine as E  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, **HP)


def run(name, M, S0, K, T, N, reps=3, sem="reference"):
    S0, K, T, N = np.broadcast_arrays(np.asarray(S0, float), np.asarray(K, float), np.asarray(T, float), np.asarray(N))
    n = S0.size
    eng.price_american_batch(model, M, S0, K, T, N, 1, "f32", E.RngSpec(seed=1), semantics=sem)  # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        p, se = eng.price_american_batch(model, M, S0, K, T, N, 1, "f32", E.RngSpec(seed=2 + r), semantics=sem)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    ps = float(np.sum(N)) * M
    kp, ks = eng.kernel_times()
    print(f"{name:48s} n={n:5d} M={M:8d} {dt * 1e3:9.2f} ms  {ps / dt / 1e9:8.1f} G path-steps/s  "
          f"{dt / n * 1e6:9.1f} us/option  kernels: paths {kp:.3f} ms sweep {ks:.3f} ms  price[0]={p[0]:.4f} +- {se[0]:.4f}", flush=True)


for n in (1, 2, 4, 8):
    run(f"config-2 options x{n}", 1_000_000, 100.0, 100.0, 1.0, np.full(n, 252))
run("config-2 x4 textbook", 1_000_000, 100.0, 100.0, 1.0, np.full(4, 252), sem="textbook")
Kg, Tg = np.meshgrid(np.linspace(70, 130, 8), np.linspace(1 / 12, 2, 7))
run("config-4 slice: 56 options x 256k", 262_144, 100.0, Kg.ravel(), Tg.ravel(), np.full(56, 252), reps=2)
days = np.arange(360, 0, -1.0)
run("curve driver: 360 points x 10k paths", 10_000, 100.0, 100.0, days / 365,
    np.maximum(10, np.minimum(130, np.ceil(days))).astype(int))
run("curve driver: 360 points x 100k paths", 100_000, 100.0, 100.0, days / 365,
    np.maximum(10, np.minimum(130, np.ceil(days))).astype(int), reps=2)

# ---- BASELINE config 1 (single GBM option, 100k x 50) and config 5 (calibration objective: 200 European options,
# 50k paths x 100 steps each, calibrator scheme hc:240-255, one fused no-store launch) -------------------------------
from options_model_b200 import _lib as L  # noqa: E402

gbm = E.gbm(100.0, 0.05, 1.0, 0.2)
for _ in range(3):
    r = eng.price_american(gbm, 100_000, 50, 100.0, "put", "f32", E.RngSpec(seed=3))
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    r = eng.price_american(gbm, 100_000, 50, 100.0, "put", "f32", E.RngSpec(seed=4 + i))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"config 1: GBM put 100k x 50 single call {dt * 1e6:8.1f} us  {100_000 * 50 / dt / 1e9:7.1f} G path-steps/s  "
      f"price {r.price:.4f} +- {r.stderr:.4f}", flush=True)

Kc, Tc = np.meshgrid(np.linspace(80, 120, 20), np.linspace(0.1, 1.0, 10))
calib = E.heston(100.0, 0.05, 1.0, scheme=L.SCHEME_HESTON_REF_CALIB, **HP)
for dtype in ("f32", "f64"):
    eng.price_european_batch(calib, 50_000, 100, Kc.ravel(), Tc.ravel(), np.zeros(200, dtype=np.int32), dtype, E.RngSpec(seed=1))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        m, se = eng.price_european_batch(calib, 50_000, 100, Kc.ravel(), Tc.ravel(), np.zeros(200, dtype=np.int32), dtype,
                                         E.RngSpec(seed=2 + i))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"config 5: 200 options x 50k x 100 ({dtype}) one objective evaluation {dt * 1e3:8.2f} ms  "
          f"{1e9 / dt / 1e9:7.1f} G path-steps/s  call[0] {m[0]:.4f} +- {se[0]:.4f}", flush=True)

# ---- global-regression LSM on a config-2 slab: two streaming passes ------------------------------------------------
S = eng.paths(model, 1_000_000, 252, "f32", E.RngSpec(seed=1))
for _ in range(2):
    g = eng.lsm_global(S, 100.0, 0.05, 1.0, "put", arrays=False)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    g = eng.lsm_global(S, 100.0, 0.05, 1.0, "put", arrays=False)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
p1, p2 = eng.kernel_times()
print(f"   pass 1 (moments + solve) {p1:.3f} ms, pass 2 (walk) {p2:.3f} ms")
print(f"global LSM (ref7 linear) 1M x 252 fp32: {ms:.3f} ms per sweep (2 slab reads = 2.02 GB -> {2.024 / ms:.2f} TB/s), "
      f"price {g['price']:.4f} rank {g['rank']} rows {g['n_rows']}", flush=True)

# ---- BASELINE config 3: per-date NN-LSM, 4M paths x 252 steps (hidden 128 on tcgen05; hidden 32 on CUDA cores) ----------
for Mn, H, sem, ep in ((1_000_000, 32, "reference", 10), (1_000_000, 128, "reference", 10), (4_000_000, 128, "reference", 10),
                       (1_000_000, 128, "textbook", 10)):
    S = eng.paths(model, Mn, 252, "f32", E.RngSpec(seed=5))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = eng.lsm_mlp(S, 100.0, 0.05, 1.0, "put", sem, hidden=H, epochs=ep, lr=1e-3, seed=1, arrays=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rows = int(r.n_itm.sum())
    print(f"config 3: NN-LSM {Mn} x 252, hidden {H}, {ep} epochs, {sem}: {dt * 1e3:9.1f} ms  price {r.price:.4f} +- {r.stderr:.4f}  "
          f"rows {rows}  {rows * ep / dt / 1e6:8.1f} M row-epochs/s  {Mn * 252 / dt / 1e9:6.2f} G path-steps/s", flush=True)
    del S
