#!/usr/bin/env python
"""bench.py -- path-steps/sec of the Heston American-put LSM hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--paths M] [--dates T]

Own arm (default): one process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A "step" is
one pass of the hot path over one batch of --batch (4) independent options of BASELINE config 2 (Heston
kappa=2 theta=0.04 xi=0.5 rho=-0.7, 252 steps, 1M paths each, quadratic-polynomial LSM): one batched K1 path
launch (Philox in-register, step-major fp32 slabs) + one grouped persistent LSM sweep launch (each option on
its own group of SMs) + the final payoff reductions, through optmc_price_american_batch.  With N GPUs every
rank prices its own batches (option-sharding, no data-path collective: weak scaling).  Timing: CUDA events on
the launching stream, barrier + synchronize on both sides, max over ranks; per-kernel times from CUDA events the
library records around its two launches.  The 4 GB of slabs are 30x the L2, so no explicit L2 flush is needed.

Reference arm (--impl reference): the reference is pure Python and cannot travel to the GPU box, so this
times its restatement (oracle/lsm_oracle.py, numpy) on the host cores, reference-style: a process pool
with one independent option per worker (options_model_3.py:1053-1056), each step a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
S0, K, R, T = 100.0, 100.0, 0.05, 1.0
METRIC = "path-steps/sec, Heston American put LSM"
UNIT = "path-steps/s"


def load_traffic(kernel_substr):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/ncu_r1_summary.json)."""
    p = os.path.join(ROOT, "profiles", "ncu_r1_summary.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for name, ks in d.get("kernels", {}).items():
            if kernel_substr in name:
                return float(ks[0]["dram_bytes"])
    except Exception:  # noqa: BLE001
        pass
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# CPU side: the oracle port (checker / baseline only)
# ------------------------------------------------------------------------------------------------------
def _cpu_price_one(args):
    seed, M, N = args
    import numpy as np

    from oracle import lsm_oracle as orc

    rng = np.random.default_rng(seed)
    Z1, Z2 = orc.draw_heston_normals(rng, N, M)
    S = orc.heston_paths_antithetic(S0, R, T, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
    return orc.lsm_sweep(S, K, R, T, "put").price


def cpu_baseline_single(M, N):
    _cpu_price_one((0, 2000, 8))  # imports + warm-up outside the timed sample
    t0 = time.perf_counter()
    price = _cpu_price_one((1, M, N))
    dt = time.perf_counter() - t0
    return {"value": M * N / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"numpy oracle (draws + Heston paths + poly2 LSM sweep), {M} paths x {N} steps, "
                      f"{dt:.1f} s, price {price:.4f}; host has {os.cpu_count()} cpus"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    cores = os.cpu_count() or 1
    M, N = args.ref_paths, args.dates
    with ProcessPoolExecutor(max_workers=cores) as ex:
        def step(i):
            return list(ex.map(_cpu_price_one, [(1000 * i + w, M, N) for w in range(cores)]))
        for i in range(args.warmup):
            step(i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(args.warmup + i)
        dt = time.perf_counter() - t0
    val = cores * M * N * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config 2 (Heston American put, poly2 LSM, {N} steps), bounded sample: "
                               f"{cores} independent options x {M} paths per step, one per worker process"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/lsm_oracle.py (numpy restatement of options_model_3.py:211-233,615-651 with the "
                                   f"8(c) polynomial regressor); {cores} worker processes x {M} paths x {N} steps per step"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.th = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.th:
            self.th.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch

    from options_model_b200 import _lib as L
    from options_model_b200 import compat
    from options_model_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    M, N, B = args.paths, args.dates, args.batch
    eng = E.Engine(local)
    stream = torch.cuda.Stream(device=local)
    model = E.heston(S0, R, T, **HP)
    b = 4  # fp32 storage
    import numpy as np

    Ns = np.full(B, N, dtype=np.int64)
    with torch.cuda.stream(stream):
        def step(i):
            # B fresh options (Philox streams) per step and rank; host scalars in, host prices out
            streams = (rank * 1_000_003 + i) * B + np.arange(B)
            return eng.price_american_batch(model, M, S0, K, T, Ns, 1, "f32", E.RngSpec(seed=42), "poly2", "reference",
                                            streams=streams)

        for i in range(args.warmup):
            step(i)
        barrier()
        l0 = eng.launch_count()
        ms_paths = ms_sweep = 0.0
        with ClockSampler(local) as clocks:
            t_start = torch.cuda.Event(enable_timing=True)
            t_end = torch.cuda.Event(enable_timing=True)
            t_start.record()
            for i in range(args.steps):
                price, se = step(args.warmup + i)
                kp, ks = eng.kernel_times()  # CUDA events around the kernels, recorded inside the library
                ms_paths += kp
                ms_sweep += ks
            t_end.record()
            barrier()
        launches = eng.launch_count() - l0
        ms_total = t_start.elapsed_time(t_end)
        ms_paths /= args.steps
        ms_sweep /= args.steps

        class _Res:
            pass

        res = _Res()
        res.price, res.stderr = float(price[0]), float(se[0])

        # end-to-end through the reference-facing call: host scalars in, host floats out, every step
        pricer = compat.AdvancedOptionPricer(K=K, r=R, sigma=None, option_type="put", rng_manager=compat.RNGManager(42),
                                             use_heston=True, heston_params=HP, use_control_variate=False,
                                             device=local)
        grid_S0, grid_T, grid_N = np.full(B, S0), np.full(B, T), Ns
        for _ in range(max(1, args.warmup // 2)):
            pricer.price_american_grid(grid_S0, grid_T, grid_N, M)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            price_e2e = float(pricer.price_american_grid(grid_S0, grid_T, grid_N, M)[0][0])
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()

    # path-sharding (second mode): ONE option of world x M paths, paths generated shard-local, the sweep
    # exchanging its per-date totals inside the kernel over NVLink peer memory.  Reported as context.
    sharded_info = None
    if world > 1 and not args.no_other_configs:
        from options_model_b200 import sharded as SH
        try:
            SH.init_peer_exchange(eng, dist)
            Mt = M * world
            off, m_loc = SH.shard_pairs(Mt, rank, world)

            def one():
                Sl = eng.paths(model, m_loc, N, "f32", E.RngSpec(seed=21, pair_offset=off))
                return SH.sweep_sharded_fused(eng, dist, Sl, Mt, K, R, T)

            with torch.cuda.stream(stream):
                for _ in range(3):
                    rs = one()
                stream.synchronize()
                barrier()
                e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0s.record(stream)
                reps = 10
                for _ in range(reps):
                    rs = one()
                e1s.record(stream)
                stream.synchronize()
            tt = torch.tensor([e0s.elapsed_time(e1s) / reps], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_sh = float(tt.item())
            sharded_info = {"workload": f"one Heston put of {Mt} paths x {N} dates path-sharded over {world} GPUs; per-date "
                                        "totals exchanged inside the persistent sweep kernel through NVLink peer memory",
                            "ms": ms_sh, "path_steps_per_s": Mt * N / ms_sh * 1e3, "price": rs.price, "collective_launches": 0}
            eng.comm_finalize()
        except Exception as e:  # noqa: BLE001 -- context only; the headline does not depend on it
            sharded_info = {"error": str(e)[:200]}
    barrier()

    # the other BASELINE configs, one short measurement each (rank 0 only; reported as context, not as `value`)
    others = {}
    if rank == 0 and not args.no_other_configs:
        def timed(fn, reps):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps, out

        with torch.cuda.stream(stream):
            gbm = E.gbm(S0, R, T, 0.2)
            dt1, r1 = timed(lambda: eng.price_american(gbm, 100_000, 50, K, "put", "f32", E.RngSpec(seed=7)), 20)
            others["config1_gbm_100k_x50"] = {"ms": dt1 * 1e3, "path_steps_per_s": 100_000 * 50 / dt1, "price": r1.price}
            # config 2 as ONE option at growing path counts (the latency -> bandwidth transition): 1 M and 4 M run the
            # persistent sweep (cash-flows in registers), 16 M exceeds its on-chip capacity and runs the split kernels
            for Mx in (1_000_000, 4_000_000, 16_000_000):
                dtx, rx = timed(lambda: eng.price_american(model, Mx, N, K, "put", "f32", E.RngSpec(seed=17)), 2)
                pk, sk = eng.kernel_times()
                others[f"config2_single_option_{Mx // 1_000_000}M_x252"] = {
                    "ms": dtx * 1e3, "path_steps_per_s": Mx * N / dtx, "paths_kernel_ms": pk, "sweep_ms": sk,
                    "sweep": "persistent" if rx.impl_used == L.SWEEP_RESIDENT else "split", "price": rx.price}
            Kg, Tg = np.meshgrid(np.linspace(70, 130, 32), np.linspace(1 / 12, 2, 32))
            n4 = 128  # one GPU's share of the 1024-option grid at 8 GPUs
            dt4, r4 = timed(lambda: eng.price_american_batch(model, 262_144, S0, Kg.ravel()[:n4], Tg.ravel()[:n4],
                                                             np.full(n4, 252), 1, "f32", E.RngSpec(seed=9)), 1)
            others["config4_128_options_x_256k_x252"] = {"ms": dt4 * 1e3, "path_steps_per_s": n4 * 262_144 * 252 / dt4,
                                                         "us_per_option": dt4 / n4 * 1e6}
            # local volatility: ImprovedIVNetwork(2 -> 64, 4 residual LayerNorm/GELU blocks) inside every step
            rs = np.random.default_rng(0)
            Hn, Ln = 64, 4
            wts = (0.1 * rs.standard_normal(3 * Hn + Ln * (Hn * Hn + 3 * Hn) + Hn + 1)).astype(np.float32)
            wts[-1] = 0.2
            ivn = dict(hidden=Hn, layers=Ln, weights=wts, m_scale=0.15, tau_scale=0.4, epsilon=1e-4)
            dlv, _ = timed(lambda: eng.paths_localvol(S0, R, T, ivn, K, 100_000, 50, "f32", E.RngSpec(seed=3)), 5)
            flop = 2.0 * (2 * Hn + Ln * Hn * Hn + Hn) * 100_000 * 50
            others["local_vol_paths_100k_x50_ivnet64x4"] = {"ms": dlv * 1e3, "path_steps_per_s": 5e6 / dlv,
                                                            "fp32_TFLOPs": flop / dlv / 1e12,
                                                            "note": "network evaluated per path per step on the CUDA cores (fp32 parity with the reference)"}
            Sg1 = eng.paths(gbm, 100_000, 50, "f32", E.RngSpec(seed=7))
            for variant in ("gpu", "cpu"):
                dtn, rn = timed(lambda: eng.lsm_gnet(Sg1, K, R, T, "put", "reference", variant=variant, epochs=25, seed=1,
                                                     arrays=False), 1)
                others[f"config1_global_network_lsm_{variant}_variant"] = {
                    "ms": dtn * 1e3, "rows": rn["n_rows"], "epochs_run": rn["epochs_run"], "price": rn["price"],
                    "note": "the reference's v3 algorithm: SingleLSMNet(7,128,3) on all dates' rows, tcgen05 forward/backward; "
                            + ("batch 8192 AdamW (om3gpu:740-798)" if variant == "gpu" else
                               "batch 256 Adam + ReduceLROnPlateau (om3:565-613; reference CPU: 174 s per epoch)")}
            del Sg1
            S4 = eng.paths(model, 4_000_000, 252, "f32", E.RngSpec(seed=11))
            dt3, r3 = timed(lambda: eng.lsm_mlp(S4, K, R, T, "put", "reference", hidden=128, epochs=10, lr=1e-3, seed=1,
                                                arrays=False), 1)
            others["config3_nn_lsm_4M_x252_hidden128_tcgen05"] = {"ms": dt3 * 1e3, "path_steps_per_s": 4_000_000 * 252 / dt3,
                                                                   "price": r3.price, "note": "sweep only (paths resident)"}
            del S4
            Kc, Tc = np.meshgrid(np.linspace(80, 120, 20), np.linspace(0.1, 1.0, 10))
            calib = E.heston(S0, R, T, scheme=L.SCHEME_HESTON_REF_CALIB, **HP)
            dt5, r5 = timed(lambda: eng.price_european_batch(calib, 50_000, 100, Kc.ravel(), Tc.ravel(),
                                                             np.zeros(200, dtype=np.int32), "f32", E.RngSpec(seed=13)), 3)
            others["config5_calibration_objective_200x50k_x100"] = {"ms": dt5 * 1e3, "path_steps_per_s": 1e9 / dt5}
            Sg = eng.paths(model, M, N, "f32", E.RngSpec(seed=15))
            dtg, rg = timed(lambda: eng.lsm_global(Sg, K, R, T, "put", arrays=False), 3)
            others["global_regression_lsm_1M_x252"] = {"ms": dtg * 1e3, "slab_GBps": 2 * b * M * (N + 1) / dtg / 1e9,
                                                       "price": rg["price"], "note": "two streaming passes over the slab"}
            del Sg
    barrier()

    if dist is not None:
        t = torch.tensor([ms_total, e2e_s * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1e3

    if rank == 0:
        import ctypes

        peak, peak_src = load_peaks()
        path_steps = world * B * M * N * args.steps
        value = path_steps / (ms_total * 1e-3)
        # algorithmic bytes: generation b per path-step (rows 0..N stored); sweep 3b per path
        # per exercise date (+ b for the terminal row); pipeline 4b per path-step.
        bytes_paths = B * b * M * (N + 1)                     # one batched launch generates B slabs
        bytes_sweep = B * (3 * b * M * (N - 1) + b * M)       # one grouped launch sweeps B options
        kern = {
            "paths_batch_kernel<f32,heston_ref_absorb,vec4>": {
                "ms": ms_paths, "alg_bytes": bytes_paths, "GBps": bytes_paths / (ms_paths * 1e-3) / 1e9,
                "dram_bytes_ncu": load_traffic("paths_batch_kernel")},
            "lsm_resident_kernel<f32,poly2,sparse>": {
                "ms": ms_sweep, "alg_bytes": bytes_sweep, "GBps": bytes_sweep / (ms_sweep * 1e-3) / 1e9,
                "dram_bytes_ncu": load_traffic("lsm_resident_kernel")},
        }
        dom = max(kern, key=lambda k: kern[k]["ms"])
        ach = kern[dom]["GBps"]
        cpu = cpu_baseline_single(args.cpu_paths, N) if world == 1 and not args.no_cpu else None
        h2d = ctypes.sizeof(L.ModelParams) + ctypes.sizeof(L.RngParams) + B * ctypes.sizeof(L.AmericanOption)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE config 2: American put, Heston (kappa=2, theta=0.04, xi=0.5, rho=-0.7), "
                                   f"{N} steps, {M} paths per option, poly2 LSM (reference semantics); a step prices a batch "
                                   f"of {B} independent options of that size per GPU in one grouped launch",
                       "batch": B,
                       "storage": "fp32 step-major slab, fp64 Gram/solve/decision", "rng": "Philox4x32-10 in-register",
                       "sharding": "options across ranks (no data-path collective)",
                       "l2": f"slabs {B * b * M * (N + 1) / 1e6:.0f} MB per step > L2 ({eng.l2_bytes / 1e6:.0f} MB): no flush needed",
                       "price": res.price, "stderr": res.stderr},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": kern[dom]["dram_bytes_ncu"], "kernel": dom, "peak_source": peak_src,
                         "note": "achieved = algorithmic bytes (SURVEY 8d: 3b per path per exercise date for the sweep, "
                                 "b per path-step for generation) / CUDA-event time of that kernel; traffic = DRAM bytes "
                                 "per launch from the committed ncu capture (profiles/)",
                         "dram_frac": (kern[dom]["dram_bytes_ncu"] / (kern[dom]["ms"] * 1e-3) / 1e9 / peak
                                       if kern[dom]["dram_bytes_ncu"] else None),
                         "dram_note": "dram_frac = measured DRAM traffic of the kernel / its time / peak: the sweep keeps the "
                                      "cash-flows in registers, so 2b of the 3b algorithmic bytes per path-date never reach HBM "
                                      "and frac (algorithmic) exceeds 1",
                         "pipeline_frac": (4 * b * B * M * N * world * args.steps) / (ms_total * 1e-3) / 1e9 / (peak * world),
                         "kernels": kern},
            "e2e": {"value": world * B * M * N * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 48 * B, "ms_per_step": e2e_ms / args.steps, "price": price_e2e,
                    "call": "compat.AdvancedOptionPricer.price_american_grid -> optmc_price_american_batch (host option "
                            "scalars in, host prices out; the path's inputs are option/model scalars, normals are "
                            "generated in-kernel)"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if sharded_info:
            others["path_sharded_single_option"] = sharded_info
        if others:
            line["other_configs"] = others
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--paths", type=int, default=1_000_000)
    ap.add_argument("--dates", type=int, default=252)
    ap.add_argument("--batch", type=int, default=4, help="independent options priced per step (one grouped launch)")
    ap.add_argument("--cpu-paths", type=int, default=1_000_000, help="CPU-baseline sample size (paths)")
    ap.add_argument("--ref-paths", type=int, default=50_000, help="reference arm: paths per worker per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of BASELINE configs 1/3/4/5")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
