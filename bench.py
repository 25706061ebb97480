#!/usr/bin/env python
"""bench.py -- path-steps/sec of the Heston American-put LSM hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--shard paths|options] [--paths M] [--dates T]

Own arm (default): one process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A "step" is one pass of
the hot path over one batch of --batch (4) options of BASELINE config 2 (Heston kappa=2 theta=0.04 xi=0.5 rho=-0.7,
252 steps, quadratic-polynomial LSM, reference semantics) through optmc_price_american_batch(_ex): one batched K1 path
launch (Philox in-register, step-major fp32 slabs) + one grouped persistent LSM sweep launch + the final reductions.
  N = 1: every option has --paths (1 M) paths.
  N > 1, --shard paths (default; the north-star split): every option has N x 1 M paths, PATH-SHARDED -- each
         rank generates and sweeps its own 1 M-path block of each of the 4 options, and the per-date Gram totals are
         exchanged INSIDE the sweep kernel over NVLink peer memory (no host-launched collective on the data path).
         Per-GPU work is that of the N = 1 step (weak scaling).  Before timing, a smaller sharded batch is checked
         against the unsharded pricing of the same options on rank 0 (bit-identical prices, identical on all ranks);
         a mismatch fails the run.
  N > 1, --shard options: every rank prices its own batch of 1 M-path options (no collective at all).
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks; per-kernel times from
CUDA events the library records around its two launches.  The 4 GB of slabs per step are 30x the L2, so no explicit L2
flush is needed.

Reference arm (--impl reference): the reference is pure Python and cannot travel to the GPU box, so this times its
restatement (oracle/lsm_oracle.py, numpy) on the host cores, reference-style: a process pool with one independent
pricing per worker (options_model_3.py:1053-1056).  Same config as the own arm at N = 1 (4 x 1 M paths x 252 steps of
work per step), priced as a bounded sample: one slice of 4 M / cores paths per worker.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
S0, K, R, T = 100.0, 100.0, 0.05, 1.0
METRIC = "path-steps/sec, Heston American put LSM"
UNIT = "path-steps/s"


def workload_string(N, M, B):
    return (f"BASELINE config 2: American put, Heston (kappa=2, theta=0.04, xi=0.5, rho=-0.7), {N} steps, {M} paths per "
            f"option per GPU, poly2 LSM (reference semantics); a step prices a batch of {B} options of that size per GPU "
            f"in one grouped launch")


def load_traffic():
    """DRAM bytes per launch of the two bench kernels from the committed ncu capture of THIS round's kernels
    (profiles/ncu_r2_summary.json, written by tools/ncu_summary.py together with the git hash of the build)."""
    for name in ("ncu_r2_summary.json", "ncu_r1_summary.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            with open(p) as f:
                d = json.load(f)
        except Exception:  # noqa: BLE001
            continue
        out = {"source": f"profiles/{name}", "git": d.get("git"), "paths": None, "sweep": None}
        for kname, ks in d.get("kernels", {}).items():
            if "paths_x2_batch_kernel" in kname or (out["paths"] is None and "paths_batch_kernel" in kname):
                out["paths"] = float(ks[0]["dram_bytes"])
            if "lsm_resident_kernel" in kname and "36" in kname or (out["sweep"] is None and "lsm_resident_kernel" in kname):
                out["sweep"] = float(ks[0]["dram_bytes"])
        return out
    return {"source": None, "git": None, "paths": None, "sweep": None}


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
            "sample": f"numpy oracle (draws + Heston paths + poly2 LSM sweep), one option of {M} paths x {N} steps, "
                      f"{dt:.1f} s, price {price:.4f}; host has {os.cpu_count()} cpus"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    cores = os.cpu_count() or 1
    N, B = args.dates, args.batch
    # the step's work (B x --paths path-steps x N) as one slice per worker
    M = max(2, (B * args.paths // cores) // 2 * 2)
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
    sample = (f"oracle/lsm_oracle.py (numpy restatement of options_model_3.py:211-233,615-651 with the 8(c) polynomial "
              f"regressor) in a reference-style process pool: per step {cores} workers x one pricing of {M} paths x {N} steps "
              f"(= the step's {B} x {args.paths} paths, sliced one per core)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(N, args.paths, B), "batch": B, "reference_sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index, active=True):
        self.index, self.samples, self.reasons, self.stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.th = None
        self.nv = None
        if not active:
            return
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
            self.stop.wait(0.02)  # a sample costs the timed loop a GIL hand-over: 50 per second are enough for a median

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


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                              timeout=5).stdout.strip() or None
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import numpy as np
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

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    M, N, B = args.paths, args.dates, args.batch
    shard = "none" if world == 1 else ("paths" if args.shard in ("auto", "paths") else "options")
    eng = E.default_engine(local)  # the engine compat's pricers use as well (one context, one set of peer mappings)
    stream = torch.cuda.Stream(device=local)
    model = E.heston(S0, R, T, **HP)
    b = 4  # fp32 storage
    Ns = np.full(B, N, dtype=np.int64)
    parity = None
    if shard == "paths":
        from options_model_b200 import sharded as SH

        SH.init_peer_exchange(eng, dist)
        # ---- correctness of the fused exchange before anything is timed: a smaller sharded batch against the unsharded
        # pricing of the SAME options (same Philox counters: pair_offset) on rank 0, and agreement across the ranks
        Ms, Nsm, Bs = 65536, 60, 2
        with torch.cuda.stream(stream):
            ps, _, _ = eng.price_american_batch(model, Ms, S0, K, T, np.full(Bs, Nsm), 1, "f32",
                                                E.RngSpec(seed=77, pair_offset=rank * (Ms // 2)), streams=np.arange(Bs) + 5,
                                                M_total=Ms * world)
            stream.synchronize()
            everyone = [None] * world
            dist.all_gather_object(everyone, [float(x) for x in ps])
            single = None
            if rank == 0:
                p1, _ = eng.price_american_batch(model, Ms * world, S0, K, T, np.full(Bs, Nsm), 1, "f32", E.RngSpec(seed=77),
                                                 streams=np.arange(Bs) + 5)
                single = [float(x) for x in p1]
            holder = [single]
            dist.broadcast_object_list(holder, src=0)
            single = holder[0]
        rel = max(abs(a - c) / abs(c) for a, c in zip(everyone[0], single))
        parity = {"rel_vs_single": rel, "identical_on_all_ranks": all(e == everyone[0] for e in everyone),
                  "checked": f"{Bs} options x {Ms * world} paths x {Nsm} dates, sharded over {world} ranks vs one GPU",
                  "sharded_prices": everyone[0], "single_gpu_prices": single}
    with torch.cuda.stream(stream):
        def step(i):
            # B fresh options (Philox streams) per step; host scalars in, host prices out
            if shard == "paths":  # the same options on every rank, each rank its own block of paths
                out = eng.price_american_batch(model, M, S0, K, T, Ns, 1, "f32", E.RngSpec(seed=42, pair_offset=rank * (M // 2)),
                                               "poly2", "reference", streams=i * B + np.arange(B), M_total=M * world)
                return out[0], out[1]
            streams = (rank * 1_000_003 + i) * B + np.arange(B)
            return eng.price_american_batch(model, M, S0, K, T, Ns, 1, "f32", E.RngSpec(seed=42), "poly2", "reference",
                                            streams=streams)

        tuned = None
        if shard == "paths" and "OPTMC_RES_SPEC" not in os.environ:
            # the sharded batch cannot time its two sweep kernels inside one call (every rank must launch the same one):
            # time both here, max over ranks, and keep the faster for the rest of the run (OPTMC_RES_SPEC is the
            # library's kernel override; the results are bit-identical either way)
            tuned = {}
            for spec in ("0", "1"):
                os.environ["OPTMC_RES_SPEC"] = spec
                step(0)
                barrier()
                t0 = time.perf_counter()
                for i in range(3):
                    step(i)
                torch.cuda.synchronize()
                tuned["speculative" if spec == "1" else "single_role"] = max_over_ranks((time.perf_counter() - t0) / 3 * 1e3)
            os.environ["OPTMC_RES_SPEC"] = "1" if tuned["speculative"] < tuned["single_role"] else "0"
            tuned["kept"] = "speculative" if os.environ["OPTMC_RES_SPEC"] == "1" else "single_role"
        for i in range(args.warmup):
            step(i)
        # which sweep kernel the batch entry point kept for this box (it times both on the first wave of a new shape)
        shape = eng.price_american_batch(model, M, S0, K, T, Ns, 1, "f32", E.RngSpec(seed=42, pair_offset=rank * (M // 2)),
                                         "poly2", "reference", streams=np.arange(B), details=True,
                                         M_total=M * world if shard == "paths" else 0)[2]["shape"]
        barrier()
        l0 = eng.launch_count()
        ms_paths = ms_sweep = 0.0
        # only rank 0 samples (its line carries `clocks`): eight processes polling NVML every 5 ms during the timed region cost
        # the 8-GPU step 5 % (3.85 vs 3.66 ms through the sampler-free e2e loop)
        with ClockSampler(local, active=(rank == 0)) as clocks:
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
        ms_total = max_over_ranks(t_start.elapsed_time(t_end))
        ms_paths /= args.steps
        ms_sweep /= args.steps
        price0, se0 = float(price[0]), float(se[0])

        # end-to-end through the reference-facing call: host scalars in, host floats out, every step
        pricer = compat.AdvancedOptionPricer(K=K, r=R, sigma=None, option_type="put", rng_manager=compat.RNGManager(42),
                                             use_heston=True, heston_params=HP, use_control_variate=False, device=local,
                                             path_shard=(rank, world) if shard == "paths" else None)
        grid_S0, grid_T, grid_N = np.full(B, S0), np.full(B, T), Ns
        for _ in range(max(1, args.warmup // 2)):
            pricer.price_american_grid(grid_S0, grid_T, grid_N, M)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            price_e2e = float(pricer.price_american_grid(grid_S0, grid_T, grid_N, M)[0][0])
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        barrier()

    others = {}
    peak, peak_src = load_peaks()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, out

    if not args.no_other_configs:
        # ---- BASELINE configs 4 and 5 as stated: sharded over the ranks by OPTION, no data-path collective; every rank
        # participates, the time is the slowest rank's ----
        with torch.cuda.stream(stream):
            Kg, Tg = np.meshgrid(np.linspace(70, 130, 32), np.linspace(1 / 12, 2, 32))
            Kg, Tg = Kg.ravel(), Tg.ravel()
            mine = np.arange(rank, 1024, world)
            barrier()
            dt4, _ = timed(lambda: eng.price_american_batch(model, 262_144, S0, Kg[mine], Tg[mine], np.full(mine.size, 252), 1,
                                                            "f32", E.RngSpec(seed=9), streams=mine), 1)
            dt4 = max_over_ranks(dt4)
            others["config4_1024_options_x_256k_x252"] = {
                "ms": dt4 * 1e3, "path_steps_per_s": 1024 * 262_144 * 252 / dt4, "us_per_option": dt4 / 1024 * 1e6 * world,
                "options_per_rank": int(mine.size), "sharding": f"options rank::{world}, no collective"}
            Kc, Tc = np.meshgrid(np.linspace(80, 120, 20), np.linspace(0.1, 1.0, 10))
            cfg = compat.CalibrationConfig(n_mc_paths=50_000, n_time_steps=100, seed=13, verbose=False, plot_results=False)
            gather = None
            if dist is not None:
                def gather(x):
                    parts = [None] * world
                    dist.all_gather_object(parts, x)
                    return parts
            obj = compat.HestonObjective(compat.HestonPricer(cfg, dtype="f32", device=local), S0, R, Kc.ravel(), Tc.ravel(),
                                         np.full(200, 0.2), common_random_numbers=True, rank=rank, world=world,
                                         all_gather=gather)
            x0 = np.array([HP["kappa"], HP["theta"], HP["xi"], HP["rho"], HP["v0"]])
            barrier()
            dt5, f5 = timed(lambda: obj(x0), 3)
            dt5 = max_over_ranks(dt5)
            others["config5_calibration_objective_200x50k_x100"] = {
                "ms": dt5 * 1e3, "path_steps_per_s": 1e9 / dt5, "objective": float(f5),
                "call": "compat.HestonObjective (heston_calibration.py:404-472), rows rank::world, one fused launch per rank + "
                        "one all_gather of the prices"}
            # ---- the reference's global network regression (om3:565-613) with the PATHS sharded over the ranks: BASELINE
            # config 1's shape per rank (100 k GBM paths x 50 dates, batch 8192 per rank), gradients exchanged through peer
            # memory inside the reduce / optimiser kernels (optmc_lsm_gnet_sharded); weak scaling of the epoch time ----
            if shard in ("none", "paths"):
                try:
                    Mg, Ng = 100_000, 50
                    Sgn = eng.paths(E.gbm(S0, R, T, 0.2), Mg, Ng, "f32", E.RngSpec(seed=7, pair_offset=rank * (Mg // 2)))
                    kwg = dict(variant="gpu", batch=8192 * world, seed=1, arrays=False, stop_patience=0)

                    def fit(ep):
                        if shard == "paths":
                            return SH.gnet_sharded(eng, dist, Sgn, Mg * world, K, R, T, "put", "reference", epochs=ep, **kwg)
                        return eng.lsm_gnet(Sgn, K, R, T, "put", "reference", epochs=ep, **kwg)

                    fit(1)
                    tg = {}
                    for ep in (1, 4):
                        barrier()
                        dtg_, rgn = timed(lambda: fit(ep), 1)
                        tg[ep] = max_over_ranks(dtg_)
                    sig = [None] * world
                    if dist is not None:
                        dist.all_gather_object(sig, (rgn["price"], rgn["best_loss"]))
                    steps_ep = -(-rgn["n_rows"] // (8192 * world))
                    others["global_network_lsm_path_sharded_100k_x50_per_rank"] = {
                        "epoch_ms": (tg[4] - tg[1]) / 3 * 1e3, "us_per_optimiser_step": (tg[4] - tg[1]) / 3 / steps_ep * 1e6,
                        "rows_all_ranks": rgn["n_rows"], "global_batch": 8192 * world, "steps_per_epoch": steps_ep,
                        "rows_per_s": rgn["n_rows"] / ((tg[4] - tg[1]) / 3), "best_loss": rgn["best_loss"], "price": rgn["price"],
                        "identical_on_all_ranks": all(x == sig[0] for x in sig) if dist is not None else True,
                        "note": "SingleLSMNet(7,128,3) on tcgen05; per step each rank pushes its 34 178-word gradient vector into "
                                "every peer's memory over NVLink and the optimiser kernel sums the ranks' vectors in rank order"}
                    del Sgn
                except Exception as e:  # noqa: BLE001 -- context only; gnet_sharded raises on all ranks or on none
                    others["global_network_lsm_path_sharded_100k_x50_per_rank"] = {"error": str(e)[:200]}
    barrier()

    # the other BASELINE configs on one GPU (rank 0 of a single-GPU run; reported as context, not as `value`)
    if rank == 0 and world == 1 and not args.no_other_configs:
        with torch.cuda.stream(stream):
            gbm = E.gbm(S0, R, T, 0.2)
            dt1, r1 = timed(lambda: eng.price_american(gbm, 100_000, 50, K, "put", "f32", E.RngSpec(seed=7)), 20)
            others["config1_gbm_100k_x50"] = {"ms": dt1 * 1e3, "path_steps_per_s": 100_000 * 50 / dt1, "price": r1.price}
            # config 2 AS STATED: one option of 1 M paths (the per-date dependency chain, not bandwidth, bounds it), both semantics;
            # then growing path counts (latency -> bandwidth): 4 M persistent, 16 M split sweep
            for Mx, sems in ((1_000_000, ("reference", "textbook")), (4_000_000, ("reference",)), (16_000_000, ("reference",))):
                for sem in sems:
                    dtx, rx = timed(lambda: eng.price_american(model, Mx, N, K, "put", "f32", E.RngSpec(seed=17), semantics=sem), 3)
                    pk, sk = eng.kernel_times()
                    others[f"config2_single_option_{Mx // 1_000_000}M_x252" + ("" if sem == "reference" else "_textbook")] = {
                        "ms": dtx * 1e3, "path_steps_per_s": Mx * N / dtx, "paths_kernel_ms": pk, "sweep_ms": sk,
                        "pipeline_frac": 4 * b * Mx * N / (dtx * 1e9) / peak,
                        "paths_frac": b * Mx * (N + 1) / (pk * 1e6) / peak if pk else None,
                        "sweep_frac": (3 * b * Mx * (N - 1) + b * Mx) / (sk * 1e6) / peak if sk else None,
                        "sweep": "persistent" if rx.impl_used == L.SWEEP_RESIDENT else "split", "price": rx.price}
            dtt, (pt, _) = timed(lambda: eng.price_american_batch(model, M, S0, K, T, Ns, 1, "f32", E.RngSpec(seed=42),
                                                                 "poly2", "textbook"), 3)
            pk, sk = eng.kernel_times()
            others["config2_batch4_textbook"] = {"ms": dtt * 1e3, "path_steps_per_s": B * M * N / dtt, "paths_kernel_ms": pk,
                                                 "sweep_ms": sk, "pipeline_frac": 4 * b * B * M * N / (dtt * 1e9) / peak,
                                                 "price": float(pt[0]),
                                                 "note": "textbook semantics (no sticky mask: ~40% of the paths in every regression, dense passes)"}
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
            dt3s, r3s = timed(lambda: eng.lsm_gnet(S4, K, R, T, "put", "reference", variant="gpu", per_date=1, epochs=10, batch=131072,
                                                   stop_patience=0, seed=1, arrays=False), 1)
            others["config3_single_lsm_net_per_date_4M_x252_tcgen05"] = {
                "ms": dt3s * 1e3, "path_steps_per_s": 4_000_000 * 252 / dt3s, "price": r3s["price"], "rows": r3s["n_rows"],
                "optimiser_steps": r3s["epochs_run"],
                "note": "a fresh SingleLSMNet(7,128,3) per exercise date (optmc_gnet_params.per_date), 10 full-batch AdamW steps per "
                        "date, in-sample decision; sweep only (paths resident)"}
            del S4
            Sg = eng.paths(model, M, N, "f32", E.RngSpec(seed=15))
            dtg, rg = timed(lambda: eng.lsm_global(Sg, K, R, T, "put", arrays=False), 3)
            others["global_regression_lsm_1M_x252"] = {"ms": dtg * 1e3, "slab_GBps": 2 * b * M * (N + 1) / dtg / 1e9,
                                                       "price": rg["price"], "note": "two streaming passes over the slab"}
            del Sg
    barrier()

    rc = 0
    if rank == 0:
        import ctypes

        traffic = load_traffic()
        path_steps = world * B * M * N * args.steps
        value = path_steps / (ms_total * 1e-3)
        # algorithmic bytes: generation b per path-step (rows 0..N stored); sweep 3b per path
        # per exercise date (+ b for the terminal row); pipeline 4b per path-step.
        bytes_paths = B * b * M * (N + 1)                     # one batched launch generates B slabs
        bytes_sweep = B * (3 * b * M * (N - 1) + b * M)       # one grouped launch sweeps B options
        kern = {
            "paths_x2_batch_kernel<heston_ref_absorb> (f32x2 packed step, 3 steps per Philox block)": {
                "ms": ms_paths, "alg_bytes": bytes_paths, "GBps": bytes_paths / (ms_paths * 1e-3) / 1e9,
                "frac": bytes_paths / (ms_paths * 1e-3) / 1e9 / peak, "dram_bytes_ncu": traffic["paths"]},
            (f"lsm_resident_spec_kernel<f32,poly2,{shape[1]},{shape[0]}> (grouped, {shape[3]} x {shape[2]} CTAs, speculative)"
             if shape[0] in (480, 736) else
             f"lsm_resident_kernel<f32,poly2,{shape[1]},{shape[0]},sparse> (grouped, {shape[3]} x {shape[2]} CTAs)"): {
                "ms": ms_sweep, "alg_bytes": bytes_sweep, "GBps": bytes_sweep / (ms_sweep * 1e-3) / 1e9,
                "frac": bytes_sweep / (ms_sweep * 1e-3) / 1e9 / peak, "dram_bytes_ncu": traffic["sweep"]},
        }
        dom = max(kern, key=lambda k: kern[k]["ms"])
        ach = kern[dom]["GBps"]
        cpu = cpu_baseline_single(args.cpu_paths, N) if world == 1 and not args.no_cpu else None
        h2d = ctypes.sizeof(L.ModelParams) + ctypes.sizeof(L.RngParams) + B * ctypes.sizeof(L.AmericanOption)
        sharding = {"none": "single GPU",
                    "paths": f"PATHS of every option across the {world} ranks (each option {world * M} paths); per-date Gram "
                             "totals exchanged inside the sweep kernel over NVLink peer memory, no collective launch",
                    "options": "options across ranks (no data-path collective)"}[shard]
        single = others.get("config2_single_option_1M_x252")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(N, M, B), "batch": B,
                       "storage": "fp32 step-major slab, fp64 Gram/solve/decision", "rng": "Philox4x32-10 in-register",
                       "sharding": sharding, "shard_mode": shard,
                       "l2": f"slabs {B * b * M * (N + 1) / 1e6:.0f} MB per step > L2 ({eng.l2_bytes / 1e6:.0f} MB): no flush needed",
                       "price": price0, "stderr": se0, "git": git_head(),
                       "sharded_sweep_kernel_ms_per_step": tuned,
                       "sweep_shape": {"threads": shape[0], "paths_per_thread": shape[1], "ctas_per_option": shape[2],
                                       "options_per_launch": shape[3]}},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": kern[dom]["dram_bytes_ncu"], "kernel": dom, "peak_source": peak_src,
                         "traffic_source": f"{traffic['source']} (ncu dram__bytes_read+write per launch, build {traffic['git']})",
                         "note": "achieved = algorithmic bytes (SURVEY 8d: 3b per path per exercise date for the sweep, "
                                 "b per path-step for generation) / CUDA-event time of that kernel, measured in this run",
                         "dram_frac": (kern[dom]["dram_bytes_ncu"] / (kern[dom]["ms"] * 1e-3) / 1e9 / peak
                                       if kern[dom]["dram_bytes_ncu"] else None),
                         "dram_note": "dram_frac = ncu DRAM traffic of the kernel / its time / peak: the sweep keeps the "
                                      "cash-flows in registers, so 2b of the 3b algorithmic bytes per path-date never reach HBM "
                                      "and frac (algorithmic) exceeds 1; the path kernel is bound by the FMA (Philox IMAD.WIDE) "
                                      "and MUFU pipes, not by HBM (profiles/paths_bench_r2.txt)",
                         "pipeline_frac": (4 * b * B * M * N * world * args.steps) / (ms_total * 1e-3) / 1e9 / (peak * world),
                         "pipeline_frac_single_option": single["pipeline_frac"] if single else None,
                         "kernels": kern},
            "e2e": {"value": world * B * M * N * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 48 * B, "ms_per_step": e2e_ms / args.steps, "price": price_e2e,
                    "call": "compat.AdvancedOptionPricer.price_american_grid -> optmc_price_american_batch(_ex) (host option "
                            "scalars in, host prices out; the path's inputs are option/model scalars, normals are "
                            "generated in-kernel)"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if parity:
            line["path_sharded_parity"] = parity
            if not (parity["rel_vs_single"] == 0.0 and parity["identical_on_all_ranks"]):
                rc = 1
        if others:
            line["other_configs"] = others
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
        if rc:
            print("bench.py: path-sharded prices differ from the single-GPU prices", file=sys.stderr, flush=True)
    if dist is not None:
        flag = torch.tensor([rc], device="cuda")
        dist.broadcast(flag, src=0)
        rc = int(flag[0])
        if shard == "paths":
            eng.comm_finalize()
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--shard", default="auto", choices=["auto", "paths", "options"],
                    help="N > 1: split every option's paths over the ranks (default) or give every rank its own options")
    ap.add_argument("--paths", type=int, default=1_000_000, help="paths per option per GPU")
    ap.add_argument("--dates", type=int, default=252)
    ap.add_argument("--batch", type=int, default=4, help="options priced per step (one grouped launch)")
    ap.add_argument("--cpu-paths", type=int, default=1_000_000, help="CPU-baseline sample size (paths)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
