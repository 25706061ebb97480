"""Path-sharded sweep with the in-kernel NVLink exchange (optmc_comm_* / optmc_lsm_poly_sharded), checked against
the single-GPU sweep and the host-loop NCCL variant, and timed.  Launch with one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_fused_check.py [--paths 2000000] [--dates 252] [--reps 20]

Exit code 0 = all checks passed.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa: E402,F401
from options_model_b200 import engine as E  # noqa: E402
from options_model_b200 import sharded  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=int, default=2_000_000)
    ap.add_argument("--dates", type=int, default=252)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--check-paths", type=int, default=200_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = E.Engine(local)
    sharded.init_peer_exchange(eng, dist)
    out = {"world": world}
    ok = True

    # ---- correctness: fp64 and fp32, both semantics, against the single-GPU sweep on the same Philox paths ----
    for dtype, tol in (("f64", 1e-11), ("f32", 1e-5)):
        for semantics in ("reference", "textbook"):
            M, N = args.check_paths, 50
            model = E.heston(100.0, 0.05, 1.0, **HP)
            off, m_loc = sharded.shard_pairs(M, rank, world)
            S_loc = eng.paths(model, m_loc, N, dtype, E.RngSpec(seed=17, pair_offset=off))
            fused = sharded.sweep_sharded_fused(eng, dist, S_loc, M, 100.0, 0.05, 1.0, "put", "poly2", semantics, arrays=True)
            host = sharded.sweep_sharded(eng, dist, S_loc, 100.0, 0.05, 1.0, "put", "poly2", semantics)
            # every rank also sweeps the whole option alone (same counters: the columns are a permutation)
            S_all = eng.paths(model, M, N, dtype, E.RngSpec(seed=17))
            single = eng.lsm(S_all, 100.0, 0.05, 1.0, "put", "poly2", semantics, impl="resident", arrays=True)
            prices = [None] * world
            dist.all_gather_object(prices, (fused.price, fused.stderr, fused.betas.tobytes()))
            same = all(p == prices[0] for p in prices)  # bit-identical on every rank
            e1 = abs(fused.price - single.price) / single.price
            e2 = abs(fused.price - host.price) / host.price
            eb = float(np.nanmax(np.abs(fused.betas - single.betas) / (1e-300 + np.abs(single.betas)))) if dtype == "f64" else 0.0
            cnt_ok = bool((fused.ex_count == single.ex_count).all()) and bool((fused.n_itm == single.n_itm).all())
            bnd_ok = bool(np.array_equal(fused.boundary, single.boundary, equal_nan=True))
            good = same and e1 < tol and e2 < tol and (dtype == "f32" or (cnt_ok and bnd_ok and eb < 1e-7))
            ok &= good
            out[f"check_{dtype}_{semantics}"] = dict(price=fused.price, rel_vs_single=e1, rel_vs_nccl_loop=e2, beta_rel=eb,
                                                     identical_on_all_ranks=same, counts_equal=cnt_ok, boundary_equal=bnd_ok,
                                                     ok=good)
            del S_loc, S_all

    # ---- timing: M_total paths split over the ranks (fp32, reference semantics) ----
    M, N = args.paths, args.dates
    model = E.heston(100.0, 0.05, 1.0, **HP)
    off, m_loc = sharded.shard_pairs(M, rank, world)
    S_loc = eng.paths(model, m_loc, N, "f32", E.RngSpec(seed=5, pair_offset=off))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            r = fn()
        ev1.record()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    ms_fused, rf = timed(lambda: sharded.sweep_sharded_fused(eng, dist, S_loc, M, 100.0, 0.05, 1.0), args.reps)
    ms_host, rh = timed(lambda: sharded.sweep_sharded(eng, dist, S_loc, 100.0, 0.05, 1.0), 2)
    out["timing"] = dict(paths_total=M, paths_per_rank=m_loc, dates=N, fused_ms=ms_fused, nccl_loop_ms=ms_host,
                         fused_path_steps_per_s=M * N / ms_fused * 1e3, price_fused=rf.price, price_nccl_loop=rh.price)
    ok &= abs(rf.price - rh.price) / rh.price < 1e-5
    eng.comm_finalize()
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        out["ok"] = bool(okt.item())
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if okt.item() else 1)


if __name__ == "__main__":
    main()
