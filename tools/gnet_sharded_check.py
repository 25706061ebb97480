"""Path-sharded global network LSM (optmc_lsm_gnet_sharded: SingleLSMNet(7,128,3) on all dates' rows, the paths split
over the ranks, gradients exchanged through peer-mapped memory inside the reduce / optimiser kernels), checked against
the single-GPU fit and timed.  Launch with one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      tools/gnet_sharded_check.py [--paths 100000] [--dates 50] [--epochs 6] [--time-paths 100000]

Checks: price / stderr / best loss / final weights bit-identical on every rank; the row count equals the single-GPU
fit's; loss within 1 % and price within 4 % of the single-GPU fit on the same paths (the estimator's own seed-to-seed spread is +-2.5 %: DESIGN 4; the ranks' mini-batches are
composed differently, so the fits agree statistically, not bit for bit).  Exit code 0 = all passed; rank 0 prints one
JSON line.
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


def run(eng, world, rank, M, N, model, epochs, batch, variant, seed=11):
    off, m_loc = sharded.shard_pairs(M, rank, world)
    S_loc = eng.paths(model, m_loc, N, "f32", E.RngSpec(seed=seed, pair_offset=off))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sharded.gnet_sharded(eng, dist, S_loc, M, 100.0, 0.05, 1.0, "put", "reference", arrays=True, variant=variant,
                               epochs=epochs, batch=batch, seed=5, return_params=True, stop_patience=0)
    dt = time.perf_counter() - t0
    return res, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=int, default=100_000)
    ap.add_argument("--dates", type=int, default=50)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--time-paths", type=int, default=0, help="per-rank paths of the weak-scaling timing (0 = skip)")
    ap.add_argument("--time-dates", type=int, default=50)
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
    M, N = args.paths, args.dates
    model = E.gbm(100.0, 0.05, 1.0, 0.2)
    for name, mdl in (("gbm", model), ("heston", E.heston(100.0, 0.05, 1.0, **HP))):
        res, _ = run(eng, world, rank, M, N, mdl, args.epochs, args.batch, "gpu")
        # the whole option on every rank alone (same Philox counters: the local columns are a subset of the global ones)
        S_all = eng.paths(mdl, M, N, "f32", E.RngSpec(seed=11))
        single = eng.lsm_gnet(S_all, 100.0, 0.05, 1.0, "put", "reference", variant="gpu", epochs=args.epochs, batch=args.batch,
                              seed=5, stop_patience=0)
        del S_all
        sig = (res["price"], res["stderr"], res["best_loss"], res["epochs_run"], res["params"].tobytes())
        sigs = [None] * world
        dist.all_gather_object(sigs, sig)
        same = all(s == sigs[0] for s in sigs)
        e_price = abs(res["price"] - single["price"]) / single["price"]
        e_loss = abs(res["best_loss"] - single["best_loss"]) / single["best_loss"]
        exc_ok = int(res["ex_count"].sum()) > 0 and abs(int(res["ex_count"].sum()) - int(single["ex_count"].sum())) < 0.05 * M
        good = same and res["n_rows"] == single["n_rows"] and res["n_paths"] == M and e_price < 0.04 and e_loss < 0.01 and exc_ok
        ok &= good
        out[f"check_{name}"] = dict(price=res["price"], single_price=single["price"], rel_price=e_price, loss=res["best_loss"],
                                    single_loss=single["best_loss"], rel_loss=e_loss, n_rows=res["n_rows"],
                                    identical_on_all_ranks=same, exercised=int(res["ex_count"].sum()),
                                    single_exercised=int(single["ex_count"].sum()), ok=good)
    if args.time_paths:
        # weak scaling: time_paths per rank, fixed epochs; per-epoch time from two runs (1 and 3 epochs)
        Mt = args.time_paths * world
        mdl = E.gbm(100.0, 0.05, 1.0, 0.2)
        ts = {}
        for ep in (1, 3):
            run(eng, world, rank, Mt, args.time_dates, mdl, ep, args.batch * world, "gpu")  # warm
            dist.barrier(); torch.cuda.synchronize()
            res, dt = run(eng, world, rank, Mt, args.time_dates, mdl, ep, args.batch * world, "gpu")
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts[ep] = float(t.item())
        steps = -(-res["n_rows"] // (args.batch * world))
        out["timing"] = dict(paths_total=Mt, dates=args.time_dates, rows=res["n_rows"], global_batch=args.batch * world,
                             steps_per_epoch=steps, epoch_ms=(ts[3] - ts[1]) / 2 * 1e3,
                             us_per_step=(ts[3] - ts[1]) / 2 / steps * 1e6, loss=res["best_loss"], price=res["price"])
    out["ok"] = bool(ok)
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok))
    if rank == 0:
        print(json.dumps(out), flush=True)
    eng.comm_finalize()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(oks) else 1)


if __name__ == "__main__":
    main()
