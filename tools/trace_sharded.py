"""Per-phase cycle trace of the path-sharded single-role sweep (one group shape of the bench launch: 37 CTAs x 27 k paths per
rank), to see what the in-kernel NVLink exchange adds to a date.  Launch with one process per GPU:

  OPTMC_RES_SPEC=0 OPTMC_RES_MAXCTAS=37 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29517 tools/trace_sharded.py

Columns (tools/trace_resident.py): 0 pass start, 1 pass done, 2 CTA totals in warp 0, 4 grid sum (local sum + exchange) complete,
5 solved."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa: E402,F401
from options_model_b200 import engine as E  # noqa: E402
from options_model_b200 import sharded  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = E.Engine(local)
sharded.init_peer_exchange(eng, dist)
M_loc, N = 1_000_000, 252
model = E.heston(100.0, 0.05, 1.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
S = eng.paths(model, M_loc, N, "f32", E.RngSpec(seed=1, pair_offset=rank * (M_loc // 2)))
for _ in range(3):
    sharded.sweep_sharded_fused(eng, dist, S, M_loc * world, 100.0, 0.05, 1.0)
dist.barrier()
out = os.path.join(ROOT, "gpurun_out", f"tr_sharded_rank{rank}.txt")
os.environ["OPTMC_TRACE"] = out
r = sharded.sweep_sharded_fused(eng, dist, S, M_loc * world, 100.0, 0.05, 1.0)
del os.environ["OPTMC_TRACE"]
dist.barrier()
rows = np.loadtxt(out, comments="#")
a = rows[rows[:, 0] == 0][:, 2:]
a = a[a[:, 5] > 0]
med = lambda x: f"median {np.median(x):7.0f} p90 {np.percentile(x, 90):7.0f}"  # noqa: E731
print(f"rank {rank}/{world} price {r.price:.6f}  per-date {med(a[1:, 0] - a[:-1, 0])} | pass {med(a[:, 1] - a[:, 0])} | reduce {med(a[:, 2] - a[:, 1])} | "
      f"grid sum + exchange {med(a[:, 4] - a[:, 2])} | solve {med(a[:, 5] - a[:, 4])} | spins {med(a[:, 7])}", flush=True)
eng.comm_finalize()
dist.barrier()
dist.destroy_process_group()
