"""Path-sharded multi-GPU sweep: one process per GPU, each rank owns a contiguous block
of antithetic pairs (Philox counters stay global through ``pair_offset``), and the only data-path
collective is the per-date all-reduce of the Gram moment vector (8 doubles for poly2) plus the final
(sum, sum^2, n) reduction.  ``engine`` is an ``engine.Engine`` (CUDA); ``dist`` is ``torch.distributed``
with NCCL on GPUs.  The host logic is backend-agnostic, so the CPU test-suite drives it over gloo with a
stand-in engine.

Option-sharding (independent options per rank, no collective at all) needs no code here: each rank
simply prices its slice of the option list (bench.py --gpus N does exactly that).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np


def shard_pairs(M: int, rank: int, world: int, granule: int = 8):
    """Contiguous block of antithetic pairs for ``rank``: returns (pair_offset, local_M).  Both members of a
    pair live on the same rank (local columns j and j + local_M/2).  Blocks are multiples of ``granule`` pairs (the
    persistent sweep stages whole 16-byte units: 4 fp32 / 2 fp64 paths) except the last rank's, which takes the
    remainder -- so every rank but possibly the last is eligible for the fused sweep, and eligibility can be agreed on
    before anything is launched (``sweep_sharded_fused``)."""
    pairs = M // 2
    per = -(-pairs // world)                   # ceil
    per = -(-per // granule) * granule         # round up to the granule
    lo = min(rank * per, pairs)
    hi = min(lo + per, pairs)
    return lo, 2 * (hi - lo)


@dataclass
class ShardedResult:
    price: float
    stderr: float
    n_paths: int
    n_collectives: int


def sweep_sharded(engine, dist, S_local, K, r, T, option_type="put", basis="poly2", semantics="reference",
                  group=None, torch_mod=None) -> ShardedResult:
    """Run the LSM sweep over this rank's slab ``S_local[(N+1), M_local]``; every rank returns the global price.

    Per date: local Gram moments (device) -> all_reduce(SUM) -> every rank solves the same p x p system
    redundantly (bit-identical beta because the reduced vector is identical) -> local update.
    """
    torch = torch_mod
    if torch is None:
        import torch  # type: ignore
    N = S_local.shape[0] - 1
    q = engine.gram_len(basis)
    gram = torch.zeros(q, dtype=torch.float64, device=S_local.device)
    sums = torch.zeros(3, dtype=torch.float64, device=S_local.device)
    engine.lsm_begin(S_local, K, r, T, option_type, basis, semantics)
    ncoll = 0
    for t in range(N - 1, 0, -1):
        engine.lsm_gram_date(t, gram)
        dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=group)
        ncoll += 1
        engine.lsm_update_date(t, gram)
    engine.lsm_finish(sums)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    ncoll += 1
    s, ss, n = (float(x) for x in sums.cpu().tolist())
    mean = s / n
    var = max((ss - n * mean * mean) / (n - 1.0), 0.0) if n > 1 else 0.0
    dt = T / N
    scale = 1.0 if semantics in ("reference", 3) else math.exp(-r * dt)
    return ShardedResult(mean * scale, math.sqrt(var / n) * scale, int(round(n)), ncoll)


def init_peer_exchange(engine, dist, group=None) -> None:
    """Wire the in-kernel exchange of ``sweep_sharded_fused``: every rank exports the CUDA-IPC handle of its
    exchange slots, the handles are all-gathered over the process group (host plumbing), every rank maps its peers.
    Raises on EVERY rank if any rank failed to map a peer (so no rank is left waiting for the others)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    err = None
    try:
        mine = engine.comm_export()
    except Exception as e:  # noqa: BLE001 -- reported collectively below
        mine, err = b"", e
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    if err is None and all(len(h) == len(mine) for h in handles):
        try:
            engine.comm_init(rank, world, handles)
        except Exception as e:  # noqa: BLE001
            err = e
    oks = [None] * world
    dist.all_gather_object(oks, err is None, group=group)  # also the barrier: nobody launches before all mapped
    if not all(oks):
        raise RuntimeError(f"peer exchange unavailable (ranks ok: {oks}): {err}")


def sweep_sharded_fused(engine, dist, S_local, M_total, K, r, T, option_type="put", basis="poly2",
                        semantics="reference", group=None, arrays=False):
    """The same sweep as ``sweep_sharded`` in ONE kernel launch per rank: the per-date totals travel between the
    GPUs inside the persistent kernel (peer stores over NVLink), so the 251 host-launched collectives of the
    NCCL variant disappear from the data path.  Requires ``init_peer_exchange`` once per process group.
    Every rank returns the global price; with ``arrays`` the per-rank exercise counts / boundaries are combined
    over the group (that combine is host plumbing after the sweep, not part of it)."""
    # Every rank must launch or none (a rank that fails validation would leave its peers spinning on the exchange, and
    # the running exchange counter would diverge): agree on the outcome of the call collectively, and re-wire the
    # exchange (fresh slots, counter reset on every rank) after any failure before raising on every rank alike.
    err = None
    res = None
    if S_local.shape[1] == 0 or (S_local.shape[1] * S_local.element_size()) % 16 != 0:
        err = ValueError(f"rank block of {S_local.shape[1]} paths is not a multiple of 16 bytes (use shard_pairs)")
    world = dist.get_world_size(group)
    if world > 1:
        pre = [None] * world
        dist.all_gather_object(pre, err is None, group=group)
        if not all(pre):
            raise RuntimeError(f"fused sharded sweep not launched: ineligible block on ranks {[i for i, ok in enumerate(pre) if not ok]}: {err}")
    elif err is not None:
        raise err
    try:
        res = engine.lsm_sharded(S_local, M_total, K, r, T, option_type, basis, semantics, arrays=arrays)
    except Exception as e:  # noqa: BLE001 -- reported collectively below
        err = e
    if world > 1:
        post = [None] * world
        dist.all_gather_object(post, None if err is None else str(err), group=group)
        if any(p is not None for p in post):
            init_peer_exchange(engine, dist, group)
            raise RuntimeError(f"fused sharded sweep failed on ranks {[i for i, p in enumerate(post) if p is not None]}: "
                               f"{[p for p in post if p is not None][0]}")
    elif err is not None:
        raise err
    if arrays and dist.get_world_size(group) > 1:
        import torch

        dev = S_local.device
        exc = torch.as_tensor(res.ex_count, device=dev)
        dist.all_reduce(exc, op=dist.ReduceOp.SUM, group=group)
        res.ex_count = exc.cpu().numpy()
        put = option_type == "put"
        bnd = torch.as_tensor(np.nan_to_num(res.boundary, nan=-np.inf if put else np.inf), device=dev)
        dist.all_reduce(bnd, op=dist.ReduceOp.MAX if put else dist.ReduceOp.MIN, group=group)
        b = bnd.cpu().numpy()
        b[~np.isfinite(b)] = np.nan
        res.boundary = b
    return res


def gnet_sharded(engine, dist, S_local, M_total, K, r, T, option_type="put", semantics="reference", group=None, arrays=False,
                 **kw):
    """The reference's global network regression (om3:482-651, ``SingleLSMNet``) with the PATHS sharded over the ranks:
    every rank collects the rows of its own paths, the per-step gradient vectors (34 177 fp32 + the batch loss) travel
    between the GPUs through peer-mapped memory inside the reduce / optimiser kernels (optmc_lsm_gnet_sharded; no
    host-launched collective on the data path), every rank ends with the same weights, loss and price.  Requires
    ``init_peer_exchange`` once per process group.  Same collective error handling as ``sweep_sharded_fused``: all ranks
    raise or none does, and the exchange is re-wired after a failure."""
    world = dist.get_world_size(group)
    err = None
    res = None
    try:
        res = engine.lsm_gnet(S_local, K, r, T, option_type, semantics, arrays=arrays, M_total=int(M_total), **kw)
    except Exception as e:  # noqa: BLE001 -- reported collectively below
        err = e
    if world > 1:
        post = [None] * world
        dist.all_gather_object(post, None if err is None else str(err), group=group)
        if any(p is not None for p in post):
            init_peer_exchange(engine, dist, group)
            raise RuntimeError(f"sharded network LSM failed on ranks {[i for i, p in enumerate(post) if p is not None]}: "
                               f"{[p for p in post if p is not None][0]}")
    elif err is not None:
        raise err
    if arrays and world > 1:
        import torch

        dev = S_local.device
        exc = torch.as_tensor(res["ex_count"], device=dev)
        dist.all_reduce(exc, op=dist.ReduceOp.SUM, group=group)
        res["ex_count"] = exc.cpu().numpy()
        put = option_type == "put"
        bnd = torch.as_tensor(np.nan_to_num(res["boundary"], nan=-np.inf if put else np.inf), device=dev)
        dist.all_reduce(bnd, op=dist.ReduceOp.MAX if put else dist.ReduceOp.MIN, group=group)
        b = bnd.cpu().numpy()
        b[~np.isfinite(b)] = np.nan
        res["boundary"] = b
    return res
