"""CPU: host arithmetic of the path-sharded network fit (optmc_gnet_shard_plan, include/optmc.h) -- which of its own
shuffled rows a rank contributes to optimiser step b, and the step's global row count -- and the collective error
handling of sharded.gnet_sharded over gloo with a stand-in engine.  No device needed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    import options_model_b200 as om

    return om.load_library()


def test_shard_plan_tiles_every_ranks_rows_exactly_once(lib):
    from options_model_b200.engine import gnet_shard_plan

    rng = np.random.default_rng(5)
    cases = [([1000], 256), ([1000, 3, 0, 517], 256), ([0, 0, 7], 8192), ([5_000_000_000, 4_999_999_999], 131072 - 1)]
    cases += [(list(rng.integers(0, 50_000, size=w)), int(b)) for w in (2, 4, 8) for b in (256, 1000, 8192)]
    for n_rows, batch in cases:
        total = sum(n_rows)
        nb = -(-total // batch)
        steps = range(nb) if nb < 2000 else list(range(50)) + list(range(nb - 50, nb))
        pos = [None] * len(n_rows)
        for b in steps:
            g_sum = 0
            for r in range(len(n_rows)):
                lo, hi, g = gnet_shard_plan(lib, n_rows, batch, b, r)
                assert 0 <= lo <= hi <= n_rows[r]
                if pos[r] is not None and b > 0 and (nb < 2000 or b != nb - 50):
                    assert lo == pos[r]  # consecutive steps tile the rank's rows
                pos[r] = hi
                g_sum += hi - lo
            assert g == g_sum                      # the normaliser is the rows actually used
            assert abs(g - min(batch, total - b * batch)) <= len(n_rows)
        assert pos == list(n_rows)                 # the last step ends at the end of every rank's rows
        lo, hi, g = gnet_shard_plan(lib, n_rows, batch, nb + 3, 0)
        assert lo == hi == n_rows[0] and g == 0    # past the epoch: empty


def test_shard_plan_one_rank_equals_the_unsharded_batches(lib):
    from options_model_b200.engine import gnet_shard_plan

    n, batch = 100_003, 8192
    for b in range(-(-n // batch)):
        lo, hi, g = gnet_shard_plan(lib, [n], batch, b, 0)
        assert (lo, hi, g) == (b * batch, min((b + 1) * batch, n), min((b + 1) * batch, n) - b * batch)


def test_shard_plan_rejects_bad_arguments(lib):
    from options_model_b200.engine import gnet_shard_plan

    for args in (([10, -1], 256, 0, 0), ([10], 0, 0, 0), ([10], 256, -1, 0), ([10], 256, 0, 1), ([1] * 9, 256, 0, 0)):
        with pytest.raises(ValueError):
            gnet_shard_plan(lib, *args)


class _FakeGnetEngine:
    """Stand-in for Engine.lsm_gnet(M_total=...) / comm_export / comm_init: one rank can be made to fail."""

    def __init__(self, rank, fail_on):
        self.rank, self.fail_on, self.rewired = rank, fail_on, 0

    def comm_export(self):
        return bytes([self.rank]) * 64

    def comm_init(self, rank, world, handles):
        self.rewired += 1

    def lsm_gnet(self, S, K, r, T, option_type, semantics, arrays=False, M_total=None, **kw):
        if self.fail_on == self.rank:
            raise RuntimeError("exchange timed out (simulated)")
        n = S.shape[0]
        return dict(price=1.25, stderr=0.01, n_paths=M_total, ex_count=np.arange(n, dtype=np.int64) * (self.rank + 1),
                    boundary=np.array([np.nan] + [90.0 + self.rank] * (n - 1)))


def _worker(rank, world, port, fail_on, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import options_model_b200  # noqa: F401
    from options_model_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeGnetEngine(rank, fail_on)
    try:
        res = sharded.gnet_sharded(eng, dist, torch.zeros(5, 8), 16, 100.0, 0.05, 1.0, "put", "reference", arrays=True, epochs=1)
        out.put((rank, "ok", res["price"], res["ex_count"].tolist(), np.nan_to_num(res["boundary"], nan=-1.0).tolist(), eng.rewired))
    except RuntimeError as e:
        out.put((rank, "raised", str(e)[:60], None, None, eng.rewired))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_on", [None, 1])
def test_gnet_sharded_host_logic_two_ranks_gloo(fail_on):
    """sharded.gnet_sharded over gloo: per-rank exercise counts are summed and boundaries max-combined over the group; if
    ANY rank's call fails EVERY rank raises, after re-wiring the peer exchange (the running tags have diverged)."""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, fail_on, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if fail_on is None:
        for rank, status, price, exc, bnd, rewired in got:
            assert status == "ok" and price == 1.25 and rewired == 0
            assert exc == [0, 3, 6, 9, 12] and bnd == [-1.0, 91.0, 91.0, 91.0, 91.0]
    else:
        assert [g[1] for g in got] == ["raised", "raised"] and all(g[5] == 1 for g in got)
