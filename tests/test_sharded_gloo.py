"""CPU: the path-sharded multi-GPU host logic (options-model_b200/sharded.py) over torch.distributed/gloo with
world_size 2.  The CUDA engine is replaced by a stand-in with the same per-date interface (lsm_begin /
lsm_gram_date / lsm_update_date / lsm_finish) built from the numpy oracle's pieces, so this checks the pair
partition, the collective sequence (one Gram all-reduce per exercise date + one final) and that every rank ends
with the single-process price.  The real engine runs the same host code on GPUs (tests/test_gpu_sweep.py, tests/test_multi_gpu.py).
"""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


class StandInEngine:
    """Per-date building blocks of include/optmc.h (optmc_lsm_begin/gram_date/update_date/finish) on the CPU,
    restated with oracle/lsm_oracle.py.  TEST INFRASTRUCTURE ONLY."""

    def __init__(self):
        from oracle import lsm_oracle as orc

        self.orc = orc

    def gram_len(self, basis="poly2"):
        return {"poly2": 8, "poly3": 11}[basis]

    def lsm_begin(self, S, K, r, T, option_type="put", basis="poly2", semantics="reference", M=None):
        self.S = S.numpy()
        self.K, self.r, self.T, self.ot, self.basis, self.sem = K, r, T, option_type, basis, semantics
        self.N = self.S.shape[0] - 1
        self.disc = np.exp(-r * T / self.N)
        self.cf = self.orc.payoff(self.S[-1], K, option_type).astype(np.float64)
        self.exercised = np.zeros(self.S.shape[1], dtype=bool)
        self.deg = 2 if basis == "poly2" else 3

    def _itm(self, t):
        itm = self.orc.payoff(self.S[t], self.K, self.ot) > 0
        if self.sem == "reference":
            itm &= ~self.exercised
        return itm

    def lsm_gram_date(self, t, gram):
        self.cf *= self.disc  # om3:620: all paths, before the mask
        itm = self._itm(t)
        x = self.S[t, itm] / self.K
        y = self.cf[itm]
        m = [np.sum(x**k) for k in range(2 * self.deg + 1)] + [np.sum(x**k * y) for k in range(self.deg + 1)]
        gram.copy_(torch.tensor(m, dtype=torch.float64))

    def lsm_update_date(self, t, gram):
        m = gram.numpy()
        p = self.deg + 1
        if m[0] < p:
            return
        G = np.array([[m[i + j] for j in range(p)] for i in range(p)])
        g = m[2 * self.deg + 1:]
        beta = self.orc.cholesky_solve_guarded(G, g)
        if beta is None:
            return
        itm = self._itm(t)
        x = self.S[t, itm] / self.K
        cont = sum(beta[i] * x**i for i in range(p))
        pay = self.orc.payoff(self.S[t, itm], self.K, self.ot)
        idx = np.where(itm)[0][pay > cont]
        self.cf[idx] = pay[pay > cont]
        self.exercised[idx] = True

    def lsm_finish(self, sums):
        sums.copy_(torch.tensor([self.cf.sum(), (self.cf**2).sum(), float(self.cf.size)], dtype=torch.float64))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, semantics, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import options_model_b200  # noqa: F401
    from options_model_b200 import sharded
    from oracle import lsm_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M, N = 6000, 12
    rng = np.random.default_rng(17)
    Z1, Z2 = orc.draw_heston_normals(rng, N, M)
    S = orc.heston_paths_antithetic(100.0, 0.05, 1.0, HP["v0"], HP["kappa"], HP["theta"], HP["xi"], HP["rho"], M, N, Z1, Z2)
    lo, m_local = sharded.shard_pairs(M, rank, world)
    h = M // 2
    cols = np.concatenate([np.arange(lo, lo + m_local // 2), h + np.arange(lo, lo + m_local // 2)])
    S_local = torch.from_numpy(np.ascontiguousarray(S[:, cols]))
    res = sharded.sweep_sharded(StandInEngine(), dist, S_local, 100.0, 0.05, 1.0, "put", "poly2", semantics,
                                torch_mod=torch)
    ref = orc.lsm_sweep(S, 100.0, 0.05, 1.0, "put", semantics=semantics)
    out.put((rank, res.price, res.stderr, res.n_paths, res.n_collectives, ref.price, ref.stderr))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("semantics", ["reference", "textbook"])
def test_path_sharded_sweep_two_ranks_gloo(semantics):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, semantics, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, price, se, n, ncoll, ref_price, ref_se in got:
        assert n == 6000
        assert ncoll == 12  # N - 1 Gram all-reduces + the final (sum, sum^2, n)
        assert price == pytest.approx(ref_price, rel=1e-12)
        assert se == pytest.approx(ref_se, rel=1e-9)


def test_shard_pairs_partition():
    sys.path.insert(0, ROOT)
    import options_model_b200  # noqa: F401
    from options_model_b200 import sharded

    for M, world in ((1_000_000, 8), (1002, 4), (6, 4), (2, 2)):
        seen = []
        for r in range(world):
            lo, m = sharded.shard_pairs(M, r, world)
            assert m % 2 == 0
            seen += list(range(lo, lo + m // 2))
        assert seen == list(range(M // 2))  # contiguous, disjoint, complete


class _FakeCommEngine:
    """Stand-in for the optmc_comm_export / optmc_comm_init pair: records what the host plumbing hands over."""

    def __init__(self, rank, fail_on=None):
        self.rank, self.fail_on, self.got = rank, fail_on, None

    def comm_export(self):
        return bytes([self.rank]) * 64

    def comm_init(self, rank, world, handles):
        if self.fail_on == rank:
            raise RuntimeError("cudaIpcOpenMemHandle failed (simulated)")
        self.got = (rank, world, list(handles))


def _peer_worker(rank, world, port, fail_on, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import options_model_b200  # noqa: F401
    from options_model_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeCommEngine(rank, fail_on)
    try:
        sharded.init_peer_exchange(eng, dist)
        out.put((rank, "ok", eng.got))
    except RuntimeError as e:
        out.put((rank, "raised", str(e)[:40]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_on", [None, 1])
def test_peer_exchange_wiring_two_ranks_gloo(fail_on):
    """Host plumbing of the in-kernel NVLink exchange (sharded.init_peer_exchange): every rank receives every rank's
    64-byte handle in rank order; if ANY rank fails to map a peer, EVERY rank raises (nobody is left spinning on a
    peer that never launches)."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, fail_on, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if fail_on is None:
        for rank, status, info in got:
            assert status == "ok" and info[0] == rank and info[1] == world
            assert info[2] == [bytes([0]) * 64, bytes([1]) * 64]
    else:
        assert [g[1] for g in got] == ["raised", "raised"]


def test_shard_pairs_blocks_are_contiguous_and_aligned():
    """Per-rank blocks tile the pairs exactly, in rank order; every block but the last non-empty one is a multiple of
    8 pairs (16-byte units of the persistent sweep)."""
    from options_model_b200 import sharded

    for M in (2, 100, 4096, 100_000, 1_000_002, 8_000_000):
        for world in (1, 2, 3, 4, 8):
            nxt, sizes = 0, []
            for r in range(world):
                off, m = sharded.shard_pairs(M, r, world)
                assert off == nxt and m % 2 == 0 and m >= 0
                nxt += m // 2
                sizes.append(m // 2)
            assert nxt == M // 2
            nonempty = [s for s in sizes if s]
            assert all(s % 8 == 0 for s in nonempty[:-1])
