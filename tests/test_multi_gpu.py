"""Two real GPUs: the path-sharded sweep with the in-kernel NVLink exchange (optmc_comm_*, optmc_lsm_poly_sharded)
against the single-GPU sweep and the host-loop NCCL variant.  Skipped on a one-GPU box; the host-side logic of the
sharding is covered on CPU by tests/test_sharded_gloo.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_sharded_sweep_two_gpus():
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_fused_check.py"), "--paths", "400000",
           "--dates", "60", "--reps", "3", "--check-paths", "100000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["world"] == 2
    for k, v in out.items():
        if k.startswith("check_"):
            assert v["ok"] and v["identical_on_all_ranks"], (k, v)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_bench_path_sharded_batch_two_gpus():
    """bench.py's multi-GPU mode (every option's paths split over the ranks, totals exchanged inside the grouped sweep
    kernel): the built-in parity check against the single-GPU pricing must report bit-identical prices."""
    port = 29900 + os.getpid() % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3",
           "--paths", "200000", "--dates", "40", "--no-cpu", "--no-other-configs"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    out = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert out["n_gpus"] == 2 and out["config"]["shard_mode"] == "paths"
    assert out["path_sharded_parity"]["rel_vs_single"] == 0.0 and out["path_sharded_parity"]["identical_on_all_ranks"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_gnet_sharded_two_gpus():
    """The global network regression with the paths split over two GPUs (gradients exchanged through peer memory inside
    the reduce / optimiser kernels): identical weights / loss / price on both ranks, statistically equal to the
    single-GPU fit on the same paths."""
    port = 29700 + os.getpid() % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "gnet_sharded_check.py"), "--paths", "100000", "--dates", "40",
           "--epochs", "5"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    out = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["world"] == 2
    for k, v in out.items():
        if k.startswith("check_"):
            assert v["ok"] and v["identical_on_all_ranks"], (k, v)
