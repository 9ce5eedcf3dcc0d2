"""Host-side multi-rank logic on CPU (gloo, world_size 2): the shard map is a partition aligned to the reduction
tiles, every rank derives the same map, and the reference arm of bench.py runs on rank 0 only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from neural_network_compression_b200 import _native as N
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for n in (1 << 20, 10 ** 6 + 7, 235200, 123457):
    b, e = N.shard_range(n, rank, world)
    t = torch.tensor([b, e], dtype=torch.int64)
    got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, t)
    # a partition of [0, n) in rank order, identical to what every other rank computes locally
    ok &= int(got[0][0]) == 0 and int(got[-1][1]) == n
    for r in range(world):
        ok &= (int(got[r][0]), int(got[r][1])) == N.shard_range(n, r, world)
        if r: ok &= int(got[r][0]) == int(got[r - 1][1])
        ok &= int(got[r][0]) %% 8 == 0   # tile offsets of NumPy's pairwise tree are multiples of 8
    tot = torch.tensor([e - b]); dist.all_reduce(tot); ok &= int(tot) == n
print("RANK", rank, "OK" if ok else "BAD", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def _torchrun(args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531"] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_shard_map_is_a_partition_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER % ROOT)
    r = _torchrun([str(w)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2


def test_reference_arm_runs_on_rank0_only():
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--cpu-sample", str(1 << 16)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["value"] > 0


INIT_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from neural_network_compression_b200 import _native as N
from neural_network_compression_b200.common import utility as U
from oracle import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ops = {"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}
def allreduce(a, op):
    t = torch.from_numpy(np.ascontiguousarray(a).copy()); dist.all_reduce(t, op=ops[op]); return t.numpy()
ok = True
for n, seed in ((300000, 5), (100003, 6)):
    w = (np.random.RandomState(seed).randn(n) * 0.05).astype(np.float32)
    w[np.abs(w) < 0.05] = 0                                   # a pruned layer: density init looks at the non-zeros only
    b, e = N.shard_range(n, rank, world)
    mine = w[b:e]
    nz = mine[mine != 0]
    # the exchange logic of utility.get_weight_distribution on shards, with NumPy standing in for the two device sweeps
    def local_minmax():
        return (nz.min(), nz.max(), nz.size) if nz.size else (np.float32(0), np.float32(0), 0)
    def local_hist(edges):
        return np.array([np.count_nonzero((nz >= edges[i]) & (nz < edges[i + 1])) for i in range(31)], dtype=np.int64)
    xnew, cdf = U.sharded_weight_distribution(local_minmax, local_hist, allreduce)
    ref = O.get_weight_distribution(w[w != 0])              # the single-rank restatement (pinned to the reference's goldens)
    ok &= xnew.tobytes() == ref[0].tobytes() and cdf.tobytes() == ref[1].tobytes()
    for bits in (2, 8):                                       # density init from the sharded CDF == from the whole tensor
        ok &= U._init_density(bits, (xnew, cdf)).tobytes() == O.init_centroids(w, bits, "density", ref).tobytes()
    # forgy: same global draws on every rank, owners supply the values
    np.random.seed(3)
    idx = np.random.randint(0, n, size=32)
    space = U.sharded_forgy_init(idx, b, e, lambda local_idx: mine[local_idx], allreduce)
    ok &= space.tobytes() == w[idx].tobytes()
print("RANK", rank, "OK" if ok else "BAD", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_density_and_forgy_init_on_shards_gloo(tmp_path):
    """density init from an all-reduce of per-rank 31-bin histograms (+ min / max), forgy through an owner
    gather: the exchange logic of utility.py under gloo on two CPU ranks equals the single-rank result bit for bit."""
    w = tmp_path / "init_worker.py"
    w.write_text(INIT_WORKER % ROOT)
    r = _torchrun([str(w)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2


def test_shard_range_small_tensors():
    from neural_network_compression_b200 import _native as N

    assert N.shard_range(100, 0, 4) == (0, 100) and N.shard_range(100, 3, 4) == (100, 100)
    b0, e0 = N.shard_range(1 << 30, 0, 8)
    assert (b0, e0) == (0, 1 << 27)
    with pytest.raises(N.NncError):
        N.shard_range(0, 0, 1)
