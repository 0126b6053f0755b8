"""World-size-2 (and 3) `gloo` tests of the multi-GPU host logic on CPU: track sharding + the single
all-reduce of tally deltas reproduce the one-process sweep.  The per-rank compute is played by
the oracle here (no GPU in this container); on the GPU box the same plumbing drives the CUDA
kernel (bench.py, tests/test_gpu_parity.py::test_sharded_runs_add_up)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import importlib
    multi = importlib.import_module("simplemoc-kernel_b200.multi")
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R, F, G, N, p, seed = case
    o = Oracle()
    src, flux0, sig = o.fill(R, F, G, seed, 0.1)
    nt = (N + p - 1) // p
    tb, te = multi.shard_tracks(nt, rank, world)
    delta = np.zeros_like(flux0)                      # zero-initialised tally deltas
    _, chk = o.run(src, delta, sig, N, p, seed, tb, te, nthreads=1)
    t = torch.from_numpy(delta)
    multi.all_reduce_tallies(t)
    c = torch.tensor([chk % 2 ** 62, multi.shard_segments(nt, p, N, rank, world)], dtype=torch.int64)
    dist.all_reduce(c)
    if rank == 0:
        np.save(os.path.join(out_dir, "flux.npy"), flux0 + t.numpy())
        np.save(os.path.join(out_dir, "meta.npy"), c.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sweep_with_gloo_allreduce(oracle, tmp_path, world):
    case = (30, 5, 32, 10_037, 100, 77)
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    R, F, G, N, p, seed = case
    src, flux0, sig = oracle.fill(R, F, G, seed, 0.1)
    want = flux0.copy()
    _, chk = oracle.run(src, want, sig, N, p, seed, nthreads=1)
    got = np.load(tmp_path / "flux.npy")
    meta = np.load(tmp_path / "meta.npy")
    assert meta[1] == N                                    # shards cover every segment once
    a, b = got.astype(np.float64), want.astype(np.float64)
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-6


def test_shard_tracks_partition():
    import importlib
    multi = importlib.import_module("simplemoc-kernel_b200.multi")
    for nt in (0, 1, 7, 1000, 10 ** 8 + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [multi.shard_tracks(nt, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nt
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multi.shard_tracks(10, 2, 2)
