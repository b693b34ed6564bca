"""Data-parallel host logic on CPU: world_size 2 over gloo.  The CPU oracle stands in for the GPU
contexts, so what is checked is the SHARDING CONTRACT the CUDA path relies on:
summing per-rank gradients of stream shards == the gradient of one context holding every stream,
the NCCL id reaches every rank, and timing is the max over ranks."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import GOLDEN


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from eigen_lstm_b200 import dp
    from oracle import oracle as orc
    text = open(os.path.join(GOLDEN, "enwik6_head.bin"), "rb").read()
    M, N, S, Bg = 256, 12, 6, 6
    params = orc.init_params(M, N, seed=3, sd=0.05, forget_bias=1.0)
    chunk = 500
    ident = dp.broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
    assert ident == bytes(range(128))
    pos = dp.stream_positions(Bg, rank, world, S, chunk)
    o = orc.Oracle(M, N, S, len(pos), "f32"); o.set_params(params); o.set_positions(pos)
    o.advance(text, S - 1)
    loss = o.forward(); o.backward()
    flat = np.concatenate([g.ravel(order="F") for g in o.grads()])
    t = torch.from_numpy(flat.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)               # what ncclAllReduce(sum) does in the library
    lt = torch.tensor([loss * len(pos)], dtype=torch.float64)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    slow = dp.max_over_ranks(dist, 1.0 + rank)
    if rank == 0:
        full = orc.Oracle(M, N, S, Bg, "f32"); full.set_params(params)
        full.set_positions(dp.stream_positions(Bg, 0, 1, S, chunk)); full.advance(text, S - 1)
        lf = full.forward(); full.backward()
        ref = np.concatenate([g.ravel(order="F") for g in full.grads()])
        out["grad_err"] = float(np.max(np.abs(t.numpy() - ref)) / np.max(np.abs(ref)))
        out["loss_err"] = abs(lt.item() / Bg - lf) / lf
        out["slow"] = slow
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_the_full_batch():
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out["grad_err"] < 1e-5, out["grad_err"]
    assert out["loss_err"] < 1e-6
    assert out["slow"] == 2.0


def test_shard_streams_partition():
    from eigen_lstm_b200 import dp
    got = np.concatenate([dp.shard_streams(16, r, 4) for r in range(4)])
    assert np.array_equal(got, np.arange(16))
    import pytest
    with pytest.raises(ValueError):
        dp.shard_streams(10, 0, 4)
