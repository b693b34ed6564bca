"""Data parallelism on real GPUs (needs >= 2): world contexts with B streams each, gradients summed by
the library's NCCL allreduce, must equal ONE context holding all world*B streams."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dtype, no_graph, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if no_graph:
        os.environ["LSTM_NO_GRAPH"] = "1"      # plain stream launches instead of the segmented graph replay (read at first use)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import eigen_lstm_b200 as el
    from eigen_lstm_b200 import dp
    from tests.conftest import GOLDEN
    text = open(os.path.join(GOLDEN, "enwik6_head.bin"), "rb").read()
    M, N, S, B = 256, 64, 6, 4
    chunk = 700
    g = el.LSTM(M, N, S, B, device=rank, dtype=dtype)
    g.dp_init(rank, world, dp.broadcast_unique_id(dist, el.dp_unique_id, rank))
    g.init_params(7, 0.05, 1.0)
    g.load_text(text); g.set_positions(dp.stream_positions(B * world, rank, world, S, chunk))
    losses = g.train_text(5, stride=S - 1, lr=0.01)   # plain, capture + replay, 3 more segmented-graph replays
    flat = np.concatenate([p.ravel(order="F") for p in g.params()])
    if rank == 0:
        one = el.LSTM(M, N, S, B * world, device=0, dtype=dtype)
        one.init_params(7, 0.05, 1.0)
        one.load_text(text); one.set_positions(dp.stream_positions(B * world, 0, 1, S, chunk))
        one.train_text(5, stride=S - 1, lr=0.01)
        ref = np.concatenate([p.ravel(order="F") for p in one.params()])
        out["err"] = float(np.max(np.abs(flat - ref)) / np.max(np.abs(ref)))
    t = torch.from_numpy(flat).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["replicas_identical"] = bool(torch.equal(lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("no_graph", [0, 1])
@pytest.mark.parametrize("dtype", [0, 1])
def test_two_gpus_equal_one_gpu_with_the_concatenated_batch(dtype, no_graph):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), dtype, no_graph, out), nprocs=2, join=True)
    assert out["replicas_identical"]
    assert out["err"] < (2e-4 if dtype == 0 else 2e-2), out["err"]
