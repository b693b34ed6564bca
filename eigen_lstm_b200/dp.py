"""Host-side helpers for data parallelism over sequences (one process per GPU).

The recurrence does not shard; the B streams do: rank r of G owns streams
[r*B, (r+1)*B) with their own text positions and carried state, the weights and Adagrad memory are
replicated, and the five gradient tensors are SUMMED over ranks (no 1/G: the reference sums over the
batch, R/lstm.cc:250-252) before the identical Adagrad update runs on every rank.  torch.distributed
is plumbing only (rendezvous, broadcasting the NCCL id, max-over-ranks timing); the gradient
allreduce itself is issued by liblstm_b200.so on its own communication stream.
"""
import numpy as np


def shard_streams(global_streams, rank, world):
    """Contiguous block of stream indices owned by `rank` (all ranks get the same count)."""
    if global_streams % world:
        raise ValueError("the global batch must divide evenly over the ranks")
    per = global_streams // world
    return np.arange(rank * per, (rank + 1) * per)


def stream_positions(global_streams, rank, world, S, chunk):
    """Start position of every local stream: stream g starts at S + g*chunk of the shared corpus."""
    return S + shard_streams(global_streams, rank, world).astype(np.int64) * int(chunk)


def broadcast_unique_id(dist, make_id, rank, src=0):
    """Create the 128-byte NCCL id on `src` and hand it to every rank through torch.distributed."""
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def max_over_ranks(dist, value, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
