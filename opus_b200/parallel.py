"""Host-side plumbing of the multi-GPU path: streams are independent, so they are sharded over ranks — one process per
GPU, no collective on the steady-state path.  The only collectives are the init-time broadcast of the workload / constant data and
the max-over-ranks of the timed region.  Backend agnostic (nccl on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def stream_range(rank, world, n_per_rank):
    """global stream ids owned by `rank` (contiguous blocks, so that a rank's host staging is one slab)"""
    if not (0 <= rank < world) or n_per_rank <= 0:
        raise ValueError('bad shard')
    return range(rank * n_per_rank, (rank + 1) * n_per_rank)


def fixture_rows(rank, world, n_per_rank, n_fixture):
    """which fixture stream each of the rank's streams replays: global id modulo the fixture size (ranks do not all start at row 0)"""
    ids = torch.arange(rank * n_per_rank, (rank + 1) * n_per_rank, dtype=torch.int64)
    return ids % n_fixture


def broadcast_fixture(packets, sizes, shape_pk, shape_sz, device, rank, world):
    """rank 0 owns the fixture (numpy arrays) and broadcasts it; other ranks pass None.  Returns (uint8 tensor, int32 tensor)."""
    if rank == 0:
        t_pk = torch.from_numpy(packets).to(device)
        t_sz = torch.from_numpy(sizes).to(device)
        assert tuple(t_pk.shape) == tuple(shape_pk) and tuple(t_sz.shape) == tuple(shape_sz)
    else:
        t_pk = torch.empty(shape_pk, dtype=torch.uint8, device=device)
        t_sz = torch.empty(shape_sz, dtype=torch.int32, device=device)
    if world > 1:
        dist.broadcast(t_pk, 0)
        dist.broadcast(t_sz, 0)
    return t_pk, t_sz


def max_over_ranks(ms, device, world):
    """the timed region of the job is the slowest rank's"""
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(n_per_rank, world, steps, ms_max):
    """whole-job frames/s: every rank decoded n_per_rank streams x steps frames within ms_max"""
    return n_per_rank * world * steps / (ms_max / 1000.0)
