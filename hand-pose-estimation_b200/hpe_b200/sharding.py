"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed): how the path shards.

  * independent poses / frames / sequences: contiguous blocks per rank, no data-path collective
    (`shard_range`);
  * one large swarm: particles block-partitioned by global index; once per generation every rank
    needs all (pcost, pbest) rows, so the owned rows are all-gathered (`allgather_rows`), NCCL over
    NVLink on the GPU box, gloo in the CPU tests;
  * timing: device time, maximum over ranks (`max_over_ranks`).

Works on CPU tensors (gloo) and CUDA tensors (nccl) alike; nothing here computes the hot path.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous block [lo, hi) of n items owned by `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allgather_rows(table, n_rows):
    """table: [n_rows, row_len] tensor (any device) whose rows shard_range(n_rows, rank, world) are
    valid on this rank; on return every row is valid on every rank.  In place."""
    w = world()
    if w == 1:
        return table
    r = rank()
    lo, hi = shard_range(n_rows, r, w)
    if n_rows % w == 0:
        dist.all_gather_into_tensor(table.view(-1), table[lo:hi].reshape(-1).clone())
        return table
    # ragged shards: pad to the largest
    per = -(-n_rows // w)
    send = torch.zeros((per, table.shape[1]), dtype=table.dtype, device=table.device)
    send[: hi - lo] = table[lo:hi]
    recv = [torch.empty_like(send) for _ in range(w)]
    dist.all_gather(recv, send)
    for k in range(w):
        klo, khi = shard_range(n_rows, k, w)
        table[klo:khi] = recv[k][: khi - klo]
    return table


def max_over_ranks(x, device="cpu"):
    if world() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device="cpu"):
    if world() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def tensor_from_device_pointer(ptr, n, dtype="<f8"):
    """torch CUDA tensor aliasing n elements at raw device pointer `ptr` (no ownership): how the
    swarm table owned by libhpe_b200 is handed to NCCL"""
    class _Arr:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": dtype, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Arr(), device="cuda")


def attach_swarm_peers(swarm):
    """exchange the CUDA IPC handles of the ranks' tables (host side, once) and map the peers"""
    w = world()
    if w == 1:
        return
    handles = [None] * w
    dist.all_gather_object(handles, swarm.ipc_handle())
    swarm.attach_peers(handles, rank())
    dist.barrier()


def run_sharded_swarm_p2p(swarm, stream=None):
    """drive one shard whose peers are attached: the rows travel by direct NVLink stores issued from the
    step kernels and the ranks meet on device-side flags; the host only enqueues launches"""
    swarm.init(stream)
    swarm.reduce(stream)
    for _ in range(swarm.maxiter - 1):
        swarm.step(stream)
        swarm.reduce(stream)
    return swarm.best()


def run_sharded_swarm(swarm, n_rows, stream=None):
    """drive one hpe_b200.Swarm shard to completion, exchanging the table with the other ranks"""
    views = {}

    def table():
        # the rows are double buffered inside the library: the current table alternates every generation
        ptr = swarm.table_ptr()
        if ptr not in views:
            views[ptr] = tensor_from_device_pointer(ptr, n_rows * swarm.ROW).view(n_rows, swarm.ROW)
        return views[ptr]

    swarm.init(stream)
    allgather_rows(table(), n_rows)
    swarm.reduce(stream)
    for _ in range(swarm.maxiter - 1):
        swarm.step(stream)
        allgather_rows(table(), n_rows)
        swarm.reduce(stream)
    return swarm.best()
