"""Concurrent host->device copy bandwidth of all ranks of one box (torchrun): the e2e leg of bench.py uploads 13.6 MB of
poses per step and rank from pinned host memory; this shows what the host side of the box sustains when N ranks do that
at once, with ordinary pinned memory and with write-combined pinned memory."""
import ctypes, glob, os, sys, time
import torch, torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 13631488
d = torch.empty(n, dtype=torch.uint8, device="cuda")
libs = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(libs[0]) if libs else None


def measure(src_ptr, label):
    st = torch.cuda.current_stream().cuda_stream
    def copy():
        rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(src_ptr), ctypes.c_size_t(n), 1, ctypes.c_void_p(st))
    for _ in range(5): copy()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 60
    e0.record()
    for _ in range(reps): copy()
    e1.record(); torch.cuda.synchronize()
    gbs = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    tot = t.clone()
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world={world} {label:28s} per rank (min) {t.item():6.1f} GB/s   all ranks {tot.item():7.1f} GB/s")


h = torch.empty(n, dtype=torch.uint8).pin_memory()
measure(h.data_ptr(), "pinned (torch pin_memory)")
if rt is not None:
    p = ctypes.c_void_p()
    if rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), 4) == 0:       # cudaHostAllocWriteCombined
        ctypes.memset(p, 1, n)
        measure(p.value, "pinned, write-combined")
if world > 1:
    dist.destroy_process_group()
