"""Diagnostic: time of the collectives the partition uses, alone, eager and inside a CUDA graph.
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/probe/nccl_allreduce.py"""
import os

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda")


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000


for mb in (0.016, 2.0, 9.0, 18.1):
    n = int(mb * 1e6 / 4)
    x = torch.ones(n, device=dev)
    eager = timed(lambda: dist.all_reduce(x))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dist.all_reduce(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(10):
                dist.all_reduce(x)
        graph = timed(g.replay, 10) / 10
    if rank == 0:
        print("all_reduce %6.3f MB x %d ranks: eager %7.1f us, in a graph %7.1f us  (%.0f GB/s algorithmic in the graph)  env %s" % (
            mb, world, eager, graph, mb * 1e6 / (graph * 1e-6) / 1e9, {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}))
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
