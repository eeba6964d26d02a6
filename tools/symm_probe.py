"""Where the multi-GPU optimiser step spends its time (development aid; run under torchrun): cross-rank barrier, the fused
all-reduce + RMSprop + broadcast kernel, both together, and NCCL all-reduce / reduce-scatter + all-gather of the same buffer —
each captured 20x in a CUDA graph and replayed, so host launch overhead is excluded like in the training step."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200 import _lib, parallel


def timed(fn, reps=20, replays=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / (reps * replays) * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, local, world = parallel.init_from_env()
    import ctypes
    import torch.distributed._symmetric_memory as symm
    lib = _lib.load()
    n = 2130000 + (-2130000) % (4 * world)
    g = symm.empty(n, dtype=torch.float32, device="cuda")
    p = symm.empty(n, dtype=torch.float32, device="cuda")
    hg, hp = symm.rendezvous(g, dist.group.WORLD), symm.rendezvous(p, dist.group.WORLD)
    g.normal_(); p.normal_()
    sq = torch.zeros(n // world, device="cuda")
    gp = (ctypes.c_uint64 * world)(*[int(x) for x in hg.buffer_ptrs])
    pp = (ctypes.c_uint64 * world)(*[int(x) for x in hp.buffer_ptrs])
    mc = (int(hg.multicast_ptr or 0), int(hp.multicast_ptr or 0))
    plain = torch.randn(n, device="cuda")
    shard = torch.zeros(n // world, device="cuda")

    def kern(mcast=True):
        _lib.check(lib.idl_rmsprop_allreduce_step(gp, pp, mc[0] if mcast else 0, mc[1] if mcast else 0, _lib.ptr(sq), n, rank, world,
                                                  1e-3, 0.99, 1e-8, 0.01, _lib.stream_ptr()))

    def full():
        hg.barrier(channel=0); kern(); hp.barrier(channel=0)

    def rs_ag():
        dist.reduce_scatter_tensor(shard, plain, op=dist.ReduceOp.AVG)
        dist.all_gather_into_tensor(plain, shard)

    res = [("barrier", timed(lambda: hg.barrier(channel=0))), ("kernel (NVLS)" if mc[0] else "kernel (peer)", timed(kern)),
           ("kernel (peer loads/stores)", timed(lambda: kern(False))), ("barrier + kernel + barrier", timed(full)),
           ("NCCL all_reduce", timed(lambda: dist.all_reduce(plain, op=dist.ReduceOp.AVG))), ("NCCL reduce_scatter + all_gather", timed(rs_ag)),
           ("local rmsprop (full)", timed(lambda: _lib.check(lib.idl_rmsprop_step(_lib.ptr(plain), _lib.ptr(plain), _lib.ptr(plain), n, 1e-3, 0.99, 1e-8, 0.01, 1.0, _lib.stream_ptr()))))]
    if rank == 0:
        print("world %d, %d floats (%.1f MB), multicast %s" % (world, n, n * 4 / 1e6, bool(mc[0])))
        for name, us in res:
            print("%-34s %8.1f us" % (name, us))
    dist.destroy_process_group()


main()
