"""Two-rank NCCL test of the sharded path (needs >= 2 GPUs; skipped otherwise): scaler
statistics merged over ranks == single-GPU statistics, sharded mimics identical to the
unsharded ones (global sequence ids in the RNG), data-parallel training step (eager and CUDA-graphed, flat-gradient
all-reduce), all-gathered predictions."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from idelucs_b200 import featurise as ft
    from idelucs_b200 import parallel
    from idelucs_b200.seqset import SeqSet
    from idelucs_b200.train import ShardedTrainer
    from idelucs_b200.utils import augment_device
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(0)
    n, k = 600, 5
    lengths = rng.integers(300, 3000, size=n)
    alph = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [alph[rng.integers(0, 4, size=int(L))].tobytes() for L in lengths]
    ranges = parallel.shard_ranges(lengths, world)
    lo, hi = ranges[rank]
    ss = SeqSet.from_sequences(seqs[lo:hi], device=dev)
    prof, sc, _ = augment_device(ss, 6, k=k, seed=42, group=dist.group.WORLD, seq_id0=lo)
    ok = True
    if rank == 0:  # the unsharded computation on one GPU
        full = SeqSet.from_sequences(seqs, device=dev)
        pf, scf, _ = augment_device(full, 6, k=k, seed=42)
        ok = ok and torch.allclose(sc.mean64, scf.mean64, rtol=1e-13, atol=0) and torch.allclose(sc.scale64, scf.scale64, rtol=1e-10)
        ok = ok and torch.allclose(prof, pf[:, lo:hi], rtol=1e-5, atol=1e-5)
        ok = ok and bool((prof == pf[:, lo:hi]).float().mean() > 0.99)
    tr = ShardedTrainer(ss, k=k, n_clusters=4, n_mimics=6, batch_sz=64, seed=3, seq_id0=lo, world=world)
    graphed = tr.enable_cuda_graph()                            # step + flat-gradient all-reduce in one CUDA graph
    losses = [float(tr.step().item()) for _ in range(4)]
    ok = ok and all(np.isfinite(losses))
    w = next(tr.net.parameters()).detach().clone()
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    ok = ok and all(torch.equal(ws[0], x) for x in ws)          # replicas stay identical (averaged gradients)
    if rank == 0:
        ret["graph"] = bool(graphed)
        ret["graph_error"] = getattr(tr, "_graph_error", "")
    tr._graph = None                                            # the same step eagerly
    losses = [float(tr.step().item()) for _ in range(3)]
    ok = ok and all(np.isfinite(losses))
    w = next(tr.net.parameters()).detach().clone()
    dist.all_gather(ws, w)
    ok = ok and all(torch.equal(ws[0], x) for x in ws)
    preds, probs, lat = tr.predict(ss, k=k)      # inference tail: all three row blocks all-gathered in rank order
    ok = ok and preds.shape[0] == n and probs.shape[0] == n and tuple(lat.shape) == (n, 64)
    if rank == 0:   # ... and equal to scoring the unsharded set with the same (replicated) network
        x = ft.Scaler.fit(ft.profiles(full, k, [ft.VariantSpec(ft.KIND_CLEAN)], out_kind=ft.OUT_FREQ_F64)[0])
        f64 = ft.profiles(full, k, [ft.VariantSpec(ft.KIND_CLEAN)], out_kind=ft.OUT_FREQ_F64)[0]
        xs = x.transform64(f64, want32=True)
        tr.net.eval()
        with torch.no_grad():
            out, lat_full = tr.net(xs)
        tr.net.train()
        ok = ok and torch.allclose(lat_full, lat, rtol=1e-4, atol=1e-4) and bool((out.argmax(1) == preds).float().mean() > 0.99)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret["ok"] = bool(t.item() == 1.0)
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_two_gpu_sharding():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 200, ret), nprocs=2, join=True)
    assert ret["ok"]
    assert ret["graph"], ret["graph_error"]
