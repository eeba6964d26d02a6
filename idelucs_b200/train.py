"""Sharded training loop on top of the hot path (BASELINE.json configs[3]: synthetic 1 M x 2 kb,
k=6, n_mimics=50, batch_sz=512, sharded over 1/2/4/8 B200).

Each rank owns a contiguous shard of the packed sequences, regenerates its pair batches on
the fly with the mimic kernel (selection mode — the 1.6 TB x_train of the reference never
exists), runs the reference's training step (idelucs/models.py:113-143: two forwards,
(1-w) InfoNCE + w IIC loss, backward, RMSprop) with the fused IIC kernel, and all-reduces the
gradients of the data-parallel MLP replicas over NCCL: the replicas' .grad tensors are views into ONE flat
buffer, averaged by a single all-reduce per step (8.5 MB at k=6) that is captured in the step's CUDA graph
together with everything else.  batch_sz is PER RANK (weak scaling in the batch, strong in the data set)."""
import torch
import torch.distributed as dist

from . import featurise as ft
from .LossFunctions import IID_loss, info_nce_loss, info_nce_loss_stacked
from .PytorchUtils import NetLinear
from .models import weights_init


class ShardedTrainer(object):
    def __init__(self, seqset, k=6, n_clusters=5, n_mimics=50, batch_sz=512, lamb=2.8, weight=0.25, lr=1e-3, seed=0,
                 seq_id0=0, world=1, materialize_bytes=0):
        from .utils import PairBatchLoader
        self.dev = seqset.device
        self.world = world
        group = dist.group.WORLD if world > 1 else None
        self.loader = PairBatchLoader(seqset, n_mimics, k=k, batch_size=batch_sz, seed=seed, group=group, seq_id0=seq_id0,
                                      materialize_bytes=materialize_bytes, drop_last=True)
        torch.manual_seed(seed)  # identical initial replicas on every rank
        net = NetLinear(4 ** k, n_clusters)
        net.apply(weights_init)
        net.to(self.dev)
        self.net = net
        self._flat_grad = None
        if world > 1:   # identical replicas (same seed); gradients live in one flat buffer = one NCCL call per step
            params = [p for p in net.parameters() if p.requires_grad]
            self._flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=self.dev)
            o = 0
            for p in params:
                p.grad = self._flat_grad[o:o + p.numel()].view_as(p)
                o += p.numel()
            for p in params:   # replicas must start identical whatever the RNG state of the rank was
                dist.broadcast(p.data, src=0)
        self.opt = torch.optim.RMSprop(self.net.parameters(), lr=lr, weight_decay=0.01, capturable=True)
        self.lamb, self.weight, self.batch_sz = lamb, weight, batch_sz
        self.gen = torch.Generator(device=self.dev).manual_seed(seed * 1000003 + seq_id0 + 1)
        self._graph = None
        self._ids = torch.zeros(batch_sz, dtype=torch.int64, device=self.dev)
        self._perm, self._cursor, self.epoch = None, 0, 0

    def enable_cuda_graph(self, warmup=11):
        """Capture featurise -> forward x2 -> losses -> backward -> RMSprop in ONE CUDA graph: the
        step is launch-bound (~100 small kernels), so replaying a graph removes the host from the
        loop.  The pair ids are the only input (static buffer filled before every replay).
        Returns False (and stays eager) if capture is not possible."""
        try:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self._step_from_ids()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            if self.world > 1:
                # the NCCL watchdog thread polls events while this thread captures: only this thread's calls may
                # invalidate the capture
                torch.cuda.synchronize(self.dev)
                with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                    self._loss = self._step_from_ids()
            else:
                with torch.cuda.graph(g, stream=s):
                    self._loss = self._step_from_ids()
            self._graph = g
        except Exception as e:  # noqa: BLE001 — eager fallback of an optimisation, not of the CUDA path
            self._graph = None
            self._graph_error = repr(e)
            torch.cuda.synchronize(self.dev)
        if self.world > 1:   # all ranks replay the graph or none does (a rank that fell back must not face captured collectives alone)
            okf = torch.tensor([1.0 if self._graph is not None else 0.0], device=self.dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if float(okf.item()) < 1.0:
                self._graph = None
        return self._graph is not None

    def _next_ids(self):
        """the next batch of a shuffled epoch over this rank's pairs: every pair once per epoch, a fresh permutation per
        epoch, the ragged tail dropped (DataLoader(shuffle=True) of idelucs/utils.py:427; the per-rank batch is fixed-size
        because it feeds a captured CUDA graph)"""
        n = self.loader.n_pairs
        if n < self.batch_sz:   # degenerate shard: sample with replacement
            return torch.randint(0, max(n, 1), (self.batch_sz,), device=self.dev, generator=self.gen)
        if self._perm is None or self._cursor + self.batch_sz > n:
            self._perm = torch.randperm(n, device=self.dev, generator=self.gen)
            self._cursor = 0
            self.epoch += 1
        ids = self._perm[self._cursor:self._cursor + self.batch_sz]
        self._cursor += self.batch_sz
        return ids

    def step(self):
        """one training step on the next batch of this rank's shuffled pairs; returns the loss tensor"""
        self._ids.copy_(self._next_ids())
        if self._graph is not None:
            self._graph.replay()
            return self._loss
        return self._step_from_ids()

    def _step_from_ids(self):
        batch = self.loader.batch(self._ids)
        if self._flat_grad is None:
            self.opt.zero_grad(set_to_none=True)
        else:
            self._flat_grad.zero_()
        if "both" in batch:   # one pass over the stacked [2B, F] batch: same per-row math as the reference's two forwards
            B = batch["true"].shape[0]   # (models.py:121-122; dropout masks are independent per row either way), half the launches
            z, h = self.net(batch["both"])
            nce = info_nce_loss_stacked(h, 0.85)
            z1, z2 = z[:B], z[B:]
        else:
            z1, h1 = self.net(batch["true"])
            z2, h2 = self.net(batch["modified"])
            nce = info_nce_loss(h1, h2, 0.85)
        loss = (1 - self.weight) * nce + self.weight * IID_loss(z1, z2, lamb=self.lamb)
        loss.backward()
        if self._flat_grad is not None:
            dist.all_reduce(self._flat_grad, op=dist.ReduceOp.AVG)
        self.opt.step()
        return loss.detach()

    @torch.no_grad()
    def predict(self, seqset, k=6, batch=4096):
        """idelucs/models.py:145-172 over the sharded sequences: (y_pred int64[N], probabilities float32[N], latent
        float32[N, 64]) of ALL sequences on every rank — each rank featurises and scores its own shard (clean float64
        profiles, one StandardScaler over all shards), then the three row blocks are all-gathered in rank order."""
        from . import parallel
        f64 = ft.profiles(seqset, k, [ft.VariantSpec(ft.KIND_CLEAN)], out_kind=ft.OUT_FREQ_F64)[0]
        sc = ft.Scaler.fit(f64, group=dist.group.WORLD if self.world > 1 else None)
        x = sc.transform64(f64, want32=True)
        net = self.net
        net.eval()
        preds, probs, lats = [], [], []
        for b in range(0, x.shape[0], batch):
            out, lat = net(x[b:b + batch])
            pr, pd = torch.max(out, 1)
            preds.append(pd); probs.append(pr); lats.append(lat)
        net.train()
        C = lats[0].shape[1] if lats else 64
        preds = torch.cat(preds) if preds else torch.zeros(0, dtype=torch.int64, device=self.dev)
        probs = torch.cat(probs) if probs else torch.zeros(0, dtype=torch.float32, device=self.dev)
        lats = torch.cat(lats) if lats else torch.zeros((0, C), dtype=torch.float32, device=self.dev)
        if self.world > 1:
            preds, probs, lats = parallel.all_gather_ragged([preds, probs, lats])
        return preds, probs, lats
