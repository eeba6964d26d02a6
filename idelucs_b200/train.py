"""Sharded training loop on top of the hot path (BASELINE.json configs[3]: synthetic 1 M x 2 kb,
k=6, n_mimics=50, batch_sz=512, sharded over 1/2/4/8 B200).

Each rank owns a contiguous shard of the packed sequences, regenerates its pair batches on
the fly with the mimic kernel (selection mode — the 1.6 TB x_train of the reference never
exists), runs the reference's training step (idelucs/models.py:113-143: two forwards,
(1-w) InfoNCE + w IIC loss, backward, RMSprop) with the fused IIC kernel, and all-reduces the
gradients of the data-parallel MLP replicas over NCCL (torch DDP).  batch_sz is PER RANK (weak
scaling in the batch, strong in the data set)."""
import torch
import torch.distributed as dist

from . import featurise as ft
from .LossFunctions import IID_loss, info_nce_loss
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
        self.net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[self.dev.index]) if world > 1 else net
        self.opt = torch.optim.RMSprop(self.net.parameters(), lr=lr, weight_decay=0.01, capturable=True)
        self.lamb, self.weight, self.batch_sz = lamb, weight, batch_sz
        self.gen = torch.Generator(device=self.dev).manual_seed(seed * 1000003 + seq_id0 + 1)
        self._graph = None
        self._ids = torch.zeros(batch_sz, dtype=torch.int64, device=self.dev)

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
            with torch.cuda.graph(g, stream=s):
                self._loss = self._step_from_ids()
            self._graph = g
            return True
        except Exception as e:  # noqa: BLE001 — eager fallback of an optimisation, not of the CUDA path
            self._graph = None
            self._graph_error = repr(e)
            torch.cuda.synchronize(self.dev)
            return False

    def step(self):
        """one training step on a random batch of this rank's pairs; returns the loss tensor"""
        self._ids.copy_(torch.randint(0, self.loader.n_pairs, (self.batch_sz,), device=self.dev, generator=self.gen))
        if self._graph is not None:
            self._graph.replay()
            return self._loss
        return self._step_from_ids()

    def _step_from_ids(self):
        batch = self.loader.batch(self._ids)
        self.opt.zero_grad(set_to_none=True)
        z1, h1 = self.net(batch["true"])
        z2, h2 = self.net(batch["modified"])
        loss = (1 - self.weight) * info_nce_loss(h1, h2, 0.85) + self.weight * IID_loss(z1, z2, lamb=self.lamb)
        loss.backward()
        self.opt.step()
        return loss.detach()

    @torch.no_grad()
    def predict(self, seqset, k=6, batch=4096):
        """cluster assignments of this rank's shard, all-gathered in rank order (models.py:145-172)"""
        from . import parallel
        f64 = ft.profiles(seqset, k, [ft.VariantSpec(ft.KIND_CLEAN)], out_kind=ft.OUT_FREQ_F64)[0]
        sc = ft.Scaler.fit(f64, group=dist.group.WORLD if self.world > 1 else None)
        x = sc.transform64(f64, want32=True)
        net = self.net.module if self.world > 1 else self.net
        net.eval()
        preds = torch.cat([torch.max(net(x[b:b + batch])[0], 1)[1] for b in range(0, x.shape[0], batch)])
        net.train()
        if self.world > 1:
            counts = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
            dist.all_gather(counts, torch.tensor([preds.shape[0]], dtype=torch.int64, device=self.dev))
            preds = parallel.all_gather_rows(preds, [int(c.item()) for c in counts])
        return preds
