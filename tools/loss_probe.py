"""Per-kernel times of the training losses (development aid): every piece replayed 50x from its own CUDA graph."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200 import _lib, LossFunctions as LF


def graph_time(fn, inner=50, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1000.0 / inner)
    return best


def main():
    lib = _lib.load()
    B, C, D = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 5, 64
    dev = torch.device("cuda")
    torch.manual_seed(0)
    z = torch.softmax(torch.randn(2 * B, C, device=dev), dim=1).contiguous()
    h = torch.randn(2 * B, D, device=dev).contiguous()
    n2 = 2 * B
    loss = torch.empty((), device=dev)
    dz = torch.empty_like(z)
    ws = LF._ws(dev, C)
    half = 4 * B * C
    st = lambda: _lib.stream_ptr()
    def iid_py():
        LF._iid_device(z[:B], z[B:], 2.8, 2.2e-16, dz1=dz[:B], dz2=dz[B:])
    def iid():
        _lib.check(lib.idl_iid_loss(_lib.ptr(z), ctypes.c_void_p(z.data_ptr() + half), B, C, 2.8, 2.2e-16, _lib.ptr(loss), None,
                                    _lib.ptr(dz), ctypes.c_void_p(dz.data_ptr() + half), _lib.ptr(ws), ws.numel(), st()))
    fn_ = torch.empty_like(h); scratch = torch.empty(3 * n2, device=dev); dh = torch.empty_like(h)
    sim = torch.empty(n2, n2, device=dev); dfn = torch.empty_like(h)
    def norm():
        _lib.check(lib.idl_nce_normalize(_lib.ptr(h), n2, D, _lib.ptr(fn_), _lib.ptr(scratch), st()))
    def mm1():
        torch.mm(fn_, fn_.t(), out=sim)
    def xent():
        _lib.check(lib.idl_nce_softmax_xent(_lib.ptr(sim), n2, 0.85, ctypes.c_void_p(scratch.data_ptr() + 4 * n2),
                                            ctypes.c_void_p(scratch.data_ptr() + 8 * n2), _lib.ptr(loss), st()))
    def mm2():
        torch.mm(sim, fn_, out=dfn)
    fn_t = fn_.t().contiguous(); dfn_t = torch.empty(D, n2, device=dev)
    def mm2_t():      # W symmetric: (fn^T W)^T
        torch.mm(fn_.t(), sim, out=dfn_t)
    def mm2_tc():
        torch.mm(fn_t, sim, out=dfn_t)
    def mm2_nt():
        torch.mm(sim, fn_t.t(), out=dfn)
    def mm2_split():  # K split in 4 by hand: [4, n2, n2/4] x [4, n2/4, D] -> sum
        a = sim.view(n2, 4, n2 // 4).transpose(0, 1)
        b = fn_.view(4, n2 // 4, D)
        torch.bmm(a, b).sum(0)
    def make_split(ns):
        a = sim.view(n2, ns, n2 // ns).transpose(0, 1)
        b = fn_.view(ns, n2 // ns, D)
        o = torch.empty(ns, n2, D, device=dev)
        return lambda: torch.bmm(a, b, out=o)
    splits = [("bmm split-K %d (no sum)" % ns, make_split(ns)) for ns in (2, 4, 8, 16, 32)]
    def lse_only():
        _lib.check(lib.idl_nce_softmax_xent(_lib.ptr(sim), n2, 0.85, ctypes.c_void_p(scratch.data_ptr() + 4 * n2),
                                            ctypes.c_void_p(scratch.data_ptr() + 8 * n2), None, st()))
    def nbwd():
        _lib.check(lib.idl_nce_normalize_backward(_lib.ptr(dfn), _lib.ptr(fn_), _lib.ptr(scratch), n2, D, _lib.ptr(dh), st()))
    zr = z.clone().requires_grad_(True); hr = h.clone().requires_grad_(True)
    def both():
        zr.grad = None; hr.grad = None
        LF.train_losses(zr, hr, 2.8, 0.25).backward()
    def empty():
        loss.add_(1.0)
    norm(); mm1()
    for name, f in splits + [("tiny torch kernel (add_)", empty), ("idl_iid_loss (fwd + grad)", iid), ("IIC through the Python layer", iid_py), ("idl_nce_normalize", norm), ("mm sim", mm1),
                    ("idl_nce_softmax_xent", xent), ("mm dfn", mm2), ("mm dfn as (fn^T W)", mm2_t), ("mm dfn as (fn_t contiguous) W", mm2_tc), ("mm dfn NT", mm2_nt), ("mm dfn split-K bmm", mm2_split), ("idl_nce_normalize_backward", nbwd), ("train_losses fwd + bwd", both)]:
        print("%-32s %7.2f us" % (name, graph_time(f)), flush=True)


main()
