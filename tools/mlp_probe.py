"""Every GEMM of one training step (strict fp32, PyTorch / cuBLAS), as autograd issues it, under both BLAS back ends PyTorch can
route to, plus hand split-K (bmm) variants of the shapes with few output tiles (development aid)."""
import os, sys
import torch


def graph_time(fn, inner=20, reps=10):
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
    dev = torch.device("cuda")
    torch.manual_seed(0)
    R = lambda *s: torch.randn(*s, device=dev)
    x, W1, a1g = R(1024, 4096), R(512, 4096), R(1024, 512)
    d1, W2, latg = R(1024, 512), R(64, 512), R(1024, 64)
    d2, W3, lg = R(1024, 64), R(5, 64), R(1024, 5)
    fn_, sim = R(1024, 64), R(1024, 1024)
    b1 = R(512)
    shapes = [("L1 fwd   addmm(b, x, W1^T)", lambda: torch.addmm(b1, x, W1.t())),
              ("L1 wgrad a1g^T x", lambda: torch.mm(a1g.t(), x)),
              ("L2 fwd   d1 W2^T", lambda: torch.mm(d1, W2.t())),
              ("L2 dgrad latg W2", lambda: torch.mm(latg, W2)),
              ("L2 wgrad latg^T d1", lambda: torch.mm(latg.t(), d1)),
              ("L3 fwd   d2 W3^T", lambda: torch.mm(d2, W3.t())),
              ("L3 dgrad lg W3", lambda: torch.mm(lg, W3)),
              ("L3 wgrad lg^T d2", lambda: torch.mm(lg.t(), d2)),
              ("NCE sim  fn fn^T", lambda: torch.mm(fn_, fn_.t())),
              ("NCE dfn  W fn", lambda: torch.mm(sim, fn_))]
    for lib in ("cublas", "cublaslt"):
        torch.backends.cuda.preferred_blas_library(lib)
        tot = 0.0
        for name, f in shapes:
            t = graph_time(f)
            tot += t
            print("%-9s %-28s %7.2f us" % (lib, name, t), flush=True)
        print("%-9s total %7.2f us" % (lib, tot))
    torch.backends.cuda.preferred_blas_library("cublas")
    def split(a, b, ns):   # a [M, K], b [K, N] -> [ns, M, N] partial products over K slices
        M, K = a.shape
        aa = a.reshape(M, ns, K // ns).transpose(0, 1) if a.is_contiguous() else a.t().reshape(ns, K // ns, M).transpose(1, 2)
        bb = b.reshape(ns, K // ns, b.shape[1]) if b.is_contiguous() else b.t().reshape(b.shape[1], ns, K // ns).permute(1, 2, 0)
        return lambda: torch.bmm(aa, bb)
    for name, a, b in (("L1 fwd", x, W1.t()), ("L2 fwd", d1, W2.t()), ("L2 wgrad", latg.t(), d1), ("L3 wgrad", lg.t(), d2), ("NCE dfn", sim, fn_)):
        for ns in (2, 4, 8, 16):
            print("split-K %-9s %2d  %7.2f us" % (name, ns, graph_time(split(a, b, ns))), flush=True)


main()
