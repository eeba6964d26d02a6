"""Loss functions with the reference's names and signatures (idelucs/LossFunctions.py).

``IID_loss`` / ``compute_joint`` run the fused sm_100a kernel K5 (forward + backward in one launch) through the C ABI.
``info_nce_loss`` runs the fused InfoNCE kernels (idl_info_nce: normalise, similarity, masked
log-softmax, cross-entropy and the gradient in three launches) on CUDA tensors."""
import sys

import torch
import torch.nn.functional as F

from . import _lib


def _ws(device, C):
    from .featurise import _workspace
    return _workspace(device, _lib.load().idl_iid_loss_workspace_bytes(C), "loss%d" % C)


class _IIDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_out, x_tf_out, lamb, EPS):
        lib = _lib.load()
        if not x_out.is_cuda:
            raise _lib.IdelucsB200Error("IID_loss: tensors must be on a CUDA device (no CPU fallback)")
        z1 = x_out.detach().contiguous().float()
        z2 = x_tf_out.detach().contiguous().float()
        B, C = z1.shape
        assert z2.shape == (B, C)
        loss = torch.empty((), dtype=torch.float32, device=z1.device)
        need_grad = x_out.requires_grad or x_tf_out.requires_grad
        dz1 = torch.empty_like(z1) if need_grad else None
        dz2 = torch.empty_like(z2) if need_grad else None
        with torch.cuda.device(z1.device):
            ws = _ws(z1.device, C)
            _lib.check(lib.idl_iid_loss(_lib.ptr(z1), _lib.ptr(z2), B, C, float(lamb), float(EPS), _lib.ptr(loss), None,
                                        _lib.ptr(dz1), _lib.ptr(dz2), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        if need_grad:
            ctx.save_for_backward(dz1, dz2)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        dz1, dz2 = ctx.saved_tensors
        return grad_out * dz1, grad_out * dz2, None, None


def IID_loss(x_out, x_tf_out, lamb=1.0, EPS=sys.float_info.epsilon):
    """idelucs/LossFunctions.py:20-46 — same arguments, returns a 0-d tensor with autograd."""
    return _IIDLoss.apply(x_out, x_tf_out, lamb, EPS)


def compute_joint(x_out, x_tf_out):
    """idelucs/LossFunctions.py:49-62 — symmetrised, normalised joint [C, C] (no autograd)."""
    lib = _lib.load()
    z1 = x_out.detach().contiguous().float()
    z2 = x_tf_out.detach().contiguous().float()
    B, C = z1.shape
    joint = torch.empty((C, C), dtype=torch.float32, device=z1.device)
    with torch.cuda.device(z1.device):
        ws = _ws(z1.device, C)
        _lib.check(lib.idl_iid_loss(_lib.ptr(z1), _lib.ptr(z2), B, C, 1.0, float(sys.float_info.epsilon), None,
                                    _lib.ptr(joint), None, None, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return joint


_targets = {}


class _InfoNCE(torch.autograd.Function):
    """fused CUDA form (idl_info_nce): value and gradient with respect to the stacked latent in three launches"""

    @staticmethod
    def forward(ctx, h, temperature):
        lib = _lib.load()
        x = h.detach().contiguous().float()
        n2, D = x.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        dh = torch.empty_like(x)
        from .featurise import _workspace
        with torch.cuda.device(x.device):
            ws = _workspace(x.device, lib.idl_info_nce_workspace_bytes(n2, D), "nce%d_%d" % (n2, D))
            _lib.check(lib.idl_info_nce(_lib.ptr(x), n2, D, float(temperature), _lib.ptr(loss), _lib.ptr(dh), _lib.ptr(ws), ws.numel(),
                                        _lib.stream_ptr()))
        ctx.save_for_backward(dh)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dh,) = ctx.saved_tensors
        return grad_out * dh, None


def _info_nce_torch(h, temperature):
    """mask-gather-free PyTorch formulation (any device / width): cross-entropy of the self-masked similarity rows against
    the index of the other view == the reference's [positive, negatives] gather with label 0"""
    n2 = h.shape[0]
    feats = F.normalize(h.float(), dim=1)
    logits = (feats @ feats.T) / temperature
    logits.fill_diagonal_(float("-inf"))      # in place: the division's backward does not need its output
    key = (n2, logits.device)
    if key not in _targets:
        idx = torch.arange(n2, device=logits.device)
        _targets[key] = (idx + n2 // 2) % n2
    return F.cross_entropy(logits, _targets[key])


def info_nce_loss_stacked(h, temperature):
    """info_nce_loss on the two views already stacked as one [2n, d] tensor (rows 0..n-1 = first view).  CUDA tensors with a
    latent width of 32 / 64 / 128 (the reference's encoders have 64) take the fused kernels; anything else the PyTorch form."""
    if h.is_cuda and h.dim() == 2 and h.shape[1] in (32, 64, 128) and h.shape[0] % 2 == 0 and h.shape[0] >= 2:
        return _InfoNCE.apply(h, temperature)
    return _info_nce_torch(h, temperature)


def info_nce_loss(z1, z2, temperature):
    """SimCLR NT-Xent (idelucs/LossFunctions.py:65-98).  The reference gathers [positive, negatives] per row with boolean
    masks and takes cross-entropy against label 0; that equals the cross-entropy of the self-masked similarity row against
    the index of the other view, which needs no mask gathers (and no host synchronisation)."""
    return info_nce_loss_stacked(torch.cat((z1, z2), 0), temperature)
