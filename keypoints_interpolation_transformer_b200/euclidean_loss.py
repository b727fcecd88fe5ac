"""Drop-in for the reference's ``euclidean_loss.py`` on the fused loss kernel (one pass over
pred/target that yields the scalar AND d loss / d pred).  ``EuclideanLoss()(output, target)``
returns a 0-dim tensor supporting ``.backward()``, ``.float()``, ``.clone().detach().cpu().numpy()``
exactly as A1_train.py:128-134 uses it."""
import torch
import torch.nn as nn

from . import _lib as K


def fused_loss(pred, target, frame_weight=None, kind=K.LOSS_EUCLID, want_grad=True, grad_scale=1.0):
    """(loss [0-dim], dpred or None).  pred/target: [..., K, 2] fp32 CUDA; frame_weight: one 0/1
    float per frame (the eval blend of A1_train.py:184) or None."""
    if not pred.is_cuda:
        raise K.KitError("fused_loss needs CUDA tensors: this build has no CPU path")
    p = pred.contiguous().float()
    t = target.to(p.device).contiguous().float()
    if p.shape != t.shape or p.shape[-1] != 2:
        raise K.KitError(f"loss shapes {tuple(p.shape)} vs {tuple(t.shape)} must match and end in [K, 2]")
    Kp = p.shape[-2]
    n_frames = p.numel() // (2 * Kp)
    w = None
    if frame_weight is not None:
        w = frame_weight.to(p.device).contiguous().float().reshape(-1)
        if w.numel() != n_frames:
            raise K.KitError(f"frame_weight has {w.numel()} entries for {n_frames} frames")
    lib = K.lib()
    partials = torch.empty(lib.kit_loss_partials(n_frames, Kp), device=p.device)
    loss = torch.empty((), device=p.device)
    dpred = torch.empty_like(p) if want_grad else None
    K.check(lib.kit_loss_fwd_bwd(K.ptr(p), K.ptr(t), K.ptr(w), n_frames, Kp, kind, float(grad_scale), K.ptr(loss),
                                 K.ptr(dpred), K.ptr(partials), K.stream_ptr()))
    return loss, dpred


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, frame_weight, kind):
        loss, dpred = fused_loss(output, target, frame_weight, kind, want_grad=output.requires_grad)
        ctx.save_for_backward(dpred) if dpred is not None else None
        ctx.shape = output.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return (dpred * g).view(ctx.shape), None, None, None


class EuclideanLoss(nn.Module):
    """euclidean_loss.py:4-17 -- mean over points of the squared euclidean distance."""

    def forward(self, output, target):
        return _LossFn.apply(output, target, None, K.LOSS_EUCLID)


class MSELoss(nn.Module):
    """A1_train.py:254 ``criterion = MSELoss()`` on the same fused kernel (= EuclideanLoss / 2)."""

    def forward(self, output, target):
        return _LossFn.apply(output, target, None, K.LOSS_MSE)


class MaskedEuclideanLoss(nn.Module):
    """The eval protocol of A1_train.py:184-186 in one pass: EuclideanLoss(pred*m + y*(1-m), y)."""

    def forward(self, output, target, frame_mask):
        return _LossFn.apply(output, target, frame_mask, K.LOSS_EUCLID)


class EuclideanDistanceLoss(nn.Module):
    """euclidean_loss.py:19-37 (A4 validation criterion): sum over points of the L2 distance."""

    def forward(self, output, target):
        return _LossFn.apply(output, target, None, K.LOSS_DISTANCE)
