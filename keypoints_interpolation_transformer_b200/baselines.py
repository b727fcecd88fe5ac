"""The evaluation's cubic-spline baseline (3_test_cubic_interpolation.py:32-58) on the device: same call as the reference's
``cubic_interpolation(data, mask)`` for one sequence, or a whole batch at once (the reference's pandas loop takes ~0.27 s per
T = 256 sequence; BASELINE configs[3] compares the model with it on 4096 of them)."""
import torch

from . import _lib as K


def cubic_interpolation(data, mask):
    """data [T+1,K,2] with mask [1,T+1] / [T+1] (the reference's arguments), or data [B,T+1,K,2] with mask [B,T+1].  Frames
    with mask == 1 and every exact 0 are missing; each (keypoint, coordinate) series is filled by the not-a-knot cubic spline
    through its remaining samples, extrapolating at both ends; a series without samples becomes 0.  Returns a new fp32 tensor
    of data's shape on data's device (CUDA required: this build has no CPU path)."""
    if not data.is_cuda:
        raise K.KitError("cubic_interpolation needs CUDA tensors: this build has no CPU path")
    single = data.dim() == 3
    d = data.detach().float().contiguous()
    if single:
        d = d.unsqueeze(0)
    B, T1, Kp, two = d.shape
    if two != 2:
        raise K.KitError(f"data must end in [K, 2], got {tuple(data.shape)}")
    m = mask.detach().to(d.device).float().reshape(-1, T1).contiguous()
    if m.shape[0] != B:
        raise K.KitError(f"mask shape {tuple(mask.shape)} does not match data {tuple(data.shape)}")
    out = torch.empty_like(d)
    K.check(K.lib().kit_cubic_interpolate(K.ptr(d), K.ptr(m), K.ptr(out), B, T1, Kp, K.stream_ptr()))
    return out[0] if single else out
