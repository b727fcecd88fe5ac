"""CPU oracle for the keypoint-interpolation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-tensor CPU restatement (torch fp32 on CPU + numpy, no nn.Transformer,
no nn.Module) of the one hot path this repository accelerates: the train / infer step of the
reference's ``KeypointCompleter``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it -- and only as the
checker or the timed CPU baseline, never as a product code path.  The product path
(``keypoints_interpolation_transformer_b200``) raises if its CUDA library is missing.

Parity status: the reference ships no tests, golden vectors or fixtures,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build
container by importing /root/reference (``tests/golden/make_golden.py``) and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against
them.

Every function cites the reference file:line it restates (paths relative to /root/reference,
``torch/`` = the PyTorch 2.11 the reference runs on).
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np
import torch

NEG_INF = float("-inf")


# --------------------------------------------------------------------------------------
# model.py
# --------------------------------------------------------------------------------------
def positional_table(max_len: int, dim: int) -> torch.Tensor:
    """model.py:34-46 -- sin/cos table, returned as [max_len, dim] (the reference buffer is
    the same numbers shaped [max_len, 1, dim])."""
    pe = torch.zeros(max_len, dim)
    pos = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
    div = torch.exp(torch.arange(0, dim, 2).float() * (-math.log(10000.0)) / dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def linear(x, w, b):
    return x @ w.t() + b


def swiglu(x, p, prefix):
    """model.py:18-22 -- fc3(fc1(x) * sigmoid(fc2(x))) (sigmoid gate, not SiLU)."""
    x1 = linear(x, p[prefix + ".fc1.weight"], p[prefix + ".fc1.bias"])
    x2 = linear(x, p[prefix + ".fc2.weight"], p[prefix + ".fc2.bias"])
    return linear(x1 * torch.sigmoid(x2), p[prefix + ".fc3.weight"], p[prefix + ".fc3.bias"])


def token_norm(x, eps: float = 1e-5):
    """model.py:124-125,150 -- nn.InstanceNorm1d(hidden) applied to [S,N,E]: PyTorch treats it
    as (batch=S, channels=N, length=E) and normalises over E per token, biased variance, no
    affine (torch/nn/modules/instancenorm.py:104-122).  Here x is [..., E]."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def layer_norm(x, w, b, eps: float = 1e-5):
    """torch/nn/modules/transformer.py:956 (post-norm LayerNorm, eps 1e-5, biased variance)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x):
    """activation="gelu" in model.py:87 -> F.gelu exact (erf) form."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def mha(xq, xkv, p, prefix, nh, bias=None):
    """torch/nn/functional.py:6244 multi_head_attention_forward, packed in-proj, dropout 0.
    xq [B,Sq,E], xkv [B,Sk,E]; bias: additive float [B,Sq,Sk] (or broadcastable) or None.
    A float key-padding mask is ADDED to the logits (functional.py:6192, :6620)."""
    B, Sq, E = xq.shape
    Sk = xkv.shape[1]
    d = E // nh
    w = p[prefix + ".in_proj_weight"]
    b = p[prefix + ".in_proj_bias"]
    q = linear(xq, w[:E], b[:E]).view(B, Sq, nh, d).transpose(1, 2)
    k = linear(xkv, w[E:2 * E], b[E:2 * E]).view(B, Sk, nh, d).transpose(1, 2)
    v = linear(xkv, w[2 * E:], b[2 * E:]).view(B, Sk, nh, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(d))
    if bias is not None:
        s = s + bias[:, None, :, :]
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Sq, E)
    return linear(o, p[prefix + ".out_proj.weight"], p[prefix + ".out_proj.bias"])


def repeat_inc_bias(frame_mask: torch.Tensor) -> torch.Tensor:
    """model.py:193-202 in closed form, batched: frame_mask [B,T] (0/1) -> [B,T,T] with
    -inf iff (j > i and frame_mask[j] == 1) else 0."""
    B, T = frame_mask.shape
    i = torch.arange(T).view(1, T, 1)
    j = torch.arange(T).view(1, 1, T)
    hit = (j > i) & (frame_mask.view(B, 1, T) == 1)
    out = torch.zeros(B, T, T)
    return out.masked_fill(hit, NEG_INF)


def get_mask(mask: torch.Tensor, size: int, matrix_type: str = "triangle") -> torch.Tensor:
    """model.py:172-209 (single sequence; mask is a 1-D 0/1 float tensor of length size)."""
    if matrix_type == "triangle":
        i = torch.arange(size).view(size, 1)
        j = torch.arange(size).view(1, size)
        return torch.zeros(size, size).masked_fill(j > i, NEG_INF)
    if matrix_type == "repeat":
        return mask.clone().float().view(1, size).repeat(size, 1)
    if matrix_type == "repeat-inc":
        return repeat_inc_bias(mask.view(1, size).float())[0]
    if matrix_type == "all":
        return torch.zeros(size, size)
    raise ValueError("Choose a correct matrixType")


def completer_forward(p, inputs, filled, nh, src_pad=None, src_bias=None, tgt_bias=None,
                      return_aux=False, cycle=False, tgt_pad=None):
    """model.py:100-170, batch-first restatement.  ``cycle=True``: KeypointCompleterCycle (model.py:212-321) -- the
    token-norm output enters the position sum TWICE (:283-284, the trig encoder already adds it) and ``tgt_pad`` [B,T]
    is handed to nn.Transformer as tgt_key_padding_mask (:294), i.e. ADDED to the decoder self-attention logits.

    inputs, filled : [B,T,K,2] fp32 (encoder input x, decoder input x_no_sota)
    src_pad        : [B,T] float key-padding mask, ADDED to encoder self-attn logits (fact 3)
    src_bias       : [B,T,T] additive encoder self-attn mask (e.g. repeat_inc_bias(x_mask))
    tgt_bias       : [B,T,T] additive decoder self-attn mask; cross-attention is unmasked and
                     tgt_pad_mask is dropped (model.py:143).
    Returns [B,T,K,2].  Calling it per sequence (B=1) is exactly the A1 un-batched call
    (A1_train.py:117-124)."""
    B, T, K, _ = inputs.shape
    H = p["input_embedding.weight"].shape[0]
    x_in = inputs.reshape(B, T, 2 * K).float()
    x_fl = filled.reshape(B, T, 2 * K).float()
    L = 0
    while f"transformer.encoder.layers.{L}.linear1.weight" in p:
        L += 1

    input_emb = linear(x_in, p["input_embedding.weight"], p["input_embedding.bias"])   # :120
    filled_emb = linear(x_fl, p["filled_embedding.weight"], p["filled_embedding.bias"])  # :121
    pe_i = p["trig_input_positional_encoder.pos_encoding"].reshape(-1, H)[:T]
    pe_f = p["trig_filled_positional_encoder.pos_encoding"].reshape(-1, H)[:T]
    ns = 2.0 if cycle else 1.0
    x = ns * token_norm(input_emb) + pe_i + p["learned_input_positional_encoder"].view(1, 1, H)   # :124-131 / :283
    y = ns * token_norm(filled_emb) + pe_f + p["learned_filled_positional_encoder"].view(1, 1, H)  # :125-132 / :284
    if cycle and tgt_pad is not None:
        tgt_bias = (tgt_bias if tgt_bias is not None else torch.zeros(B, T, T)) + tgt_pad.float().view(B, 1, T)
    x = swiglu(x, p, "swiGlu_input_prev")      # :136
    y = swiglu(y, p, "swiGlu_filled_prev")     # :137

    enc_bias = None
    if src_bias is not None or src_pad is not None:
        enc_bias = torch.zeros(B, T, T)
        if src_bias is not None:
            enc_bias = enc_bias + src_bias
        if src_pad is not None:
            enc_bias = enc_bias + src_pad.float().view(B, 1, T)

    for l in range(L):   # torch/nn/modules/transformer.py:956 (post-norm encoder layer)
        pre = f"transformer.encoder.layers.{l}"
        x = layer_norm(x + mha(x, x, p, pre + ".self_attn", nh, enc_bias),
                       p[pre + ".norm1.weight"], p[pre + ".norm1.bias"])
        ff = linear(gelu_erf(linear(x, p[pre + ".linear1.weight"], p[pre + ".linear1.bias"])),
                    p[pre + ".linear2.weight"], p[pre + ".linear2.bias"])
        x = layer_norm(x + ff, p[pre + ".norm2.weight"], p[pre + ".norm2.bias"])
    mem = layer_norm(x, p["transformer.encoder.norm.weight"], p["transformer.encoder.norm.bias"])

    for l in range(L):   # transformer.py:1147-1153 (post-norm decoder layer)
        pre = f"transformer.decoder.layers.{l}"
        y = layer_norm(y + mha(y, y, p, pre + ".self_attn", nh, tgt_bias),
                       p[pre + ".norm1.weight"], p[pre + ".norm1.bias"])
        y = layer_norm(y + mha(y, mem, p, pre + ".multihead_attn", nh, None),
                       p[pre + ".norm2.weight"], p[pre + ".norm2.bias"])
        ff = linear(gelu_erf(linear(y, p[pre + ".linear1.weight"], p[pre + ".linear1.bias"])),
                    p[pre + ".linear2.weight"], p[pre + ".linear2.bias"])
        y = layer_norm(y + ff, p[pre + ".norm3.weight"], p[pre + ".norm3.bias"])
    dec = layer_norm(y, p["transformer.decoder.norm.weight"], p["transformer.decoder.norm.bias"])

    dec = swiglu(dec, p, "swiGlu_decoded")              # :147
    dec = token_norm(dec + filled_emb)                   # :150 (pre-norm embedding residual)
    dec = dec * torch.sigmoid(dec)                       # :152
    out = linear(dec, p["fc_final.weight"], p["fc_final.bias"])   # :155
    out = out.view(B, T, K, 2)                            # :160-167 generalised 54 -> K (fact 1)
    if return_aux:
        return out, {"memory": mem, "filled_emb": filled_emb}
    return out


# --------------------------------------------------------------------------------------
# euclidean_loss.py / A1_train.py step glue
# --------------------------------------------------------------------------------------
def euclidean_loss(output, target):
    """euclidean_loss.py:8-17 -- mean over points of sum_xy (o-t)^2."""
    o = output.reshape(-1, 2)
    t = target.reshape(-1, 2)
    return torch.sum((o - t) ** 2, dim=1).mean()


def euclidean_distance_loss(output, target):
    """euclidean_loss.py:19-37 (A4 validation criterion) -- SUM over points of the L2 distance."""
    return torch.norm(output.reshape(-1, 2) - target.reshape(-1, 2), dim=1).sum()


def cubic_interpolation(data, mask):
    """3_test_cubic_interpolation.py:32-58 -- data [T+1,K,2] (numpy or tensor), mask [T+1] or [1,T+1] 0/1.  Masked frames are
    zeroed, every exact 0 becomes missing, and each (keypoint, coordinate) series is filled by what pandas'
    ``interpolate(method="cubicspline", limit_direction="both")`` computes: scipy.interpolate.CubicSpline (not-a-knot,
    extrapolating) through the remaining samples; all-missing series become 0 (np.nan_to_num).  A series with ONE sample
    makes the reference raise (CubicSpline needs two points); here it is filled with that sample.  Returns float32
    [T+1,K,2]."""
    from scipy.interpolate import CubicSpline
    d = np.array(torch.as_tensor(data).detach().cpu().numpy(), dtype=np.float32, copy=True)
    m = np.asarray(torch.as_tensor(mask).detach().cpu().numpy()).reshape(-1)
    T1, K, _ = d.shape
    d[m == 1] = 0.0
    out = np.zeros_like(d)
    t = np.arange(T1, dtype=np.float64)
    for k in range(K):
        for c in range(2):
            y = d[:, k, c].astype(np.float64)
            ok = (y != 0) & ~np.isnan(y)
            n = int(ok.sum())
            if n == 0:
                continue
            if n == 1:
                out[:, k, c] = y[ok][0]
                continue
            filled = y.copy()
            filled[~ok] = CubicSpline(t[ok], y[ok])(t[~ok])
            out[:, k, c] = filled.astype(np.float32)
    return out


def mse_loss(output, target):
    """A1_train.py:254 criterion = MSELoss() (= euclidean_loss / 2)."""
    return ((output - target) ** 2).mean()


def eval_blend(pred, y, y_mask):
    """A1_train.py:184 -- pred*m + y*(1-m), m broadcast over [K,2]; y_mask [B,T] or [T]."""
    m = y_mask.float()[..., None, None]
    return pred * m + y * (1.0 - m)


def step_slices(inputs, sota, mask, zero_masked=False):
    """A1_train.py:93-100 (+ A4_train_with_pretrained.py:107-108 when zero_masked).
    inputs [B,T+1,K,2], sota [B,T,K,2], mask [B,T+1] -> x, x_no_sota, y, x_mask, y_mask."""
    x = inputs[:, :-1].float()
    x_no_sota = inputs[:, 1:].float()
    x_mask = mask[:, :-1].float()
    y_mask = mask[:, 1:].float()
    if zero_masked:
        x = torch.where(x_mask.bool()[:, :, None, None], torch.zeros_like(x), x)
    return x, x_no_sota, sota.float(), x_mask, y_mask


def train_forward_loss(p, inputs, sota, mask, nh, criterion="mse", zero_masked=False):
    """A1_train.py:91-128 batched: returns (loss, pred)."""
    x, xf, y, xm, ym = step_slices(inputs, sota, mask, zero_masked)
    pred = completer_forward(p, x, xf, nh, src_pad=xm, src_bias=repeat_inc_bias(xm),
                             tgt_bias=repeat_inc_bias(ym))
    loss = mse_loss(pred, y) if criterion == "mse" else euclidean_loss(pred, y)
    return loss, pred


def eval_forward_loss(p, inputs, sota, mask, nh, zero_masked=False):
    """A1_train.py:149-186 batched: returns (loss, blended pred)."""
    x, xf, y, xm, ym = step_slices(inputs, sota, mask, zero_masked)
    pred = completer_forward(p, x, xf, nh, src_pad=xm, src_bias=repeat_inc_bias(xm),
                             tgt_bias=repeat_inc_bias(ym))
    pred = eval_blend(pred, y, ym)
    return euclidean_loss(pred, y), pred


# --------------------------------------------------------------------------------------
# dataloader.py: normalisation, missing blocks, SOS
# --------------------------------------------------------------------------------------
def normalize_pose(data: np.ndarray, left_shoulder: int, right_shoulder: int, right_eye: int):
    """dataloader.py:71-140 -- float32 arithmetic in the reference's operation order.
    data [T,K,2] float32; returns a new array (the reference works in place)."""
    d = np.array(data, dtype=np.float32, copy=True)
    T = d.shape[0]
    f32 = np.float32
    have = False
    sx = sy = ex = ey = f32(0)
    for t in range(T):
        ls = d[t, left_shoulder]
        rs = d[t, right_shoulder]
        if ls[0] == 0.0 or rs[0] == 0.0:
            if not have:
                continue                      # :83-85 frame left untouched
        else:
            dist = f32(f32(f32(ls[0] - rs[0]) ** 2 + f32(ls[1] - rs[1]) ** 2) ** f32(0.5))
            hm = f32(dist / f32(2))                                  # :110
            sx = f32(f32(0.5) - f32(f32(3) * hm))                    # :120
            sy = f32(d[t, right_eye, 1] - f32(hm / f32(2)))
            ex = f32(f32(0.5) + f32(f32(3) * hm))                    # :121
            ey = f32(f32(0.5) + f32(f32(3.5) * hm))
            have = True
        wx = f32(ex - sx)
        wy = f32(sy - ey)
        for k in range(d.shape[1]):
            if d[t, k, 0] == 0:               # :129 skip only on x == 0
                continue
            nx = f32(f32(d[t, k, 0] - sx) / wx)
            ny = f32(f32(d[t, k, 1] - ey) / wy)
            d[t, k, 0] = nx
            d[t, k, 1] = f32(f32(1) - ny)
    return d


def missing_blocks_from_config(T: int, config: dict, rng=_pyrandom, nprng=np.random):
    """dataloader.py:364-419 -- block placement policy (non-random, per-dataset mode).
    Consumes the RNG streams in the reference's order.  Returns [(start, end), ...]."""
    block_limit = [np.percentile(nprng.normal(config["mean_consecutive_missing"],
                                              config["std_consecutive_missing"],
                                              config["samples"]), q) for q in [25, 75]]
    block_size = [np.percentile(nprng.normal(config["mean_number_missing_blocks"],
                                             config["std_number_missing_blocks"],
                                             config["samples"]), q) for q in [25, 75]]
    num_blocks_min = max(math.floor(block_limit[0]), 1)
    num_blocks_max = math.ceil(block_limit[1])
    block_size_min = max(math.floor(block_size[0]), 1)
    block_size_max = math.ceil(block_size[1])
    num_blocks = rng.randint(num_blocks_min, num_blocks_max)
    section = max(1, T // num_blocks)
    rest = T % num_blocks
    if section < block_size_max + 4:
        section = max(block_size_max + 4, 1)
        num_blocks = max(1, T // section)
        rest = T % num_blocks
    blocks = []
    for r in range(num_blocks):
        n0 = min(rng.randint(block_size_min, block_size_max), section)
        _rest = rest if r == num_blocks - 1 else 0
        off = rng.randint(0, _rest + section - n0)
        a = section * r + off
        blocks.append((a, min(a + n0, T - 1)))
    return blocks


def hold_fill_sources(T: int, blocks):
    """dataloader.py:421-434 as an index map: frame t of the output is frame src[t] of the
    input.  Block 0 copies the frame AFTER it, later blocks the frame BEFORE (already
    possibly overwritten -- the sequential semantics are preserved by chasing indices)."""
    src = list(range(T))
    mask = [0.0] * T
    for n, (a, b) in enumerate(blocks):
        ref = b if n == 0 else a - 1
        for t in range(a, b):
            src[t] = src[ref]
            mask[t] = 1.0
    return np.asarray(src, dtype=np.int32), np.asarray(mask, dtype=np.float32)


def random_missing_sources(T: int, rng=_pyrandom):
    """dataloader.py:320-334 (is_random_missing=True): 60 % draws with replacement, frames are
    ZEROED (src = -1)."""
    n = int(T * (60 / 100))
    picks = rng.choices(range(T), k=n)
    src = np.arange(T, dtype=np.int32)
    mask = np.zeros(T, dtype=np.float32)
    for t in picks:
        src[t] = -1
        mask[t] = 1.0
    return src, mask


def apply_sources_add_sos(video: np.ndarray, src: np.ndarray, mask: np.ndarray):
    """dataloader.py:421-434 + add_sos :482-493 -> (inputs [T+1,K,2], mask [T+1])."""
    T, K, C = video.shape
    out = np.empty((T + 1, K, C), dtype=np.float32)
    out[0] = 1.0
    for t in range(T):
        out[t + 1] = 0.0 if src[t] < 0 else video[src[t]]
    return out, np.concatenate([np.zeros(1, np.float32), mask.astype(np.float32)])


# --------------------------------------------------------------------------------------
# augmentation.py
# --------------------------------------------------------------------------------------
def _rotate_pts(x, y, ox, oy, angle):
    """augmentation.py:65-80 -- python doubles, result stored back into a float32 tensor."""
    c, s = math.cos(angle), math.sin(angle)
    qx = ox + c * (x - ox) - s * (y - oy)
    qy = oy + s * (x - ox) + c * (y - oy)
    return qx, qy


def augment_rotate(sign: np.ndarray, angle: float, body_ids, hand_ids):
    """augmentation.py:121-142 -- BODY ids then HAND ids again (hands rotate twice)."""
    out = np.array(sign, dtype=np.float32, copy=True)
    for ids in (body_ids, hand_ids):
        sub = out[:, ids, :].astype(np.float64)
        qx, qy = _rotate_pts(sub[..., 0], sub[..., 1], 0.5, 0.5, angle)
        out[:, ids, 0] = qx.astype(np.float32)
        out[:, ids, 1] = qy.astype(np.float32)
    return out


def perspective_matrix(kind: str, a: float, b: float = 0.0, left: bool = True) -> np.ndarray:
    """augmentation.py:165-187 -- the 3x3 homography cv2.getPerspectiveTransform returns for
    the reference's src/dest quads, solved in float64 (src/dest are float32 like the
    reference's arrays).  kind 'squeeze': a=move_left, b=move_right; 'perspective':
    a=move_ratio, left = outcome of the 0.5 coin (:180)."""
    src = np.array(((0, 1), (1, 1), (0, 0), (1, 0)), dtype=np.float32)
    if kind == "squeeze":
        dst = np.array(((0 + a, 1), (1 - b, 1), (0 + a, 0), (1 - b, 0)), dtype=np.float32)
    elif left:
        dst = np.array(((0 + a, 1 - a), (1, 1), (0 + a, 0 + a), (1, 0)), dtype=np.float32)
    else:
        dst = np.array(((0, 1), (1 - a, 1 - a), (0, 0), (1 - a, 0 + a)), dtype=np.float32)
    A = np.zeros((8, 8), dtype=np.float64)
    rhs = np.zeros(8, dtype=np.float64)
    for i in range(4):
        x, y = float(src[i, 0]), float(src[i, 1])
        u, v = float(dst[i, 0]), float(dst[i, 1])
        A[i] = [x, y, 1, 0, 0, 0, -x * u, -y * u]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * v, -y * v]
        rhs[i] = u
        rhs[i + 4] = v
    h = np.linalg.solve(A, rhs)
    return np.append(h, 1.0).reshape(3, 3)


def augment_shear(sign: np.ndarray, mtx: np.ndarray, body_ids):
    """augmentation.py:194-199 -- cv2.perspectiveTransform (float64 matrix on float32 points,
    rounded to float32) then every COORDINATE equal to the image of (0,0) is reset to 0."""
    out = np.array(sign, dtype=np.float32, copy=True)
    sub = out[:, body_ids, :].astype(np.float64)
    x, y = sub[..., 0], sub[..., 1]
    w = mtx[2, 0] * x + mtx[2, 1] * y + mtx[2, 2]
    w = np.where(np.abs(w) > np.finfo(np.float64).eps, 1.0 / w, 0.0)
    px = ((mtx[0, 0] * x + mtx[0, 1] * y + mtx[0, 2]) * w).astype(np.float32)
    py = ((mtx[1, 0] * x + mtx[1, 1] * y + mtx[1, 2]) * w).astype(np.float32)
    w0 = mtx[2, 2]
    w0 = 1.0 / w0 if abs(w0) > np.finfo(np.float64).eps else 0.0
    zx = np.float32(mtx[0, 2] * w0)
    zy = np.float32(mtx[1, 2] * w0)
    px = np.where(px == zx, np.float32(0), px)
    py = np.where(py == zy, np.float32(0), py)
    out[:, body_ids, 0] = px
    out[:, body_ids, 1] = py
    return out


def augment_arm_joint_rotate(sign: np.ndarray, arm_chains, angles):
    """augmentation.py:206-233 -- arm_chains = [[chest, shoulder, elbow, wrist], ...];
    angles[c][j] = angle in radians or None when the 0.5 coin (:223) failed.  Sequential
    along the chain, per frame, through float32 storage."""
    out = np.array(sign, dtype=np.float32, copy=True)
    for c, chain in enumerate(arm_chains):
        for j, origin in enumerate(chain):
            ang = angles[c][j]
            if ang is None:
                continue
            for tgt in chain[j + 1:]:
                ox = out[:, origin, 0].astype(np.float64)
                oy = out[:, origin, 1].astype(np.float64)
                qx, qy = _rotate_pts(out[:, tgt, 0].astype(np.float64),
                                     out[:, tgt, 1].astype(np.float64), ox, oy, ang)
                out[:, tgt, 0] = qx.astype(np.float32)
                out[:, tgt, 1] = qy.astype(np.float32)
    return out


# --------------------------------------------------------------------------------------
# deterministic parameters / synthetic data shared by golden generation, tests and bench
# --------------------------------------------------------------------------------------
def state_dict_schema(input_size, hidden, layers, ff=2048, max_len=2048):
    """Names and shapes of the reference state_dict, in the order the
    reference module registers them."""
    H = hidden
    s = [("learned_input_positional_encoder", (1, 1, H)),
         ("learned_filled_positional_encoder", (1, 1, H)),
         ("input_embedding.weight", (H, input_size)), ("input_embedding.bias", (H,)),
         ("filled_embedding.weight", (H, input_size)), ("filled_embedding.bias", (H,)),
         ("trig_input_positional_encoder.pos_encoding", (max_len, 1, H)),
         ("trig_filled_positional_encoder.pos_encoding", (max_len, 1, H))]
    for g in ("swiGlu_input_prev", "swiGlu_filled_prev"):
        for f in ("fc1", "fc2", "fc3"):
            s += [(f"{g}.{f}.weight", (H, H)), (f"{g}.{f}.bias", (H,))]
    for l in range(layers):
        pre = f"transformer.encoder.layers.{l}"
        s += [(pre + ".self_attn.in_proj_weight", (3 * H, H)), (pre + ".self_attn.in_proj_bias", (3 * H,)),
              (pre + ".self_attn.out_proj.weight", (H, H)), (pre + ".self_attn.out_proj.bias", (H,)),
              (pre + ".linear1.weight", (ff, H)), (pre + ".linear1.bias", (ff,)),
              (pre + ".linear2.weight", (H, ff)), (pre + ".linear2.bias", (H,)),
              (pre + ".norm1.weight", (H,)), (pre + ".norm1.bias", (H,)),
              (pre + ".norm2.weight", (H,)), (pre + ".norm2.bias", (H,))]
    s += [("transformer.encoder.norm.weight", (H,)), ("transformer.encoder.norm.bias", (H,))]
    for l in range(layers):
        pre = f"transformer.decoder.layers.{l}"
        s += [(pre + ".self_attn.in_proj_weight", (3 * H, H)), (pre + ".self_attn.in_proj_bias", (3 * H,)),
              (pre + ".self_attn.out_proj.weight", (H, H)), (pre + ".self_attn.out_proj.bias", (H,)),
              (pre + ".multihead_attn.in_proj_weight", (3 * H, H)), (pre + ".multihead_attn.in_proj_bias", (3 * H,)),
              (pre + ".multihead_attn.out_proj.weight", (H, H)), (pre + ".multihead_attn.out_proj.bias", (H,)),
              (pre + ".linear1.weight", (ff, H)), (pre + ".linear1.bias", (ff,)),
              (pre + ".linear2.weight", (H, ff)), (pre + ".linear2.bias", (H,)),
              (pre + ".norm1.weight", (H,)), (pre + ".norm1.bias", (H,)),
              (pre + ".norm2.weight", (H,)), (pre + ".norm2.bias", (H,)),
              (pre + ".norm3.weight", (H,)), (pre + ".norm3.bias", (H,))]
    s += [("transformer.decoder.norm.weight", (H,)), ("transformer.decoder.norm.bias", (H,))]
    for f in ("fc1", "fc2", "fc3"):
        s += [(f"swiGlu_decoded.{f}.weight", (H, H)), (f"swiGlu_decoded.{f}.bias", (H,))]
    s += [("fc_final.weight", (input_size, H)), ("fc_final.bias", (input_size,))]
    return s


def deterministic_state_dict(input_size, hidden, layers, ff=2048, max_len=2048, gain=1.0):
    """Closed-form pseudo-random weights (no torch RNG) so golden fixtures do not have to
    store megabytes of parameters: w[i] = a * sin(0.37 i + 1.3 n + 0.11 i^2 mod 7) with a
    fan-in scaled amplitude; norm weights near 1; the PE buffers are the real tables."""
    sd = {}
    for n, (name, shape) in enumerate(state_dict_schema(input_size, hidden, layers, ff, max_len)):
        if name.endswith("pos_encoding"):
            sd[name] = positional_table(max_len, hidden).view(max_len, 1, hidden).clone()
            continue
        numel = int(np.prod(shape))
        i = np.arange(numel, dtype=np.float64)
        base = np.sin(0.37 * i + 1.3 * n + 0.11 * np.mod(i * i, 7.0))
        if name.startswith("learned_"):
            v = 0.5 + 0.5 * base
        elif "norm" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * base
        elif name.endswith("bias"):
            v = 0.05 * base
        else:
            v = gain * base * math.sqrt(3.0 / shape[-1])
        sd[name] = torch.from_numpy(v.astype(np.float32)).view(*shape).clone()
    return sd


AUTSL_CONFIG = {   # dataset_config.json:2-9
    "mean_consecutive_missing": 5.28, "std_consecutive_missing": 4.15, "samples": 491,
    "mean_number_missing_blocks": 4.18, "std_number_missing_blocks": 1.78,
}


def synthetic_batch(B, T, K, seed=42, zero_frac=0.02, smooth=True):
    """SURVEY.md section 8(d) synthetic inputs: (inputs [B,T+1,K,2], gt [B,T,K,2],
    mask [B,T+1]) with AUTSL block statistics, hold-filled exactly as dataloader.py:421-434."""
    rs = np.random.RandomState(seed)
    pr = _pyrandom.Random(seed)
    t = np.arange(T, dtype=np.float64).reshape(1, T, 1, 1) / T
    if smooth:
        f = rs.uniform(0.5, 3.0, size=(B, 1, K, 2))
        ph = rs.uniform(0.0, 1.0, size=(B, 1, K, 2))
        gt = (0.5 + 0.3 * np.sin(2 * np.pi * (f * t + ph))).astype(np.float32)
    else:
        gt = rs.uniform(0.0, 1.0, size=(B, T, K, 2)).astype(np.float32)
    if zero_frac > 0:
        z = rs.uniform(size=(B, T, K)) < zero_frac
        gt[z] = 0.0
    inputs = np.empty((B, T + 1, K, 2), dtype=np.float32)
    mask = np.empty((B, T + 1), dtype=np.float32)
    for b in range(B):
        blocks = missing_blocks_from_config(T, AUTSL_CONFIG, rng=pr, nprng=rs)
        src, m = hold_fill_sources(T, blocks)
        inputs[b], mask[b] = apply_sources_add_sos(gt[b], src, m)
    return torch.from_numpy(inputs), torch.from_numpy(gt), torch.from_numpy(mask)
