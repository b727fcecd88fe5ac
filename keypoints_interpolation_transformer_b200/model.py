"""Drop-in for the reference's ``model.py`` (KeypointCompleter, SwiGLU/PositionalEncoding semantics,
get_mask) whose forward and backward run on the sm_100a step engine.

Same constructor (model.py:61), same ``forward`` signature (model.py:100), same ``get_mask``
(model.py:172), same ``state_dict`` keys and shapes -- a reference ``.pth``
loads with ``load_state_dict`` and vice versa.  All parameters are views into ONE fp32 arena
(``flat_params``); their gradients live in ``flat_grads`` (same layout) so the optimiser and the
data-parallel all-reduce work on a few large contiguous ranges.

There is no CPU path: ``forward`` requires the module to be on a CUDA device.
"""
import collections
import math
import os

import torch
import torch.nn as nn

from . import _lib as K
from .engine import ModelLayout, StepEngine, make_mask


class _Node(nn.Module):
    """Name-space container so that state_dict keys match the reference module tree."""


def positional_table(max_len, dim):
    """model.py:34-46 (same operations, so the buffer is bit-identical to the reference's)."""
    pe = torch.zeros(max_len, dim)
    pos = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
    div = torch.exp(torch.arange(0, dim, 2).float() * (-math.log(10000.0)) / dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0).transpose(0, 1)


class _SlotLease:
    """Marks one engine slot busy between a grad-enabled forward and its backward (or the release of its autograd graph):
    a second forward of the same shape before that backward gets its own engine instead of overwriting the first one's
    saved activations (gradient accumulation, ``(l1 + l2).backward()``)."""

    def __init__(self, module, key, slot):
        self.busy, self.key, self.slot = module._busy_slots, key, slot
        self.busy.setdefault(key, set()).add(slot)

    def release(self):
        s = self.busy.get(self.key)
        if s is not None:
            s.discard(self.slot)

    def __del__(self):
        self.release()


class _CompleterFn(torch.autograd.Function):
    """Autograd bridge of the compatibility path: one node for the whole model."""

    @staticmethod
    def forward(ctx, module, engine, lease, x_enc, xes, x_dec, xds, enc_mask, dec_mask, zero_masked, out_shape, *params):
        pred = torch.empty(out_shape, dtype=torch.float32, device=x_enc.device)
        engine.forward(x_enc, xes, x_dec, xds, enc_mask, dec_mask, pred, zero_masked)
        engine._generation = getattr(engine, "_generation", 0) + 1
        ctx.module, ctx.engine, ctx.lease, ctx.generation = module, engine, lease, engine._generation
        return pred

    @staticmethod
    def backward(ctx, dpred):
        module, engine = ctx.module, ctx.engine
        if engine._generation != ctx.generation:
            raise K.KitError("KeypointCompleter.backward: the engine that holds this forward's activations has run another "
                             "forward since (its slot lease was released early); re-run the forward")
        dpred = dpred.contiguous().float()
        # The engine accumulates into the model's gradient arena, and after attach_flat_grads() every p.grad is a VIEW of that
        # arena: the arena's content is set aside, this backward's gradients are taken out as a fresh tensor (autograd owns what
        # we return and adds it to p.grad itself), and the arena is restored -- nothing is wiped, nothing is counted twice.
        keep = module.flat_grads.clone()
        module.flat_grads.zero_()
        engine.backward(dpred)
        snap = module.flat_grads.clone()
        module.flat_grads.copy_(keep)
        ctx.lease.release()
        grads = tuple(snap[o:o + n].view(s) for (o, n, s) in module._param_slices)
        return (None,) * 11 + grads


class KeypointCompleter(nn.Module):
    FF = 2048          # nn.Transformer default dim_feedforward (model.py:84-90)
    MAX_LEN = 512 * 4  # model.py:74-75
    VARIANT = K.MODEL_COMPLETER
    USES_TGT_PAD = False   # model.py:143: tgt_pad_mask never reaches nn.Transformer

    def __init__(self, input_size, hidden_dim, num_layers, num_heads):
        super().__init__()
        self.input_size, self.hidden_dim, self.num_layers, self.num_heads = input_size, hidden_dim, num_layers, num_heads
        self.layout = ModelLayout(input_size, hidden_dim, num_layers, num_heads, self.FF, self.MAX_LEN, self.VARIANT)
        flat = torch.zeros(self.layout.total)
        self._param_names, self._param_slices, self._buffer_names = [], [], []
        object.__setattr__(self, "flat_params", flat)
        object.__setattr__(self, "flat_grads", None)
        self._engines = collections.OrderedDict()   # (batch, seq_len, training, slot) -> StepEngine, least recently used first
        self._busy_slots = {}                       # (batch, seq_len) -> slots whose forward awaits its backward
        self._dirty = True
        # register the reference's tree of names; tensors are views of the arena
        order = sorted(self.layout.entries.items(), key=lambda kv: kv[1][0])
        for name, (off, numel, rows, cols, is_buffer) in order:
            shape = self.layout.shape_of(name)
            parent = self
            parts = name.split(".")
            for p in parts[:-1]:
                if p not in parent._modules:
                    parent.add_module(p, _Node())
                parent = parent._modules[p]
            view = flat[off:off + numel].view(shape)
            if is_buffer:
                parent.register_buffer(parts[-1], view)
                self._buffer_names.append((name, off, numel, shape))
            else:
                parent.register_parameter(parts[-1], nn.Parameter(view))
        # slices in nn.Module traversal order (what self.parameters() yields)
        for name, _ in self.named_parameters():
            off, numel = self.layout.entries[name][:2]
            self._param_names.append(name)
            self._param_slices.append((off, numel, self.layout.shape_of(name)))
        self.reset_parameters()
        self.register_load_state_dict_post_hook(lambda m, k: m.mark_dirty())

    # ------------------------------------------------------------------ parameters
    def reset_parameters(self):
        """Initialisation of the reference: nn.Linear defaults everywhere, xavier_uniform_ on every
        >1-D parameter inside nn.Transformer (torch/nn/modules/transformer.py:313-317), zero MHA
        biases, unit LayerNorms, torch.rand learned PEs (model.py:77-78)."""
        sd = dict(self.named_parameters())
        with torch.no_grad():
            for name, p in sd.items():
                in_tr = name.startswith("transformer.")
                if name.startswith("learned_"):
                    p.copy_(torch.rand(p.shape))
                elif "norm" in name:
                    p.fill_(1.0 if name.endswith("weight") else 0.0)
                elif p.dim() > 1:
                    if in_tr:
                        nn.init.xavier_uniform_(p)
                    else:
                        nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                elif name.endswith("in_proj_bias") or name.endswith("out_proj.bias"):
                    p.zero_()
                else:  # Linear bias: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                    w = sd[name[:-4] + "weight"]
                    bound = 1.0 / math.sqrt(w.shape[1])
                    p.uniform_(-bound, bound)
            for name, off, numel, shape in self._buffer_names:
                self.flat_params[off:off + numel].view(shape).copy_(positional_table(self.MAX_LEN, self.hidden_dim))
        self.mark_dirty()

    def mark_dirty(self):
        """Tell the engine the fp32 parameters changed (bf16 GEMM operands must be re-derived)."""
        self._dirty = True

    def _rebind(self, flat):
        object.__setattr__(self, "flat_params", flat)
        object.__setattr__(self, "flat_grads", None)
        mods = dict(self.named_modules())
        for name, (off, numel, shape) in zip(self._param_names, self._param_slices):
            parent, _, leaf = name.rpartition(".")
            p = mods[parent]._parameters[leaf]
            p.data = flat[off:off + numel].view(shape)
            p.grad = None
        for name, off, numel, shape in self._buffer_names:
            parent, _, leaf = name.rpartition(".")
            mods[parent]._buffers[leaf] = flat[off:off + numel].view(shape)
        self._engines = collections.OrderedDict()
        self._busy_slots = {}
        self._dirty = True

    def _apply(self, fn, recurse=True):
        new_flat = fn(self.flat_params)
        if new_flat.dtype != torch.float32:
            raise K.KitError("KeypointCompleter keeps fp32 master parameters; bf16 is used inside the kernels")
        self._rebind(new_flat.contiguous())
        return self

    def train(self, mode=True):
        if not mode:
            self._dirty = True
        return super().train(mode)

    def ensure_flat_grads(self):
        if self.flat_grads is None or self.flat_grads.device != self.flat_params.device:
            object.__setattr__(self, "flat_grads", torch.zeros(self.layout.trainable, device=self.flat_params.device))
            self._engines = collections.OrderedDict()
        return self.flat_grads

    def attach_flat_grads(self):
        """Make every ``p.grad`` a view of ``flat_grads`` (what the fused train step and FlatAdam use)."""
        g = self.ensure_flat_grads()
        for p, (off, numel, shape) in zip(self.parameters(), self._param_slices):
            p.grad = g[off:off + numel].view(shape)
        return g

    # ------------------------------------------------------------------ engines
    ENGINE_CACHE = int(os.environ.get("KIT_ENGINE_CACHE", "8"))

    def engine_for(self, batch, seq_len, training, slot=0, pin=False):
        """One engine (activation workspace + kernel plans) per (batch, seq_len, training, slot).  ``slot``: independent
        engines of the same shape (concurrent streams of ``train.TrainStep(streams=...)``, outstanding autograd forwards).
        The cache is bounded (``KIT_ENGINE_CACHE`` engines, least recently used evicted first): the reference feeds
        variable-length videos at batch 1 (A1_train.py:244), and every new length would otherwise keep a workspace forever.
        ``pin``: never evict (engines a captured CUDA graph points into)."""
        if not self.flat_params.is_cuda:
            raise K.KitError("KeypointCompleter.forward needs a CUDA device: this build has no CPU path")
        key = (batch, seq_len, bool(training), slot)
        eng = self._engines.get(key)
        if eng is None:
            grads = self.ensure_flat_grads() if training else None
            with torch.cuda.device(self.flat_params.device):
                eng = StepEngine(self.layout, batch, seq_len, self.flat_params, grads, training=training)
            eng._fresh = False
            eng._pinned = False
            self._engines[key] = eng
            self._evict()
        else:
            self._engines.move_to_end(key)
        if pin:
            eng._pinned = True
        if self.training or self._dirty or not eng._fresh:
            eng.refresh_weights()
            eng._fresh = True
            if not self.training:
                self._dirty = False
                for other in self._engines.values():
                    if other is not eng:
                        other._fresh = False
        return eng

    def _evict(self):
        if len(self._engines) <= self.ENGINE_CACHE:
            return
        newest = next(reversed(self._engines))
        for key in list(self._engines):
            if len(self._engines) <= self.ENGINE_CACHE:
                break
            eng = self._engines[key]
            busy = key[3] in self._busy_slots.get((key[0], key[1]), ())
            if key == newest or eng._pinned or busy:
                continue
            torch.cuda.current_stream(self.flat_params.device).synchronize()   # its kernels may still be running
            del self._engines[key]

    # ------------------------------------------------------------------ forward (model.py:100-170)
    def forward(self, inputs, filled=None, src_pad_mask=None, tgt_pad_mask=None, src_mask=None, tgt_mask=None,
                frame_masks=None, zero_masked=False):
        """Reference semantics (model.py:100-170).  ``inputs``/``filled``: [T,K,2] (one sequence) or
        [B,T,K,2].  ``src_pad_mask`` [B,T] float is ADDED to the encoder logits; ``src_mask`` /
        ``tgt_mask`` are additive [T,T] (or the reference's batched [T, B*heads, T]) float masks;
        ``tgt_pad_mask`` is accepted and ignored (model.py:143).

        Superset: ``frame_masks=(x_mask, y_mask)`` ([B,T] or [T] 0/1 floats) synthesises the A1 masks
        in-kernel -- encoder "repeat-inc"(x_mask) + x_mask as key padding, decoder "repeat-inc"(y_mask)
        (A1_train.py:117-124) -- without materialising any [T,T] tensor."""
        if filled is None:
            raise TypeError("filled is required (model.py:106 calls filled.flatten)")
        unbatched = inputs.dim() == 3
        dev = self.flat_params.device
        x = inputs.to(dev).float()
        xf = filled.to(dev).float()
        if unbatched:
            x, xf = x.unsqueeze(0), xf.unsqueeze(0)
        B, T = x.shape[0], x.shape[1]
        K2 = x.shape[2] * x.shape[3]
        if K2 != self.input_size:
            raise K.KitError(f"inputs carry {K2} values per frame, model was built with input_size={self.input_size}")
        x = x.reshape(B, T, K2).contiguous()
        xf = xf.reshape(B, T, K2).contiguous()
        H, NH = self.hidden_dim, self.num_heads

        def as_bias(m):
            if m is None:
                return None, 0, 0
            m = m.to(dev)
            if m.dtype == torch.bool:
                m = torch.zeros(m.shape, device=dev).masked_fill(m, float("-inf"))
            m = m.float()
            if m.dim() == 2:
                return m.contiguous(), 0, 0
            if m.dim() == 3:     # reference batched layout [T, B*NH, T] (model.py:109-110 permutes it)
                m = m.permute(1, 0, 2).contiguous()
                if m.shape[0] == B * NH:
                    return m, NH * T * T, T * T
                if m.shape[0] == B:
                    return m, T * T, 0
            raise K.KitError(f"unsupported attention-mask shape {tuple(m.shape)}")

        def as_frame(m):
            m = m.to(dev).float()
            if m.dim() == 1:
                m = m.unsqueeze(0)
            if m.shape != (B, T):
                raise K.KitError(f"frame mask shape {tuple(m.shape)} != ({B}, {T})")
            return m.contiguous()

        if frame_masks is not None:
            xm, ym = as_frame(frame_masks[0]), as_frame(frame_masks[1])
            enc_mask = make_mask(xm, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
            dec_mask = make_mask(ym, K.MASK_REPEAT_INC | (K.MASK_KEYPAD_ADD if self.USES_TGT_PAD else 0))
        else:
            sb, sbb, sbh = as_bias(src_mask)
            tb, tbb, tbh = as_bias(tgt_mask)

            def as_pad(pad):
                if pad is None:
                    return None
                pad = pad.to(dev)
                if pad.dtype == torch.bool:
                    pad = torch.zeros(pad.shape, device=dev).masked_fill(pad, float("-inf"))
                return as_frame(pad)

            pad = as_pad(src_pad_mask)
            tpad = as_pad(tgt_pad_mask) if self.USES_TGT_PAD else None
            enc_mask = make_mask(pad, K.MASK_KEYPAD_ADD if pad is not None else 0, sb, sbb, sbh)
            dec_mask = make_mask(tpad, K.MASK_KEYPAD_ADD if tpad is not None else 0, tb, tbb, tbh)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        out_shape = (B, T, K2 // 2, 2)
        if need_grad:
            busy = self._busy_slots.get((B, T), ())
            slot = next(i for i in range(len(busy) + 1) if i not in busy)      # the lowest slot with no forward awaiting its backward
            eng = self.engine_for(B, T, training=True, slot=slot)
            lease = _SlotLease(self, (B, T), slot)
            pred = _CompleterFn.apply(self, eng, lease, x, T * K2, xf, T * K2, enc_mask, dec_mask, bool(zero_masked), out_shape,
                                      *self.parameters())
        else:
            eng = self.engine_for(B, T, training=False)
            pred = torch.empty(out_shape, dtype=torch.float32, device=dev)
            eng.forward(x, T * K2, xf, T * K2, enc_mask, dec_mask, pred, bool(zero_masked))
        return pred[0] if unbatched else pred

    # ------------------------------------------------------------------ model.py:172-209
    def get_mask(self, mask, size, matrixType="triangle") -> torch.Tensor:
        """Vectorised ``get_mask``: same values as the reference's Python double loop, bit-exact
        ({0, 1, -inf}); runs as one kernel when ``mask`` is on the GPU."""
        if matrixType not in K.MATRIX_TYPES:
            raise K.KitError("Choose a correct matrixType - model.py")
        mtype = K.MATRIX_TYPES[matrixType]
        if mask is not None and torch.is_tensor(mask) and mask.is_cuda and matrixType in ("repeat", "repeat-inc"):
            fm = mask.reshape(-1).float().contiguous()
            out = torch.empty(size, size, device=mask.device)
            K.check(K.lib().kit_get_mask(K.ptr(fm), size, mtype, K.ptr(out), K.stream_ptr()))
            return out
        i = torch.arange(size).view(size, 1)
        j = torch.arange(size).view(1, size)
        if matrixType == "triangle":
            return torch.zeros(size, size).masked_fill(j > i, float("-inf"))
        if matrixType == "all":
            return torch.zeros(size, size)
        fm = mask.reshape(1, size).float().cpu()
        rep = fm.repeat(size, 1)
        if matrixType == "repeat":
            return rep
        out = torch.where(rep == 1, torch.tensor(float("-inf")), rep)
        return out.masked_fill(j <= i, 0.0)


class KeypointCompleterCycle(KeypointCompleter):
    """model.py:212-321 -- the second-stage model of A2_train_cycle.py.  Same parameters and kernels as
    ``KeypointCompleter`` with three differences: the trig tables have 512 rows (:226-227); the token-norm output
    enters the position sum twice (:279-284: ``PositionalEncoding`` already returns ``norm + pe``, then ``norm`` is
    added again); and ``tgt_pad_mask`` IS handed to nn.Transformer (:294), i.e. added to the decoder self-attention
    logits.  A2_train_cycle.py:111-115 calls it with "all" masks and all-ones pad masks (a constant shift of every
    logit), and feeds it the first model's prediction as ``filled``; only this model's parameters are stepped
    (A2_train_cycle.py:241), so no gradient flows back into ``inputs`` / ``filled`` here."""
    MAX_LEN = 512
    VARIANT = K.MODEL_CYCLE
    USES_TGT_PAD = True
