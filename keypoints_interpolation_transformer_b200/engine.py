"""Thin Python handle on the C-ABI step engine (include/kit.h: kit_layout_*, kit_engine_*)."""
import ctypes as C

import torch

from . import _lib as K


class ModelLayout:
    """Parameter-arena layout of one model configuration (names = the reference's state_dict keys,
    SURVEY.md section 8b)."""

    def __init__(self, input_size, hidden, layers, heads, ff=2048, max_len=2048, variant=0):
        self.cfg = K.KitModelConfig(input_size, hidden, layers, heads, ff, max_len, variant, 0)
        lib = K.lib()
        n = lib.kit_layout_num_entries(C.byref(self.cfg))
        if n < 0:
            raise K.KitError(lib.kit_last_error().decode())
        self.entries = {}
        name = C.create_string_buffer(256)
        for i in range(n):
            off, numel, rows, cols, isbuf = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
            K.check(lib.kit_layout_entry(C.byref(self.cfg), i, name, 256, C.byref(off), C.byref(numel),
                                         C.byref(rows), C.byref(cols), C.byref(isbuf)))
            self.entries[name.value.decode()] = (off.value, numel.value, rows.value, cols.value, bool(isbuf.value))
        self.trainable = lib.kit_layout_trainable_floats(C.byref(self.cfg))
        self.total = lib.kit_layout_total_floats(C.byref(self.cfg))
        self.buckets = []
        for b in range(lib.kit_layout_num_buckets(C.byref(self.cfg))):
            lo, hi = C.c_int64(), C.c_int64()
            K.check(lib.kit_layout_bucket(C.byref(self.cfg), b, C.byref(lo), C.byref(hi)))
            self.buckets.append((lo.value, hi.value))

    def shape_of(self, name):
        off, numel, rows, cols, _ = self.entries[name]
        H = self.cfg.hidden
        if name.startswith("learned_"):
            return (1, 1, H)
        if name.endswith("pos_encoding"):
            return (rows, 1, cols)
        if rows == 1:
            return (cols,)
        return (rows, cols)


def make_mask(frame_mask=None, flags=0, bias=None, bias_stride_b=0, bias_stride_h=0):
    """Builds a KitAttnMask; keeps references to the tensors so they outlive the async kernels."""
    m = K.KitAttnMask()
    keep = []
    if frame_mask is not None:
        assert frame_mask.dtype == torch.float32 and frame_mask.dim() == 2 and frame_mask.stride(1) == 1
        m.frame_mask = frame_mask.data_ptr()
        m.frame_mask_stride = frame_mask.stride(0)
        m.flags = flags
        keep.append(frame_mask)
    else:
        m.flags = flags & K.MASK_TRIANGLE
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        m.bias = bias.data_ptr()
        m.bias_stride_b = bias_stride_b
        m.bias_stride_h = bias_stride_h
        keep.append(bias)
    m._keep = keep
    return m


class StepEngine:
    """One (batch, seq_len) instance of the forward/backward engine bound to a parameter arena."""

    def __init__(self, layout: ModelLayout, batch, seq_len, params, grads, training=True):
        assert params.is_cuda and params.dtype == torch.float32 and params.numel() >= layout.total
        # The library keeps its one-time state (shared-memory attributes, SM count) for the device the process uses and launches
        # on the CURRENT device's stream (one process per GPU, as under torchrun): a model on another device is an error, not a
        # silent launch on the wrong stream.
        if params.device.index is not None and params.device.index != torch.cuda.current_device():
            raise K.KitError(f"the model lives on cuda:{params.device.index} but the current device is cuda:{torch.cuda.current_device()}: "
                             "call torch.cuda.set_device() first (one process per GPU)")
        self.layout = layout
        self.batch, self.seq_len, self.training = batch, seq_len, training
        self._h = C.c_void_p()
        lib = K.lib()
        K.check(lib.kit_engine_create(C.byref(layout.cfg), batch, seq_len, 1 if training else 0, C.byref(self._h)))
        nbytes = lib.kit_engine_workspace_bytes(self._h)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=params.device)
        self.params, self.grads = params, grads
        K.check(lib.kit_engine_bind(self._h, K.ptr(params), K.ptr(grads), K.ptr(self.workspace), nbytes))
        self._masks = None
        self.fwd_launches = self.bwd_launches = 0

    def __del__(self):
        try:
            if self._h:
                K.lib().kit_engine_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def refresh_weights(self):
        K.check(K.lib().kit_engine_refresh_weights(self._h, K.stream_ptr()))

    def forward(self, x_enc, x_enc_batch_stride, x_dec, x_dec_batch_stride, enc_mask, dec_mask, pred,
                zero_masked_enc=False):
        """x_enc / x_dec: fp32 CUDA tensors (any view whose frames are contiguous rows of input_size)."""
        self._masks = (enc_mask, dec_mask, x_enc, x_dec)      # keep alive until backward
        K.check(K.lib().kit_engine_forward(
            self._h, K.ptr(x_enc), x_enc_batch_stride, K.ptr(x_dec), x_dec_batch_stride,
            C.byref(enc_mask) if enc_mask is not None else None,
            C.byref(dec_mask) if dec_mask is not None else None,
            1 if zero_masked_enc else 0, K.ptr(pred), K.stream_ptr()))
        self.fwd_launches = K.lib().kit_engine_last_launches(self._h)

    def operand_ptrs(self):
        """(x_enc, x_dec) device pointers of the engine's own bf16 operand buffers [B*T, k2p] and k2p (kit_engine_operands): the
        pre-pass writes into them and ``forward(None, 0, None, 0, ...)`` skips the packing pass."""
        xe, xd, k2p = C.c_void_p(), C.c_void_p(), C.c_int32()
        K.check(K.lib().kit_engine_operands(self._h, C.byref(xe), C.byref(xd), C.byref(k2p)))
        return xe, xd, k2p.value

    def backward(self, dpred, bucket_callback=None):
        cb = K.BUCKET_CALLBACK(lambda b, _u: bucket_callback(b)) if bucket_callback is not None else None
        K.check(K.lib().kit_engine_backward(self._h, K.ptr(dpred), cb, None, K.stream_ptr()))
        self.bwd_launches = K.lib().kit_engine_last_launches(self._h)

    PROF_CATEGORIES = ("gemm_tn", "gemm_wgrad", "attn_fwd", "attn_bwd", "ffn")

    def set_profiling(self, on):
        K.check(K.lib().kit_engine_set_profiling(self._h, 1 if on else 0))

    def profile(self):
        """{category: (ms, launches, flops)} of the last forward (+ backward); synchronises."""
        out = {}
        for i, name in enumerate(self.PROF_CATEGORIES):
            ms, n, fl = C.c_float(), C.c_int64(), C.c_double()
            K.check(K.lib().kit_engine_profile_read(self._h, i, C.byref(ms), C.byref(n), C.byref(fl)))
            out[name] = (ms.value, n.value, fl.value)
        return out

    def debug_read(self, name):
        lib = K.lib()
        out = torch.empty(self._buf_elems(name), device=self.params.device)
        K.check(lib.kit_engine_debug_read(self._h, name.encode(), K.ptr(out), out.numel(), K.stream_ptr()))
        return out

    def _buf_elems(self, name):
        M, c = self.batch * self.seq_len, self.layout.cfg
        H, FF = c.hidden, c.ff
        k2p = (c.input_size + 7) // 8 * 8
        base = name.split(".")[-1]
        if base in ("qkv", "gqkv"):
            return M * 3 * H
        if base in ("kv_all", "gkv"):
            return M * 2 * H * c.layers
        if base == "kv_bias":
            return 2 * H * c.layers
        if base in ("si12", "sf12", "sd12", "g2h"):
            return M * 2 * H
        if base in ("z", "hh", "gff"):
            return M * FF
        if base in ("xe", "xd", "dp"):
            return M * k2p
        if base.startswith("st"):
            return 2 * M
        if base.startswith("lse"):
            return self.batch * c.heads * self.seq_len
        return M * H
