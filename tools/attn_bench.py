"""Times kit_attention_fwd (and bwd) on the long-sequence shapes: configs[3] (T = 256, d = 32) and configs[4] (T = 512, d = 64).
KIT_ATTN_TC=0 selects the mma.sync streaming kernel, default the tcgen05 kernel.  usage: attn_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402


def bench(B, NH, S, d, flags, n=10):
    H = NH * d
    dev = "cuda"
    qkv = torch.randn(B * S, 3 * H, device=dev).to(torch.bfloat16)
    fm = (torch.rand(B, S, device=dev) < 0.4).float()
    mask = K.KitAttnMask()
    mask.frame_mask = fm.data_ptr()
    mask.frame_mask_stride = S
    mask.flags = flags
    out = torch.empty(B * S, H, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, NH, S, device=dev)
    lib, sp = K.lib(), K.stream_ptr()
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]

    def f():
        K.check(lib.kit_attention_fwd(K.ptr(q), 3 * H, K.ptr(k), 3 * H, K.ptr(v), 3 * H, K.ptr(out), H, K.ptr(lse), B, NH, S, S, d,
                                      C.byref(mask), sp))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    fl = 4.0 * B * NH * S * S * d
    print(f"attention fwd B={B} NH={NH} S={S} d={d}: {us:9.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  (KIT_ATTN_TC={os.environ.get('KIT_ATTN_TC', '1')})")


def bench_bwd(B, NH, S, d, flags, n=10):
    H = NH * d
    dev = "cuda"
    qkv = torch.randn(B * S, 3 * H, device=dev).to(torch.bfloat16)
    dout = torch.randn(B * S, H, device=dev).to(torch.bfloat16)
    fm = (torch.rand(B, S, device=dev) < 0.4).float()
    mask = K.KitAttnMask()
    mask.frame_mask = fm.data_ptr()
    mask.frame_mask_stride = S
    mask.flags = flags
    out = torch.empty(B * S, H, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, NH, S, device=dev)
    dq = torch.empty(B * S, 3 * H, dtype=torch.bfloat16, device=dev)
    acc = torch.empty(B * S * H + B * NH * S, device=dev) if S > 64 else None
    lib, sp = K.lib(), K.stream_ptr()
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    K.check(lib.kit_attention_fwd(K.ptr(q), 3 * H, K.ptr(k), 3 * H, K.ptr(v), 3 * H, K.ptr(out), H, K.ptr(lse), B, NH, S, S, d,
                                  C.byref(mask), sp))

    def f():
        K.check(lib.kit_attention_bwd(K.ptr(q), 3 * H, K.ptr(k), 3 * H, K.ptr(v), 3 * H, K.ptr(out), H, K.ptr(dout), H, K.ptr(lse),
                                      K.ptr(dq[:, :H]), 3 * H, K.ptr(dq[:, H:2 * H]), 3 * H, K.ptr(dq[:, 2 * H:]), 3 * H, K.ptr(acc), B, NH,
                                      S, S, d, C.byref(mask), sp))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    fl = 10.0 * B * NH * S * S * d
    print(f"attention bwd B={B} NH={NH} S={S} d={d}: {us:9.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
          f"(KIT_ATTN_T64={os.environ.get('KIT_ATTN_T64', '1')} KIT_ATTN_TC={os.environ.get('KIT_ATTN_TC', '1')})")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "t64":     # the train-step shape (BASELINE configs[1]): B = 256, T = 64, 8 heads of 32
        for env in ("1", "0"):
            os.environ["KIT_ATTN_T64"] = env
            print(f"KIT_ATTN_T64={env}")
            bench(256, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, n=50)
            bench_bwd(256, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, n=50)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "long":    # long-sequence backward: tcgen05 (attention_tcb.cu) against the mma.sync streaming kernel
        for env in ("1", "0"):
            os.environ["KIT_ATTN_TC"] = env
            bench_bwd(64, 8, 512, 64, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
            bench_bwd(256, 8, 256, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
            bench_bwd(16, 8, 2048, 64, K.MASK_REPEAT_INC)
        sys.exit(0)
    bench(64, 8, 512, 64, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
    bench(1024, 8, 256, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
    bench(16, 8, 2048, 64, K.MASK_REPEAT_INC)
