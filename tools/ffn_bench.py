"""Times the fused feed-forward kernel (kit_ffn_fwd) against the two-GEMM path it replaces.  usage: ffn_bench.py [M] [FF]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    FF = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    H, dev = 256, "cuda"
    bf = torch.bfloat16
    nbuf = 6   # rotate buffers so that the activations are not simply L2-resident
    xs = [torch.randn(M, H, device=dev).to(bf) for _ in range(nbuf)]
    w1 = (torch.randn(FF, H, device=dev) / math.sqrt(H)).to(bf)
    w2 = (torch.randn(H, FF, device=dev) / math.sqrt(FF)).to(bf)
    b1, b2 = torch.randn(FF, device=dev), torch.randn(H, device=dev)
    gamma, beta = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    zs = [torch.empty(M, FF, dtype=bf, device=dev) for _ in range(nbuf)]
    hs = [torch.empty(M, FF, dtype=bf, device=dev) for _ in range(nbuf)]
    s, y = torch.empty(M, H, dtype=bf, device=dev), torch.empty(M, H, dtype=bf, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    lib, sp = K.lib(), K.stream_ptr()
    i = [0]

    def fused(store):
        def f():
            j = i[0] % nbuf
            i[0] += 1
            K.check(lib.kit_ffn_fwd(K.ptr(xs[j]), K.ptr(w1), K.ptr(w2), K.ptr(b1), K.ptr(b2), K.ptr(gamma), K.ptr(beta),
                                    K.ptr(zs[j]), K.ptr(hs[j]), K.ptr(s), K.ptr(y), K.ptr(mean), K.ptr(rstd), M, H, FF, store, sp))
        return f

    def two_gemm():
        j = i[0] % nbuf
        i[0] += 1
        K.check(lib.kit_gemm_bf16(0, K.ptr(xs[j]), H, K.ptr(w1), H, K.ptr(hs[j]), FF, M, FF, H, K.ptr(b1), None, 0, K.OUT_BF16,
                                  K.ACT_GELU, K.ptr(zs[j]), FF, 1, sp))
        K.check(lib.kit_gemm_bf16(0, K.ptr(hs[j]), FF, K.ptr(w2), FF, K.ptr(s), H, M, H, FF, K.ptr(b2), K.ptr(xs[j]), H, K.OUT_BF16,
                                  K.ACT_NONE, None, 0, 1, sp))

    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
    gs = [torch.randn(M, H, device=dev).to(bf) for _ in range(nbuf)]
    dzs = [torch.empty(M, FF, dtype=bf, device=dev) for _ in range(nbuf)]
    dx = torch.empty(M, H, dtype=bf, device=dev)
    for zt in zs:
        zt.copy_(torch.randn(M, FF, device=dev))

    def bwd():
        j = i[0] % nbuf
        i[0] += 1
        K.check(lib.kit_ffn_bwd(K.ptr(gs[j]), K.ptr(w2t), K.ptr(w1t), K.ptr(zs[j]), K.ptr(dzs[j]), K.ptr(dx), M, H, FF, sp))

    fl = 4.0 * M * H * FF
    if os.environ.get("FFN_DBG_SWEEP"):
        for dbg, what in ((0, "all on"), (1, "no GELU"), (2, "no GEMM2 MMAs"), (4, "no GEMM1 MMAs"), (6, "no MMAs"), (8, "no h sts"),
                          (16, "no fence.proxy.async"), (25, "no GELU / sts / fence"), (31, "everything off")):
            os.environ["KIT_FFN_DBG"] = str(dbg)
            print(f"dbg={dbg:2d} {what:28s}: {timeit(fused(0)):8.1f} us")
        return
    for name, fn in (("fused train (z, h stored)", fused(1)), ("fused inference", fused(0)), ("fused backward (z read, dz stored)", bwd),
                     ("two GEMMs (gelu + residual, no LN)", two_gemm)):
        us = timeit(fn)
        print(f"{name:40s} M={M} FF={FF}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")


def trace(M, FF, store):
    import ctypes as C
    os.environ["KIT_FFN_TRACE"] = "1"
    H, dev, bf = 256, "cuda", torch.bfloat16
    x = torch.randn(M, H, device=dev).to(bf)
    w1 = (torch.randn(FF, H, device=dev) / math.sqrt(H)).to(bf)
    w2 = (torch.randn(H, FF, device=dev) / math.sqrt(FF)).to(bf)
    b1, b2 = torch.randn(FF, device=dev), torch.randn(H, device=dev)
    gamma, beta = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    z, hh = torch.empty(M, FF, dtype=bf, device=dev), torch.empty(M, FF, dtype=bf, device=dev)
    s, y = torch.empty(M, H, dtype=bf, device=dev), torch.empty(M, H, dtype=bf, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    lib = K.lib()
    for _ in range(3):
        K.check(lib.kit_ffn_fwd(K.ptr(x), K.ptr(w1), K.ptr(w2), K.ptr(b1), K.ptr(b2), K.ptr(gamma), K.ptr(beta), K.ptr(z),
                                K.ptr(hh), K.ptr(s), K.ptr(y), K.ptr(mean), K.ptr(rstd), M, H, FF, store, K.stream_ptr()))
    torch.cuda.synchronize()
    buf = (C.c_longlong * 256)()
    lib.kit_ffn_trace_read.argtypes = [C.c_void_p]
    assert lib.kit_ffn_trace_read(buf) == 0
    t0 = buf[0]
    NC = FF // 128
    print(f"trace store={store} (cycles from kernel entry; CTA 0, item 0)")
    print("  c: mma1_start  g2_hfull | epi: acc1_full math_done h_empty_ok arrived")
    for c in range(NC):
        print(f"  {c:2d}: {buf[1 + c] - t0:8d} {buf[17 + c] - t0:8d} | {buf[33 + c] - t0:8d} {buf[49 + c] - t0:8d} "
              f"{buf[65 + c] - t0:8d} {buf[81 + c] - t0:8d} w17 {buf[101 + c] - t0:8d} | peer(+{buf[128] - t0}): "
              f"{buf[128 + 33 + c] - t0:8d} {buf[128 + 49 + c] - t0:8d} {buf[128 + 81 + c] - t0:8d} w17 {buf[128 + 101 + c] - t0:8d}")
    print(f"  acc2_full {buf[97] - t0}  epi_done {buf[98] - t0}  exit {buf[99] - t0}")
    del os.environ["KIT_FFN_TRACE"]


if __name__ == "__main__":
    if os.environ.get("FFN_TRACE"):
        trace(16384, 2048, 0)
        trace(16384, 2048, 1)
        sys.exit(0)
    main()
