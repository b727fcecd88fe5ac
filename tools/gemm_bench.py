"""Micro-benchmark of the tcgen05 GEMM at the shapes the train step uses (CUDA events, L2-cold
rotation of operands).  python tools/gemm_bench.py [--iters N] [--only NAME]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402

M = 16384
SHAPES = [  # name, mode, M, N, K, bias, addend, act
    ("ff1_gelu", 0, M, 2048, 256, True, False, K.ACT_GELU),
    ("ff2_resid", 0, M, 256, 2048, True, True, K.ACT_NONE),
    ("ff2_dgrad_gelubwd", 0, M, 2048, 256, False, False, K.ACT_GELU_BWD),
    ("ff1_dgrad_resid", 0, M, 256, 2048, False, True, K.ACT_NONE),
    ("qkv", 0, M, 768, 256, True, False, K.ACT_NONE),
    ("outproj_resid", 0, M, 256, 256, True, True, K.ACT_NONE),
    ("plain_256", 0, M, 256, 256, False, False, K.ACT_NONE),
    ("swiglu12", 0, M, 512, 256, True, False, K.ACT_NONE),
    ("wgrad_ff1", 1, 2048, 256, M, False, False, K.ACT_NONE),
    ("wgrad_ff2", 1, 256, 2048, M, False, False, K.ACT_NONE),
    ("wgrad_256", 1, 256, 256, M, False, False, K.ACT_NONE),
    ("wgrad_qkv", 1, 768, 256, M, False, False, K.ACT_NONE),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default=None)
    ap.add_argument("--warm", action="store_true", help="no L2 flush between iterations (operands L2-resident, as inside the step)")
    args = ap.parse_args()
    dev = "cuda"
    lib = K.lib()
    sp = K.stream_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, mode, m, n, k, has_bias, has_add, act in SHAPES:
        if args.only and args.only != name:
            continue
        if mode == 0:
            a = torch.randn(m, k, device=dev).to(torch.bfloat16)
            b = (torch.randn(n, k, device=dev) / k ** 0.5).to(torch.bfloat16)
            c = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            out_kind = K.OUT_BF16
            lda, ldb = k, k
        else:
            a = torch.randn(k, m, device=dev).to(torch.bfloat16)
            b = (torch.randn(k, n, device=dev) / k ** 0.5).to(torch.bfloat16)
            c = torch.zeros(m, n, device=dev)
            out_kind = K.OUT_F32_ATOMIC
            lda, ldb = m, n
        bias = torch.randn(n, device=dev) if has_bias else None
        add = torch.randn(m, n, device=dev).to(torch.bfloat16) if has_add else None
        aux = torch.randn(m, n, device=dev).to(torch.bfloat16) if act != K.ACT_NONE else None

        def run():
            K.check(lib.kit_gemm_bf16(mode, K.ptr(a), lda, K.ptr(b), ldb, K.ptr(c), n, m, n, k, K.ptr(bias), K.ptr(add), n,
                                      out_kind, act, K.ptr(aux), n, 0 if mode == 1 else 1, sp))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            if not args.warm:
                flush.zero_()                   # evict L2 between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        med = times[len(times) // 2]
        fl = 2.0 * m * n * k
        if os.environ.get("KIT_GEMM_TRACE"):
            import ctypes
            buf = (ctypes.c_longlong * 16)()
            lib.kit_gemm_trace_read.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
            if lib.kit_gemm_trace_read(buf) == 0:
                t0 = buf[0]
                names = ["entry", "setup_done", "pdl_done", "tma_first", "full_first", "mma_item0_done", "epi_tmem_full",
                         "epi_ld_done", "epi_store_issued", "epi_loop_end", "epi_store_drained", "exit", "bar_init_done", "tmem_alloc_done",
                         "cta_synced"]
                print("   trace(cycles): " + " ".join(f"{n}={buf[i] - t0}" for i, n in enumerate(names) if buf[i] > 0))
        print(f"{name:20s} M={m:6d} N={n:5d} K={k:6d}  {med * 1e3:8.1f} us  {fl / (med * 1e-3) / 1e12:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
