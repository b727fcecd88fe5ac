"""HBM roofline of the fused per-frame passes (pre-pass, loss) at BASELINE config 4 size
(B=4096, T=256, K=71) and config 2 size (B=256, T=64).  CUDA events, inputs >> L2 at config 4."""
import json
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import euclidean_loss, missing  # noqa: E402
from keypoints_interpolation_transformer_b200 import preprocess as PP  # noqa: E402


def timeit(fn, iters=10, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def measure(shapes=((4096, 256, 71), (256, 64, 71)), peak=None, dev="cuda"):
    """[{kernel, B, T, K, ms, algorithmic_bytes, GBps, frac_of_measured_hbm}] for the pre-pass and the two loss passes."""
    if peak is None:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for (B, T, K) in shapes:
        raw = torch.rand(B, T, K, 2, device=dev)
        rs = np.random.RandomState(0)
        pr = random.Random(0)
        src = np.empty((B, T), np.int32)
        msk = np.empty((B, T), np.float32)
        for b in range(min(B, 64)):
            src[b], msk[b] = missing.draw_sources(T, False, "AUTSL", pr, rs, missing.DATASET_CONFIG)
        for b in range(64, B):
            src[b], msk[b] = src[b % 64], msk[b % 64]
        src_d, msk_d = torch.from_numpy(src).to(dev), torch.from_numpy(msk).to(dev)
        body = list(range(K))
        pp = PP.Prepass(K, dev, body, list(range(33, K)), 5, 6, 2, [[0, 5, 7, 9], [0, 6, 8, 10]])
        augs = [PP.aug_rotate(0.1) if (b % 2) else PP.aug_none() for b in range(B)]
        # keep the aug table on the device across iterations: time the kernel, not the host->device upload
        import ctypes as C
        from keypoints_interpolation_transformer_b200 import _lib as KL
        arr = (KL.KitSeqAug * B)(*augs)
        aug_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        cfg = KL.KitPrepassConfig()
        cfg.B, cfg.T, cfg.K, cfg.normalize = B, T, K, 1
        cfg.left_shoulder, cfg.right_shoulder, cfg.right_eye = 5, 6, 2
        cfg.n_body, cfg.n_hand = pp.body.numel(), pp.hand.numel()
        for c in range(2):
            for j in range(4):
                cfg.arm_chain[c * 4 + j] = pp.arm_chains[c][j]
        cfg.zero_masked_enc, cfg.k2p = 0, pp.k2p
        y = torch.empty_like(raw)
        inputs = torch.empty(B, T + 1, K, 2, device=dev)
        mask = torch.empty(B, T + 1, device=dev)
        xe = torch.empty(B * T, pp.k2p, dtype=torch.bfloat16, device=dev)
        xd = torch.empty(B * T, pp.k2p, dtype=torch.bfloat16, device=dev)

        def run_pre():
            KL.check(KL.lib().kit_prepass(C.byref(cfg), KL.ptr(raw), KL.ptr(src_d), KL.ptr(msk_d), KL.ptr(aug_dev), KL.ptr(pp.body),
                                          KL.ptr(pp.hand), KL.ptr(y), KL.ptr(inputs), KL.ptr(mask), KL.ptr(xe), KL.ptr(xd),
                                          KL.stream_ptr()))
        ms = timeit(run_pre, flush=flush)
        by = B * (T * K * 8 + T * K * 8 + (T + 1) * K * 8 + (T + 1) * 4 + 2 * T * pp.k2p * 2 + T * 8)
        out.append({"kernel": "prepass_kernel (stand-alone: y + fp32 inputs + mask + both bf16 operands)", "B": B, "T": T, "K": K, "ms": ms,
                    "algorithmic_bytes": by, "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak})

        # the in-step form (train.RawTrainStep): the fp32 `inputs` tensor nobody reads is not written; fraction on SURVEY.md
        # 8(d)'s algorithmic bytes: read (T+1) K 8 + (T+1) 4, write 2 T K 8  (= 437 820 B per sequence at T = 256, K = 71)
        def run_pre_step():
            KL.check(KL.lib().kit_prepass(C.byref(cfg), KL.ptr(raw), KL.ptr(src_d), KL.ptr(msk_d), KL.ptr(aug_dev), KL.ptr(pp.body),
                                          KL.ptr(pp.hand), KL.ptr(y), None, KL.ptr(mask), KL.ptr(xe), KL.ptr(xd), KL.stream_ptr()))
        ms = timeit(run_pre_step, flush=flush)
        by = B * ((T + 1) * K * 8 + (T + 1) * 4 + 2 * T * K * 8)
        traffic = B * (T * K * 8 + T * K * 8 + (T + 1) * 4 + 2 * T * pp.k2p * 2 + T * 8)
        out.append({"kernel": "prepass_kernel (in-step: y + mask + bf16 operands)", "B": B, "T": T, "K": K, "ms": ms,
                    "algorithmic_bytes": by, "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak,
                    "bytes_actually_moved": traffic, "bytes_definition": "SURVEY.md 8(d): read (T+1)K8 + (T+1)4, write 2TK8 per sequence"})
        pred = torch.rand(B, T, K, 2, device=dev)

        def run_loss():
            euclidean_loss.fused_loss(pred, y, None, 0, want_grad=True)
        ms = timeit(run_loss, flush=flush)
        by = B * (3 * T * K * 8)
        out.append({"kernel": "loss_kernel (fwd+grad)", "B": B, "T": T, "K": K, "ms": ms, "algorithmic_bytes": by,
                    "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak})
        fm = (torch.rand(B, T, device=dev) < 0.3).float()

        def run_loss_eval():
            euclidean_loss.fused_loss(pred, y, fm, 0, want_grad=False)
        ms = timeit(run_loss_eval, flush=flush)
        by = B * (2 * T * K * 8 + T * 4)
        out.append({"kernel": "loss_kernel (masked eval, fwd only)", "B": B, "T": T, "K": K, "ms": ms, "algorithmic_bytes": by,
                    "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak})
    return out


def main():
    for o in measure():
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
