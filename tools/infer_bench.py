"""BASELINE configs[3]: inference (A1_train.py:149-186 eval step) on long sequences, B=4096 x T=256 x K=71, through the public
EvalStep (forward + blend + masked EuclideanLoss), CUDA events.  Prints one JSON line.
    python tools/infer_bench.py [--batch 4096] [--seq 256] [--steps 5]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, synthetic, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seq", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    K, H, L, NH = 71, 256, 6, 8
    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * K, H, L, NH).to(dev)
    m.eval()
    chunk = min(args.batch, 512)
    parts = [synthetic.synthetic_batch(chunk, args.seq, K, seed=42 + i, smooth=True) for i in range((args.batch + chunk - 1) // chunk)]
    inputs, gt, mask = (torch.cat([p[j] for p in parts])[:args.batch].to(dev) for j in range(3))
    ev = train.EvalStep(m)
    for _ in range(2):
        loss, _ = ev(inputs, gt, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = ev(inputs, gt, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    S, ff = args.seq, 2048
    fwd_flops = 2 * S * (2 * 2 * K * H + 9 * H * H + 2 * K * H) + L * (8 * S * H * H + 4 * S * S * H + 4 * S * H * ff) \
        + L * (16 * S * H * H + 8 * S * S * H + 4 * S * H * ff)
    print(json.dumps({"workload": f"eval step (forward + blend + masked EuclideanLoss), B={args.batch} x T={args.seq} x K={K}, "
                                  f"H={H} L={L}+{L}, BASELINE configs[3]",
                      "sequences_per_s": args.batch / (ms * 1e-3), "ms_per_step": ms, "masked_loss": float(loss),
                      "model_tflops": fwd_flops * args.batch / (ms * 1e-3) / 1e12,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
