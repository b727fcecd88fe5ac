"""BASELINE configs[4]: the scaled model (d_model = 512, 8 + 8 layers, T = 512) train step on one GPU, CUDA events, with the
engine's per-category breakdown.  Prints one JSON line.
    python tools/scaled_bench.py [--batch 32] [--seq 512] [--hidden 512] [--layers 8] [--steps 5]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, synthetic, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    K, H, L, NH, S, ff = 71, args.hidden, args.layers, args.heads, args.seq, 2048
    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * K, H, L, NH).to(dev)
    m.train()
    step = train.TrainStep(m, optim.FlatAdam(m, lr=5e-6), criterion="mse")
    inputs, gt, mask = (t.to(dev) for t in synthetic.synthetic_batch(args.batch, S, K, seed=42, smooth=True))
    for _ in range(3):
        loss = step(inputs, gt, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(inputs, gt, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    eng = m.engine_for(args.batch, S, training=True)
    eng.set_profiling(True)
    step(inputs, gt, mask)
    torch.cuda.synchronize()
    prof = {k: {"ms": v[0], "launches": v[1], "tflops": (v[2] / (v[0] * 1e-3) / 1e12) if v[0] > 0 else None}
            for k, v in eng.profile().items()}
    eng.set_profiling(False)
    fwd = 2 * S * (2 * 2 * K * H + 9 * H * H + 2 * K * H) + L * (8 * S * H * H + 4 * S * S * H + 4 * S * H * ff) \
        + L * (16 * S * H * H + 8 * S * S * H + 4 * S * H * ff)
    print(json.dumps({"workload": f"train step (fwd+loss+bwd+Adam), B={args.batch} x T={S} x K={K}, H={H} L={L}+{L} heads={NH}, "
                                  "BASELINE configs[4] shape on one GPU",
                      "sequences_per_s": args.batch / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss),
                      "model_tflops": 3 * fwd * args.batch / (ms * 1e-3) / 1e12, "launches": step.last_launches,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "breakdown": prof}))


if __name__ == "__main__":
    main()
