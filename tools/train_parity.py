"""Train the default model for a few hundred steps on synthetic sequences with the fused kernels and with the plain two-GEMM /
ln_bwd paths, from the same initial weights and batches, and compare the loss curves and the held-out interpolation loss
(A1_train.py:184-186 protocol: EuclideanLoss on the blended prediction), next to the hold-frame and cubic-spline baselines.
    python tools/train_parity.py [--steps 300]"""
import argparse
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(steps, lr):
    from keypoints_interpolation_transformer_b200 import baselines, euclidean_loss, model, optim, synthetic, train
    dev = torch.device("cuda", 0)
    torch.manual_seed(123)
    m = model.KeypointCompleter(142, 256, 6, 8).to(dev)
    m.train()
    step = train.TrainStep(m, optim.FlatAdam(m, lr=lr), criterion="mse")
    ring = [tuple(t.to(dev) for t in synthetic.synthetic_batch(256, 64, 71, seed=1000 + i, smooth=True)) for i in range(8)]
    curve = []
    for i in range(steps):
        loss = step(*ring[i % len(ring)])
        if i % 25 == 0 or i == steps - 1:
            curve.append(round(loss.item(), 6))
    m.eval()
    held = tuple(t.to(dev) for t in synthetic.synthetic_batch(512, 64, 71, seed=777, smooth=True))
    ev, _ = train.EvalStep(m)(*held)
    inputs, gt, mask = held
    crit = euclidean_loss.MaskedEuclideanLoss()
    hold = crit(inputs[:, 1:].contiguous(), gt, mask[:, 1:].contiguous())
    cubic = crit(baselines.cubic_interpolation(inputs, mask)[:, 1:].contiguous(), gt, mask[:, 1:].contiguous())
    return {"curve": curve, "heldout_masked_euclid": ev.item(), "hold_frame_baseline": hold.item(), "cubic_baseline": cubic.item()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        print(json.dumps(run(args.steps, args.lr)))
        return
    out = {}
    for name, env in (("fused (default)", {}), ("plain (KIT_FUSE_FFN=0)", {"KIT_FUSE_FFN": "0"}),
                      ("all fused (KIT_FUSE_LNBWD=1)", {"KIT_FUSE_LNBWD": "1"})):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, __file__, "--child", "--steps", str(args.steps), "--lr", str(args.lr)], env=e,
                           capture_output=True, text=True, check=True)
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
