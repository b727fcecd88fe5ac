"""Exactly one kernel-by-kernel train step of the BENCHMARKED form (BASELINE configs[1] shape, from raw keypoints: device-drawn
policy + fused pre-pass + forward + loss + backward + Adam = train.RawTrainStep, what bench.py times) between
cudaProfilerStart / Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/one_step.py
Prints the number of launches the step claims (TrainStep.last_launches).  `one_step.py batch` runs the dataloader-layout step
(train.TrainStep: pre-processing outside the step) instead."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, synthetic, train  # noqa: E402
from keypoints_interpolation_transformer_b200 import preprocess as PP  # noqa: E402

KP = 71


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * KP, 256, 6, 8).to(dev)
    m.train()
    raw_mode = not (len(sys.argv) > 1 and sys.argv[1] == "batch")
    opt = optim.FlatAdam(m, lr=5e-6)
    batches = []
    for i in range(2):
        parts = synthetic.synthetic_batch(256, 64, KP, seed=42 + i, smooth=True)
        if raw_mode:   # as bench.py: the synthetic ground truth with a plausible shoulder line / eye height
            raw = parts[1].clone()
            raw[:, :, 5, 0] = 0.40 + 0.02 * raw[:, :, 5, 0]
            raw[:, :, 6, 0] = 0.60 + 0.02 * raw[:, :, 6, 0]
            raw[:, :, 2, 1] = 0.30 + 0.02 * raw[:, :, 2, 1]
            parts = (raw.contiguous(),)
        batches.append(tuple(t.to(dev) for t in parts))
    if raw_mode:
        pp = PP.Prepass(KP, dev, list(range(KP)), list(range(29, KP)), 5, 6, 2, [[0, 5, 7, 9], [0, 6, 8, 10]])
        pol = PP.DevicePolicy("AUTSL", seed=42, have_augmentation=True, augmentations_prob=0.5, has_arms=True, device=dev)
        step = train.RawTrainStep(m, pp, pol, opt, criterion="mse", normalize=True)
    else:
        step = train.TrainStep(m, opt, criterion="mse")
    for i in range(4):
        step(*batches[i % 2])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(*batches[0])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("claimed launches per step:", step.last_launches)


if __name__ == "__main__":
    main()
