"""Exactly one kernel-by-kernel train step (BASELINE configs[1] shape) between cudaProfilerStart / Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/one_step.py
Prints the number of launches the step claims (TrainStep.last_launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, synthetic, train  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    m = model.KeypointCompleter(142, 256, 6, 8).to(dev)
    m.train()
    step = train.TrainStep(m, optim.FlatAdam(m, lr=5e-6), criterion="mse")
    batches = [tuple(t.to(dev) for t in synthetic.synthetic_batch(256, 64, 71, seed=42 + i, smooth=True)) for i in range(2)]
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
