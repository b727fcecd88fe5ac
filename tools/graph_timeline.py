"""In-graph timeline of the benchmarked train step: the step (train.RawTrainStep, BASELINE configs[1] shape) is captured as one CUDA
graph and replayed under torch.profiler (CUPTI kernel activity records, no replay / serialisation as under ncu), then every
kernel's start and duration inside ONE replay is listed: busy time per kernel class, the gaps between consecutive kernels and
how much consecutive kernels overlap (programmatic dependent launch).  usage: python tools/graph_timeline.py [out.json]"""
import collections
import json
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, synthetic, train  # noqa: E402
from keypoints_interpolation_transformer_b200 import preprocess as PP  # noqa: E402

KP = 71


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void kit::", "").replace("kit::", "")
    return name[:70]


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * KP, 256, 6, 8).to(dev)
    m.train()
    opt = optim.FlatAdam(m, lr=5e-6, capturable=True)
    batches = []
    for i in range(2):
        parts = synthetic.synthetic_batch(256, 64, KP, seed=42 + i, smooth=True)
        raw = parts[1].clone()
        raw[:, :, 5, 0] = 0.40 + 0.02 * raw[:, :, 5, 0]
        raw[:, :, 6, 0] = 0.60 + 0.02 * raw[:, :, 6, 0]
        raw[:, :, 2, 1] = 0.30 + 0.02 * raw[:, :, 2, 1]
        batches.append((raw.contiguous().to(dev),))
    pp = PP.Prepass(KP, dev, list(range(KP)), list(range(29, KP)), 5, 6, 2, [[0, 5, 7, 9], [0, 6, 8, 10]])
    pol = PP.DevicePolicy("AUTSL", seed=42, have_augmentation=True, augmentations_prob=0.5, has_arms=True, device=dev)
    step = train.RawTrainStep(m, pp, pol, opt, criterion="mse", normalize=True, use_graph=True)
    for i in range(8):
        step(*batches[i % 2])
    torch.cuda.synchronize()
    assert step.use_graph and len(step._graphs) == 2
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(4):
            step(*batches[i % 2])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start
           and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    n_per = len(evs) // 4
    one = evs[2 * n_per:3 * n_per]            # the third replay
    t0 = one[0].time_range.start
    span = one[-1].time_range.end - t0
    busy = collections.defaultdict(lambda: [0, 0.0])
    gaps = collections.defaultdict(lambda: [0, 0.0])    # idle time BEFORE a kernel of this class (previous kernel ended, this one not started)
    overlap = 0.0
    covered_until = one[0].time_range.start
    idle = 0.0
    rows = []
    for i, e in enumerate(one):
        s, t = e.time_range.start, e.time_range.end
        busy[short(e.name)][0] += 1
        busy[short(e.name)][1] += t - s
        if s > covered_until:
            idle += s - covered_until
            gaps[short(e.name)][0] += 1
            gaps[short(e.name)][1] += s - covered_until
        else:
            overlap += min(covered_until, t) - s
        covered_until = max(covered_until, t)
        rows.append((round(s - t0, 2), round(t - s, 2), short(e.name)))
    total_busy = sum(v[1] for v in busy.values())
    print(f"{len(one)} kernels in one replay, span {span:.1f} us, sum of kernel durations {total_busy:.1f} us, "
          f"idle (no kernel running) {idle:.1f} us, overlapped (two kernels running) {overlap:.1f} us")
    print("| kernel | launches | busy us | share of span | avg us | idle before it us |\n|---|---|---|---|---|---|")
    for name, (c, t) in sorted(busy.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {c} | {t:.0f} | {t / span * 100:.1f}% | {t / c:.1f} | {gaps[name][1]:.0f} |")
    if len(sys.argv) > 1:
        json.dump({"span_us": span, "busy_us": total_busy, "idle_us": idle, "overlap_us": overlap, "kernels": rows}, open(sys.argv[1], "w"))


if __name__ == "__main__":
    main()
