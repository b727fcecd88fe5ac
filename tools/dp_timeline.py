"""Timeline of the data-parallel train step on rank 0 (torchrun --nproc-per-node N tools/dp_timeline.py): CUPTI kernel records of a
few replayed steps -- where the NCCL all-reduce kernels run relative to backward, how long the tail after the last compute kernel
of backward is, how much idle time the graph chain leaves."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, parallel, synthetic, train  # noqa: E402
from keypoints_interpolation_transformer_b200 import preprocess as PP  # noqa: E402

KP = 71


def main():
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * KP, 256, 6, 8).to(dev)
    m.train()
    m.ensure_flat_grads()
    reducer = parallel.BucketReducer(m.flat_grads, m.layout.buckets) if world > 1 else None
    if world > 1:
        dist.broadcast(m.flat_params, src=0)
    opt = optim.FlatAdam(m, lr=5e-6, capturable=True)
    batches = []
    for i in range(2):
        parts = synthetic.synthetic_batch(256, 64, KP, seed=42 + 97 * rank + i, smooth=True)
        raw = parts[1].clone()
        raw[:, :, 5, 0] = 0.40 + 0.02 * raw[:, :, 5, 0]
        raw[:, :, 6, 0] = 0.60 + 0.02 * raw[:, :, 6, 0]
        raw[:, :, 2, 1] = 0.30 + 0.02 * raw[:, :, 2, 1]
        batches.append((raw.contiguous().to(dev),))
    pp = PP.Prepass(KP, dev, list(range(KP)), list(range(29, KP)), 5, 6, 2, [[0, 5, 7, 9], [0, 6, 8, 10]])
    pol = PP.DevicePolicy("AUTSL", seed=42 + rank, have_augmentation=True, augmentations_prob=0.5, has_arms=True, device=dev)
    step = train.RawTrainStep(m, pp, pol, opt, criterion="mse", normalize=True, reducer=reducer, use_graph=True)
    for i in range(10):
        step(*batches[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(6):
            step(*batches[i % 2])
        torch.cuda.synchronize()
    if rank == 0:
        try:
            prof.export_chrome_trace("gpurun_out/dp_trace_rank0.json")
            import json
            tr = json.load(open("gpurun_out/dp_trace_rank0.json"))
            seen = set()
            for ev in tr.get("traceEvents", []):
                if "nccl" in str(ev.get("name", "")).lower() and ev.get("cat") == "kernel":
                    key = (tuple(ev["args"].get("grid", [])), tuple(ev["args"].get("block", [])), round(ev.get("dur", 0), -1))
                    if key not in seen:
                        seen.add(key)
                        print("nccl kernel grid", ev["args"].get("grid"), "block", ev["args"].get("block"), "dur", ev.get("dur"))
            os.remove("gpurun_out/dp_trace_rank0.json")
        except Exception as exc:   # noqa: BLE001
            print("trace export failed:", exc)
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start
               and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
        evs.sort(key=lambda e: e.time_range.start)
        # steps are delimited by the policy kernel
        starts = [i for i, e in enumerate(evs) if "policy_kernel" in e.name]
        for si in range(2, min(5, len(starts) - 1)):
            one = evs[starts[si]:starts[si + 1]]
            t0 = one[0].time_range.start
            nxt = evs[starts[si + 1]].time_range.start
            print(f"--- step {si}: {len(one)} kernels, policy to next policy {nxt - t0:.1f} us")
            covered, idle = t0, 0.0
            for e in one:
                s, t = e.time_range.start, e.time_range.end
                name = e.name.replace("void kit::", "").replace("kit::", "")[:60]
                is_nccl = "nccl" in e.name.lower()
                if not is_nccl:
                    if s > covered + 2.0:
                        print(f"    idle {s - covered:7.1f} us before {name} (at {s - t0:8.1f})")
                        idle += s - covered
                    covered = max(covered, t)
                if is_nccl or "adam" in e.name.lower() or "wgrad_group" in e.name or "fill" in e.name.lower():
                    print(f"  {s - t0:8.1f} .. {t - t0:8.1f}  ({t - s:7.1f} us)  {name}")
            print(f"    idle on the compute chain {idle:.1f} us; last kernel ends at {covered - t0:.1f} us")
    if world > 1:
        step.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
