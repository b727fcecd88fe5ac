"""Benchmark of the hot path: the KeypointCompleter train step (BASELINE.json metric
"train sequences/sec (T=64, K=71 synthetic)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config train|infer|scaled]

--config train (default, BASELINE configs[1]): one step = forward + loss + backward + Adam over one synthetic batch of
B=256 sequences x T=64 frames x K=71 keypoints per GPU; under torchrun the batches are sharded across ranks (weak scaling)
with a bucketed NCCL gradient all-reduce.  --config scaled (configs[4]): the same step for d_model=512, 8+8 layers, T=512,
B=64 per GPU.  --config infer (configs[3]): the eval step (forward + blend + masked EuclideanLoss, A1_train.py:149-186) on
B=4096 x T=256 per GPU, with the model-vs-cubic-vs-hold-frame table of A1_train.py:184-195 from the same batch.
Rank 0 prints ONE JSON line.  Beside the kit numbers the line carries

  * ``gpu_baseline``: stock PyTorch (nn.Transformer, cuBLAS + SDPA) running the same module from the same weights on the
    same GPU and batch -- fp32, tf32 and bf16 autocast (oracle/torch_reference.py) -- the comparator SURVEY.md 8(d) names;
  * ``cpu_baseline``: the same stock module on the host cores (the reference's CPU path; /root/reference does not travel to
    the GPU box): (ii) batched train step = ``value``; (i) batched forward + loss; (iii) the A1-faithful batch-1 loop with the
    Python-loop get_mask (A1_train.py:117-135).

``--impl reference`` times (ii) alone on a bounded sample and prints it as the reference arm.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train sequences/sec (T=64, K=71 synthetic)"
UNIT = "sequences/s"
KP = 71
CONFIGS = {
    # name: per-GPU batch, frames, hidden, layers, heads, mode, BASELINE.json configs index
    "train": dict(B=256, T=64, H=256, L=6, NH=8, mode="train", idx=1),
    "scaled": dict(B=64, T=512, H=512, L=8, NH=8, mode="train", idx=4),
    "infer": dict(B=4096, T=256, H=256, L=6, NH=8, mode="eval", idx=3),
}
CPU_SAMPLE_B = 32


def flops_per_seq_fwd(S, Hd, layers, kp=KP, ff=2048):
    """BASELINE.md section 3 / SURVEY.md 8(d)."""
    return 2 * S * (2 * 2 * kp * Hd + 9 * Hd * Hd + 2 * kp * Hd) + layers * (8 * S * Hd * Hd + 4 * S * S * Hd + 4 * S * Hd * ff) \
        + layers * (16 * S * Hd * Hd + 8 * S * S * Hd + 4 * S * Hd * ff)


def workload_name(c, world=1):
    what = "A1 train step (fwd+loss+bwd+Adam)" if c["mode"] == "train" else "A1 eval step (fwd + blend + masked EuclideanLoss)"
    return (f"{what}, B={c['B']}/GPU x T={c['T']} x K={KP}, H={c['H']} L={c['L']}+{c['L']} heads={c['NH']} ff=2048, "
            f"random missing blocks (AUTSL statistics), BASELINE configs[{c['idx']}]")


# --------------------------------------------------------------------------------------------
# CPU legs: the stock nn.Transformer module (oracle/torch_reference.py) on all host threads
# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, c, batch=CPU_SAMPLE_B, detail=False):
    import torch
    from oracle import kit_oracle as ko
    from oracle import torch_reference as tr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T, H, L, NH = c["T"], c["H"], c["L"], c["NH"]
    m = tr.StockCompleter(2 * KP, H, L, NH)
    m.load_state_dict(ko.deterministic_state_dict(2 * KP, H, L))
    inputs, gt, mask = ko.synthetic_batch(batch, T, KP, seed=42)
    out = {"cores": cores}
    if c["mode"] == "train":
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=5e-6)
        for _ in range(warmup):
            tr.train_step(m, opt, inputs, gt, mask)
        t0 = time.perf_counter()
        for _ in range(steps):
            tr.train_step(m, opt, inputs, gt, mask)
        dt = time.perf_counter() - t0
        out.update(value=batch * steps / dt, ms_per_step=1e3 * dt / steps,
                   sample=f"(ii) {steps} steps of fwd+MSELoss+bwd+Adam on B={batch} x T={T} x K={KP}, fp32, stock nn.Transformer "
                          f"module on CPU, {cores} threads (oracle/torch_reference.py = model.py:100-170 + A1_train.py:117-135)")

    def fwd_loss():
        with torch.no_grad():
            xm, ym = mask[:, :-1], mask[:, 1:]
            pred = m(inputs[:, :-1], inputs[:, 1:], src_pad_mask=xm, src_mask=tr.repeat_inc_masks(xm, NH),
                     tgt_mask=tr.repeat_inc_masks(ym, NH))
            blend = pred * ym[:, :, None, None] + gt * (1 - ym)[:, :, None, None]
            return ((blend - gt) ** 2).sum(-1).mean()

    if c["mode"] == "eval" or detail:
        m.eval()
        fwd_loss()
        n = max(1, steps)
        t0 = time.perf_counter()
        for _ in range(n):
            fwd_loss()
        dt = time.perf_counter() - t0
        fl = {"value": batch * n / dt, "ms_per_step": 1e3 * dt / n,
              "sample": f"(i) {n} x batched forward + blend + EuclideanLoss, B={batch} x T={T}, no_grad"}
        if c["mode"] == "eval":
            out.update(value=fl["value"], ms_per_step=fl["ms_per_step"],
                       sample=fl["sample"] + f", stock nn.Transformer module on CPU, {cores} threads")
        else:
            out["fwd_loss"] = fl
    if detail and c["mode"] == "train":
        # (iii) A1_train.py:89-135 as written: one sequence per iteration, two Python-loop get_mask calls, unbatched step
        m.train()
        opt1 = torch.optim.Adam(m.parameters(), lr=5e-6)
        n_seq = 4

        def one(b):
            x1, g1, m1 = inputs[b:b + 1], gt[b:b + 1], mask[b:b + 1]
            xm, ym = m1[:, :-1], m1[:, 1:]
            sm = tr.get_mask_loop(xm[0], T).unsqueeze(0).repeat(NH, 1, 1)
            tm = tr.get_mask_loop(ym[0], T).unsqueeze(0).repeat(NH, 1, 1)
            pred = m(x1[:, :-1], x1[:, 1:], src_pad_mask=xm, src_mask=sm, tgt_mask=tm)
            loss = torch.nn.functional.mse_loss(pred, g1)
            opt1.zero_grad()
            loss.backward()
            opt1.step()
            return float(loss.detach())     # A1_train.py:131 reads the loss back every step

        one(0)
        t0 = time.perf_counter()
        for b in range(n_seq):
            one(1 + b)
        dt = time.perf_counter() - t0
        out["a1_faithful"] = {"value": n_seq / dt, "ms_per_step": 1e3 * dt / n_seq,
                              "sample": f"(iii) {n_seq} sequences, batch 1, two Python-loop get_mask calls per step (model.py:193-202)"}
    return out


def reference_arm(args, rank, c):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 2))
    r = cpu_reference_run(steps, warmup, c)
    line = {"metric": METRIC if args.config == "train" else f"{c['mode']} sequences/sec (T={c['T']}, K={KP} synthetic)",
            "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(c),
                       "sample": f"each timed step is a bounded sample of that workload: B={CPU_SAMPLE_B} sequences",
                       "note": "the reference is pure Python/PyTorch and /root/reference does not travel to the GPU box; timed as "
                               "the stock nn.Transformer restatement of its module (oracle/torch_reference.py, pinned to the "
                               "reference's own outputs by tests/test_oracle_golden.py) on all host cores"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# stock PyTorch on the same GPU (the comparator SURVEY.md 8(d) names)
# --------------------------------------------------------------------------------------------
def gpu_baseline_run(kit_model, c, batch, dev, steps=5, warmup=2):
    """The stock nn.Transformer module with the kit model's weights, same batch, same step: {mode: sequences/s}."""
    import torch
    from oracle import torch_reference as tr
    sd = {k: v.detach().clone() for k, v in kit_model.state_dict().items()}
    inputs, gt, mask = batch
    B = inputs.shape[0]
    out = {"unit": UNIT, "what": "oracle/torch_reference.StockCompleter (nn.Transformer: cuBLAS GEMMs + SDPA + ATen elementwise) from the "
                               "same state_dict, same device-resident batch, "
                               + ("fwd + MSELoss + bwd + torch.optim.Adam" if c["mode"] == "train" else "fwd + blend + EuclideanLoss, no_grad"),
           "steps": steps}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for name, tf32, ac in (("fp32", False, None), ("tf32", True, None), ("bf16_autocast", True, torch.bfloat16)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            m = tr.StockCompleter(2 * KP, c["H"], c["L"], c["NH"]).to(dev)
            m.load_state_dict(sd)
            if c["mode"] == "train":
                m.train()
                opt = torch.optim.Adam(m.parameters(), lr=5e-6)

                def step():
                    return tr.train_step(m, opt, inputs, gt, mask, autocast_dtype=ac)
            else:
                m.eval()

                def step():
                    with torch.no_grad():
                        xm, ym = mask[:, :-1], mask[:, 1:]
                        ctx = torch.autocast("cuda", dtype=ac) if ac is not None else torch.autocast("cuda", enabled=False)
                        with ctx:
                            pred = m(inputs[:, :-1], inputs[:, 1:], src_pad_mask=xm, src_mask=tr.repeat_inc_masks(xm, c["NH"]),
                                     tgt_mask=tr.repeat_inc_masks(ym, c["NH"])).float()
                        blend = pred * ym[:, :, None, None] + gt * (1 - ym)[:, :, None, None]
                        return ((blend - gt) ** 2).sum(-1).mean()
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": B / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss)}
            del m
            torch.cuda.empty_cache()
        except Exception as exc:   # noqa: BLE001 -- a baseline that cannot run (memory) must not cost the headline line
            out[name] = {"error": repr(exc)[:200]}
            torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    return out


def dp_grad_check(rank, world, dev, c):
    """Sharded gradients (bucketed all-reduce, 1/world) == the gradient of the global batch computed on one GPU
    (SURVEY.md section 4 (iv)); relative error, identical on every rank."""
    import torch
    import torch.distributed as dist
    from keypoints_interpolation_transformer_b200 import model, optim, parallel, synthetic, train
    B_local = 16
    torch.manual_seed(7)
    m = model.KeypointCompleter(2 * KP, c["H"], c["L"], c["NH"]).to(dev)
    m.train()
    dist.broadcast(m.flat_params, src=0)
    inputs, gt, mask = (t.to(dev) for t in synthetic.synthetic_batch(B_local * world, min(c["T"], 64), KP, seed=99, smooth=True))
    ref_step = train.TrainStep(m, optim.FlatAdam(m, lr=0.0), criterion="mse")
    ref_step.forward_backward(inputs, gt, mask)
    torch.cuda.synchronize()
    ref = m.flat_grads.clone()
    lo, hi = parallel.shard_batch(B_local * world, rank, world)
    reducer = parallel.BucketReducer(m.flat_grads, m.layout.buckets)
    dp_step = train.TrainStep(m, optim.FlatAdam(m, lr=0.0), criterion="mse", reducer=reducer)
    dp_step.forward_backward(inputs[lo:hi].contiguous(), gt[lo:hi].contiguous(), mask[lo:hi].contiguous())
    torch.cuda.synchronize()
    rel = (((m.flat_grads / world) - ref).norm() / ref.norm()).reshape(1)
    dist.all_reduce(rel, op=dist.ReduceOp.MAX)
    return {"rel_err_sharded_vs_global_batch": float(rel), "local_batch": B_local, "global_batch": B_local * world,
            "what": "|allreduce(shard grads)/world - grad(global batch on one GPU)| / |grad|, max over ranks"}


# --------------------------------------------------------------------------------------------
# GPU leg
# --------------------------------------------------------------------------------------------
def in_graph_timeline(run, batches, replays=4):
    """Critical-path attribution of ONE replayed step from CUPTI kernel activity records (torch.profiler; no serialisation or
    cache flush as under ncu, no launch gaps as in the kernel-by-kernel leg): with programmatic dependent launch every kernel of
    the chain starts early and waits for its predecessor, so a kernel's contribution to the step is its END minus the previous
    END; the contributions add up to the span of the replay.  Returns None when CUPTI is not available."""
    try:
        import torch
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(replays):
                run(batches[i % len(batches)])
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start
               and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
        if len(evs) < replays or len(evs) % replays:
            return None
        evs.sort(key=lambda e: e.time_range.start)
        n_per = len(evs) // replays
        one = sorted(evs[(replays - 2) * n_per:(replays - 1) * n_per], key=lambda e: e.time_range.end)
        prev = min(e.time_range.start for e in one)
        t0, agg = prev, {}
        for e in one:
            name = re.sub(r"\(.*", "", e.name).replace("void kit::", "").replace("kit::", "")[:64]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += max(0.0, e.time_range.end - prev)
            prev = max(prev, e.time_range.end)
        span = prev - t0
        return {"span_us": span, "kernels": n_per,
                "by_kernel": {k: {"launches": v[0], "us": round(v[1], 1), "share": round(v[1] / span, 4)}
                              for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                "method": "CUPTI kernel records of one CUDA-graph replay; per kernel: end - previous end (sums to span_us)"}
    except Exception as exc:   # noqa: BLE001 -- a supplementary leg must not cost the bench line
        return {"unavailable": repr(exc)[:200]}


def in_graph_roofline(timeline, flops_per_launch, peak_tf):
    """The dominant kernel's rate INSIDE the replayed step (in_graph_timeline: end - previous end per launch, so neither the launch
    gaps of the kernel-by-kernel leg nor the cold caches of ncu): supplementary to `achieved`, which keeps the CUDA-event figure."""
    if not timeline or "by_kernel" not in timeline or not flops_per_launch:
        return None
    us = sum(v["us"] for k, v in timeline["by_kernel"].items() if k.startswith("ffn_kernel"))
    n = sum(v["launches"] for k, v in timeline["by_kernel"].items() if k.startswith("ffn_kernel"))
    if n == 0 or us <= 0:
        return None
    tf = flops_per_launch / (us / n * 1e-6) / 1e12
    return {"avg_launch_us": us / n, "launches": n, "achieved": tf, "frac": tf / peak_tf if peak_tf else None,
            "share_of_step": us / timeline["span_us"]}


def gpu_run(args, c):
    import torch
    import torch.distributed as dist
    from keypoints_interpolation_transformer_b200 import dataloader, model, optim, parallel, synthetic, train

    # keep stdout clean for the ONE JSON line (NCCL prints its version banner there)
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    B, T, H, L, NH = c["B"], c["T"], c["H"], c["L"], c["NH"]
    is_train = c["mode"] == "train"
    fwd_flops = flops_per_seq_fwd(T, H, L)
    seq_flops = 3 * fwd_flops if is_train else fwd_flops

    dp_check = dp_grad_check(rank, world, dev, c) if (world > 1 and is_train) else None

    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * KP, H, L, NH).to(dev)
    reducer = None
    step = None
    if is_train:
        m.train()
        use_graph = not args.no_graph     # world > 1: a chain of graphs cut at the all-reduce buckets (train.TrainStep)
        opt = optim.FlatAdam(m, lr=5e-6, capturable=use_graph)
        if world > 1:
            m.ensure_flat_grads()
            reducer = parallel.BucketReducer(m.flat_grads, m.layout.buckets)
            dist.broadcast(m.flat_params, src=0)
        # the step is captured once per input slot as a CUDA graph (a chain of graphs under data parallelism) and replayed
        raw_mode = args.config == "train" and args.streams == 1 and not args.batch_layout
        if raw_mode:
            # BASELINE configs[1] from RAW keypoints: the random policy (augmentation + missing blocks, dataloader.py:649-675),
            # the fused pre-pass (normalize_pose, augmentation, hold-fill, SOS, A1 slices) and the train step, all inside the
            # timed step (train.RawTrainStep)
            from keypoints_interpolation_transformer_b200 import preprocess as PP
            pp = PP.Prepass(KP, dev, list(range(KP)), list(range(29, KP)), 5, 6, 2, [[0, 5, 7, 9], [0, 6, 8, 10]])
            pol = PP.DevicePolicy("AUTSL", seed=42 + rank, have_augmentation=True, augmentations_prob=0.5, has_arms=True, device=dev)
            step = train.RawTrainStep(m, pp, pol, opt, criterion="mse", normalize=True, reducer=reducer, use_graph=use_graph)
        else:
            step = train.TrainStep(m, opt, criterion="mse", reducer=reducer, use_graph=use_graph, streams=args.streams)

        def run(batch):
            return step(*batch)
    else:
        m.eval()
        use_graph = False
        raw_mode = False
        ev = train.EvalStep(m)

        def run(batch):
            return ev(*batch)[0]

    # a ring of distinct synthetic batches (pinned host copies + device-resident copies)
    ring = 4 if args.config == "train" else 2
    gen_b = min(B, 512)        # host generation is a Python loop per sequence: larger batches tile 512 distinct sequences
    host, devb = [], []
    for i in range(ring):
        parts = synthetic.synthetic_batch(gen_b, T, KP, seed=42 + 97 * rank + i, smooth=True)
        if gen_b < B:
            parts = tuple(t.repeat((B + gen_b - 1) // gen_b, *([1] * (t.dim() - 1)))[:B].contiguous() for t in parts)
        if raw_mode:      # the raw keypoints [B,T,K,2]: the synthetic ground truth with a plausible shoulder line / eye height
            raw = parts[1].clone()
            raw[:, :, 5, 0] = 0.40 + 0.02 * raw[:, :, 5, 0]
            raw[:, :, 6, 0] = 0.60 + 0.02 * raw[:, :, 6, 0]
            raw[:, :, 2, 1] = 0.30 + 0.02 * raw[:, :, 2, 1]
            parts = (raw.contiguous(),)
        hb = tuple(t.pin_memory() for t in parts)
        host.append(hb)
        devb.append(tuple(t.to(dev, non_blocking=True) for t in hb))
    torch.cuda.synchronize()
    h2d = sum(t.numel() * 4 for t in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    if use_graph:                     # capture every input slot outside the timed region (two eager steps come first)
        for i in range(ring + 2):
            run(devb[i % ring])
    for i in range(args.warmup):
        run(devb[i % ring])
    barrier()
    if is_train and use_graph and not step.use_graph:
        raise RuntimeError("bench: CUDA-graph capture of the train step fell back to kernel-by-kernel launches; "
                           "re-run with --no-graph to measure that mode on purpose")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = run(devb[i % ring])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    eng = m.engine_for(B, T, training=is_train)
    launches = step.last_launches if is_train else eng.fwd_launches + 1
    # ---- end-to-end through the public API with HOST buffers ("e2e"): dataloader.DevicePrefetcher copies batch i+1 from
    # pinned host memory on a side stream while step i computes.  Steady state: the feed is one batch longer than the loop,
    # so the K timed steps contain exactly K host->device batch copies (the first timed batch was copied during warm-up,
    # the batch after the last one is copied during the last timed step) and K device->host reads of the loss.
    n_warm = min(3, args.warmup)
    feed = dataloader.DevicePrefetcher((host[i % ring] for i in range(n_warm + args.steps + 1)), dev)
    for _ in range(n_warm):
        run(next(feed)).item()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    for _ in range(args.steps):
        last = run(next(feed)).item()    # device->host read of the loss every step
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    timeline = in_graph_timeline(run, devb) if (is_train and use_graph and world == 1 and rank == 0) else None

    # ---- per-kernel-class timing: CUDA events around every launch of the engine, kernel by kernel on one stream (no graph, no
    # overlap between consecutive kernels), plus events around the whole eager step so that the shares add up
    eng.set_profiling(True)
    prof_steps = 3
    acc = {}
    eager_ms = 0.0
    if is_train:
        prof_step = step if args.streams == 1 else train.TrainStep(m, opt, criterion="mse", reducer=reducer, use_graph=False)
    for i in range(prof_steps):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        if is_train:
            prof_step._eager(*devb[i % ring])
        else:
            run(devb[i % ring])
        g1.record()
        torch.cuda.synchronize()
        eager_ms += g0.elapsed_time(g1)
        for k, (pms, n, fl) in eng.profile().items():
            a = acc.setdefault(k, [0.0, 0, 0.0])
            a[0] += pms
            a[1] += n
            a[2] += fl
    eng.set_profiling(False)
    eager_ms /= prof_steps
    if rank != 0:
        if world > 1:
            if step is not None and hasattr(step, "release_graphs"):
                step.release_graphs()
            dist.barrier()
            dist.destroy_process_group()
        return
    # tensor-core family = plain GEMMs + grouped weight gradients + the fused feed-forward kernels + attention
    fam = [acc[k] for k in ("gemm_tn", "gemm_wgrad", "ffn") if k in acc]
    gemm_ms, gemm_n, gemm_fl = (sum(a[i] for a in fam) for i in range(3))
    burst_tf = peaks.get("bf16_tflops") or 1590.0
    sustained_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    family_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # the dominant kernel of the step (largest share of the ncu launch list, profiles/): ffn_kernel
    ffn = acc.get("ffn", [0.0, 0, 0.0])
    dom_ms, dom_n, dom_fl = ffn if ffn[1] > 0 else (gemm_ms, gemm_n, gemm_fl)
    achieved_tf = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic = None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get("ffn_kernel_avg_bytes_per_launch") if (ffn[1] > 0 and args.config == "train") else None
    except Exception:
        pass
    step_ms = ms / args.steps
    seqs = B * world * args.steps
    value = seqs / (ms * 1e-3)
    e2e = seqs / (ms_e2e * 1e-3)
    cpu = cpu_reference_run(steps=2, warmup=1, c=c, detail=True) if (world == 1 and not args.no_cpu_baseline) else None
    base_batch = devb[0]
    if raw_mode:           # the stock module takes the dataloader layout: one batch through the stand-alone pre-pass (untimed)
        src, miss, aug = pol.draw(B, T)
        res = pp(devb[0][0], src, miss, normalize=True, aug_dev=aug)
        base_batch = (res["inputs"], res["y"], res["mask"])
    gpu_base = gpu_baseline_run(m, c, base_batch, dev) if (world == 1 and not args.no_gpu_baseline) else None
    # ---- HBM roofline of the fused per-frame passes (pre-pass, loss) at BASELINE configs[3] size, measured live
    frame = None
    if world == 1 and not args.no_framepass and args.config == "train":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import framepass_bench
            frame = [{"kernel": r["kernel"], "bound": "hbm", "workload": f"B={r['B']} x T={r['T']} x K={r['K']}",
                      "achieved": r["GBps"], "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                      "frac": r["frac_of_measured_hbm"], "algorithmic_bytes": r["algorithmic_bytes"], "ms": r["ms"]}
                     for r in framepass_bench.measure(shapes=((4096, 256, 71),), peak=peaks.get("hbm_gbs", 6650.0))]
        except Exception as exc:   # the headline line must still be printed
            frame = {"error": repr(exc)}
    interp = None
    if not is_train and world == 1:
        interp = interpolation_table(m, devb[0], dev)
    breakdown = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] // prof_steps,
                     "share_of_eager_step": (v[0] / prof_steps) / eager_ms if eager_ms > 0 else None,
                     "tflops": (v[2] / (v[0] * 1e-3) / 1e12) if v[0] > 0 else None} for k, v in acc.items()}
    breakdown["_note"] = ("CUDA events around every engine launch in a kernel-by-kernel step on one stream (no graph replay, no "
                          f"programmatic overlap of consecutive kernels): that step takes {eager_ms:.3f} ms against {step_ms:.3f} ms "
                          "replayed, so per-kernel TFLOP/s are conservative and shares are of the EAGER step")
    metric = METRIC if args.config == "train" else \
        (f"train sequences/sec (T={T}, K={KP} synthetic, d_model={H}, {L}+{L} layers)" if is_train
         else f"eval sequences/sec (T={T}, K={KP} synthetic)")
    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(c), "parallelism": f"dp{world}", "global_batch": B * world,
                   "cuda_graph": bool(is_train and step.use_graph), "streams": args.streams,
                   "step_input": ("raw keypoints [B,T,K,2]: device-drawn policy (augmentation p=0.5 + missing blocks) + fused pre-pass "
                                  "(normalize_pose, augmentation, hold-fill, SOS, bf16 operands) run INSIDE every timed step "
                                  "(train.RawTrainStep; dataloader.py:623-686 + A1_train.py:89-135)") if raw_mode
                   else "dataloader layout (inputs [B,T+1,K,2], sota, mask): pre-processing outside the step",
                   "l2": "no flush needed: each step streams GBs of activations/weights (>> 126 MB L2); a ring of "
                         f"{ring} distinct device-resident batches",
                   "model_flops_per_seq": seq_flops, "model_tflops": seq_flops * value / 1e12,
                   "model_frac_of_burst_peak": seq_flops * value / 1e12 / burst_tf / world},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor",
                     "kernel": "ffn_kernel<BWD> (fused feed-forward block, forward and input-gradient pass)" if ffn[1] > 0
                     else "gemm_tcgen05_kernel (TN + wgrad)",
                     "achieved": achieved_tf, "peak": burst_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / burst_tf if burst_tf else None, "traffic": traffic,
                     "frac_of_sustained_peak": achieved_tf / sustained_tf if sustained_tf else None,
                     "algorithmic_flops_per_launch": dom_fl / dom_n if dom_n else None,
                     "avg_launch_us": 1e3 * dom_ms / dom_n if dom_n else None,
                     "launches_per_step": dom_n // prof_steps,
                     "ms_per_step": dom_ms / prof_steps, "share_of_eager_step": (dom_ms / prof_steps) / eager_ms if eager_ms else None,
                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst: the step is a few ms at full clocks; "
                                     "frac_of_sustained_peak beside it)") if peaks else "fallback",
                     "in_graph": in_graph_roofline(timeline, dom_fl / dom_n if dom_n else None, burst_tf) if ffn[1] > 0 else None,
                     "tensor_family": {"what": "plain GEMMs + grouped weight gradients + ffn_kernel (attention listed in breakdown)",
                                       "achieved": family_tf, "frac": family_tf / burst_tf if burst_tf else None,
                                       "launches_per_step": gemm_n // prof_steps, "ms_per_step": gemm_ms / prof_steps,
                                       "share_of_eager_step": (gemm_ms / prof_steps) / eager_ms if eager_ms else None}},
        "breakdown": breakdown,
        "in_graph": timeline,
        "framepass_roofline": frame,
        "gpu_baseline": gpu_base,
        "cpu_baseline": None if cpu is None else {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                                                 "sample": cpu["sample"],
                                                 "fwd_loss": cpu.get("fwd_loss"), "a1_faithful": cpu.get("a1_faithful")},
    }
    if gpu_base is not None:
        best = max((v["value"] for v in gpu_base.values() if isinstance(v, dict) and "value" in v), default=None)
        line["vs_stock_pytorch"] = {"e2e_over_best_stock_mode": (e2e / best) if best else None,
                                    "e2e_over_fp32": (e2e / gpu_base["fp32"]["value"]) if "value" in gpu_base.get("fp32", {}) else None}
    if dp_check is not None:
        line["dp_grad_check"] = dp_check
    if interp is not None:
        line["interpolation"] = interp
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        if step is not None and hasattr(step, "release_graphs"):
            step.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


def interpolation_table(m, batch, dev, cpu_cubic_seqs=4):
    """A1_train.py:184-195 on one batch: masked EuclideanLoss of the model's blend, of the hold-frame input and of the cubic
    spline (device kernel, checked against the pandas / scipy restatement on a few sequences)."""
    import torch
    from keypoints_interpolation_transformer_b200 import baselines, train
    from keypoints_interpolation_transformer_b200 import _lib as K
    from keypoints_interpolation_transformer_b200.euclidean_loss import fused_loss
    inputs, gt, mask = batch
    y_mask = mask[:, 1:].contiguous()
    loss_model, _ = train.EvalStep(m)(inputs, gt, mask)
    hold = inputs[:, 1:].contiguous()
    loss_hold, _ = fused_loss(hold, gt, y_mask, K.LOSS_EUCLID, want_grad=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cub = baselines.cubic_interpolation(inputs, mask)
    e1.record()
    torch.cuda.synchronize()
    loss_cubic, _ = fused_loss(cub[:, 1:].contiguous(), gt, y_mask, K.LOSS_EUCLID, want_grad=False)
    out = {"masked_euclidean_loss": {"model_random_init": float(loss_model), "hold_frame": float(loss_hold), "cubic_spline": float(loss_cubic)},
           "cubic_gpu_ms": e0.elapsed_time(e1), "sequences": int(inputs.shape[0]),
           "note": "weights are random-init (no checkpoint or dataset in this environment): the model column shows the protocol runs, "
                   "not interpolation quality; interpolation-MSE parity with the reference is tests/test_model_gpu.py"}
    try:
        from oracle import kit_oracle as ko
        t0 = time.perf_counter()
        worst = 0.0
        for b in range(cpu_cubic_seqs):
            ref = ko.cubic_interpolation(inputs[b].cpu(), mask[b:b + 1].cpu())
            worst = max(worst, float((torch.as_tensor(ref) - cub[b].cpu()).abs().max()))
        out["cubic_cpu_ms_per_seq"] = 1e3 * (time.perf_counter() - t0) / cpu_cubic_seqs
        out["cubic_gpu_vs_cpu_max_abs"] = worst
    except Exception as exc:   # noqa: BLE001
        out["cubic_cpu_check"] = repr(exc)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="kit", choices=["kit", "reference"])
    ap.add_argument("--config", default="train", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the configuration's)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs under ncu)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-on-GPU leg")
    ap.add_argument("--no-framepass", action="store_true", help="skip the pre-pass / loss HBM roofline leg")
    ap.add_argument("--streams", type=int, default=1, help="concurrent sub-batch chains per step (train.TrainStep streams)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--batch-layout", action="store_true",
                    help="train config: feed pre-processed (inputs, sota, mask) batches instead of raw keypoints (round-1 behaviour)")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.batch:
        c["B"] = args.batch
    if args.steps is None:
        args.steps = 100 if args.config == "train" else 10
    if args.warmup is None:
        args.warmup = 10 if args.config == "train" else 3
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":          # rank 0 alone runs it; other ranks exit 0 without work
        reference_arm(args, int(os.environ.get("RANK", "0")), c)
        return
    gpu_run(args, c)


if __name__ == "__main__":
    main()
