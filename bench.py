"""Benchmark of the hot path: the KeypointCompleter train step (BASELINE.json metric
"train sequences/sec (T=64, K=71 synthetic)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = forward + masked/unmasked loss + backward + Adam over one synthetic batch of
B=256 sequences x T=64 frames x K=71 keypoints per GPU (BASELINE.json configs[1]); under torchrun
the batches are sharded across ranks (weak scaling) with a bucketed NCCL gradient all-reduce.
Rank 0 prints ONE JSON line.  ``--impl reference`` times the CPU restatement of the reference
(oracle/kit_oracle.py -- the reference itself is Python and /root/reference does not exist on the
GPU box) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train sequences/sec (T=64, K=71 synthetic)"
UNIT = "sequences/s"
B_PER_GPU, T, KP = 256, 64, 71
H, L, NH = 256, 6, 8
CPU_SAMPLE_B = 32


def flops_per_seq_train(S=T, Hd=H, layers=L, kp=KP, ff=2048):
    """BASELINE.md section 3."""
    fwd = 2 * S * (2 * 2 * kp * Hd + 9 * Hd * Hd + 2 * kp * Hd) + layers * (8 * S * Hd * Hd + 4 * S * S * Hd + 4 * S * Hd * ff) \
        + layers * (16 * S * Hd * Hd + 8 * S * S * Hd + 4 * S * Hd * ff)
    return 3 * fwd


# --------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference step, all host threads
# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, batch=CPU_SAMPLE_B):
    import torch
    from oracle import kit_oracle as ko
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = ko.deterministic_state_dict(2 * KP, H, L)
    params = {k: v.clone().requires_grad_(not k.endswith("pos_encoding")) for k, v in sd.items()}
    trainable = [p for p in params.values() if p.requires_grad]
    opt = torch.optim.Adam(trainable, lr=5e-6)
    inputs, gt, mask = ko.synthetic_batch(batch, T, KP, seed=42)

    def step():
        opt.zero_grad()
        loss, _ = ko.train_forward_loss(params, inputs, gt, mask, NH, criterion="mse")
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "ms_per_step": 1e3 * dt / steps, "cores": cores,
            "sample": f"{steps} steps of fwd+loss+bwd+Adam on B={batch} x T={T} x K={KP}, fp32, torch CPU, "
                      f"{cores} threads (oracle/kit_oracle.py restatement of A1_train.py:117-135)"}


def reference_arm(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 2))
    r = cpu_reference_run(steps, warmup)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"A1 train step (fwd+loss+bwd+Adam), B={B_PER_GPU}/GPU x T={T} x K={KP}, H={H} L={L}+{L} "
                                   f"heads={NH} ff=2048, random missing blocks (AUTSL stats), BASELINE configs[1]",
                       "sample": f"each timed step is a bounded sample of that workload: B={CPU_SAMPLE_B} sequences",
                       "note": "the reference is pure Python/PyTorch and /root/reference does not travel to the GPU box; "
                               "timed as its CPU restatement (oracle port) on all host cores"},
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
# GPU leg
# --------------------------------------------------------------------------------------------
def gpu_run(args):
    import torch
    import torch.distributed as dist
    from keypoints_interpolation_transformer_b200 import model, optim, parallel, synthetic, train

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

    torch.manual_seed(42)
    m = model.KeypointCompleter(2 * KP, H, L, NH).to(dev)
    m.train()
    use_graph = not args.no_graph     # world > 1: a chain of graphs cut at the all-reduce buckets (train.TrainStep)
    opt = optim.FlatAdam(m, lr=5e-6, capturable=use_graph)
    reducer = None
    if world > 1:
        m.ensure_flat_grads()
        reducer = parallel.BucketReducer(m.flat_grads, m.layout.buckets)
        opt.grad_scale = 1.0 / world
        dist.broadcast(m.flat_params, src=0)
    # the step is captured once per input slot as a CUDA graph (a chain of graphs under data parallelism) and replayed
    step = train.TrainStep(m, opt, criterion="mse", reducer=reducer, use_graph=use_graph, streams=args.streams)

    # a ring of distinct synthetic batches (pinned host copies + device-resident copies)
    ring = 4
    host, devb = [], []
    for i in range(ring):
        inputs, gt, mask = synthetic.synthetic_batch(B_PER_GPU, T, KP, seed=42 + 97 * rank + i, smooth=True)
        hb = tuple(t.pin_memory() for t in (inputs, gt, mask))
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
            step(*devb[i % ring])
    for i in range(args.warmup):
        step(*devb[i % ring])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(*devb[i % ring])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = step.last_launches
    # ---- end-to-end through the public API with HOST buffers ("e2e"): dataloader.DevicePrefetcher copies batch i+1 from
    # pinned host memory on a side stream while step i computes.  Steady state: the feed is one batch longer than the loop,
    # so the K timed steps contain exactly K host->device batch copies (the first timed batch was copied during warm-up,
    # the batch after the last one is copied during the last timed step) and K device->host reads of the loss.
    from keypoints_interpolation_transformer_b200 import dataloader
    n_warm = min(3, args.warmup)
    feed = dataloader.DevicePrefetcher((host[i % ring] for i in range(n_warm + args.steps + 1)), dev)
    for _ in range(n_warm):
        step(*next(feed)).item()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    for _ in range(args.steps):
        last = step(*next(feed)).item()    # device->host read of the loss every step
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel (tcgen05 GEMM), CUDA events inside the engine, same process
    eng = m.engine_for(B_PER_GPU, T, training=True)
    eng.set_profiling(True)
    prof_steps = 3
    acc = {}
    prof_step = step if args.streams == 1 else train.TrainStep(m, opt, criterion="mse", reducer=reducer, use_graph=False)
    for i in range(prof_steps):
        prof_step._eager(*devb[i % ring])   # kernel by kernel, one stream: the events sit between the launches
        torch.cuda.synchronize()
        for k, (pms, n, fl) in eng.profile().items():
            a = acc.setdefault(k, [0.0, 0, 0.0])
            a[0] += pms
            a[1] += n
            a[2] += fl
    eng.set_profiling(False)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # tensor-core family = plain GEMMs + grouped weight gradients + the fused feed-forward kernels
    fam = [acc[k] for k in ("gemm_tn", "gemm_wgrad", "ffn") if k in acc]
    gemm_ms, gemm_n, gemm_fl = (sum(a[i] for a in fam) for i in range(3))
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    family_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # the dominant kernel of the step (largest share of the ncu launch list, profiles/r01c_summary.md): ffn_kernel
    ffn = acc.get("ffn", [0.0, 0, 0.0])
    dom_ms, dom_n, dom_fl = ffn if ffn[1] > 0 else (gemm_ms, gemm_n, gemm_fl)
    achieved_tf = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic = None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get("ffn_kernel_avg_bytes_per_launch") if ffn[1] > 0 else None
    except Exception:
        pass
    step_ms = ms / args.steps
    seqs = B_PER_GPU * world * args.steps
    value = seqs / (ms * 1e-3)
    e2e = seqs / (ms_e2e * 1e-3)
    cpu = cpu_reference_run(steps=2, warmup=1) if (world == 1 and not args.no_cpu_baseline) else None
    # ---- HBM roofline of the fused per-frame passes (pre-pass, loss) at BASELINE configs[3] size, measured live
    frame = None
    if world == 1 and not args.no_framepass:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import framepass_bench
            frame = [{"kernel": r["kernel"], "bound": "hbm", "workload": f"B={r['B']} x T={r['T']} x K={r['K']}",
                      "achieved": r["GBps"], "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                      "frac": r["frac_of_measured_hbm"], "algorithmic_bytes": r["algorithmic_bytes"], "ms": r["ms"]}
                     for r in framepass_bench.measure(shapes=((4096, 256, 71),), peak=peaks.get("hbm_gbs", 6650.0))]
        except Exception as exc:   # the headline line must still be printed
            frame = {"error": repr(exc)}
    breakdown = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] // prof_steps,
                     "tflops": (v[2] / (v[0] * 1e-3) / 1e12) if v[0] > 0 else None} for k, v in acc.items()}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"A1 train step (fwd+loss+bwd+Adam), B={B_PER_GPU}/GPU x T={T} x K={KP}, H={H} L={L}+{L} "
                               f"heads={NH} ff=2048, random missing blocks (AUTSL stats), BASELINE configs[1]",
                   "parallelism": f"dp{world}", "global_batch": B_PER_GPU * world, "cuda_graph": bool(use_graph), "streams": args.streams,
                   "l2": "no flush needed: each step streams ~3 GB of activations/weights (>> 126 MB L2); a ring of "
                         f"{ring} distinct device-resident batches",
                   "model_flops_per_seq": flops_per_seq_train(),
                   "model_tflops": flops_per_seq_train() * value / 1e12},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor",
                     "kernel": "ffn_kernel<BWD> (fused feed-forward block, forward and input-gradient pass)" if ffn[1] > 0
                     else "gemm_tcgen05_kernel (TN + wgrad)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                     "algorithmic_flops_per_launch": dom_fl / dom_n if dom_n else None,
                     "avg_launch_us": 1e3 * dom_ms / dom_n if dom_n else None,
                     "launches_per_step": dom_n // prof_steps,
                     "ms_per_step": dom_ms / prof_steps, "share_of_step": (dom_ms / prof_steps) / step_ms,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                     if peaks else "fallback",
                     "tensor_family": {"what": "all tcgen05 kernels of the step: plain GEMMs + grouped weight gradients + ffn_kernel",
                                       "achieved": family_tf, "frac": family_tf / peak_tf if peak_tf else None,
                                       "launches_per_step": gemm_n // prof_steps, "ms_per_step": gemm_ms / prof_steps,
                                       "share_of_step": (gemm_ms / prof_steps) / step_ms}},
        "breakdown": breakdown,
        "framepass_roofline": frame,
        "cpu_baseline": None if cpu is None else {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                                                 "sample": cpu["sample"]},
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="kit", choices=["kit", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs under ncu)")
    ap.add_argument("--no-framepass", action="store_true", help="skip the pre-pass / loss HBM roofline leg")
    ap.add_argument("--streams", type=int, default=1, help="concurrent sub-batch chains per step (train.TrainStep streams)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":          # rank 0 alone runs it; other ranks exit 0 without work
        reference_arm(args, int(os.environ.get("RANK", "0")))
        return
    gpu_run(args)


if __name__ == "__main__":
    main()
