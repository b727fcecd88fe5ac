"""Data-parallel gradients == single-process gradients on the same global batch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
Every rank computes the gradient of the GLOBAL batch alone (reference), then the ranks run the sharded step with the bucketed
NCCL all-reduce (parallel.BucketReducer, overlapped with backward) and compare: (sum of shard gradients) / world == global."""
import os
import sys

os.environ.setdefault("KIT_DP_COMPRESS", "none")    # this check is about the fp32 path: sharded == global to reduction-order noise

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, parallel, synthetic, train  # noqa: E402


def main():
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    Kp, H, L, NH, T, B_local = 71, 256, 6, 8, 64, 32
    torch.manual_seed(7)
    m = model.KeypointCompleter(2 * Kp, H, L, NH).to(dev)
    m.train()
    dist.broadcast(m.flat_params, src=0)
    inputs, gt, mask = (t.to(dev) for t in synthetic.synthetic_batch(B_local * world, T, Kp, seed=99, smooth=True))
    # reference: the whole global batch on this GPU
    ref_step = train.TrainStep(m, optim.FlatAdam(m, lr=0.0), criterion="mse")
    ref_loss = ref_step.forward_backward(inputs, gt, mask)
    torch.cuda.synchronize()
    ref = m.flat_grads.clone()
    # data parallel: this rank's shard, bucketed all-reduce overlapped with backward, 1 / world applied like FlatAdam does
    lo, hi = parallel.shard_batch(B_local * world, rank, world)
    reducer = parallel.BucketReducer(m.flat_grads, m.layout.buckets)
    dp_step = train.TrainStep(m, optim.FlatAdam(m, lr=0.0), criterion="mse", reducer=reducer)
    loss = dp_step.forward_backward(inputs[lo:hi].contiguous(), gt[lo:hi].contiguous(), mask[lo:hi].contiguous())
    torch.cuda.synchronize()
    got = m.flat_grads / world
    rel = ((got - ref).norm() / ref.norm()).item()
    losses = torch.tensor([loss.item()], device=dev)
    dist.all_reduce(losses)
    mean_loss = losses.item() / world
    ok = rel < 2e-3 and abs(mean_loss - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    # the chained-graph data-parallel step (train.TrainStep(use_graph=True) with a reducer) == the kernel-by-kernel one:
    # six Adam steps each from the same weights on this rank's shard
    shard = tuple(t[lo:hi].contiguous() for t in (inputs, gt, mask))

    def six_steps(use_graph):
        torch.manual_seed(7)
        m2 = model.KeypointCompleter(2 * Kp, H, L, NH).to(dev)
        m2.train()
        dist.broadcast(m2.flat_params, src=0)
        start = m2.flat_params.clone()
        m2.ensure_flat_grads()
        red = parallel.BucketReducer(m2.flat_grads, m2.layout.buckets)
        opt = optim.FlatAdam(m2, lr=1e-4, capturable=use_graph)
        opt.grad_scale = 1.0 / world
        st = train.TrainStep(m2, opt, criterion="mse", reducer=red, use_graph=use_graph)
        ls = [st(*shard).item() for _ in range(6)]
        torch.cuda.synchronize()
        return m2.flat_params[:m2.layout.trainable].clone(), start[:m2.layout.trainable], ls, st

    p_eager, p0, l_eager, _ = six_steps(False)
    p_graph, _, l_graph, st_g = six_steps(True)
    chained = st_g.use_graph and len(st_g._graphs) == 1
    moved = (p_eager - p0).norm().item()
    rel_g = ((p_graph - p_eager).norm() / (p_eager - p0).norm()).item()
    ok_g = chained and rel_g < 1e-2 and all(abs(a - b) < 1e-4 * abs(a) for a, b in zip(l_eager, l_graph))
    print(f"rank {rank}/{world}: chained-graph step vs kernel by kernel after 6 Adam steps: |dparams| / |movement| = {rel_g:.3e} "
          f"(movement {moved:.3e}), losses {l_eager[-1]:.6f} vs {l_graph[-1]:.6f}, graph chain captured: {chained}  "
          f"{'OK' if ok_g else 'MISMATCH'}", flush=True)
    ok = ok and ok_g
    print(f"rank {rank}/{world}: |dp - global| / |global| = {rel:.3e}  mean shard loss {mean_loss:.6f} vs global {ref_loss.item():.6f}  "
          f"bytes reduced {reducer.bytes_reduced}  {'OK' if ok else 'MISMATCH'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
