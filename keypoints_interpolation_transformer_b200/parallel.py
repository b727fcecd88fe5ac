"""Data parallelism for the train step: one process per GPU (torchrun), batches sharded across ranks,
ONE exchange per iteration -- a sum all-reduce of the flat gradient arena in a few large contiguous
buckets, each launched as soon as backward has finished that range (``kit_engine_backward``'s
bucket callback) on a side stream so NCCL traffic over NVLink overlaps the rest of backward.
The reference has no distributed code (A1_train.py:244 batch_size=1, single device); equal local
batch sizes make mean-of-means equal to the global mean, so gradients are summed and Adam applies
1/world (FlatAdam.grad_scale)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(n_items, rank, world):
    """Contiguous equal shards (the remainder is dropped so every rank has the same local batch)."""
    per = n_items // world
    return rank * per, (rank + 1) * per


class BucketReducer:
    """All-reduces ``flat_grads[lo:hi]`` for each bucket in the order backward completes them."""

    def __init__(self, flat_grads, buckets, group=None):
        self.flat = flat_grads
        self.buckets = list(buckets)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grads.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.works = []
        self.bytes_reduced = 0

    def begin(self):
        self.works = []

    def bucket_ready(self, b):
        if self.world == 1:
            return
        lo, hi = self.buckets[b]
        chunk = self.flat[lo:hi]
        self.bytes_reduced += chunk.numel() * 4
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()                               # backward kernels enqueued so far
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in self.works:
            w.wait()                                  # makes the current stream wait for the collective
        if self.cuda and self.works:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.works = []

    def reduce_all(self):
        """Non-overlapped path (used by tests): every bucket after backward."""
        self.begin()
        for b in range(len(self.buckets)):
            self.bucket_ready(b)
        self.finish()
