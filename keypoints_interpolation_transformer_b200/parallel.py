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
            # The all-reduce kernels that run BESIDE backward use the SMs the compute kernels leave free
            # (reserve_sms_for_collectives): the default communicator is capped at that many CTAs (ncclConfig maxCTAs) so the two
            # never compete for an SM.  The cap is set on the communicator, not through NCCL_MAX_CTAS: the all-reduce of the LAST
            # bucket runs after backward has ended, on its own uncapped communicator (BucketReducer.tail_group).
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = COLLECTIVE_SMS
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local), pg_options=opts)
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


# SMs left to NCCL's kernels (and NCCL_MAX_CTAS): 8 on 2 GPUs, 16 from 4 GPUs up (8 GPUs: 4.30 ms per step against 4.35 with 8,
# profiles/r02_dp.md)
COLLECTIVE_SMS = int(os.environ.get("KIT_COLLECTIVE_SMS", "16" if int(os.environ.get("WORLD_SIZE", "1")) >= 4 else "8"))


def reserve_sms_for_collectives(n=None):
    """Every persistent compute kernel of the library sizes its grid to (SM count - n): NCCL's CTAs then have SMs of their own
    instead of waiting for a compute CTA to retire and pushing the next kernel's CTAs into a second wave.  Must run before the
    first engine is created (plans cache their grids); ``BucketReducer`` calls it."""
    from . import _lib as K
    n = COLLECTIVE_SMS if n is None else n
    if os.environ.get("KIT_SM_RESERVE") is None:      # an explicit environment setting wins (A/B measurements)
        K.check(K.lib().kit_set_sm_reserve(int(n)))
    return n


def shard_batch(n_items, rank, world):
    """Contiguous equal shards (the remainder is dropped so every rank has the same local batch)."""
    per = n_items // world
    return rank * per, (rank + 1) * per


class BucketReducer:
    """All-reduces ``flat_grads[lo:hi]`` for each bucket in the order backward completes them, on a side stream.

    ``compress="bf16"`` (``KIT_DP_COMPRESS=bf16``; default is fp32): the bucket is cast to bf16, summed over the ranks in bf16
    and cast back -- half the NVLink bytes, but two more passes over the bucket; on NVSwitch the fp32 sum measured FASTER
    (2 GPUs: 4.26 ms per step against 4.31) and exact to reduction order (``bench.py``'s ``dp_grad_check``: 2e-7 against 2e-3),
    so it stays the default.

    Every bucket records an event when its reduced gradients are back in the arena: ``wait_bucket(b)`` lets the optimiser
    step that range while later buckets are still on the wire (``train.TrainStep`` / ``FlatAdam.step_range``)."""

    def __init__(self, flat_grads, buckets, group=None, compress=None):
        self.flat = flat_grads
        self.buckets = list(buckets)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grads.is_cuda
        if self.cuda and self.world > 1:
            reserve_sms_for_collectives()
        self.stream = torch.cuda.Stream() if self.cuda else None
        # The last bucket's all-reduce starts when backward has ended: nothing competes for SMs any more, and every microsecond of
        # it is exposed (2 GPUs: 5.9 MB took 110 us on the 8-CTA communicator).  It gets a communicator of its own without the cap.
        self.tail_group = None
        if self.cuda and self.world > 1 and group is None and dist.get_backend() == "nccl" and os.environ.get("KIT_DP_TAIL_GROUP", "1") == "1":
            self.tail_group = dist.new_group(backend="nccl")
        if compress is None:
            compress = os.environ.get("KIT_DP_COMPRESS", "none")
        self.compress = compress if (self.cuda and compress == "bf16") else "none"
        self.stage = torch.empty(flat_grads.numel(), dtype=torch.bfloat16, device=flat_grads.device) if self.compress == "bf16" else None
        self.works = []
        self.events = {}
        self.bytes_reduced = 0

    def begin(self):
        self.works = []
        self.events = {}

    def bucket_ready(self, b):
        if self.world == 1 or os.environ.get("KIT_DP_NO_ALLREDUCE") == "1":    # (timing experiments: the step without its collectives)
            return
        lo, hi = self.buckets[b]
        chunk = self.flat[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()                               # backward kernels enqueued so far
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                if self.compress == "bf16":
                    wire = self.stage[lo:hi]
                    wire.copy_(chunk)                 # fp32 -> bf16
                    dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()
                    chunk.copy_(wire)                 # back into the arena
                    self.bytes_reduced += wire.numel() * 2
                else:
                    grp = self.tail_group if (self.tail_group is not None and b == len(self.buckets) - 1) else self.group
                    dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=grp, async_op=True).wait()
                    self.bytes_reduced += chunk.numel() * 4
                done = torch.cuda.Event()
                done.record(self.stream)              # (work.wait() made the side stream wait for NCCL's stream)
                self.events[b] = done
        else:
            self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.bytes_reduced += chunk.numel() * 4

    def wait_bucket(self, b):
        """The current stream waits until bucket b's reduced gradients are in the arena."""
        ev = self.events.get(b)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def finish(self):
        for w in self.works:
            w.wait()
        if self.cuda:
            for b in sorted(self.events):
                self.wait_bucket(b)
        self.works = []

    def reduce_all(self):
        """Non-overlapped path (used by tests): every bucket after backward."""
        self.begin()
        for b in range(len(self.buckets)):
            self.bucket_ready(b)
        self.finish()
