"""The train / eval step of the reference (A1_train.py:84-218) on the sm_100a engine.

Two surfaces:

* ``train_epoch`` / ``eval_epoch`` keep the reference's signatures and loop structure
  (one dataloader item per iteration, ``(inputs, sota, mask)`` tuples, a list of per-step losses
  returned) so A1_train.py's ``train()`` can call them unchanged.
* ``TrainStep`` / ``EvalStep`` are the batched fast path: the A1 slices are pointer offsets into the
  dataloader tensor, the masks are synthesised in-kernel, loss + dLoss/dpred is one fused pass, the
  gradients land in the flat arena (all-reduced per bucket under data parallelism) and Adam is one
  kernel.  No autograd graph, no per-parameter Python work, no host synchronisation.
"""
import torch

from . import _lib as K
from .engine import make_mask
from .euclidean_loss import EuclideanLoss, fused_loss
from .optim import FlatAdam


def fused_criterion_kind(criterion):
    """The fused-loss kind that computes exactly ``criterion(pred, y)``, or None when the criterion must be called itself."""
    from .euclidean_loss import EuclideanDistanceLoss, MSELoss
    if type(criterion) is EuclideanLoss:
        return "euclid"
    if type(criterion) is EuclideanDistanceLoss:
        return "distance"
    if type(criterion) is MSELoss or (type(criterion) is torch.nn.MSELoss and criterion.reduction == "mean"):
        return "mse"
    return None


class TrainStep:
    """fwd + loss + bwd + (all-reduce) + Adam for one batch laid out as the dataloader yields it:
    inputs [B,T+1,K,2] (SOS + hold-filled frames), sota [B,T,K,2], mask [B,T+1] (A1_train.py:91-135).

    criterion: "mse" (A1_train.py:254) or "euclid" (A4_train_with_pretrained.py:259).
    zero_masked: A4_train_with_pretrained.py:107-108.  reducer: parallel.BucketReducer or None.

    streams > 1: the batch is cut into that many equal sub-batches, each running forward + loss + backward through its own
    engine on its own CUDA stream, all accumulating into the one gradient arena (the weight gradients are reduce-adds);
    Adam follows the join.  Sequences are independent, so the result is the single-stream step's (up to the order of the
    fp32 gradient sums).  At B = 256, T = 64 most kernels are one 128-row tile per SM and latency-bound (launch, first TMA
    tile, tear-down): two half-size chains in flight fill each other's bubbles."""

    def __init__(self, model, optimizer=None, criterion="mse", zero_masked=False, reducer=None, lr=5e-6, use_graph=False,
                 streams=1):
        self.model = model
        self.streams = int(streams)
        self._side = None
        # data parallel: a graph holding NCCL collectives hung process-group teardown on 2 x B200, so the step is captured
        # as a CHAIN of graphs cut at the bucket boundaries of backward and the collectives are launched between them
        self.use_graph = use_graph and (reducer is None or self.streams == 1)
        self.optimizer = optimizer if optimizer is not None else FlatAdam(model, lr=lr, capturable=self.use_graph)
        if self.use_graph and not getattr(self.optimizer, "capturable", False):
            raise ValueError("use_graph=True needs FlatAdam(capturable=True): the step count must live on the device")
        self._graphs = {}        # (input pointers, shape) -> (CUDAGraph, static loss)
        self._eager_calls = 0
        # KIT_DP_SINGLE_GRAPH=1: capture the data-parallel step as ONE graph holding the NCCL all-reduces (side-stream fork /
        # join inside the capture) instead of the chain.  release_graphs() must run before destroy_process_group().
        import os
        self.single_graph = os.environ.get("KIT_DP_SINGLE_GRAPH", "0") == "1"
        self.kind = {"mse": K.LOSS_MSE, "euclid": K.LOSS_EUCLID, "distance": K.LOSS_DISTANCE}[criterion]
        self.zero_masked = zero_masked
        self.reducer = reducer
        if reducer is not None:
            # gradients are SUMMED over the ranks: Adam applies 1 / world (equal local batches: mean of means = global mean)
            scale = 1.0 / max(1, reducer.world)
            cur = getattr(self.optimizer, "grad_scale", None)
            if cur is None:
                raise ValueError("TrainStep(reducer=...) needs an optimizer with a grad_scale attribute (optim.FlatAdam)")
            if cur not in (1.0, scale):
                raise ValueError(f"optimizer.grad_scale = {cur} but the reducer spans {reducer.world} ranks (expected {scale})")
            self.optimizer.grad_scale = scale
        self.pred = None
        self.last_launches = 0
        self._defer_finish = False    # _eager / the graph chain: the collectives are waited for bucket by bucket by _optimizer_step

    def _optimizer_step(self):
        """A1_train.py:135.  Under data parallelism with a range-capable optimiser, every gradient bucket is stepped as soon
        as ITS all-reduce has completed (in the order backward finished them), so Adam on the early buckets runs while the
        last ones are still on the wire."""
        red = self.reducer
        if red is None or self.streams > 1 or not getattr(self.optimizer, "capturable", False) or not hasattr(self.optimizer, "step_range"):
            if red is not None:
                red.finish()
            self.optimizer.step()
            self.last_launches += 3      # adam + weight refresh and cross-attention bias gather at the next forward
            return
        for b, (lo, hi) in enumerate(red.buckets):
            red.wait_bucket(b)
            self.optimizer.step_range(lo, hi, first=(b == 0))
        self.last_launches += 2 + len(red.buckets)

    def _forward_backward_split(self, inputs, sota, mask):
        """The ``streams > 1`` step: fork the current stream into ``streams`` side streams, one sub-batch each, join."""
        model, S = self.model, self.streams
        B, T1 = inputs.shape[0], inputs.shape[1]
        T = T1 - 1
        K2 = inputs.shape[2] * inputs.shape[3]
        Bs = B // S
        if Bs * S != B:
            raise K.KitError(f"TrainStep(streams={S}): batch {B} is not divisible by the number of streams")
        grads = model.ensure_flat_grads()
        if self._side is None:
            self._side = [torch.cuda.Stream(device=inputs.device) for _ in range(S)]
        if self.pred is None or self.pred.shape != (B, T, K2 // 2, 2):
            self.pred = torch.empty(B, T, K2 // 2, 2, device=inputs.device)
        main = torch.cuda.current_stream()
        grads.zero_()                                            # optimizer.zero_grad()  (A1_train.py:133)
        fork = torch.cuda.Event()
        fork.record(main)
        losses, keep = [], []
        launches = 1
        pending = {}       # bucket -> events of the sub-streams that have finished it (data parallel)

        def bucket_ready(b):
            ev = torch.cuda.Event()
            ev.record()
            pending.setdefault(b, []).append(ev)
            if len(pending[b]) == S:
                for e in pending[b]:
                    main.wait_event(e)
                with torch.cuda.stream(main):
                    self.reducer.bucket_ready(b)

        if self.reducer is not None:
            self.reducer.begin()
        for s, side in enumerate(self._side):
            lo, hi = s * Bs, (s + 1) * Bs
            side.wait_event(fork)
            with torch.cuda.stream(side):
                eng = model.engine_for(Bs, T, training=True, slot=s, pin=self.use_graph)
                inp, msk = inputs[lo:hi], mask[lo:hi]
                enc_mask = make_mask(msk[:, :-1], K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
                dec_mask = make_mask(msk[:, 1:], K.MASK_REPEAT_INC)
                pred = self.pred[lo:hi]
                eng.forward(inp, T1 * K2, inp[:, 1:], T1 * K2, enc_mask, dec_mask, pred, self.zero_masked)
                loss, dpred = fused_loss(pred, sota[lo:hi], None, self.kind, want_grad=True, grad_scale=1.0 / S)
                eng.backward(dpred, bucket_ready if self.reducer is not None else None)
                losses.append(loss)
                keep.append((dpred, enc_mask, dec_mask))
                launches += eng.fwd_launches + eng.bwd_launches + 3
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
        total = torch.stack(losses).sum() / S                    # mean over the batch = mean of the equal sub-batch means
        if self.reducer is not None:
            self.reducer.finish()
        self._keep = keep
        self.last_launches = launches + 2
        return total

    def forward_backward(self, inputs, sota, mask, _bucket_cb=None):
        model = self.model
        B, T1 = inputs.shape[0], inputs.shape[1]
        T = T1 - 1
        K2 = inputs.shape[2] * inputs.shape[3]
        assert inputs.is_cuda and inputs.dtype == torch.float32 and inputs.is_contiguous()
        assert mask.dtype == torch.float32 and mask.is_contiguous() and sota.is_contiguous()
        if self.streams > 1:
            return self._forward_backward_split(inputs, sota, mask)
        eng = model.engine_for(B, T, training=True, pin=self.use_graph)   # a captured graph points into its workspace
        grads = model.ensure_flat_grads()
        x_dec = inputs[:, 1:]                                   # A1_train.py:94 -- a pointer offset
        enc_mask = make_mask(mask[:, :-1], K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)     # x_mask  (A1_train.py:99,117,121)
        dec_mask = make_mask(mask[:, 1:], K.MASK_REPEAT_INC)                          # y_mask  (A1_train.py:100,118)
        if self.pred is None or self.pred.shape != (B, T, K2 // 2, 2):
            self.pred = torch.empty(B, T, K2 // 2, 2, device=inputs.device)
        eng.forward(inputs, T1 * K2, x_dec, T1 * K2, enc_mask, dec_mask, self.pred, self.zero_masked)
        loss, dpred = fused_loss(self.pred, sota, None, self.kind, want_grad=True)
        grads.zero_()                                            # optimizer.zero_grad()  (A1_train.py:133)
        if _bucket_cb is not None:                               # segmented capture: the caller owns the collectives
            eng.backward(dpred, _bucket_cb)
        elif self.reducer is not None:
            self.reducer.begin()
            eng.backward(dpred, self.reducer.bucket_ready)
            if not self._defer_finish:
                self.reducer.finish()
        else:
            eng.backward(dpred)
        self.last_launches = eng.fwd_launches + eng.bwd_launches + 3
        return loss

    def _capture_chain(self, *batch):
        """Data-parallel capture: [graph | bucket 0 ready | graph | bucket 1 ready | ... | last bucket ready | wait for the
        collectives | graph (Adam)].  The cuts are made from inside ``kit_engine_backward``'s bucket callback -- the capture of the running
        graph ends, a new one begins on the same stream and shares the memory pool -- so each all-reduce is enqueued on
        the reducer's side stream exactly where the kernel-by-kernel step enqueues it and still overlaps the rest of
        backward, while the ~240 kernels of the step are replayed instead of launched."""
        pool = torch.cuda.graph_pool_handle()
        chain, cur = [], [None]

        def begin():
            cur[0] = torch.cuda.CUDAGraph()
            cur[0].capture_begin(pool=pool, capture_error_mode="thread_local")

        def end():
            cur[0].capture_end()
            chain.append(("graph", cur[0]))
            cur[0] = None

        n_buckets = len(self.reducer.buckets)

        def cut(b):
            end()
            chain.append(("bucket", b))
            if b < n_buckets - 1:         # backward ends with its last bucket: nothing is left to capture
                begin()

        side = torch.cuda.Stream(device=batch[0].device)
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                begin()
                loss = self.forward_backward(*batch, _bucket_cb=cut)
                if cur[0] is not None:     # (a backward that reported fewer buckets than the layout lists)
                    end()
                chain.append(("adam", None))     # stepped bucket by bucket between the replays (_optimizer_step)
        except Exception:
            if cur[0] is not None:
                try:
                    cur[0].capture_end()
                except Exception:   # noqa: BLE001 -- the capture is already invalid
                    pass
            raise
        torch.cuda.current_stream().wait_stream(side)
        return chain, loss

    def _replay_chain(self, chain):
        self.reducer.begin()
        for kind, x in chain:
            if kind == "graph":
                x.replay()
            elif kind == "bucket":
                self.reducer.bucket_ready(x)
            else:
                self._optimizer_step()

    def _eager(self, *batch):
        self._defer_finish = self.reducer is not None and self.streams == 1
        try:
            loss = self.forward_backward(*batch)
        finally:
            self._defer_finish = False
        self._optimizer_step()                                   # A1_train.py:135
        return loss

    def __call__(self, *batch):
        """``batch`` = (inputs, sota, mask) as the dataloader yields them (``RawTrainStep``: the raw keypoint tensor).
        use_graph: the whole step (weight refresh, forward, loss, zero grads, backward, Adam -- ~280 kernel launches
        with their programmatic-dependency edges) is captured once per distinct set of input buffers (e.g. the two slots of
        ``dataloader.DevicePrefetcher``) and replayed; the returned loss is the graph's static 0-dim tensor.  With a
        reducer (data parallelism) the capture is a chain of graphs cut at backward's bucket boundaries and the
        all-reduces are launched between the replays (``_capture_chain``)."""
        if not self.use_graph:
            return self._eager(*batch)
        key = tuple(t.data_ptr() for t in batch) + (tuple(batch[0].shape),)
        entry = self._graphs.get(key)
        if entry is None:
            if self._eager_calls < 2 or len(self._graphs) >= 8:   # plans / attributes are set up by eager steps first
                self._eager_calls += 1
                return self._eager(*batch)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            steps_before = self.optimizer.step_count
            try:
                if self.reducer is not None and not self.single_graph:
                    graph, loss = self._capture_chain(*batch)
                else:
                    with torch.cuda.graph(graph):
                        loss = self._eager(*batch)
            except Exception as exc:   # noqa: BLE001 -- a failed capture must not cost the step: fall back to launches
                import warnings
                warnings.warn(f"CUDA-graph capture of the train step failed ({exc!r}); continuing kernel by kernel")
                self.use_graph = False
                self.optimizer.step_count = steps_before
                torch.cuda.synchronize()
                return self._eager(*batch)
            self.optimizer.step_count = steps_before              # capture enqueues nothing
            entry = (graph, loss, batch)                          # keep the buffers alive
            self._graphs[key] = entry
        self.optimizer.sync_host_values()
        if isinstance(entry[0], list):
            self._replay_chain(entry[0])          # (the optimiser is stepped between the replays and counts its own steps)
        else:
            entry[0].replay()
            self.optimizer.step_count += 1
        return entry[1]

    def release_graphs(self):
        """Drops the captured graphs (call before ``destroy_process_group()`` when a graph holds collectives)."""
        torch.cuda.synchronize()
        self._graphs.clear()
        torch.cuda.synchronize()


class RawTrainStep(TrainStep):
    """The train step FROM RAW KEYPOINTS: what ``LSP_Dataset.__getitem__`` (dataloader.py:623-686) and ``train_epoch``
    (A1_train.py:89-135) do per sequence, for a whole batch on the device and inside the step -- the random policy
    (which augmentation with which parameters, which frames go missing: ``preprocess.DevicePolicy`` = kit_draw_policy), the
    fused pre-pass (normalize_pose, augmentation, hold-fill, SOS, the A1 slices as the engine's bf16 operands, the target
    ``sota``, the frame mask: kit_prepass writing straight into the engine's operand buffers), forward, loss, backward, Adam.
    ``step(raw)`` with raw [B,T,K,2] fp32 on the device; with ``use_graph`` the whole sequence replays as one CUDA graph
    (the policy's Philox offset is a device counter, so every replay draws fresh values)."""

    def __init__(self, model, prepass, policy, optimizer=None, criterion="mse", normalize=True, zero_masked=False, reducer=None,
                 lr=5e-6, use_graph=False):
        super().__init__(model, optimizer, criterion=criterion, zero_masked=zero_masked, reducer=reducer, lr=lr, use_graph=use_graph)
        self.prepass, self.policy, self.normalize = prepass, policy, normalize
        self.y = self.mask = None

    def forward_backward(self, raw, _bucket_cb=None):
        model = self.model
        B, T, Kp = raw.shape[0], raw.shape[1], raw.shape[2]
        assert raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous()
        eng = model.engine_for(B, T, training=True, pin=self.use_graph)
        grads = model.ensure_flat_grads()
        if self.pred is None or self.pred.shape != (B, T, Kp, 2):
            self.pred = torch.empty(B, T, Kp, 2, device=raw.device)
            self.y = torch.empty(B, T, Kp, 2, device=raw.device)
            self.mask = torch.empty(B, T + 1, device=raw.device)
        src, miss, aug = self.policy.draw(B, T)
        xe, xd, k2p = eng.operand_ptrs()
        self.prepass.run(raw, src, miss, aug, self.y, self.mask, xe, xd, k2p, normalize=self.normalize,
                         zero_masked_enc=self.zero_masked)
        enc_mask = make_mask(self.mask[:, :-1], K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)     # x_mask  (A1_train.py:99,117,121)
        dec_mask = make_mask(self.mask[:, 1:], K.MASK_REPEAT_INC)                          # y_mask  (A1_train.py:100,118)
        eng.forward(None, 0, None, 0, enc_mask, dec_mask, self.pred, False)               # operands are in place
        loss, dpred = fused_loss(self.pred, self.y, None, self.kind, want_grad=True)
        grads.zero_()
        if _bucket_cb is not None:
            eng.backward(dpred, _bucket_cb)
        elif self.reducer is not None:
            self.reducer.begin()
            eng.backward(dpred, self.reducer.bucket_ready)
            if not self._defer_finish:
                self.reducer.finish()
        else:
            eng.backward(dpred)
        self.last_launches = eng.fwd_launches + eng.bwd_launches + 3 + 2 + (1 if B <= 512 else 2)   # + policy (2) and pre-pass (1; 2 beyond 512 sequences) launches
        return loss


class EvalStep:
    """A1_train.py:149-186 batched: forward, blend pred*m + y*(1-m), EuclideanLoss -- masks and blend
    fused into the kernels.  Returns (loss, pred_blended?)."""

    def __init__(self, model, zero_masked=False):
        self.model = model
        self.zero_masked = zero_masked
        self.pred = None

    @torch.no_grad()
    def __call__(self, inputs, sota, mask, blend_output=False):
        model = self.model
        B, T1 = inputs.shape[0], inputs.shape[1]
        T = T1 - 1
        K2 = inputs.shape[2] * inputs.shape[3]
        eng = model.engine_for(B, T, training=False)
        enc_mask = make_mask(mask[:, :-1], K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD)
        dec_mask = make_mask(mask[:, 1:], K.MASK_REPEAT_INC)
        if self.pred is None or self.pred.shape != (B, T, K2 // 2, 2):
            self.pred = torch.empty(B, T, K2 // 2, 2, device=inputs.device)
        eng.forward(inputs, T1 * K2, inputs[:, 1:], T1 * K2, enc_mask, dec_mask, self.pred, self.zero_masked)
        y_mask = mask[:, 1:].contiguous()
        loss, _ = fused_loss(self.pred, sota, y_mask, K.LOSS_EUCLID, want_grad=False)
        if blend_output:
            m = y_mask[:, :, None, None]
            return loss, self.pred * m + sota * (1 - m)
        return loss, self.pred


# ---------------------------------------------------------------------------------------------
# Reference-shaped epoch loops (A1_train.py:84-137, :139-218)
# ---------------------------------------------------------------------------------------------
def train_epoch(model, dataloader, criterion, optimizer, device):
    """Same contract as A1_train.py:84: iterates ``dataloader`` yielding (inputs, sota, mask) with a
    leading batch dimension (1 in the reference, any B here), returns the list of per-step losses
    (numpy scalars).  With a ``FlatAdam`` optimizer the fused step runs; with any other
    ``torch.optim`` optimizer the autograd-compatible module path runs, call for call like A1."""
    model.train()
    losses = []
    # The fused step evaluates the loss inside the engine: only criteria it implements exactly are routed there.  Anything
    # else (WeightedMSELoss, nn.L1Loss, an MSELoss with a non-default reduction, a user criterion) is called as the reference
    # calls it -- criterion(pred, y) through the autograd-compatible module path -- whatever the optimizer.
    kind = fused_criterion_kind(criterion)
    fused = isinstance(optimizer, FlatAdam) and kind is not None
    step = None
    if fused:
        step = TrainStep(model, optimizer, criterion=kind)
    elif isinstance(optimizer, FlatAdam):
        model.attach_flat_grads()      # loss.backward() accumulates into views of the arena FlatAdam steps on
    for i, data in enumerate(dataloader):
        inputs, sota, mask = data
        inputs = inputs.to(device).float().contiguous()
        sota = sota.to(device).float().contiguous()
        mask = mask.to(device).float().contiguous()
        if fused:
            loss = step(inputs, sota, mask)
        else:
            x = inputs[:, :-1]
            x_no_sota = inputs[:, 1:]
            x_mask = mask[:, :-1]
            y_mask = mask[:, 1:]
            pred = model(x, x_no_sota, frame_masks=(x_mask, y_mask))
            loss = criterion(pred, sota).float()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            model.mark_dirty()
        losses.append(loss.clone().detach().cpu().numpy())
    return losses


def eval_epoch(model, dataloader, criterion, epoch, device):
    """Same contract as A1_train.py:139 (losses list; the wandb/cubic extras stay caller-side)."""
    model.eval()
    losses = []
    step = EvalStep(model)
    last = None
    for i, data in enumerate(dataloader):
        inputs, sota, mask = data
        inputs = inputs.to(device).float().contiguous()
        sota = sota.to(device).float().contiguous()
        mask = mask.to(device).float().contiguous()
        if isinstance(criterion, EuclideanLoss):
            loss, pred = step(inputs, sota, mask, blend_output=(i == 1))
        else:
            _, pred = step(inputs, sota, mask, blend_output=True)
            loss = criterion(pred, sota)
        losses.append(loss.clone().detach().cpu().numpy())
        if i == 1:
            x_mask = mask[:, :-1]
            last = {"inputs": inputs[:, :-1] * (1 - x_mask)[:, :, None, None], "prediction": pred, "sota": sota,
                    "epoch": epoch}
    return losses, last


def lr_lambda(current_step, lr, optim):
    """A1_train.py:42-54: write the per-epoch learning rate into the optimiser."""
    for group in optim.param_groups:
        group["lr"] = lr[current_step]
    return optim
