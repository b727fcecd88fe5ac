"""Adam over the flat parameter arena: one kernel for the whole model (A1_train.py:135,256 uses
torch.optim.Adam with default betas/eps, no weight decay).  ``param_groups[0]['lr']`` is honoured at
every step so the reference's per-epoch ``lr_lambda`` write (A1_train.py:42-54) keeps working."""
import torch

from . import _lib as K


class FlatAdam:
    """``capturable=True`` keeps the step count and the learning rate in a 16-byte device record (kit_adam_step_dev) so
    that ``train.TrainStep(use_graph=True)`` can replay the whole step as one CUDA graph; ``param_groups[0]['lr']`` is
    still honoured (copied to the device when it changes)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        self.model = model
        self.capturable = capturable
        self._dev_state = None
        self._dev_lr = None
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "params": list(model.parameters())}]
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self.grad_scale = 1.0        # 1/world_size under data parallelism (sum all-reduce)

    def _state(self):
        n = self.model.layout.trainable
        dev = self.model.flat_params.device
        if self.exp_avg is None or self.exp_avg.device != dev:
            self.exp_avg = torch.zeros(n, device=dev)
            self.exp_avg_sq = torch.zeros(n, device=dev)
        return n

    def zero_grad(self, set_to_none=False):
        g = self.model.ensure_flat_grads()
        g.zero_()

    def _device_state(self):
        """[completed_steps:int32, lr, step_size, inv_sqrt_bc2] on the device; lr re-uploaded when the host value changed."""
        dev = self.model.flat_params.device
        if self._dev_state is None or self._dev_state.device != dev:
            self._dev_state = torch.zeros(4, dtype=torch.float32, device=dev)
            self._dev_state.view(torch.int32)[0] = self.step_count
            self._dev_lr = None
        lr = float(self.param_groups[0]["lr"])
        if self._dev_lr != lr:
            self._dev_state[1] = lr
            self._dev_lr = lr
        return self._dev_state

    def sync_host_values(self):
        """Push lr (and nothing else) to the device record; call before replaying a captured step."""
        if self.capturable:
            self._device_state()

    @torch.no_grad()
    def step(self):
        n = self._state()
        g = self.model.ensure_flat_grads()
        grp = self.param_groups[0]
        st = self._device_state() if self.capturable else None   # (created from the count BEFORE this step)
        self.step_count += 1
        if self.capturable:
            K.check(K.lib().kit_adam_step_dev(K.ptr(self.model.flat_params), K.ptr(g), K.ptr(self.exp_avg),
                                              K.ptr(self.exp_avg_sq), n, K.ptr(st), float(grp["betas"][0]),
                                              float(grp["betas"][1]), float(grp["eps"]), float(self.grad_scale),
                                              K.stream_ptr()))
            self.model.mark_dirty()
            return
        K.check(K.lib().kit_adam_step(K.ptr(self.model.flat_params), K.ptr(g), K.ptr(self.exp_avg),
                                      K.ptr(self.exp_avg_sq), n, float(grp["lr"]), float(grp["betas"][0]),
                                      float(grp["betas"][1]), float(grp["eps"]), self.step_count,
                                      float(self.grad_scale), K.stream_ptr()))
        self.model.mark_dirty()

    @torch.no_grad()
    def step_range(self, lo, hi, first):
        """Adam over arena range [lo, hi) (``capturable`` mode): data-parallel training steps each gradient bucket as soon as its
        all-reduce has completed.  ``first``: the first range of this optimiser step (advances the step count)."""
        if not self.capturable:
            raise K.KitError("FlatAdam.step_range needs capturable=True (step count on the device)")
        self._state()
        g = self.model.ensure_flat_grads()
        grp = self.param_groups[0]
        st = self._device_state()
        if first:
            self.step_count += 1
        K.check(K.lib().kit_adam_step_dev_range(K.ptr(self.model.flat_params[lo:hi]), K.ptr(g[lo:hi]), K.ptr(self.exp_avg[lo:hi]),
                                                K.ptr(self.exp_avg_sq[lo:hi]), hi - lo, K.ptr(st), float(grp["betas"][0]),
                                                float(grp["betas"][1]), float(grp["eps"]), float(self.grad_scale),
                                                1 if first else 0, K.stream_ptr()))
        self.model.mark_dirty()

    def state_dict(self):
        self._state()
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self._state()
        self.step_count = int(sd["step"])
        if self._dev_state is not None:
            self._dev_state.view(torch.int32)[0] = self.step_count
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])
