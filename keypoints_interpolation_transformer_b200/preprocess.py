"""Python handle on the fused per-frame pre-pass (include/kit.h: kit_prepass): pose normalisation,
augmentation, hold-fill of missing blocks, SOS frame and the A1 input slices in ONE pass over the
[B,T,K,2] keypoint tensor."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as K


def perspective_matrix(src, dst):
    """The 3x3 homography cv2.getPerspectiveTransform(src, dst) returns (augmentation.py:173,187),
    solved in float64 from the float32 corner arrays."""
    A = np.zeros((8, 8), dtype=np.float64)
    rhs = np.zeros(8, dtype=np.float64)
    for i in range(4):
        x, y = float(src[i][0]), float(src[i][1])
        u, v = float(dst[i][0]), float(dst[i][1])
        A[i] = [x, y, 1, 0, 0, 0, -x * u, -y * u]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * v, -y * v]
        rhs[i], rhs[i + 4] = u, v
    return np.append(np.linalg.solve(A, rhs), 1.0).reshape(3, 3)


def aug_none():
    a = K.KitSeqAug()
    a.kind = K.AUG_NONE
    return a


def aug_rotate(angle):
    a = K.KitSeqAug()
    a.kind = K.AUG_ROTATE
    a.cos_t, a.sin_t = math.cos(angle), math.sin(angle)
    return a


def aug_shear(mtx):
    a = K.KitSeqAug()
    a.kind = K.AUG_SHEAR
    for i, v in enumerate(np.asarray(mtx, dtype=np.float64).reshape(9)):
        a.mtx[i] = float(v)
    w0 = mtx[2][2]
    w0 = 1.0 / w0 if abs(w0) > np.finfo(np.float64).eps else 0.0
    a.zero_x = float(np.float32(mtx[0][2] * w0))
    a.zero_y = float(np.float32(mtx[1][2] * w0))
    return a


def aug_arm(angles):
    """angles[c][j]: radians or None (coin failed) for chain c (2) and joint j (4)."""
    a = K.KitSeqAug()
    a.kind = K.AUG_ARM_ROTATE
    for c in range(2):
        for j in range(4):
            ang = angles[c][j] if c < len(angles) and j < len(angles[c]) else None
            a.arm_cos[c * 4 + j] = 2.0 if ang is None else math.cos(ang)
            a.arm_sin[c * 4 + j] = 0.0 if ang is None else math.sin(ang)
    return a


class DevicePolicy:
    """The per-batch random policy of ``LSP_Dataset.__getitem__`` (dataloader.py:649-675) drawn ON THE DEVICE (kit_draw_policy):
    augmentation selection and parameters + missing blocks, Philox streams (seed, device counter) -- the reference's
    distributions, not its draws; ``dataloader.KeypointBatcher`` keeps the host path in the reference's RNG order."""

    def __init__(self, dataset_name="AUTSL", seed=0, have_augmentation=True, augmentations_prob=0.5, has_arms=True, device="cuda",
                 config=None):
        from . import missing
        cfg = (config or missing.DATASET_CONFIG)[dataset_name]
        self.stats = K.KitMissingStats(cfg["mean_consecutive_missing"], cfg["std_consecutive_missing"],
                                       cfg["mean_number_missing_blocks"], cfg["std_number_missing_blocks"], int(cfg["samples"]))
        self.aug = K.KitAugPolicy(augmentations_prob, 15.0, 0.15, 0.5, 1 if has_arms else 0) if have_augmentation else None
        self.seed = int(seed)
        self.device = torch.device(device)
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)      # Philox offset, advanced by every draw
        self._bufs = {}

    def draw(self, B, T, want_draws=False):
        """-> (src_index [B,T] int32, frame_missing [B,T] fp32, aug [B * sizeof(KitSeqAug)] uint8 or None[, draws [B,12] fp64])."""
        key = (B, T)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(B, T, dtype=torch.int32, device=self.device),
                               torch.empty(B, T, dtype=torch.float32, device=self.device),
                               torch.empty(B * C.sizeof(K.KitSeqAug), dtype=torch.uint8, device=self.device) if self.aug else None)
        src, miss, aug = self._bufs[key]
        draws = torch.empty(B, 12, dtype=torch.float64, device=self.device) if (want_draws and self.aug) else None
        K.check(K.lib().kit_draw_policy(C.byref(self.stats), C.byref(self.aug) if self.aug else None, B, T, self.seed,
                                        K.ptr(self.counter), K.ptr(src), K.ptr(miss), K.ptr(aug), K.ptr(draws), K.stream_ptr()))
        return (src, miss, aug, draws) if want_draws else (src, miss, aug)


class Prepass:
    """Static description of the skeleton (which keypoints are body / hands / shoulders / arms)."""

    def __init__(self, K_points, device, body_ids=None, hand_ids=None, left_shoulder=0, right_shoulder=0, right_eye=0,
                 arm_chains=None):
        self.K, self.device = K_points, torch.device(device)
        if self.device.type != "cuda":
            raise K.KitError("the pre-pass runs on a CUDA device: this build has no CPU path")
        self.body = None if body_ids is None else torch.tensor(list(body_ids), dtype=torch.int32, device=device)
        self.hand = None if hand_ids is None else torch.tensor(list(hand_ids), dtype=torch.int32, device=device)
        self.ids = (left_shoulder, right_shoulder, right_eye)
        self.arm_chains = arm_chains if arm_chains is not None else [[0, 0, 0, 0], [0, 0, 0, 0]]
        self.k2p = (2 * K_points + 7) // 8 * 8

    def _config(self, B, T, normalize, zero_masked_enc, k2p):
        cfg = K.KitPrepassConfig()
        cfg.B, cfg.T, cfg.K = B, T, self.K
        cfg.normalize = 1 if normalize else 0
        cfg.left_shoulder, cfg.right_shoulder, cfg.right_eye = self.ids
        cfg.n_body = 0 if self.body is None else self.body.numel()
        cfg.n_hand = 0 if self.hand is None else self.hand.numel()
        for c in range(2):
            for j in range(4):
                cfg.arm_chain[c * 4 + j] = int(self.arm_chains[c][j])
        cfg.zero_masked_enc = 1 if zero_masked_enc else 0
        cfg.k2p = k2p
        return cfg

    def run(self, raw, src_index, frame_missing, aug_dev, y, mask, x_enc_ptr, x_dec_ptr, k2p, normalize=False, zero_masked_enc=False):
        """The in-step form: every buffer is the caller's (no allocation), the bf16 operands go to raw device pointers (the
        engine's own operand buffers), the fp32 ``inputs`` tensor nobody reads is not written.  ``aug_dev``: uint8 tensor
        holding [B] KitSeqAug records (DevicePolicy.draw) or None."""
        B, T = raw.shape[0], raw.shape[1]
        cfg = self._config(B, T, normalize, zero_masked_enc, k2p)
        K.check(K.lib().kit_prepass(C.byref(cfg), K.ptr(raw), K.ptr(src_index), K.ptr(frame_missing), K.ptr(aug_dev),
                                    K.ptr(self.body), K.ptr(self.hand), K.ptr(y), None, K.ptr(mask), x_enc_ptr, x_dec_ptr,
                                    K.stream_ptr()))

    def __call__(self, raw, src_index=None, frame_missing=None, augs=None, normalize=False, zero_masked_enc=False,
                 want_inputs=True, want_bf16=False, aug_dev=None):
        """raw [B,T,K,2] fp32 CUDA.  Returns dict(y, inputs, mask, x_enc, x_dec).  ``augs``: list of [B] KitSeqAug (host), or
        ``aug_dev``: the records already on the device (DevicePolicy.draw)."""
        raw = raw.to(self.device).float().contiguous()
        B, T, Kp, _ = raw.shape
        assert Kp == self.K
        if src_index is None:
            src_index = torch.arange(T, dtype=torch.int32, device=self.device).repeat(B, 1)
        if frame_missing is None:
            frame_missing = torch.zeros(B, T, device=self.device)
        src_index = src_index.to(self.device, torch.int32).contiguous()
        frame_missing = frame_missing.to(self.device, torch.float32).contiguous()
        cfg = K.KitPrepassConfig()
        cfg.B, cfg.T, cfg.K = B, T, Kp
        cfg.normalize = 1 if normalize else 0
        cfg.left_shoulder, cfg.right_shoulder, cfg.right_eye = self.ids
        cfg.n_body = 0 if self.body is None else self.body.numel()
        cfg.n_hand = 0 if self.hand is None else self.hand.numel()
        for c in range(2):
            for j in range(4):
                cfg.arm_chain[c * 4 + j] = int(self.arm_chains[c][j])
        cfg.zero_masked_enc = 1 if zero_masked_enc else 0
        cfg.k2p = self.k2p if want_bf16 else 0
        if augs is not None:
            assert len(augs) == B
            arr = (K.KitSeqAug * B)(*augs)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            aug_dev = host.to(self.device)
        y = torch.empty_like(raw)
        inputs = torch.empty(B, T + 1, Kp, 2, device=self.device) if want_inputs else None
        mask = torch.empty(B, T + 1, device=self.device) if want_inputs else None
        xe = torch.empty(B * T, self.k2p, dtype=torch.bfloat16, device=self.device) if want_bf16 else None
        xd = torch.empty(B * T, self.k2p, dtype=torch.bfloat16, device=self.device) if want_bf16 else None
        K.check(K.lib().kit_prepass(C.byref(cfg), K.ptr(raw), K.ptr(src_index), K.ptr(frame_missing), K.ptr(aug_dev),
                                    K.ptr(self.body), K.ptr(self.hand), K.ptr(y), K.ptr(inputs), K.ptr(mask), K.ptr(xe),
                                    K.ptr(xd), K.stream_ptr()))
        return {"y": y, "inputs": inputs, "mask": mask, "x_enc": xe, "x_dec": xd}
