"""Drop-ins for the per-batch pieces of the reference's ``dataloader.py`` that sit on the hot path:
``normalize_pose`` (:71-140), ``put_missing_frames`` (:314-436), ``add_sos`` (:482-493) and a
GPU-resident batch source replacing ``LSP_Dataset.__getitem__`` (:623-686) for fixed-length
sequences.  The random POLICY (which frames go missing, which augmentation with which parameters) is
drawn on the host exactly as the reference draws it; the per-element work runs in the fused pre-pass
kernel.  HDF5 loading (:227-279) is out of scope (no h5py / dataset here)."""
import math
import random

import numpy as np
import torch

from . import missing
from . import preprocess as PP


def _as_cuda(x, device="cuda"):
    t = torch.as_tensor(x)
    return t.to(device).float()


def normalize_pose(data, body_dict, device="cuda"):
    """dataloader.py:71-140.  ``data`` [T,K,2] (numpy or tensor) is normalised IN PLACE like the
    reference and returned; ``body_dict`` maps 'pose_left_shoulder', 'pose_right_shoulder',
    'pose_right_eye' to keypoint indices."""
    t = _as_cuda(data, device)
    pp = PP.Prepass(t.shape[1], device, left_shoulder=body_dict['pose_left_shoulder'],
                    right_shoulder=body_dict['pose_right_shoulder'], right_eye=body_dict['pose_right_eye'])
    out = pp(t.unsqueeze(0), normalize=True, want_inputs=False)["y"][0]
    if isinstance(data, np.ndarray):
        data[...] = out.cpu().numpy()
        return data
    data.copy_(out.to(data.device))
    return data


def put_missing_frames(video, is_random_missing, dataset_name, device="cuda"):
    """dataloader.py:314-436: returns (video_with_missing [T,K,2], mask [T]); ``video`` is modified in
    place like the reference.  RNG: python ``random`` and ``numpy.random`` global streams, consumed in
    the reference's order."""
    T = video.shape[0]
    src, mask = missing.draw_sources(T, is_random_missing, dataset_name)
    pp = PP.Prepass(video.shape[1], device)
    res = pp(_as_cuda(video, device).unsqueeze(0), torch.from_numpy(src).unsqueeze(0), torch.from_numpy(mask).unsqueeze(0))
    out = res["inputs"][0, 1:]
    video.copy_(out.to(video.device))
    return video, res["mask"][0, 1:].to(video.device)


def add_sos(video, mask=None):
    """dataloader.py:482-493."""
    sos = torch.ones(1, video.shape[1], video.shape[2], device=video.device, dtype=video.dtype)
    video = torch.cat((sos, video), dim=0)
    if mask is not None:
        return video, torch.cat((torch.zeros(1, device=mask.device, dtype=mask.dtype), mask))
    return video


class KeypointBatcher:
    """GPU-resident replacement of ``LSP_Dataset`` + ``DataLoader`` for equal-length sequences: holds
    raw keypoints [N,T,K,2] on the device and yields ``(inputs [B,T+1,K,2], sota [B,T,K,2],
    mask [B,T+1])`` batches -- normalisation, augmentation (dataloader.py:649-663 dispatch, p=0.5,
    uniform choice of 4), missing blocks and SOS in one fused pass per batch.

    Documented deviation: the reference's augmentations mutate the dataset in
    place and accumulate over epochs; here they are applied functionally per batch."""

    def __init__(self, raw, body_type_identifiers=None, body_section_dict=None, dataset_name="AUTSL", normalize=True,
                 have_augmentation=True, augmentations_prob=0.5, is_random_missing=False, device="cuda", seed=None,
                 device_policy=False):
        self.raw = _as_cuda(raw, device).contiguous()
        self.N, self.T, self.K = self.raw.shape[:3]
        self.dataset_name, self.normalize = dataset_name, normalize
        self.have_augmentation, self.augmentations_prob = have_augmentation, augmentations_prob
        self.is_random_missing = is_random_missing
        self.rng = random.Random(seed) if seed is not None else random
        self.nprng = np.random.RandomState(seed) if seed is not None else np.random
        ids = body_type_identifiers or {"pose": list(range(self.K)), "left_hand": [], "rigth_hand": []}
        bd = body_section_dict or {}
        body = ids["pose"] + ids["left_hand"] + ids["rigth_hand"]
        hand = ids["left_hand"] + ids["rigth_hand"]
        chains = [[bd.get(n, 0) for n in ("pose_chest_middle_up", f"pose_{s}_shoulder", f"pose_{s}_elbow", f"pose_{s}_wrist")]
                  for s in ("left", "right")]
        self.has_arms = all(n in bd for n in ("pose_chest_middle_up", "pose_left_wrist", "pose_right_wrist"))
        self.pp = PP.Prepass(self.K, device, body, hand, bd.get("pose_left_shoulder", 0), bd.get("pose_right_shoulder", 0),
                             bd.get("pose_right_eye", 0), chains)
        self.device = device
        # device_policy: the missing blocks of a batch are drawn by kit_draw_missing (Philox) instead of one host draw per
        # sequence in the reference's RNG order (0.8 ms per sequence: 40x the train step at B=256)
        self.device_policy = device_policy and not is_random_missing and dataset_name != "all"
        self._policy_seed = seed if seed is not None else random.getrandbits(48)
        self._policy_calls = 0

    def _draw_aug(self):
        """dataloader.py:649-663 with the draws of augmentation.py:132,166-185,223-224."""
        r = self.rng
        if not (self.have_augmentation and r.random() < self.augmentations_prob):
            return PP.aug_none()
        sel = r.randrange(4)
        src = np.array(((0, 1), (1, 1), (0, 0), (1, 0)), dtype=np.float32)
        if sel == 0:
            return PP.aug_rotate(math.radians(r.uniform(-15, 15)))
        if sel == 1:
            a = r.uniform(-0.15, 0.15)
            if r.random() < 0.5:
                dst = np.array(((0 + a, 1 - a), (1, 1), (0 + a, 0 + a), (1, 0)), dtype=np.float32)
            else:
                dst = np.array(((0, 1), (1 - a, 1 - a), (0, 0), (1 - a, 0 + a)), dtype=np.float32)
            return PP.aug_shear(PP.perspective_matrix(src, dst))
        if sel == 2:
            ml, mr = r.uniform(-0.15, 0.15), r.uniform(-0.15, 0.15)
            dst = np.array(((0 + ml, 1), (1 - mr, 1), (0 + ml, 0), (1 - mr, 0)), dtype=np.float32)
            return PP.aug_shear(PP.perspective_matrix(src, dst))
        if not self.has_arms:
            return PP.aug_none()
        angles = [[math.radians(r.uniform(-15, 15)) if r.random() < 0.5 else None for _ in range(4)] for _ in range(2)]
        return PP.aug_arm(angles)

    def batch(self, indices, want_bf16=False):
        """One batch.  Host policy: per sequence, in the reference's order -- the augmentation draws of
        ``__getitem__`` (dataloader.py:649-663), then the missing-block draws of ``put_missing_frames`` (:674) -- so a
        seeded batch consumes the RNG streams exactly like the reference's DataLoader iterating the same indices."""
        idx = torch.as_tensor(indices, device=self.device, dtype=torch.long)
        B = idx.numel()
        if self.device_policy:
            augs = [self._draw_aug() for _ in range(B)]
            src_t, msk_t = missing.draw_sources_device(B, self.T, self.dataset_name, self._policy_seed,
                                                       offset=self._policy_calls, device=self.device,
                                                       config=missing.DATASET_CONFIG)
            self._policy_calls += 1
        else:
            augs = []
            src = np.empty((B, self.T), dtype=np.int32)
            msk = np.empty((B, self.T), dtype=np.float32)
            for b in range(B):
                augs.append(self._draw_aug())
                src[b], msk[b] = missing.draw_sources(self.T, self.is_random_missing, self.dataset_name, self.rng, self.nprng,
                                                      missing.DATASET_CONFIG)
            src_t, msk_t = torch.from_numpy(src), torch.from_numpy(msk)
        res = self.pp(self.raw[idx], src_t, msk_t, augs, normalize=self.normalize, want_bf16=want_bf16)
        return res["inputs"], res["y"], res["mask"]

    def __iter__(self):
        order = list(range(self.N))
        self.rng.shuffle(order)
        return iter(order)


class DevicePrefetcher:
    """Feeds ``(inputs, sota, mask)`` tuples of PINNED host tensors (what ``DataLoader(pin_memory=True)`` yields for the
    reference's dataset, A1_train.py:244) to the device one batch ahead: the host->device copy of batch i+1 runs on a side
    stream while the train step of batch i computes, so the step never waits for PCIe.  Two device-side slots are re-used;
    a slot is overwritten only after the step that consumed it has finished (event recorded by ``__next__``).

        for inputs, sota, mask in DevicePrefetcher(loader, device):
            loss = step(inputs, sota, mask)
    """

    def __init__(self, iterable, device="cuda"):
        self.it = iter(iterable)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.ready = [None, None]          # copy finished
        self.released = [None, None]       # consumer finished with the slot
        self.cur = 0
        self.h2d_bytes = 0
        self._pending = self._issue(0)

    def _issue(self, slot):
        try:
            host = next(self.it)
        except StopIteration:
            return False
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self.released[slot] is not None:
                self.copy_stream.wait_event(self.released[slot])
            if self.slots[slot] is None or any(d.shape != h.shape for d, h in zip(self.slots[slot], host)):
                self.copy_stream.wait_stream(main)
                self.slots[slot] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            for d, h in zip(self.slots[slot], host):
                d.copy_(h, non_blocking=True)
            self.h2d_bytes = sum(h.numel() * h.element_size() for h in host)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.ready[slot] = ev
        return True

    def __iter__(self):
        return self

    def __next__(self):
        if not self._pending:
            raise StopIteration
        slot = self.cur
        main = torch.cuda.current_stream(self.device)
        # everything enqueued on the main stream so far used the OTHER slot: it may be refilled once that work is done
        other = slot ^ 1
        ev = torch.cuda.Event()
        ev.record(main)
        self.released[other] = ev
        self._pending = self._issue(other)
        main.wait_event(self.ready[slot])
        self.cur = other
        return self.slots[slot]
