"""Host-side policy of the reference's missing-frame generator (dataloader.py:314-436): WHICH frames
go missing is drawn on the host with the same RNG calls in the same order as the reference; the
EFFECT (hold-fill, zeroing, SOS, mask) is applied on the device by the fused pre-pass, driven by the
index map this module produces."""
import json
import math
import random as _pyrandom

import numpy as np

# dataset_config.json:2-28 (the statistics, not the file paths)
DATASET_CONFIG = {
    "AUTSL": {"mean_consecutive_missing": 5.28, "std_consecutive_missing": 4.15, "samples": 491,
              "mean_number_missing_blocks": 4.18, "std_number_missing_blocks": 1.78},
    "AEC": {"mean_consecutive_missing": 3.25, "std_consecutive_missing": 3.09, "samples": 267,
            "mean_number_missing_blocks": 1.92, "std_number_missing_blocks": 1.1},
    "PUCP_PSL_DGI305": {"mean_consecutive_missing": 4.04, "std_consecutive_missing": 5.63, "samples": 185,
                        "mean_number_missing_blocks": 1.66, "std_number_missing_blocks": 1.11},
}


def load_configuration(name):
    """utils.py:113-118: read ``<name>.json`` if it exists, else the built-in statistics."""
    try:
        with open(f"{name}.json", "r") as fh:
            return json.load(fh)
    except OSError:
        return DATASET_CONFIG


def draw_blocks(T, dataset_name, rng=_pyrandom, nprng=np.random, config=None):
    """dataloader.py:337-419 -> list of (start, end) missing blocks."""
    blocks = []
    if dataset_name == "all":                                       # :337-361
        num_blocks = rng.randint(4, 7)
        section = T // num_blocks
        rest = T % num_blocks
        for r in range(num_blocks):
            n0 = min(rng.randint(3, 8), section)
            _rest = rest if r == num_blocks - 1 else 0
            off = rng.randint(0, min(0, _rest + section - n0))
            a = section * r + off
            blocks.append((a, min(a + n0, T - 1)))
        return blocks
    cfg = (config or load_configuration("dataset_config"))[dataset_name]
    lim = [np.percentile(nprng.normal(cfg["mean_consecutive_missing"], cfg["std_consecutive_missing"],
                                      cfg["samples"]), q) for q in (25, 75)]
    size = [np.percentile(nprng.normal(cfg["mean_number_missing_blocks"], cfg["std_number_missing_blocks"],
                                       cfg["samples"]), q) for q in (25, 75)]
    nb_min, nb_max = max(math.floor(lim[0]), 1), math.ceil(lim[1])
    bs_min, bs_max = max(math.floor(size[0]), 1), math.ceil(size[1])
    num_blocks = rng.randint(nb_min, nb_max)
    section = max(1, T // num_blocks)
    rest = T % num_blocks
    if section < bs_max + 4:
        section = max(bs_max + 4, 1)
        num_blocks = max(1, T // section)
        rest = T % num_blocks
    for r in range(num_blocks):
        n0 = min(rng.randint(bs_min, bs_max), section)
        _rest = rest if r == num_blocks - 1 else 0
        off = rng.randint(0, _rest + section - n0)
        a = section * r + off
        blocks.append((a, min(a + n0, T - 1)))
    return blocks


def blocks_to_sources(T, blocks):
    """dataloader.py:421-434 as an index map (frame t <- frame src[t]) plus the 0/1 mask.  The copy
    is sequential in the reference (block 0 holds the frame AFTER it, later blocks the frame BEFORE,
    which an earlier block may already have overwritten), so indices are chased in that order."""
    src = np.arange(T, dtype=np.int32)
    mask = np.zeros(T, dtype=np.float32)
    for n, (a, b) in enumerate(blocks):
        ref = b if n == 0 else a - 1
        for t in range(a, b):
            src[t] = src[ref]
            mask[t] = 1.0
    return src, mask


def random_sources(T, rng=_pyrandom):
    """dataloader.py:320-334 (is_random_missing): 60 % draws with replacement; frames zeroed (-1)."""
    picks = rng.choices(range(T), k=int(T * (60 / 100)))
    src = np.arange(T, dtype=np.int32)
    mask = np.zeros(T, dtype=np.float32)
    for t in picks:
        src[t] = -1
        mask[t] = 1.0
    return src, mask


def draw_sources(T, is_random_missing, dataset_name, rng=_pyrandom, nprng=np.random, config=None):
    if is_random_missing:
        return random_sources(T, rng)
    return blocks_to_sources(T, draw_blocks(T, dataset_name, rng, nprng, config))


def draw_sources_device(B, T, dataset_name, seed, offset=0, device="cuda", config=None, return_blocks=False):
    """The non-random policy of ``draw_sources`` for a whole batch on the device (kit_draw_missing): the same procedure
    (quartiles of ``samples`` normals per statistic, block count / lengths / offsets, hold-fill chase) on a Philox stream --
    the reference's distribution, not its draws.  ~20 us per batch instead of ~0.8 ms per SEQUENCE on the host.
    Returns (src [B,T] int32, mask [B,T] float32[, blocks [B,64,2] int32, n_blocks [B] int32]) on ``device``."""
    import ctypes as C

    import torch

    from . import _lib as K
    cfg = (config or load_configuration("dataset_config"))[dataset_name]
    st = K.KitMissingStats(cfg["mean_consecutive_missing"], cfg["std_consecutive_missing"],
                           cfg["mean_number_missing_blocks"], cfg["std_number_missing_blocks"], int(cfg["samples"]))
    src = torch.empty(B, T, dtype=torch.int32, device=device)
    mask = torch.empty(B, T, dtype=torch.float32, device=device)
    blocks = torch.empty(B, 64, 2, dtype=torch.int32, device=device) if return_blocks else None
    nb = torch.empty(B, dtype=torch.int32, device=device) if return_blocks else None
    K.check(K.lib().kit_draw_missing(C.byref(st), B, T, int(seed), int(offset), K.ptr(src), K.ptr(mask), K.ptr(blocks),
                                     K.ptr(nb), K.stream_ptr()))
    return (src, mask, blocks, nb) if return_blocks else (src, mask)
