"""Times the in-step pre-processing pieces at the train-step shape (B = 256, T = 64, K = 71): kit_draw_policy and kit_prepass."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import preprocess as PP  # noqa: E402


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


if __name__ == "__main__":
    dev = "cuda"
    B, T, K = 256, 64, 71
    raw = torch.rand(B, T, K, 2, device=dev)
    pp = PP.Prepass(K, dev, list(range(K)), list(range(29, K)), 5, 6, 2, [[0, 5, 7, 9], [0, 6, 8, 10]])
    for aug in (True, False):
        pol = PP.DevicePolicy("AUTSL", seed=1, have_augmentation=aug, device=dev)
        print(f"kit_draw_policy (augmentation={aug}): {t(lambda: pol.draw(B, T)):.1f} us")
    pol = PP.DevicePolicy("AUTSL", seed=1, device=dev)
    src, miss, aug = pol.draw(B, T)
    y = torch.empty(B, T, K, 2, device=dev)
    mask = torch.empty(B, T + 1, device=dev)
    xe = torch.empty(B * T, 144, dtype=torch.bfloat16, device=dev)
    xd = torch.empty(B * T, 144, dtype=torch.bfloat16, device=dev)
    import ctypes as C
    run = lambda a: pp.run(raw, src, miss, a, y, mask, C.c_void_p(xe.data_ptr()), C.c_void_p(xd.data_ptr()), 144, normalize=True)
    print(f"kit_prepass with device-drawn augmentation: {t(lambda: run(aug)):.1f} us")
    print(f"kit_prepass without augmentation:           {t(lambda: run(None)):.1f} us")
