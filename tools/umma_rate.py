"""Issue / completion rate of tcgen05.mma (kind::f16, one CTA) by shape, measured with clock64 inside one thread
(csrc/probe.cu umma_rate_kernel): cycles per instruction for `reps` back-to-back MMAs."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402
from umma_probe import idesc  # noqa: E402

if __name__ == "__main__":
    fn = K.lib().kit_umma_rate
    fn.restype = C.c_int
    fn.argtypes = [C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    reps = 64
    print(f"{'shape':34s} {'issue cyc/MMA':>14s} {'complete cyc/MMA':>17s}")
    for (M, N, a_mn, b_mn, a_tmem) in [(64, 64, 0, 0, 0), (64, 32, 0, 1, 0), (64, 32, 1, 1, 0), (64, 64, 0, 0, 1), (64, 32, 0, 1, 1),
                                        (128, 64, 0, 0, 0), (128, 128, 0, 0, 0), (128, 32, 0, 1, 0), (128, 32, 1, 1, 0), (128, 64, 1, 1, 0),
                                        (128, 128, 0, 0, 1), (128, 32, 0, 1, 1), (128, 256, 0, 0, 0), (128, 192, 0, 0, 0),
                                        (128, 256, 0, 0, 1), (128, 64, 0, 0, 1), (128, 192, 0, 0, 1)]:
        for _ in range(2):
            K.check(fn(idesc(M, N, bool(a_mn), bool(b_mn)), reps, a_mn, b_mn, a_tmem, out.data_ptr(), None))
            torch.cuda.synchronize()
        o = out.cpu().tolist()
        name = f"M={M} N={N} A={'tmem' if a_tmem else ('mn' if a_mn else 'k')}-major B={'mn' if b_mn else 'k'}-major"
        print(f"{name:34s} {o[0] / reps:14.1f} {o[1] / reps:17.1f}")
