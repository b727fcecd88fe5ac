"""Clock trace of CTA 0 of attn64_fwd_kernel (KIT_A64_TRACE=1): cycles from kernel entry for the producer warp, the MMA thread and
the first softmax warp, per unit.  usage: KIT_A64_TRACE=1 python tools/attn_trace.py"""
import ctypes as C
import os
import sys

os.environ["KIT_A64_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import attn_bench as ab  # noqa: E402
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402

if __name__ == "__main__":
    ab.bench(256, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, n=5)
    buf = (C.c_longlong * 256)()
    fn = K.lib().kit_a64_trace_read
    fn.restype = C.c_int
    assert fn(buf) == 0
    t0 = buf[0]
    names = {0: "entry", 1: "setup done (after __syncthreads)", 2: "pdl_wait passed", 120: "softmax warp done", 121: "exit"}
    for i in range(4):
        names[8 + i] = f"producer: unit {i} slot free, TMA issued"
        for g in range(2):
            names[32 + 2 * i + g] = f"mma: S(unit {i}, head {g}) issued"
            names[48 + 2 * i + g] = f"mma: PV(unit {i}, head {g}) issued"
        names[64 + 8 * i] = f"softmax: unit {i} s_full passed"
        names[65 + 8 * i] = f"softmax: unit {i} scores loaded, s_free"
        names[66 + 8 * i] = f"softmax: unit {i} exponentials done"
        names[67 + 8 * i] = f"softmax: unit {i} P stored, p_full"
        names[68 + 8 * i] = f"softmax: unit {i} previous output drained"
    rows = sorted((buf[i] - t0, names.get(i, str(i))) for i in range(256) if buf[i] != 0)
    for c, n in rows:
        print(f"{c:8d}  {n}")
