"""A few launches of the T = 64 attention kernels (forward + backward) at the train-step shape, for ncu captures."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import attn_bench as ab
from keypoints_interpolation_transformer_b200 import _lib as K

if __name__ == "__main__":
    ab.bench(256, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, n=3)
    ab.bench_bwd(256, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, n=3)
