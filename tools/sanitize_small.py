"""Small invocations of the kernels added late in round 1, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`
(one tool per gpurun call, B200_PROFILING.md)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402
from keypoints_interpolation_transformer_b200 import baselines, model, optim, synthetic, train  # noqa: E402


def main():
    dev, bf = "cuda", torch.bfloat16
    lib, sp = K.lib(), K.stream_ptr()
    H = 256
    for M, FF in ((512, 256), (300, 128)):
        x = torch.randn(M, H, device=dev).to(bf)
        w1 = (torch.randn(FF, H, device=dev) / 16).to(bf)
        w2 = (torch.randn(H, FF, device=dev) / math.sqrt(FF)).to(bf)
        b1, b2 = torch.randn(FF, device=dev), torch.randn(H, device=dev)
        gamma, beta = torch.ones(H, device=dev), torch.zeros(H, device=dev)
        z, hh = torch.empty(M, FF, dtype=bf, device=dev), torch.empty(M, FF, dtype=bf, device=dev)
        s, y = torch.empty(M, H, dtype=bf, device=dev), torch.empty(M, H, dtype=bf, device=dev)
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        for store in (1, 0):
            K.check(lib.kit_ffn_fwd(K.ptr(x), K.ptr(w1), K.ptr(w2), K.ptr(b1), K.ptr(b2), K.ptr(gamma), K.ptr(beta), K.ptr(z),
                                    K.ptr(hh), K.ptr(s), K.ptr(y), K.ptr(mean), K.ptr(rstd), M, H, FF, store, sp))
        dz, dx = torch.empty(M, FF, dtype=bf, device=dev), torch.empty(M, H, dtype=bf, device=dev)
        K.check(lib.kit_ffn_bwd(K.ptr(x), K.ptr(w2.t().contiguous()), K.ptr(w1.t().contiguous()), K.ptr(z), K.ptr(dz), K.ptr(dx),
                                M, H, FF, sp))
        dg, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
        K.check(lib.kit_gemm_lnbwd(K.ptr(x), K.ptr(w2[:, :H].contiguous()), K.ptr(s), K.ptr(s), K.ptr(gamma), K.ptr(mean), K.ptr(rstd),
                                   K.ptr(dx), K.ptr(dg), K.ptr(db), M, H, sp))
    torch.cuda.synchronize()
    print("ffn fwd / bwd / gemm+lnbwd ok")
    inputs, gt, mask = synthetic.synthetic_batch(3, 40, 7, seed=1, smooth=True)
    baselines.cubic_interpolation(inputs.cuda(), mask.cuda())
    torch.cuda.synchronize()
    print("cubic ok")
    for env in ({}, {"KIT_FUSE_LNBWD": "1"}):       # a small train step with H = 256 (fused FFN), T = 96 (streaming attention backward)
        os.environ.pop("KIT_FUSE_LNBWD", None)
        os.environ.update(env)
        m = model.KeypointCompleter(2 * 7, H, 1, 8).cuda().train()
        step = train.TrainStep(m, optim.FlatAdam(m, lr=1e-3), criterion="mse")
        inputs, gt, mask = (t.cuda() for t in synthetic.synthetic_batch(3, 96, 7, seed=2, smooth=True))
        print("train step loss", step(inputs, gt, mask).item(), env)
    mc = model.KeypointCompleterCycle(2 * 7, 64, 1, 4).cuda().eval()
    with torch.no_grad():
        mc(inputs[:, :-1], inputs[:, 1:], frame_masks=(mask[:, :-1], mask[:, 1:]))
    torch.cuda.synchronize()
    print("all ok")


if __name__ == "__main__":
    main()
