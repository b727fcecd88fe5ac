"""GPU parity tests of the individual kernels, called through the C ABI (include/kit.h) and
compared with a plain fp32 PyTorch restatement of the same op on the same seeded inputs.

Tolerances: bf16 tensor-core kernels 2e-2 relative (north_star); fp32 elementwise 1e-5 relative;
masks / indices bit-exact."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from keypoints_interpolation_transformer_b200 import _lib as K

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sp():
    return K.stream_ptr()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def _bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------- GEMM (tcgen05)
@pytest.mark.parametrize("M,N,Kd", [(128, 128, 64), (256, 128, 256), (300, 200, 144), (1000, 768, 256),
                                    (16384, 256, 256), (2048, 256, 2048), (1024, 2048, 256), (515, 142, 256),
                                    # fp32 rows of 142 floats (pitch 568 B): even M takes the row-pair TMA epilogue, odd M the scalar one
                                    (512, 142, 256), (16384, 142, 256), (130, 70, 64)])
def test_gemm_tn_bias(M, N, Kd):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = _bf(torch.randn(M, Kd, generator=g)).to(DEV)
    b = _bf(torch.randn(N, Kd, generator=g) / math.sqrt(Kd)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = a.float() @ b.float().t() + bias
    # bf16 out
    if N % 8 == 0:
        c = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        K.check(K.lib().kit_gemm_bf16(0, K.ptr(a), Kd, K.ptr(b), Kd, K.ptr(c), N, M, N, Kd, K.ptr(bias), None, 0,
                                      K.OUT_BF16, K.ACT_NONE, None, 0, 1, _sp()))
        torch.cuda.synchronize()
        assert _rel(c, ref) < 5e-3
        assert (c.float() - ref).abs().max().item() < 0.05 * max(1.0, ref.abs().max().item())
    # fp32 out, unaligned leading dimension allowed
    c32 = torch.full((M, N), float("nan"), device=DEV)
    K.check(K.lib().kit_gemm_bf16(0, K.ptr(a), Kd, K.ptr(b), Kd, K.ptr(c32), N, M, N, Kd, K.ptr(bias), None, 0,
                                  K.OUT_F32, K.ACT_NONE, None, 0, 1, _sp()))
    torch.cuda.synchronize()
    assert torch.isfinite(c32).all()
    assert _rel(c32, ref) < 1e-4          # fp32 accumulate of exact bf16 products


def test_gemm_tn_epilogues():
    M, N, Kd = 640, 512, 256
    g = torch.Generator(device="cpu").manual_seed(5)
    a = _bf(torch.randn(M, Kd, generator=g)).to(DEV)
    b = _bf(torch.randn(N, Kd, generator=g) / math.sqrt(Kd)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    add = _bf(torch.randn(M, N, generator=g)).to(DEV)
    pre = a.float() @ b.float().t() + bias
    c = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    lib = K.lib()
    # bias + residual addend
    K.check(lib.kit_gemm_bf16(0, K.ptr(a), Kd, K.ptr(b), Kd, K.ptr(c), N, M, N, Kd, K.ptr(bias), K.ptr(add), N,
                              K.OUT_BF16, K.ACT_NONE, None, 0, 1, _sp()))
    torch.cuda.synchronize()
    assert _rel(c, pre + add.float()) < 5e-3
    # in-place accumulate (addend == output)
    acc = add.clone()
    K.check(lib.kit_gemm_bf16(0, K.ptr(a), Kd, K.ptr(b), Kd, K.ptr(acc), N, M, N, Kd, K.ptr(bias), K.ptr(acc), N,
                              K.OUT_BF16, K.ACT_NONE, None, 0, 1, _sp()))
    torch.cuda.synchronize()
    assert _rel(acc, pre + add.float()) < 5e-3
    # gelu (exact erf) with the pre-activation saved
    z = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    K.check(lib.kit_gemm_bf16(0, K.ptr(a), Kd, K.ptr(b), Kd, K.ptr(c), N, M, N, Kd, K.ptr(bias), None, 0,
                              K.OUT_BF16, K.ACT_GELU, K.ptr(z), N, 1, _sp()))
    torch.cuda.synchronize()
    assert _rel(z, pre) < 5e-3
    assert _rel(c, torch.nn.functional.gelu(pre)) < 5e-3
    # multiply by gelu'(z)
    zz = _bf(torch.randn(M, N, generator=g)).to(DEV)
    K.check(lib.kit_gemm_bf16(0, K.ptr(a), Kd, K.ptr(b), Kd, K.ptr(c), N, M, N, Kd, None, None, 0,
                              K.OUT_BF16, K.ACT_GELU_BWD, K.ptr(zz), N, 1, _sp()))
    torch.cuda.synchronize()
    zf = zz.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).sum().backward()
    assert _rel(c, (pre - bias) * zf.grad) < 5e-3


@pytest.mark.parametrize("Kr,M,N,ldc", [(64, 128, 128, 128), (1024, 256, 256, 256), (16384, 256, 256, 256),
                                        (4096, 2048, 256, 256), (4096, 256, 2048, 2048), (500, 142, 256, 256),
                                        (1000, 256, 142, 142), (8192, 768, 256, 256)])
def test_gemm_wgrad_accumulates(Kr, M, N, ldc):
    g = torch.Generator(device="cpu").manual_seed(Kr + M)
    lda = (M + 7) // 8 * 8
    ldb = (N + 7) // 8 * 8
    a = torch.zeros(Kr, lda, dtype=torch.bfloat16)
    b = torch.zeros(Kr, ldb, dtype=torch.bfloat16)
    a[:, :M] = _bf(torch.randn(Kr, M, generator=g))
    b[:, :N] = _bf(torch.randn(Kr, N, generator=g) / math.sqrt(Kr))
    a, b = a.to(DEV), b.to(DEV)
    c0 = torch.randn(M, ldc, generator=g).to(DEV)
    c = c0.clone()
    K.check(K.lib().kit_gemm_bf16(1, K.ptr(a), lda, K.ptr(b), ldb, K.ptr(c), ldc, M, N, Kr, None, None, 0,
                                  K.OUT_F32_ATOMIC, K.ACT_NONE, None, 0, 0, _sp()))
    torch.cuda.synchronize()
    ref = c0.clone()
    ref[:, :N] += a[:, :M].float().t() @ b[:, :N].float()
    assert _rel(c, ref) < 1e-4
    if ldc > N:
        assert torch.equal(c[:, N:], c0[:, N:])      # columns outside N untouched


def test_gelu_matches_erf_form():
    """The kernels' GELU (csrc/common.cuh: 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))), MUFU.TANH) against the reference's erf form
    (nn.Transformer(activation="gelu"), model.py:87), value and derivative, in fp32 through the generic epilogue of the GEMM
    (A = x, B = identity, so the accumulator is x exactly).  Bound = the fit (2.6e-5 / 1.1e-4, stated in kit.h) + the 2^-11
    relative error of tanh.approx; the bf16 rounding of every stored activation (2^-9 relative) is 8 x larger."""
    M, N = 4096, 64
    x = _bf(torch.linspace(-12.0, 12.0, M * N).view(M, N)).to(DEV)
    eye = _bf(torch.eye(N)).to(DEV)
    z = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    out = torch.empty(M, N, device=DEV)
    K.check(K.lib().kit_gemm_bf16(0, K.ptr(x), N, K.ptr(eye), N, K.ptr(out), N, M, N, N, None, None, 0, K.OUT_F32, K.ACT_GELU,
                                  K.ptr(z), N, 1, _sp()))
    xd = x.double()
    ref = 0.5 * xd * (1 + torch.erf(xd / math.sqrt(2.0)))
    assert torch.equal(z, x)
    err = (out.double() - ref).abs()
    assert (err <= 3e-5 + 2.6e-4 * xd.abs()).all(), float((err - 2.6e-4 * xd.abs()).max())
    assert float(err[xd.abs() < 0.5].max()) < 1e-4
    # derivative: out = acc * gelu'(aux) with acc = aux = x
    K.check(K.lib().kit_gemm_bf16(0, K.ptr(x), N, K.ptr(eye), N, K.ptr(out), N, M, N, N, None, None, 0, K.OUT_F32, K.ACT_GELU_BWD,
                                  K.ptr(x), N, 1, _sp()))
    dref = 0.5 * (1 + torch.erf(xd / math.sqrt(2.0))) + xd * torch.exp(-xd * xd / 2) / math.sqrt(2 * math.pi)
    derr = (out.double() - xd * dref).abs() / xd.abs().clamp_min(1e-3)
    # the fit contributes 1.1e-4; the rest is tanh.approx's 2^-11 entering through 1 - t^2 (times 0.5 x u'(x) <= ~2.5)
    assert float(derr.max()) < 3e-3, float(derr.max())


# ------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, bias):
    d = q.shape[-1]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if bias is not None:
        s = s + bias
    return torch.softmax(s, dim=-1) @ v


def _mask_bias(fm, flags, Sq, Sk):
    B = fm.shape[0]
    i = torch.arange(Sq, device=fm.device).view(1, Sq, 1)
    j = torch.arange(Sk, device=fm.device).view(1, 1, Sk)
    bias = torch.zeros(B, Sq, Sk, device=fm.device)
    if flags & K.MASK_REPEAT_INC:
        bias = bias.masked_fill((j > i) & (fm.view(B, 1, Sk) == 1), float("-inf"))
    if flags & K.MASK_TRIANGLE:
        bias = bias.masked_fill(j > i, float("-inf"))
    if flags & K.MASK_KEYPAD_ADD:
        bias = bias + fm.view(B, 1, Sk)
    return bias


@pytest.mark.parametrize("B,NH,S,d,flags,explicit", [
    (3, 4, 64, 32, 0, False), (2, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, False),
    (2, 8, 64, 32, K.MASK_REPEAT_INC, False), (2, 4, 12, 16, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, False),
    (2, 2, 100, 64, K.MASK_TRIANGLE, False), (1, 4, 256, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, False),
    (3, 8, 200, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, False), (2, 8, 512, 64, K.MASK_REPEAT_INC, False),
    (2, 4, 130, 32, 0, False), (2, 4, 40, 32, 0, True), (2, 4, 150, 32, 0, True),
    # attention_t64.cu (tcgen05, S <= 64, d = 32): odd batch, short sequences, every mask kind, the bench shape
    (5, 8, 40, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, False), (1, 2, 7, 32, K.MASK_REPEAT_INC, False),
    (3, 4, 64, 32, K.MASK_TRIANGLE, False), (256, 8, 64, 32, K.MASK_REPEAT_INC | K.MASK_KEYPAD_ADD, False),
    (301, 8, 64, 32, 0, False)])
def test_attention_fwd_bwd(B, NH, S, d, flags, explicit, monkeypatch):
    _attention_case(B, NH, S, d, flags, explicit)
    if S <= 64 and d == 32 and not explicit:   # short sequences take attention_t64.cu by default: the mma.sync tile kernels too
        monkeypatch.setenv("KIT_ATTN_T64", "0")
        _attention_case(B, NH, S, d, flags, explicit)
    if S > 64 and not explicit:          # long sequences take the tcgen05 forward by default: the mma.sync streaming kernel too
        monkeypatch.setenv("KIT_ATTN_TC", "0")
        _attention_case(B, NH, S, d, flags, explicit)


def _attention_case(B, NH, S, d, flags, explicit):
    H = NH * d
    g = torch.Generator(device="cpu").manual_seed(B * 100 + S)
    qkv = _bf(torch.randn(B * S, 3 * H, generator=g)).to(DEV)
    dout = _bf(torch.randn(B * S, H, generator=g)).to(DEV)
    fm = (torch.rand(B, S, generator=g) < 0.4).float().to(DEV)
    fm[:, 0] = 0
    mask = K.KitAttnMask()
    bias = None
    if flags:
        mask.frame_mask = fm.data_ptr()
        mask.frame_mask_stride = S
        mask.flags = flags
        bias = _mask_bias(fm, flags, S, S)[:, None]
    eb = None
    if explicit:
        eb = torch.randn(B, S, S, generator=g).to(DEV)
        eb[:, :, -3:] = float("-inf")
        mask.bias = eb.data_ptr()
        mask.bias_stride_b = S * S
        mask.bias_stride_h = 0
        bias = eb[:, None]
    out = torch.empty(B * S, H, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, NH, S, device=DEV)
    lib = K.lib()
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    K.check(lib.kit_attention_fwd(K.ptr(q), 3 * H, K.ptr(k), 3 * H, K.ptr(v), 3 * H, K.ptr(out), H, K.ptr(lse), B, NH, S,
                                  S, d, C.byref(mask), _sp()))
    dq = torch.empty(B * S, 3 * H, dtype=torch.bfloat16, device=DEV)
    dq_acc = torch.empty(B * S * H + B * NH * S, device=DEV) if S > 64 else None      # dQ partial sums + rowsum(dO * O)
    K.check(lib.kit_attention_bwd(K.ptr(q), 3 * H, K.ptr(k), 3 * H, K.ptr(v), 3 * H, K.ptr(out), H, K.ptr(dout), H,
                                  K.ptr(lse), K.ptr(dq[:, :H]), 3 * H, K.ptr(dq[:, H:2 * H]), 3 * H, K.ptr(dq[:, 2 * H:]),
                                  3 * H, K.ptr(dq_acc), B, NH, S, S, d, C.byref(mask), _sp()))
    torch.cuda.synchronize()

    def heads(x):
        return x.float().view(B, S, NH, d).transpose(1, 2)
    qf, kf, vf = (heads(t).detach().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qf, kf, vf, bias)
    ref.backward(heads(dout))
    got = heads(out)
    assert _rel(got, ref) < 1e-2
    assert _rel(heads(dq[:, :H]), qf.grad) < 2e-2
    assert _rel(heads(dq[:, H:2 * H]), kf.grad) < 2e-2
    assert _rel(heads(dq[:, 2 * H:]), vf.grad) < 2e-2


@pytest.mark.parametrize("B,NH,Sq,Sk,d,flags", [
    (2, 4, 130, 200, 64, K.MASK_KEYPAD_ADD), (3, 8, 256, 100, 32, 0), (2, 2, 70, 300, 64, K.MASK_KEYPAD_ADD),
    (2, 4, 300, 129, 32, K.MASK_KEYPAD_ADD)])
def test_attention_cross_lengths(B, NH, Sq, Sk, d, flags, monkeypatch):
    """Query and key sequences of different lengths (the cross-attention call shape of the decoder layers, model.py:141-145 ->
    torch/nn/modules/transformer.py:1147, when encoder and decoder inputs differ in length): the tcgen05 long-sequence forward
    (attention_tc.cu) and backward (attention_tcb.cu), then the mma.sync kernels, against fp32 torch."""
    for tc in ("1", "0"):
        monkeypatch.setenv("KIT_ATTN_TC", tc)
        H = NH * d
        g = torch.Generator(device="cpu").manual_seed(B * 1000 + Sq + Sk)
        q = _bf(torch.randn(B * Sq, H, generator=g)).to(DEV)
        kv = _bf(torch.randn(B * Sk, 2 * H, generator=g)).to(DEV)
        dout = _bf(torch.randn(B * Sq, H, generator=g)).to(DEV)
        fm = (torch.rand(B, Sk, generator=g) < 0.4).float().to(DEV)
        mask = K.KitAttnMask()
        bias = None
        if flags:
            mask.frame_mask = fm.data_ptr()
            mask.frame_mask_stride = Sk
            mask.flags = flags
            bias = _mask_bias(fm, flags, Sq, Sk)[:, None]
        out = torch.empty(B * Sq, H, dtype=torch.bfloat16, device=DEV)
        lse = torch.empty(B, NH, Sq, device=DEV)
        lib = K.lib()
        k, v = kv[:, :H], kv[:, H:]
        K.check(lib.kit_attention_fwd(K.ptr(q), H, K.ptr(k), 2 * H, K.ptr(v), 2 * H, K.ptr(out), H, K.ptr(lse), B, NH, Sq, Sk, d,
                                      C.byref(mask), _sp()))
        dq = torch.empty(B * Sq, H, dtype=torch.bfloat16, device=DEV)
        dkv = torch.empty(B * Sk, 2 * H, dtype=torch.bfloat16, device=DEV)
        dq_acc = torch.empty(B * Sq * H + B * NH * Sq, device=DEV)
        K.check(lib.kit_attention_bwd(K.ptr(q), H, K.ptr(k), 2 * H, K.ptr(v), 2 * H, K.ptr(out), H, K.ptr(dout), H, K.ptr(lse),
                                      K.ptr(dq), H, K.ptr(dkv[:, :H]), 2 * H, K.ptr(dkv[:, H:]), 2 * H, K.ptr(dq_acc), B, NH, Sq, Sk, d,
                                      C.byref(mask), _sp()))
        torch.cuda.synchronize()

        def heads(x, S):
            return x.float().view(B, S, NH, d).transpose(1, 2)
        qf = heads(q, Sq).detach().requires_grad_(True)
        kf, vf = (heads(t, Sk).detach().requires_grad_(True) for t in (k, v))
        ref = _attn_ref(qf, kf, vf, bias)
        ref.backward(heads(dout, Sq))
        assert _rel(heads(out, Sq), ref) < 1e-2, tc
        assert _rel(heads(dq, Sq), qf.grad) < 2e-2, tc
        assert _rel(heads(dkv[:, :H], Sk), kf.grad) < 2e-2, tc
        assert _rel(heads(dkv[:, H:], Sk), vf.grad) < 2e-2, tc


# ------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("M,H", [(1000, 256), (333, 64), (512, 512), (64, 1024), (16384, 256), (40000, 256)])
def test_add_layernorm_fwd_bwd(M, H):
    g = torch.Generator(device="cpu").manual_seed(M + H)
    a = _bf(torch.randn(M, H, generator=g)).to(DEV)
    b = _bf(torch.randn(M, H, generator=g)).to(DEV)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(H, generator=g)).to(DEV)
    s = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    y = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    mean = torch.empty(M, device=DEV)
    rstd = torch.empty(M, device=DEV)
    lib = K.lib()
    K.check(lib.kit_add_layernorm_fwd(K.ptr(a), K.ptr(b), K.ptr(gamma), K.ptr(beta), K.ptr(s), K.ptr(y), K.ptr(mean),
                                      K.ptr(rstd), M, H, _sp()))
    torch.cuda.synchronize()
    sref = (a.float() + b.float())
    assert _rel(s, sref) < 4e-3
    sx = s.float().detach().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    yref = torch.nn.functional.layer_norm(sx, (H,), gr, br, 1e-5)
    assert _rel(y, yref) < 4e-3
    assert torch.allclose(mean, sx.mean(-1), atol=1e-5, rtol=1e-5)
    dy = _bf(torch.randn(M, H, generator=g)).to(DEV)
    add = _bf(torch.randn(M, H, generator=g)).to(DEV)
    yref.backward(dy.float())
    dx = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    dgamma = torch.zeros(H, device=DEV)
    dbeta = torch.zeros(H, device=DEV)
    K.check(lib.kit_layernorm_bwd(K.ptr(dy), K.ptr(s), K.ptr(mean), K.ptr(rstd), K.ptr(gamma), K.ptr(add), K.ptr(dx),
                                  K.ptr(dgamma), K.ptr(dbeta), M, H, _sp()))
    torch.cuda.synchronize()
    assert _rel(dx, sx.grad + add.float()) < 5e-3
    assert _rel(dgamma, gr.grad) < 1e-4
    assert _rel(dbeta, br.grad) < 1e-4
    # without an addend (the form the engine uses; H = 256 takes ln_bwd_rows_kernel: every row of a warp in flight at once,
    # M = 40000 walks more than one chunk of rows per warp), accumulating into non-zero dgamma / dbeta
    dx2 = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    dgamma2 = torch.ones(H, device=DEV)
    dbeta2 = torch.full((H,), -2.0, device=DEV)
    K.check(lib.kit_layernorm_bwd(K.ptr(dy), K.ptr(s), K.ptr(mean), K.ptr(rstd), K.ptr(gamma), None, K.ptr(dx2),
                                  K.ptr(dgamma2), K.ptr(dbeta2), M, H, _sp()))
    torch.cuda.synchronize()
    assert _rel(dx2, sx.grad) < 5e-3
    assert _rel(dgamma2 - 1.0, gr.grad) < 1e-4
    assert _rel(dbeta2 + 2.0, br.grad) < 1e-4


# ------------------------------------------------------------------------------- fused feed-forward block
@pytest.mark.parametrize("M,FF,store", [(256, 128, 1), (512, 2048, 1), (16384, 2048, 1), (1000, 512, 0), (1000, 1024, 1),
                                        (40000, 2048, 0), (40000, 512, 1)])
def test_ffn_fused_fwd(M, FF, store):
    """kit_ffn_fwd: s = x + linear2(gelu(linear1(x))), y = LN(s), z / h saved for the backward -- against fp32 torch on the
    same bf16 operands (the hidden activation rounded to bf16 between the GEMMs, as the tensor core reads it).  Covers one
    item per CTA pair, a ragged last item, and the persistent loop (more items than CTA pairs)."""
    H = 256
    g = torch.Generator(device="cpu").manual_seed(M + FF)
    x = _bf(torch.randn(M, H, generator=g)).to(DEV)
    w1 = _bf(torch.randn(FF, H, generator=g) / math.sqrt(H)).to(DEV)
    w2 = _bf(torch.randn(H, FF, generator=g) / math.sqrt(FF)).to(DEV)
    b1, b2 = (0.1 * torch.randn(FF, generator=g)).to(DEV), (0.1 * torch.randn(H, generator=g)).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV), (0.1 * torch.randn(H, generator=g)).to(DEV)
    z = torch.full((M, FF), float("nan"), dtype=torch.bfloat16, device=DEV)
    hh = torch.full((M, FF), float("nan"), dtype=torch.bfloat16, device=DEV)
    s = torch.full((M, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    y = torch.full((M, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    K.check(K.lib().kit_ffn_fwd(K.ptr(x), K.ptr(w1), K.ptr(w2), K.ptr(b1), K.ptr(b2), K.ptr(gamma), K.ptr(beta), K.ptr(z),
                                K.ptr(hh), K.ptr(s), K.ptr(y), K.ptr(mean), K.ptr(rstd), M, H, FF, store, _sp()))
    torch.cuda.synchronize()
    z_ref = x.float() @ w1.float().t() + b1
    h_ref = torch.nn.functional.gelu(z_ref)                 # erf form (model.py:87); the kernel's tanh form is within 5e-4
    s_ref = x.float() + _bf(h_ref).float() @ w2.float().t() + b2
    sb = _bf(s_ref).float()
    y_ref = torch.nn.functional.layer_norm(sb, (H,), gamma, beta, 1e-5)
    assert torch.isfinite(s.float()).all() and torch.isfinite(y.float()).all()
    assert _rel(s, s_ref) < 5e-3
    assert _rel(y, y_ref) < 1e-2
    assert (mean - sb.mean(1)).abs().max().item() < 2e-2
    assert _rel(rstd, (sb.var(1, unbiased=False) + 1e-5).rsqrt()) < 1e-2
    if store:
        assert _rel(z, z_ref) < 5e-3
        assert _rel(hh, h_ref) < 5e-3
        assert (hh.float() - h_ref).abs().max().item() < 0.03 * max(1.0, h_ref.abs().max().item())
    else:
        assert torch.isnan(z.float()).all() and torch.isnan(hh.float()).all()     # inference writes neither


@pytest.mark.parametrize("M,FF", [(256, 128), (512, 2048), (16384, 2048), (1000, 512), (40000, 1024)])
def test_ffn_fused_bwd(M, FF):
    """kit_ffn_bwd: dz = (g W2) * gelu'(z), dx = dz W1 + g against fp32 torch on the same bf16 operands (dz rounded to bf16
    between the GEMMs, as the tensor core reads it and as the weight gradients read it)."""
    H = 256
    gen = torch.Generator(device="cpu").manual_seed(M * 3 + FF)
    g = _bf(torch.randn(M, H, generator=gen)).to(DEV)
    w1 = _bf(torch.randn(FF, H, generator=gen) / math.sqrt(H)).to(DEV)
    w2 = _bf(torch.randn(H, FF, generator=gen) / math.sqrt(FF)).to(DEV)
    z = _bf(torch.randn(M, FF, generator=gen)).to(DEV)
    w2t, w1t = w2.t().contiguous(), w1.t().contiguous()
    dz = torch.full((M, FF), float("nan"), dtype=torch.bfloat16, device=DEV)
    dx = torch.full((M, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    K.check(K.lib().kit_ffn_bwd(K.ptr(g), K.ptr(w2t), K.ptr(w1t), K.ptr(z), K.ptr(dz), K.ptr(dx), M, H, FF, _sp()))
    torch.cuda.synchronize()
    zf = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).sum().backward()           # gelu'(z), erf form
    dz_ref = (g.float() @ w2.float()) * zf.grad
    dx_ref = _bf(dz_ref).float() @ w1.float() + g.float()
    assert torch.isfinite(dz.float()).all() and torch.isfinite(dx.float()).all()
    assert _rel(dz, dz_ref) < 5e-3
    assert _rel(dx, dx_ref) < 5e-3


@pytest.mark.parametrize("M,Kd", [(128, 256), (16384, 768), (1000, 2048), (18944, 256)])
def test_gemm_layernorm_backward_epilogue(M, Kd):
    """EPI_ADD_LNBWD: dy = A B^T + addend stays on chip and the LayerNorm backward (row statistics across the four column-group
    warps, dgamma / dbeta partials by warp transpose-reduce) runs in the epilogue -- against fp32 torch autograd."""
    H = 256
    g = torch.Generator(device="cpu").manual_seed(M + Kd)
    a = _bf(torch.randn(M, Kd, generator=g)).to(DEV)
    b = _bf(torch.randn(H, Kd, generator=g) / math.sqrt(Kd)).to(DEV)
    add = _bf(torch.randn(M, H, generator=g)).to(DEV)
    s = _bf(torch.randn(M, H, generator=g) * 2 + 0.3).to(DEV)
    gamma = (1 + 0.2 * torch.randn(H, generator=g)).to(DEV)
    sf = s.float()
    mean, var = sf.mean(1), sf.var(1, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    dx = torch.full((M, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    dgamma, dbeta = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    K.check(K.lib().kit_gemm_lnbwd(K.ptr(a), K.ptr(b), K.ptr(add), K.ptr(s), K.ptr(gamma), K.ptr(mean.contiguous()),
                                   K.ptr(rstd.contiguous()), K.ptr(dx), K.ptr(dgamma), K.ptr(dbeta), M, Kd, _sp()))
    torch.cuda.synchronize()
    dy = _bf(a.float() @ b.float().t() + add.float()).float()           # the kernel rounds dy to bf16 like the unfused path
    sr = sf.clone().requires_grad_(True)
    gp = gamma.clone().requires_grad_(True)
    bp = torch.zeros(H, device=DEV, requires_grad=True)
    torch.nn.functional.layer_norm(sr, (H,), gp, bp, 1e-5).backward(dy)
    assert torch.isfinite(dx.float()).all()
    assert _rel(dx, sr.grad) < 5e-3
    assert _rel(dgamma, gp.grad) < 2e-3
    assert _rel(dbeta, bp.grad) < 2e-3
