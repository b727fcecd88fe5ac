"""Probe of the tcgen05 operand layouts through kit_umma_probe (csrc/probe.cu): each case lays out operand tiles byte by
byte, runs one MMA sequence and compares the accumulator with a numpy product.  Prints one line per case and writes
gpurun_out/umma_probe.json.  tests/test_umma_probe_gpu.py asserts the cases the attention kernels rely on."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keypoints_interpolation_transformer_b200 import _lib as K  # noqa: E402


def idesc(M, N, a_mn=False, b_mn=False):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def bf16_bits(x):
    """float32 array -> uint16 bf16 bit patterns (values are chosen exactly representable)."""
    return (x.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def sw_tile(mat_u16, row_bytes=128):
    """[rows, row_bytes/2] uint16 -> bytes of the swizzled tile: 16-byte chunk c of row r lands at chunk c ^ f(r) of that row,
    f = r & 7 for 128-byte rows (SWIZZLE_128B), (r >> 1) & 3 for 64-byte rows (SWIZZLE_64B)."""
    rows, cols = mat_u16.shape
    assert cols * 2 == row_bytes
    nchunk = row_bytes // 16
    src = mat_u16.reshape(rows, nchunk, 8)
    out = np.zeros_like(src)
    for r in range(rows):
        f = (r & 7) if row_bytes == 128 else ((r >> 1) & 3)
        for c in range(nchunk):
            out[r, c ^ f] = src[r, c]
    return out.reshape(-1).view(np.uint8)


def small_ints(rng, shape):
    return rng.integers(-4, 5, size=shape).astype(np.float32)


def run(a_bytes, b_bytes, args, n_cols):
    dev = torch.device("cuda")
    a = torch.from_numpy(np.ascontiguousarray(a_bytes)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(b_bytes)).to(dev)
    out = torch.empty(128, n_cols, device=dev)
    arr = (C.c_int32 * 17)(*[int(v) if v < 2 ** 31 else int(v) - 2 ** 32 for v in args])
    K.check(K.lib().kit_umma_probe(K.ptr(a), a.numel(), K.ptr(b), b.numel(), arr, K.ptr(out), K.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def args(idesc_v, steps, a, b, n_cols, a_tmem=(0, 0, 0), d_lane=0):
    # a, b = (off, step, lbo, sbo, layout)
    return [idesc_v, steps, *a, *b, *a_tmem, d_lane, n_cols]


def cases():
    rng = np.random.default_rng(0)
    res = {}

    def report(name, got, want, lanes=None):
        ok = bool(np.array_equal(got, want))
        res[name] = {"ok": ok, "max_abs_diff": float(np.abs(got - want).max())}
        if lanes is not None:
            res[name]["lanes_written"] = lanes
        print(f"{name}: {'OK' if ok else 'MISMATCH'} (max |diff| {res[name]['max_abs_diff']})", flush=True)

    # 1. K-major A [128 x 64], K-major B [128 x 64], SW128: D = A B^T
    A = small_ints(rng, (128, 64)); B = small_ints(rng, (128, 64))
    got = run(sw_tile(bf16_bits(A)), sw_tile(bf16_bits(B)), args(idesc(128, 128), 4, (0, 32, 0, 1024, 2), (0, 32, 0, 1024, 2), 128), 128)
    report("kmajor_128x128x64", got, A @ B.T)

    # 2. K-major, a head's 32-column half selected by a +64 byte start offset (K = 32)
    got = run(sw_tile(bf16_bits(A)), sw_tile(bf16_bits(B)), args(idesc(128, 128), 2, (64, 32, 0, 1024, 2), (64, 32, 0, 1024, 2), 128), 128)
    report("kmajor_half_offset64", got, A[:, 32:] @ B[:, 32:].T)

    # 3. MN-major A (M = 128 = two [K rows x 64] tiles 16 KB apart), MN-major B N = 64: D[m][n] = sum_k At[k][m] Bt[k][n], K = 128
    At = small_ints(rng, (128, 128)); Bt = small_ints(rng, (128, 64))
    a_bytes = np.concatenate([sw_tile(bf16_bits(At[:, :64])), sw_tile(bf16_bits(At[:, 64:]))])
    got = run(a_bytes, sw_tile(bf16_bits(Bt)), args(idesc(128, 64, True, True), 8, (0, 2048, 16384, 1024, 2), (0, 2048, 8192, 1024, 2), 64), 64)
    report("mnmajor_A128_B64_K128", got, At.T @ Bt)

    # 4. MN-major B restricted to N = 32: first half (offset 0) and second half (start offset +64 bytes)
    for off, sl in ((0, slice(0, 32)), (64, slice(32, 64))):
        got = run(a_bytes, sw_tile(bf16_bits(Bt)), args(idesc(128, 32, True, True), 8, (0, 2048, 16384, 1024, 2), (off, 2048, 8192, 1024, 2), 32), 32)
        report(f"mnmajor_B_N32_off{off}", got, At.T @ Bt[:, sl])

    # 5. K-major A with MN-major B N = 32 at +64 (dQ = dS K_h : A = dS [q][keys], B = K tile [key][64 cols])
    A2 = small_ints(rng, (128, 128))     # [q][key], two k-blocks of 64 keys
    a2 = np.concatenate([sw_tile(bf16_bits(A2[:, :64])), sw_tile(bf16_bits(A2[:, 64:]))])
    Kt = small_ints(rng, (128, 64))      # [key][64 cols]
    for off, sl in ((0, slice(0, 32)), (64, slice(32, 64))):
        # k-step s covers keys 16 s .. 16 s + 15: A start = tile (s // 4) + (s % 4) * 32 -> two probe runs of 4 steps accumulate
        # is not expressible with one a_step; run the two k-blocks separately and add
        tot = np.zeros((128, 32), np.float32)
        for kb in range(2):
            got = run(a2[kb * 16384:(kb + 1) * 16384], sw_tile(bf16_bits(Kt))[kb * 8192:(kb + 1) * 8192],
                      args(idesc(128, 32, False, True), 4, (0, 32, 0, 1024, 2), (off, 2048, 8192, 1024, 2), 32), 32)
            tot += got
        report(f"kmajorA_mnmajorB_N32_off{off}", tot, A2 @ Kt[:, sl])

    # 6. M = 64: which lanes hold the accumulator rows
    A64 = small_ints(rng, (64, 64)); B64 = small_ints(rng, (64, 64))
    for d_lane in (0, 16):
        got = run(sw_tile(bf16_bits(A64)), sw_tile(bf16_bits(B64)), args(idesc(64, 64), 4, (0, 32, 0, 1024, 2), (0, 32, 0, 1024, 2), 64, d_lane=d_lane), 64)
        written = [int(l) for l in range(128) if not np.all(got[l] == -12345.0)]
        want = np.full((128, 64), -12345.0, np.float32)
        D = A64 @ B64.T
        for i in range(64):
            want[(i % 16) + 32 * (i // 16) + d_lane] = D[i]
        report(f"m64_lanes_dlane{d_lane}", got, want, lanes=written)

    # 7. A operand from tensor memory: row m in lane m, bf16 pairs (k = 2c, 2c + 1) in 32-bit column c
    A3 = small_ints(rng, (128, 64))
    a_words = bf16_bits(A3).reshape(128, 32, 2)
    a_u32 = (a_words[:, :, 0].astype(np.uint32) | (a_words[:, :, 1].astype(np.uint32) << 16)).astype(np.uint32)
    got = run(a_u32.reshape(-1).view(np.uint8), sw_tile(bf16_bits(B)), args(idesc(128, 128), 4, (0, 0, 0, 0, 0), (0, 32, 0, 1024, 2), 128, a_tmem=(1, 32, 8)), 128)
    report("a_from_tmem_kmajorB", got, A3 @ B.T)
    # ... with an MN-major B (O = P V: B = V tile [key][64 cols]), K = 64 keys
    Vt = small_ints(rng, (64, 64))
    got = run(a_u32.reshape(-1).view(np.uint8), sw_tile(bf16_bits(Vt)), args(idesc(128, 64, False, True), 4, (0, 0, 0, 0, 0), (0, 2048, 8192, 1024, 2), 64, a_tmem=(1, 32, 8)), 64)
    report("a_from_tmem_mnmajorB", got, A3 @ Vt)

    # 8. SWIZZLE_64B K-major tiles with 64-byte rows (one head of d = 32 per tile), K = 32
    A4 = small_ints(rng, (128, 32)); B4 = small_ints(rng, (128, 32))
    got = run(sw_tile(bf16_bits(A4), 64), sw_tile(bf16_bits(B4), 64), args(idesc(128, 128), 2, (0, 32, 0, 512, 4), (0, 32, 0, 512, 4), 128), 128)
    report("sw64_kmajor_K32", got, A4 @ B4.T)
    # ... and MN-major B N = 32 from a [K rows x 32] SW64 tile
    Bt4 = small_ints(rng, (128, 32))
    got = run(a_bytes, sw_tile(bf16_bits(Bt4), 64), args(idesc(128, 32, True, True), 8, (0, 2048, 16384, 1024, 2), (0, 1024, 0, 512, 4), 32), 32)
    report("sw64_mnmajor_B_N32", got, At.T @ Bt4)
    return res


if __name__ == "__main__":
    r = cases()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(r, open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w"), indent=1)
