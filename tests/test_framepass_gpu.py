"""GPU parity of the HBM-bound per-frame passes (pre-pass, loss, get_mask, Adam) against the
reference-generated golden fixtures and the CPU oracle.  fp32 elementwise tolerance 1e-5 relative;
masks, indices, zero patterns and hold-fill copies bit-exact."""
import math
import os
import random

import numpy as np
import pytest
import torch

from keypoints_interpolation_transformer_b200 import _lib as K
from keypoints_interpolation_transformer_b200 import augmentation, dataloader, euclidean_loss, model, optim
from keypoints_interpolation_transformer_b200 import preprocess as PP
from oracle import kit_oracle as ko

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def test_normalize_pose_matches_reference(golden_dir):
    g = _load(golden_dir, "normalize_pose")
    ls, rs, re = (int(v) for v in g["ids"])
    body = {"pose_left_shoulder": ls, "pose_right_shoulder": rs, "pose_right_eye": re}
    for n in range(3):
        data = g[f"in{n}"].copy()
        out = dataloader.normalize_pose(data, body)
        assert out is data                                    # in place, like the reference
        ref = g[f"out{n}"]
        assert np.array_equal(out == 0, ref == 0)
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-7)


def test_augmentations_match_reference(golden_dir):
    g = _load(golden_dir, "augmentation")
    ids = {"pose": list(g["pose"]), "left_hand": list(g["left_hand"]), "rigth_hand": list(g["right_hand"])}
    body = {"pose_chest_middle_up": 0, "pose_left_shoulder": 5, "pose_left_elbow": 7, "pose_left_wrist": 9,
            "pose_right_shoulder": 6, "pose_right_elbow": 8, "pose_right_wrist": 10}
    aug = augmentation.augmentation(ids, body)
    assert np.array_equal(np.array(aug.ARM_IDENTIFIERS_ORDER), g["arm_chains"])
    base = g["base"]
    for n in range(3):
        random.seed(100 + n)
        s = torch.from_numpy(base.copy())
        out = aug.augment_rotate(s, (-15, 15))
        assert out is s
        np.testing.assert_allclose(out.numpy(), g[f"rotate{n}"], rtol=1e-5, atol=1e-6)
        random.seed(200 + n)
        out = aug.augment_shear(torch.from_numpy(base.copy()), "squeeze", (-0.15, 0.15)).numpy()
        np.testing.assert_allclose(out, g[f"squeeze{n}"], rtol=1e-5, atol=1e-6)
        assert np.array_equal(out == 0, g[f"squeeze{n}"] == 0)
        random.seed(300 + n)
        out = aug.augment_shear(torch.from_numpy(base.copy()), "perspective", (-0.15, 0.15)).numpy()
        np.testing.assert_allclose(out, g[f"persp{n}"], rtol=1e-5, atol=1e-6)
        assert np.array_equal(out == 0, g[f"persp{n}"] == 0)
        random.seed(400 + n)
        out = aug.augment_arm_joint_rotate(torch.from_numpy(base.copy()), 0.5, (-15, 15)).numpy()
        np.testing.assert_allclose(out, g[f"arm{n}"], rtol=1e-5, atol=1e-6)


def test_put_missing_frames_and_sos_bit_exact(golden_dir):
    g = _load(golden_dir, "missing_frames")
    for n in range(int(g["count"])):
        T, seed = (int(v) for v in g[f"meta{n}"])
        random.seed(seed)
        np.random.seed(seed)
        video = torch.arange(T, dtype=torch.float32).view(T, 1, 1).repeat(1, 3, 2).clone()
        v2, mask = dataloader.put_missing_frames(video, False, str(g[f"ds{n}"]))
        assert v2 is video
        assert np.array_equal(v2[:, 0, 0].numpy().astype(np.int32), g[f"src{n}"])
        assert np.array_equal(mask.cpu().numpy(), g[f"mask{n}"])
        v3, m3 = dataloader.add_sos(v2, mask.cpu())
        assert np.array_equal(v3.numpy(), g[f"sos_video{n}"])
        assert np.array_equal(m3.numpy(), g[f"sos_mask{n}"])
    random.seed(77)
    video = (1 + torch.arange(30, dtype=torch.float32)).view(30, 1, 1).repeat(1, 2, 2).clone()
    v2, mask = dataloader.put_missing_frames(video, True, "AUTSL")
    assert np.array_equal(v2.numpy(), g["rand_video"])
    assert np.array_equal(mask.cpu().numpy(), g["rand_mask"])


@pytest.mark.parametrize("B,T,Kp,zero_masked", [(5, 64, 71, False), (3, 33, 54, True), (2, 256, 71, True)])
def test_fused_prepass_slices_and_bf16_operands(B, T, Kp, zero_masked):
    """The whole pre-pass against the oracle: hold-fill + SOS + A1 slices (+ A4 zeroing), bit exact."""
    rs = np.random.RandomState(B * T)
    raw = rs.uniform(0.05, 0.95, size=(B, T, Kp, 2)).astype(np.float32)
    raw[rs.uniform(size=(B, T, Kp)) < 0.03] = 0.0
    pr = random.Random(5)
    src = np.empty((B, T), np.int32)
    msk = np.empty((B, T), np.float32)
    for b in range(B):
        blocks = ko.missing_blocks_from_config(T, ko.AUTSL_CONFIG, rng=pr, nprng=rs)
        src[b], msk[b] = ko.hold_fill_sources(T, blocks)
    pp = PP.Prepass(Kp, DEV)
    res = pp(torch.from_numpy(raw), torch.from_numpy(src), torch.from_numpy(msk), zero_masked_enc=zero_masked, want_bf16=True)
    torch.cuda.synchronize()
    assert torch.equal(res["y"].cpu(), torch.from_numpy(raw))
    k2p = pp.k2p
    for b in range(B):
        ref_in, ref_mask = ko.apply_sources_add_sos(raw[b], src[b], msk[b])
        assert np.array_equal(res["inputs"][b].cpu().numpy(), ref_in)
        assert np.array_equal(res["mask"][b].cpu().numpy(), ref_mask)
        x, xf, _, xm, _ = ko.step_slices(torch.from_numpy(ref_in)[None], torch.from_numpy(raw[b])[None],
                                         torch.from_numpy(ref_mask)[None], zero_masked=zero_masked)
        xe = res["x_enc"].view(B, T, k2p)[b].float().cpu()
        xd = res["x_dec"].view(B, T, k2p)[b].float().cpu()
        assert torch.equal(xe[:, :2 * Kp], x[0].reshape(T, 2 * Kp).to(torch.bfloat16).float())
        assert torch.equal(xd[:, :2 * Kp], xf[0].reshape(T, 2 * Kp).to(torch.bfloat16).float())
        assert (xe[:, 2 * Kp:] == 0).all() and (xd[:, 2 * Kp:] == 0).all()


def test_get_mask_bit_exact(golden_dir):
    g = _load(golden_dir, "get_mask")
    m = model.KeypointCompleter(108, 64, 1, 4)
    for n, T in enumerate([1, 2, 7, 16, 33]):
        fm = torch.from_numpy(g[f"mask{n}"])
        for typ in ["triangle", "repeat", "repeat-inc", "all"]:
            key = f"{typ}{n}"
            if key not in g.files:
                continue
            for dev in ("cpu", DEV):
                got = m.get_mask(fm.to(dev), T, typ).cpu().numpy()
                assert got.shape == g[key].shape and np.array_equal(got, g[key]), (key, dev)


def test_loss_matches_reference_and_is_deterministic(golden_dir):
    g = _load(golden_dir, "loss")
    o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g["t"]).to(DEV)
    loss = euclidean_loss.EuclideanLoss()(o, t)
    assert loss.dim() == 0
    assert abs(loss.item() - float(g["euclid"])) <= 1e-5 * float(g["euclid"])
    (loss * 3.0).backward()
    ref = o.detach().clone().requires_grad_(True)
    (ko.euclidean_loss(ref, t) * 3.0).backward()
    assert torch.allclose(o.grad, ref.grad, rtol=1e-5, atol=1e-9)
    mse = euclidean_loss.MSELoss()(o.detach(), t)
    assert abs(mse.item() - float(g["mse"])) <= 1e-5 * float(g["mse"])
    again = euclidean_loss.EuclideanLoss()(o.detach(), t)
    assert again.item() == loss.item()                       # fixed reduction order
    # A1 usage: .float(), .clone().detach().cpu().numpy()
    assert np.isfinite(loss.float().clone().detach().cpu().numpy())
    # eval blend (A1_train.py:184-186)
    fm = (torch.rand(3, 17, device=DEV) < 0.4).float()
    got = euclidean_loss.MaskedEuclideanLoss()(o.detach(), t, fm)
    want = ko.euclidean_loss(ko.eval_blend(o.detach().cpu(), t.cpu(), fm.cpu()), t.cpu())
    assert abs(got.item() - want.item()) <= 1e-5 * want.item()
    d = euclidean_loss.EuclideanDistanceLoss()(o.detach()[0], t[0])
    assert abs(d.item() - float(g["euclid_dist"])) <= 1e-5 * float(g["euclid_dist"])


def test_loss_at_baseline_inference_size_properties():
    """BASELINE config 4 size (B=4096, T=256, K=71): size-independent properties."""
    B, T, Kp = 4096, 256, 71
    gen = torch.Generator(device=DEV).manual_seed(0)
    p = torch.rand(B, T, Kp, 2, device=DEV, generator=gen)
    t = torch.rand(B, T, Kp, 2, device=DEV, generator=gen)
    l1, d1 = euclidean_loss.fused_loss(p, t)
    l0, _ = euclidean_loss.fused_loss(p, p)
    assert l0.item() == 0.0
    ref = ((p - t) ** 2).sum(-1).mean()
    assert abs(l1.item() - ref.item()) <= 1e-5 * ref.item()
    assert torch.allclose(d1, 2 * (p - t) / (B * T * Kp), rtol=1e-5, atol=1e-12)
    fm = (torch.rand(B, T, device=DEV, generator=gen) < 0.3).float()
    lm, dm = euclidean_loss.fused_loss(p, t, fm)
    lc, _ = euclidean_loss.fused_loss(p, t, 1 - fm)
    assert abs(lm.item() + lc.item() - l1.item()) <= 1e-5 * l1.item()        # masked + complement = total
    assert (dm[fm == 0] == 0).all()


def test_flat_adam_matches_torch_adam():
    n = 4096 * 3
    gen = torch.Generator(device="cpu").manual_seed(1)
    p0 = torch.randn(n, generator=gen)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=3e-4)
    p = p0.clone().to(DEV)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    for step in range(1, 6):
        gsrc = torch.randn(n, generator=gen)
        ref.grad = gsrc.clone()
        opt.step()
        K.check(K.lib().kit_adam_step(K.ptr(p), K.ptr(gsrc.to(DEV)), K.ptr(m), K.ptr(v), n, 3e-4, 0.9, 0.999, 1e-8, step,
                                      1.0, K.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-5, atol=1e-7)


def test_device_prefetcher_yields_batches_in_order():
    """dataloader.DevicePrefetcher: batches arrive on the device unchanged and in order while two slots are recycled."""
    from keypoints_interpolation_transformer_b200 import dataloader
    host = [tuple(t.pin_memory() for t in (torch.full((4, 9, 5, 2), float(i)), torch.full((4, 8, 5, 2), float(-i)),
                                           torch.full((4, 9), float(i % 2)))) for i in range(7)]
    seen = []
    for inputs, sota, mask in dataloader.DevicePrefetcher(host, "cuda"):
        assert inputs.is_cuda and sota.is_cuda and mask.is_cuda
        seen.append((inputs.sum().item(), sota.sum().item(), mask.sum().item()))   # consumed before the slot is refilled
    want = [(h[0].sum().item(), h[1].sum().item(), h[2].sum().item()) for h in host]
    assert seen == want


def test_device_missing_block_policy_structure_and_statistics():
    """kit_draw_missing (dataloader.py:364-434 on a Philox stream): the index map / mask are EXACTLY what the reference's
    sequential hold-fill makes of the drawn blocks, the blocks obey the policy's constraints, draws are reproducible, and
    block count / masked-frame statistics agree with the host policy (reference RNG) within sampling error."""
    from keypoints_interpolation_transformer_b200 import missing
    B, T = 1024, 64
    src, mask, blocks, nb = missing.draw_sources_device(B, T, "AUTSL", seed=7, device=DEV, config=missing.DATASET_CONFIG,
                                                        return_blocks=True)
    src2, mask2 = missing.draw_sources_device(B, T, "AUTSL", seed=7, device=DEV, config=missing.DATASET_CONFIG)
    assert torch.equal(src, src2) and torch.equal(mask, mask2)                    # same (seed, offset) -> same draws
    src3, _ = missing.draw_sources_device(B, T, "AUTSL", seed=7, offset=1, device=DEV, config=missing.DATASET_CONFIG)
    assert not torch.equal(src, src3)
    src, mask, blocks, nb = src.cpu().numpy(), mask.cpu().numpy(), blocks.cpu().numpy(), nb.cpu().numpy()
    for b in range(B):
        bl = [(int(a), int(e)) for a, e in blocks[b, :nb[b]]]
        assert all(0 <= a < T and a <= e <= T - 1 for a, e in bl)
        assert all(bl[i][0] >= bl[i - 1][0] for i in range(1, len(bl)))         # one block per section, in order
        ref_src, ref_mask = missing.blocks_to_sources(T, bl)
        assert np.array_equal(src[b], ref_src) and np.array_equal(mask[b], ref_mask)
    pr, rs = random.Random(3), np.random.RandomState(3)
    host = [missing.draw_blocks(T, "AUTSL", pr, rs, missing.DATASET_CONFIG) for _ in range(400)]
    host_nb = np.array([len(h) for h in host], dtype=np.float64)
    host_masked = np.array([missing.blocks_to_sources(T, h)[1].sum() for h in host])
    assert abs(nb.mean() - host_nb.mean()) < 0.35, (nb.mean(), host_nb.mean())
    assert abs(mask.sum(1).mean() - host_masked.mean()) < 2.0, (mask.sum(1).mean(), host_masked.mean())
    assert set(np.unique(nb)) <= set(range(1, 65)) and nb.min() >= host_nb.min() - 1 and nb.max() <= host_nb.max() + 1


def test_keypoint_batcher_device_policy_feeds_the_step():
    from keypoints_interpolation_transformer_b200 import dataloader as dl
    raw = torch.rand(8, 32, 12, 2)
    bt = dl.KeypointBatcher(raw, have_augmentation=False, normalize=False, seed=5, device_policy=True)
    inputs, sota, mask = bt.batch(list(range(8)))
    assert inputs.shape == (8, 33, 12, 2) and sota.shape == (8, 32, 12, 2) and mask.shape == (8, 33)
    assert torch.equal(sota.cpu(), raw)
    m = mask[:, 1:].bool().cpu()
    assert m.any() and (inputs[:, 0] == 1).all() and (mask[:, 0] == 0).all()
    assert torch.equal(inputs[:, 1:].cpu()[~m], raw[~m])                         # untouched frames pass through


def test_keypoint_batcher_with_augmentation_follows_the_reference_dispatch():
    """KeypointBatcher with augmentation ON against LSP_Dataset.__getitem__'s sequence of calls (dataloader.py:649-675) made
    with the drop-in functions that are themselves pinned to the reference's outputs (test_augmentations_match_reference,
    test_missing_frames_and_sos_match_reference): same global RNG streams, same per-sequence order (augmentation draws, then
    the missing-block draws), so a seeded batch equals the reference's DataLoader over the same indices."""
    import random
    from keypoints_interpolation_transformer_b200 import augmentation
    from keypoints_interpolation_transformer_b200 import dataloader as dl
    Kp, T, N = 21, 40, 12
    ids = {"pose": list(range(0, 13)), "left_hand": list(range(13, 17)), "rigth_hand": list(range(17, 21))}
    body = {"pose_chest_middle_up": 0, "pose_left_shoulder": 1, "pose_left_elbow": 2, "pose_left_wrist": 3,
            "pose_right_shoulder": 4, "pose_right_elbow": 5, "pose_right_wrist": 6, "pose_right_eye": 7}
    g = torch.Generator().manual_seed(11)
    raw = torch.rand(N, T, Kp, 2, generator=g) * 0.8 + 0.1
    raw[torch.rand(N, T, Kp, generator=g) < 0.03] = 0.0
    bt = dl.KeypointBatcher(raw, ids, body, dataset_name="AUTSL", normalize=False, have_augmentation=True, augmentations_prob=0.5)
    random.seed(123)
    np.random.seed(123)
    inputs, sota, mask = bt.batch(list(range(N)))
    # the reference's __getitem__ over the same indices with the same global streams
    random.seed(123)
    np.random.seed(123)
    aug = augmentation.augmentation(ids, body)
    kinds = set()
    for b in range(N):
        x = raw[b].clone()
        if random.random() < 0.5:
            sel = random.randrange(4)
            kinds.add(sel)
            if sel == 0:
                aug.augment_rotate(x, angle_range=(-15, 15))
            if sel == 1:
                aug.augment_shear(x, "perspective", squeeze_ratio=(-0.15, 0.15))
            if sel == 2:
                aug.augment_shear(x, "squeeze", squeeze_ratio=(-0.15, 0.15))
            if sel == 3:
                aug.augment_arm_joint_rotate(x, 0.5, angle_range=(-15, 15))
        miss, m = dl.put_missing_frames(x.clone().detach(), False, "AUTSL")
        miss, m = dl.add_sos(miss, m)
        assert (sota[b].cpu() - x).abs().max().item() <= 1e-6, b
        assert torch.equal(mask[b].cpu(), m.cpu()), b                                  # bit exact
        assert (inputs[b].cpu() - miss.cpu()).abs().max().item() <= 1e-6, b
        assert torch.equal(sota[b].cpu() == 0, x == 0), b                              # zero pattern exact
    assert len(kinds) >= 3      # the seed exercises the dispatch


def test_device_policy_draws_the_reference_distribution():
    """kit_draw_policy: the augmentation records written on the device equal what the host path builds from the SAME scalar
    draws (preprocess.aug_* = augmentation.py:132,166-187,221-224 incl. the float64 homography), the selection frequencies are
    those of dataloader.py:649-651 (p = 0.5, then uniform over 4), parameters stay in their ranges, the device counter
    advances by one per call and consecutive calls draw different values."""
    import math
    from keypoints_interpolation_transformer_b200 import preprocess as PP
    B, T = 4096, 64
    pol = PP.DevicePolicy("AUTSL", seed=11, have_augmentation=True, augmentations_prob=0.5, has_arms=True, device=DEV)
    src, miss, aug, draws = pol.draw(B, T, want_draws=True)
    torch.cuda.synchronize()
    assert int(pol.counter.item()) == 1
    rec = np.frombuffer(aug.cpu().numpy().tobytes(), dtype=np.dtype(K.KitSeqAug))
    d = draws.cpu().numpy()
    sel = d[:, 0].astype(int)
    freq = [(sel == k).mean() for k in (-1, 0, 1, 2, 3)]
    assert abs(freq[0] - 0.5) < 0.03 and all(abs(f - 0.125) < 0.02 for f in freq[1:]), freq
    src_c = np.array(((0, 1), (1, 1), (0, 0), (1, 0)), dtype=np.float32)
    checked = {0: 0, 1: 0, 2: 0, 3: 0}
    for b in range(0, B, 7):
        r, k = rec[b], sel[b]
        if k == -1:
            assert r["kind"] == K.AUG_NONE
            continue
        if k == 0:
            assert abs(d[b, 1]) <= math.radians(15) and r["kind"] == K.AUG_ROTATE
            ref = PP.aug_rotate(d[b, 1])
            assert abs(r["cos_t"] - ref.cos_t) <= 1e-7 and abs(r["sin_t"] - ref.sin_t) <= 1e-7
        elif k in (1, 2):
            assert r["kind"] == K.AUG_SHEAR
            if k == 1:
                a = d[b, 1]
                assert abs(a) <= 0.15
                dst = (np.array(((0 + a, 1 - a), (1, 1), (0 + a, 0 + a), (1, 0)), dtype=np.float32) if d[b, 3] == 1.0 else
                       np.array(((0, 1), (1 - a, 1 - a), (0, 0), (1 - a, 0 + a)), dtype=np.float32))
            else:
                ml, mr = d[b, 1], d[b, 2]
                assert abs(ml) <= 0.15 and abs(mr) <= 0.15
                dst = np.array(((0 + ml, 1), (1 - mr, 1), (0 + ml, 0), (1 - mr, 0)), dtype=np.float32)
            ref = PP.aug_shear(PP.perspective_matrix(src_c, dst))
            assert np.allclose(np.array(r["mtx"]), np.array(list(ref.mtx)), rtol=0, atol=1e-12)
            assert r["zero_x"] == np.float32(ref.zero_x) and r["zero_y"] == np.float32(ref.zero_y)
        else:
            assert r["kind"] == K.AUG_ARM_ROTATE
            ang = [[None if math.isnan(d[b, 4 + c * 4 + j]) else d[b, 4 + c * 4 + j] for j in range(4)] for c in range(2)]
            ref = PP.aug_arm(ang)
            assert np.allclose(np.array(r["arm_cos"]), np.array(list(ref.arm_cos)), atol=1e-7)
            assert np.allclose(np.array(r["arm_sin"]), np.array(list(ref.arm_sin)), atol=1e-7)
        checked[k] += 1
    assert all(v > 10 for v in checked.values()), checked
    arm = d[sel == 3, 4:]
    assert abs(np.isnan(arm).mean() - 0.5) < 0.05 and np.nanmax(np.abs(arm)) <= math.radians(15)
    # missing blocks: same kernel as kit_draw_missing (statistics checked in test_device_missing_policy...)
    m = miss.cpu().numpy()
    assert set(np.unique(m)) <= {0.0, 1.0} and 0.2 < m.mean() < 0.7 and (src.cpu().numpy() < T).all()
    src2, miss2, aug2, draws2 = pol.draw(B, T, want_draws=True)
    torch.cuda.synchronize()
    assert int(pol.counter.item()) == 2
    assert not np.array_equal(np.nan_to_num(draws2.cpu().numpy()), np.nan_to_num(d))


def test_raw_train_step_equals_prepass_plus_train_step():
    """train.RawTrainStep (policy + pre-pass writing the engine's bf16 operands + step, A1_train.py:89-135 from raw keypoints)
    against the same policy draws pushed through the stand-alone pre-pass and the batch-layout TrainStep: same loss and
    gradients (the operands are the same bf16 roundings of the same fp32 values); and as a CUDA graph every replay draws a
    fresh policy (device counter)."""
    from keypoints_interpolation_transformer_b200 import model, optim, train
    from keypoints_interpolation_transformer_b200 import preprocess as PP
    Kp, H, L, NH, B, T = 21, 64, 2, 2, 6, 40
    body, hand = list(range(0, 21)), list(range(13, 21))
    chains = [[0, 1, 2, 3], [0, 4, 5, 6]]
    torch.manual_seed(3)
    m = model.KeypointCompleter(2 * Kp, H, L, NH).to(DEV)
    m.train()
    g = torch.Generator().manual_seed(5)
    raw = (torch.rand(B, T, Kp, 2, generator=g) * 0.8 + 0.1).to(DEV)
    raw[:, :, 9] = 0.0
    pp = PP.Prepass(Kp, DEV, body, hand, 1, 4, 7, chains)
    for zero_masked in (False, True):
        pol = PP.DevicePolicy("AUTSL", seed=77, device=DEV)
        step = train.RawTrainStep(m, pp, pol, optim.FlatAdam(m, lr=0.0), criterion="mse", normalize=True, zero_masked=zero_masked)
        loss = step.forward_backward(raw)
        torch.cuda.synchronize()
        got = m.flat_grads.clone()
        pol.counter.zero_()                                       # the same draws again
        src, miss, aug = pol.draw(B, T)
        res = pp(raw, src, miss, normalize=True, aug_dev=aug)
        ref_step = train.TrainStep(m, optim.FlatAdam(m, lr=0.0), criterion="mse", zero_masked=zero_masked)
        ref_loss = ref_step.forward_backward(res["inputs"], res["y"], res["mask"])
        torch.cuda.synchronize()
        assert torch.equal(step.y, res["y"]) and torch.equal(step.mask, res["mask"])
        assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
        assert ((got - m.flat_grads).norm() / m.flat_grads.norm()).item() < 1e-3
    pol = PP.DevicePolicy("AUTSL", seed=5, device=DEV)
    gstep = train.RawTrainStep(m, pp, pol, optim.FlatAdam(m, lr=1e-4, capturable=True), criterion="mse", use_graph=True)
    losses = [gstep(raw).item() for _ in range(6)]
    assert gstep.use_graph and len(gstep._graphs) == 1
    assert int(pol.counter.item()) == 6 and len(set(round(v, 7) for v in losses[2:])) > 1     # fresh draws at every replay


def test_cubic_interpolation_kernel(golden_dir):
    """kit_cubic_interpolate (the evaluation's cubic-spline baseline, 3_test_cubic_interpolation.py:32-58) against outputs of
    the reference's own function (extrapolation at both ends, exact zeros as missing, 3- / 2-sample and empty series), the
    reference call signature, and a batch against the scipy oracle at BASELINE configs[3] length (T = 256)."""
    from keypoints_interpolation_transformer_b200 import baselines
    g = np.load(os.path.join(golden_dir, "cubic.npz"))
    for n in range(int(g["count"])):
        data, mask, ref = (torch.from_numpy(g[f"{k}{n}"]) for k in ("data", "mask", "out"))
        got = baselines.cubic_interpolation(data.to(DEV), mask.unsqueeze(0).to(DEV)).cpu()     # reference call: mask [1, T+1]
        assert got.shape == ref.shape
        keep = (mask == 0)[:, None, None] & (data != 0)
        assert torch.equal(got[keep], data[keep])                                               # samples bit-exact
        scale = ref.abs().max().item()
        assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, scale), n
    inputs, gt, mask = ko.synthetic_batch(6, 256, 71, seed=5)
    inputs[:, :, 3] = 0.0                                        # a keypoint the detector never saw
    got = baselines.cubic_interpolation(inputs.to(DEV), mask.to(DEV)).cpu()
    for b in range(6):
        ref = torch.from_numpy(ko.cubic_interpolation(inputs[b], mask[b]))
        assert (got[b] - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), b
    assert torch.count_nonzero(got[:, :, 3]) == 0
