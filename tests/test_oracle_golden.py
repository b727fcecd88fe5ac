"""Pin the CPU oracle (oracle/kit_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import kit_oracle as ko


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", ["completer_small_k54", "completer_small_k71", "completer_default_k54"])
def test_completer_forward_loss_and_grads(golden_dir, name):
    g = _load(golden_dir, name)
    K, H, L, NH, B = (int(g[k]) for k in ("K", "H", "L", "NH", "B"))
    sd = ko.deterministic_state_dict(2 * K, H, L)
    params = {k: v.clone().requires_grad_(not k.endswith("pos_encoding")) for k, v in sd.items()}
    inputs, gt, mask = (torch.from_numpy(g[k]) for k in ("inputs", "gt", "mask"))
    loss, pred = ko.train_forward_loss(params, inputs, gt, mask, NH, criterion="mse")
    # fp32 CPU vs fp32 CPU with a different op order: 1e-4 relative to the output scale
    ref_pred = torch.from_numpy(g["pred"])
    assert (pred - ref_pred).abs().max().item() <= 1e-4 * max(1.0, ref_pred.abs().max().item())
    assert abs(loss.item() - float(g["loss_mse"].mean())) <= 1e-5 * max(1.0, abs(loss.item()))
    # eval protocol (A1_train.py:184-186)
    with torch.no_grad():
        x, xf, y, xm, ym = ko.step_slices(inputs, gt, mask)
        for b in range(B):
            e = ko.euclidean_loss(ko.eval_blend(pred[b], y[b], ym[b]), y[b]).item()
            assert abs(e - float(g["loss_euclid_eval"][b])) <= 2e-5 * max(1.0, abs(e))
            e2 = ko.euclidean_loss(pred[b], y[b]).item()
            assert abs(e2 - float(g["loss_euclid_train"][b])) <= 2e-5 * max(1.0, abs(e2))
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    norms = g["grad_norms"]
    for n, ref_norm in zip(names, norms):
        got = params[n].grad.norm().item()
        assert abs(got - ref_norm) <= 2e-4 * max(ref_norm, 1e-3), (n, got, ref_norm)
    for key in g.files:
        if key.startswith("grad::"):
            n = key[6:]
            ref = torch.from_numpy(g[key])
            err = (params[n].grad - ref).abs().max().item()
            assert err <= 2e-4 * max(ref.abs().max().item(), 1e-3), (n, err)


def test_cycle_model_matches_reference(golden_dir):
    """KeypointCompleterCycle (model.py:212-321) as A2_train_cycle.py calls it: second model ("all" masks, all-ones pad masks)
    and first-model style (repeat-inc masks + frame pad masks on BOTH stacks); parameters' and inputs' gradients."""
    g = _load(golden_dir, "cycle_small_k54")
    K, H, L, NH, B, T = (int(g[k]) for k in ("K", "H", "L", "NH", "B", "T"))
    sd = ko.deterministic_state_dict(2 * K, H, L)
    sd = {k: (v[:512] if k.endswith("pos_encoding") else v) for k, v in sd.items()}
    inputs, gt, mask = (torch.from_numpy(g[k]) for k in ("inputs", "gt", "mask"))
    names = [str(n) for n in g["grad_names"]]
    for mode in ("second", "first"):
        params = {k: v.clone().requires_grad_(not k.endswith("pos_encoding")) for k, v in sd.items()}
        x = inputs[:, :-1].clone().requires_grad_(True)
        xf = inputs[:, 1:].clone().requires_grad_(True)
        xm, ym = mask[:, :-1], mask[:, 1:]
        if mode == "second":
            pred = ko.completer_forward(params, x, xf, NH, src_pad=torch.ones_like(xm), tgt_pad=torch.ones_like(ym), cycle=True)
        else:
            pred = ko.completer_forward(params, x, xf, NH, src_pad=xm, src_bias=ko.repeat_inc_bias(xm),
                                        tgt_bias=ko.repeat_inc_bias(ym), tgt_pad=ym, cycle=True)
        ref_pred = torch.from_numpy(g[mode + "_pred"])
        assert (pred - ref_pred).abs().max().item() <= 1e-4 * max(1.0, ref_pred.abs().max().item())
        loss = sum(ko.mse_loss(pred[b], gt[b]) / B for b in range(B))       # make_golden: per-sequence MSE / B
        assert abs(loss.item() - float(g[mode + "_loss"])) <= 1e-5 * max(1.0, abs(loss.item()))
        loss.backward()
        for n, ref_norm in zip(names, g[mode + "_grad_norms"]):
            got = params[n].grad.norm().item()
            assert abs(got - ref_norm) <= 2e-4 * max(ref_norm, 1e-3), (mode, n, got, ref_norm)
        for key in g.files:
            if key.startswith(mode + "_grad::"):
                n = key.split("::", 1)[1]
                ref = torch.from_numpy(g[key])
                assert (params[n].grad - ref).abs().max().item() <= 2e-4 * max(ref.abs().max().item(), 1e-3), (mode, n)
        for got, key in ((x.grad, "_dinputs"), (xf.grad, "_dfilled")):
            ref = torch.from_numpy(g[mode + key])
            assert (got - ref).abs().max().item() <= 2e-4 * max(ref.abs().max().item(), 1e-6), (mode, key)


def test_get_mask_bit_exact(golden_dir):
    g = _load(golden_dir, "get_mask")
    for n, T in enumerate([1, 2, 7, 16, 33]):
        fm = torch.from_numpy(g[f"mask{n}"])
        for typ in ["triangle", "repeat", "repeat-inc", "all"]:
            key = f"{typ}{n}"
            if key not in g.files:
                continue
            got = ko.get_mask(fm, T, typ).numpy()
            assert got.shape == g[key].shape
            assert np.array_equal(got, g[key]), key       # {0, 1, -inf}: bit exact


def test_normalize_pose_matches_reference(golden_dir):
    g = _load(golden_dir, "normalize_pose")
    ls, rs, re = (int(v) for v in g["ids"])
    for n in range(3):
        got = ko.normalize_pose(g[f"in{n}"], ls, rs, re)
        ref = g[f"out{n}"]
        assert np.array_equal(got == 0, ref == 0)
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)


def test_augmentations_match_reference(golden_dir):
    g = _load(golden_dir, "augmentation")
    base = g["base"]
    pose, lh, rh = list(g["pose"]), list(g["left_hand"]), list(g["right_hand"])
    body_ids = pose + lh + rh
    hand_ids = lh + rh
    chains = [list(c) for c in g["arm_chains"]]
    for n in range(3):
        ang = math.radians(float(g[f"rotate{n}_u"][0]))
        np.testing.assert_allclose(ko.augment_rotate(base, ang, body_ids, hand_ids), g[f"rotate{n}"],
                                   rtol=1e-6, atol=1e-7)
        ml, mr = (float(v) for v in g[f"squeeze{n}_u"])
        got = ko.augment_shear(base, ko.perspective_matrix("squeeze", ml, mr), body_ids)
        np.testing.assert_allclose(got, g[f"squeeze{n}"], rtol=1e-6, atol=1e-7)
        assert np.array_equal(got == 0, g[f"squeeze{n}"] == 0)
        a = float(g[f"persp{n}_u"][0])
        left = float(g[f"persp{n}_r"][0]) < 0.5
        got = ko.augment_shear(base, ko.perspective_matrix("perspective", a, left=left), body_ids)
        np.testing.assert_allclose(got, g[f"persp{n}"], rtol=1e-6, atol=1e-7)
        assert np.array_equal(got == 0, g[f"persp{n}"] == 0)
        # arm-joint rotate: replay the coin/angle tape in the reference's draw order
        coins = list(g[f"arm{n}_r"])
        us = list(g[f"arm{n}_u"])
        angles = []
        for chain in chains:
            row = []
            for _ in chain:
                if coins.pop(0) < 0.5:
                    row.append(math.radians(us.pop(0)))
                else:
                    row.append(None)
            angles.append(row)
        assert not coins and not us
        np.testing.assert_allclose(ko.augment_arm_joint_rotate(base, chains, angles), g[f"arm{n}"],
                                   rtol=1e-6, atol=1e-7)


def test_missing_frames_policy_and_hold_fill_bit_exact(golden_dir):
    g = _load(golden_dir, "missing_frames")
    import json
    cfgs = {"AUTSL": ko.AUTSL_CONFIG,
            "AEC": dict(mean_consecutive_missing=3.25, std_consecutive_missing=3.09, samples=267,
                        mean_number_missing_blocks=1.92, std_number_missing_blocks=1.1),
            "PUCP_PSL_DGI305": dict(mean_consecutive_missing=4.04, std_consecutive_missing=5.63, samples=185,
                                    mean_number_missing_blocks=1.66, std_number_missing_blocks=1.11)}
    for n in range(int(g["count"])):
        T, seed = (int(v) for v in g[f"meta{n}"])
        ds = str(g[f"ds{n}"])
        random.seed(seed)
        np.random.seed(seed)
        blocks = ko.missing_blocks_from_config(T, cfgs[ds], rng=random, nprng=np.random)
        src, mask = ko.hold_fill_sources(T, blocks)
        assert np.array_equal(src, g[f"src{n}"]), (n, blocks)
        assert np.array_equal(mask, g[f"mask{n}"])
        video = np.arange(T, dtype=np.float32).reshape(T, 1, 1).repeat(3, 1).repeat(2, 2)
        v, m = ko.apply_sources_add_sos(video, src, mask)
        assert np.array_equal(v, g[f"sos_video{n}"])
        assert np.array_equal(m, g[f"sos_mask{n}"])
    random.seed(77)
    src, mask = ko.random_missing_sources(30, rng=random)
    video = (1 + np.arange(30, dtype=np.float32)).reshape(30, 1, 1).repeat(2, 1).repeat(2, 2)
    v, _ = ko.apply_sources_add_sos(video, src, mask)
    assert np.array_equal(v[1:], g["rand_video"])
    assert np.array_equal(mask, g["rand_mask"])


def test_losses_match_reference(golden_dir):
    g = _load(golden_dir, "loss")
    o, t = torch.from_numpy(g["o"]), torch.from_numpy(g["t"])
    assert abs(ko.euclidean_loss(o, t).item() - float(g["euclid"])) <= 1e-6 * float(g["euclid"])
    assert abs(ko.mse_loss(o, t).item() - float(g["mse"])) <= 1e-6 * float(g["mse"])
    assert abs(ko.euclidean_loss(o, t).item() - 2 * ko.mse_loss(o, t).item()) < 1e-5
    if "distance" in g.files:      # EuclideanDistanceLoss (euclidean_loss.py:19-37) and its gradient
        oo = torch.from_numpy(g["distance_o"]).requires_grad_(True)
        d = ko.euclidean_distance_loss(oo, t[0])
        assert abs(d.item() - float(g["distance"])) <= 1e-6 * float(g["distance"])
        d.backward()
        np.testing.assert_allclose(oo.grad.numpy(), g["distance_grad"], rtol=1e-5, atol=1e-7)


def test_cubic_interpolation_matches_reference(golden_dir):
    """oracle cubic_interpolation against outputs of the reference's own function (3_test_cubic_interpolation.py:32-58)."""
    g = _load(golden_dir, "cubic")
    for n in range(int(g["count"])):
        got = ko.cubic_interpolation(g[f"data{n}"], g[f"mask{n}"])
        ref = g[f"out{n}"]
        assert got.shape == ref.shape and not np.isnan(got).any()
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)
        keep = (g[f"mask{n}"] == 0)[:, None, None] & (g[f"data{n}"] != 0)
        assert np.array_equal(got[keep], g[f"data{n}"][keep])          # samples are kept bit-exactly


@pytest.mark.parametrize("name", ["completer_small_k54", "completer_small_k71"])
def test_torch_reference_matches_reference_goldens(golden_dir, name):
    """oracle/torch_reference.StockCompleter (the nn.Transformer restatement bench.py times as gpu_baseline / cpu_baseline) loads
    the reference's state_dict keys and reproduces the reference's own outputs, loss and gradient norms."""
    from oracle import torch_reference as tr
    g = _load(golden_dir, name)
    K, H, L, NH = (int(g[k]) for k in ("K", "H", "L", "NH"))
    m = tr.StockCompleter(2 * K, H, L, NH)
    missing, unexpected = m.load_state_dict(ko.deterministic_state_dict(2 * K, H, L), strict=True)
    assert not missing and not unexpected
    inputs, gt, mask = (torch.from_numpy(g[k]) for k in ("inputs", "gt", "mask"))
    x_mask, y_mask = mask[:, :-1], mask[:, 1:]
    pred = m(inputs[:, :-1], inputs[:, 1:], src_pad_mask=x_mask, src_mask=tr.repeat_inc_masks(x_mask, NH),
             tgt_mask=tr.repeat_inc_masks(y_mask, NH))
    ref_pred = torch.from_numpy(g["pred"])
    assert (pred - ref_pred).abs().max().item() <= 1e-4 * max(1.0, ref_pred.abs().max().item())
    loss = torch.nn.functional.mse_loss(pred, gt)
    assert abs(loss.item() - float(g["loss_mse"].mean())) <= 1e-5 * max(1.0, abs(loss.item()))
    loss.backward()
    grads = dict(m.named_parameters())
    for n, ref_norm in zip([str(n) for n in g["grad_names"]], g["grad_norms"]):
        got = grads[n].grad.norm().item()
        assert abs(got - ref_norm) <= 2e-4 * max(ref_norm, 1e-3), (n, got, ref_norm)
    # the one-sequence Python-loop mask of the A1-faithful baseline equals the vectorised one
    assert torch.equal(tr.get_mask_loop(x_mask[0], x_mask.shape[1]), tr.repeat_inc_masks(x_mask[:1], 1)[0])
