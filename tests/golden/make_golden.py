"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (it imports /root/reference read-only; nothing is copied):

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so tests read the committed ``*.npz`` files instead.
Weights are closed-form (oracle.kit_oracle.deterministic_state_dict) so fixtures stay small.
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, REF)
for _stub in ("h5py", "matplotlib", "matplotlib.pyplot", "wandb", "tqdm"):   # imported, never used here
    try:
        __import__(_stub)
    except Exception:
        sys.modules[_stub] = types.ModuleType(_stub)

import model as ref_model                      # noqa: E402  /root/reference/model.py
import euclidean_loss as ref_loss              # noqa: E402
import augmentation as ref_aug                 # noqa: E402
import dataloader as ref_dl                    # noqa: E402
from oracle import kit_oracle as ko            # noqa: E402

torch.set_num_threads(8)


def ref_forward_unbatched(m, x, xf, x_mask, y_mask, K):
    """Exactly the A1_train.py:117-124 call for one sequence; K != 54 through the fc_final hook
     because model.py:167 hard-codes 54."""
    T = x.shape[0]
    src_mask = m.get_mask(x_mask, T, "repeat-inc")
    tgt_mask = m.get_mask(y_mask, T, "repeat-inc")
    cap = {}
    h = m.fc_final.register_forward_hook(lambda mod, i, o: cap.__setitem__("o", o))
    try:
        out = m(x, xf, src_pad_mask=x_mask.unsqueeze(0), tgt_pad_mask=y_mask.unsqueeze(0),
                src_mask=src_mask, tgt_mask=tgt_mask)
    except RuntimeError:
        out = None
    finally:
        h.remove()
    hooked = cap["o"].reshape(T, K, 2)
    if K == 54:
        assert out is not None and torch.equal(out, hooked)
    return hooked


def completer_case(name, K, H, L, NH, T, B, seed, grads_full):
    sd = ko.deterministic_state_dict(2 * K, H, L)
    m = ref_model.KeypointCompleter(input_size=2 * K, hidden_dim=H, num_layers=L, num_heads=NH)
    missing = m.load_state_dict(sd, strict=True)
    inputs, gt, mask = ko.synthetic_batch(B, T, K, seed=seed)
    preds, losses_mse, losses_euc_eval = [], [], []
    m.zero_grad()
    total = 0.0
    for b in range(B):
        x = inputs[b, :-1]
        xf = inputs[b, 1:]
        xm = mask[b, :-1].clone()
        ym = mask[b, 1:].clone()
        pred = ref_forward_unbatched(m, x, xf, xm, ym, K)
        preds.append(pred.detach())
        loss = torch.nn.MSELoss()(pred, gt[b])                       # A1_train.py:254,128
        losses_mse.append(loss.detach())
        total = total + loss / B                                      # batched mean == mean of means
        blended = pred * ym[:, None, None] + gt[b] * (1 - ym)[:, None, None]     # A1_train.py:184
        losses_euc_eval.append(ref_loss.EuclideanLoss()(blended, gt[b]).detach())
    total.backward()
    out = {
        "K": K, "H": H, "L": L, "NH": NH, "T": T, "B": B, "seed": seed,
        "inputs": inputs.numpy(), "gt": gt.numpy(), "mask": mask.numpy(),
        "pred": torch.stack(preds).numpy(),
        "loss_mse": torch.stack(losses_mse).numpy(),
        "loss_euclid_eval": torch.stack(losses_euc_eval).numpy(),
        "loss_euclid_train": np.array([ref_loss.EuclideanLoss()(p, g).item()
                                       for p, g in zip(preds, gt)], dtype=np.float32),
    }
    names = [n for n, _ in m.named_parameters()]
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([p.grad.norm().item() for _, p in m.named_parameters()], dtype=np.float64)
    for n, p in m.named_parameters():
        if n in grads_full:
            out["grad::" + n] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "pred", out["pred"].shape, "loss", out["loss_mse"])


def cycle_case(name="cycle_small_k54", K=54, H=64, L=2, NH=4, T=12, B=2, seed=9):
    """KeypointCompleterCycle (model.py:212-321) called like the second model of A2_train_cycle.py:111-115 ("all" masks,
    all-ones pad masks) and like its first model (:105-109, repeat-inc masks + frame pad masks); gradients w.r.t. the
    parameters AND the two inputs (the cycle feeds one model's output into the other)."""
    sd = ko.deterministic_state_dict(2 * K, H, L)
    sd = {k: (v[:512] if k.endswith("pos_encoding") else v) for k, v in sd.items()}       # max_len = 512 (:226-227)
    m = ref_model.KeypointCompleterCycle(input_size=2 * K, hidden_dim=H, num_layers=L, num_heads=NH)
    m.load_state_dict(sd, strict=True)
    gm = ref_model.KeypointCompleter(input_size=2 * K, hidden_dim=16, num_layers=1, num_heads=2).get_mask   # A2:99-103
    inputs, gt, mask = ko.synthetic_batch(B, T, K, seed=seed)
    out = {"K": K, "H": H, "L": L, "NH": NH, "T": T, "B": B, "seed": seed,
           "inputs": inputs.numpy(), "gt": gt.numpy(), "mask": mask.numpy()}
    for mode in ("second", "first"):
        m.zero_grad()
        preds, dxs, dfs, total = [], [], [], 0.0
        for b in range(B):
            x = inputs[b, :-1].clone().requires_grad_(True)
            xf = inputs[b, 1:].clone().requires_grad_(True)
            xm, ym = mask[b, :-1].clone(), mask[b, 1:].clone()
            if mode == "second":
                pred = m(x, xf, src_pad_mask=torch.ones_like(xm.unsqueeze(0)), tgt_pad_mask=torch.ones_like(ym.unsqueeze(0)),
                         src_mask=gm(xm, T, "all"), tgt_mask=gm(ym, T, "all"))
            else:
                pred = m(x, xf, src_pad_mask=xm.unsqueeze(0), tgt_pad_mask=ym.unsqueeze(0),
                         src_mask=gm(xm, T, "repeat-inc"), tgt_mask=gm(ym, T, "repeat-inc"))
            loss = torch.nn.MSELoss()(pred, gt[b]) / B
            loss.backward()
            total += loss.item()
            preds.append(pred.detach())
            dxs.append(x.grad.clone())
            dfs.append(xf.grad.clone())
        out[mode + "_pred"] = torch.stack(preds).numpy()
        out[mode + "_loss"] = np.float32(total)
        out[mode + "_dinputs"] = torch.stack(dxs).numpy()
        out[mode + "_dfilled"] = torch.stack(dfs).numpy()
        names = [n for n, _ in m.named_parameters()]
        out["grad_names"] = np.array(names)
        out[mode + "_grad_norms"] = np.array([p.grad.norm().item() for _, p in m.named_parameters()], dtype=np.float64)
        for n in ("input_embedding.weight", "fc_final.weight", "transformer.decoder.layers.0.self_attn.in_proj_weight",
                  "learned_filled_positional_encoder"):
            out[mode + "_grad::" + n] = dict(m.named_parameters())[n].grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "losses", out["second_loss"], out["first_loss"])


def get_mask_case():
    m = ref_model.KeypointCompleter(input_size=108, hidden_dim=16, num_layers=1, num_heads=2)
    rs = np.random.RandomState(3)
    out = {}
    for n, T in enumerate([1, 2, 7, 16, 33]):
        fm = torch.from_numpy((rs.uniform(size=T) < 0.4).astype(np.float32))
        out[f"mask{n}"] = fm.numpy()
        for typ in ["triangle", "repeat", "repeat-inc", "all"]:
            if T == 1 and typ in ("repeat", "repeat-inc"):
                continue            # reference squeeze() collapses [1,1] to a scalar
            out[f"{typ}{n}"] = m.get_mask(fm, T, typ).numpy()
    np.savez_compressed(os.path.join(HERE, "get_mask.npz"), **out)


def normalize_case():
    rs = np.random.RandomState(5)
    out = {}
    body = {"pose_left_shoulder": 5, "pose_right_shoulder": 6, "pose_right_eye": 2}
    for n, (T, K) in enumerate([(9, 12), (24, 54), (40, 71)]):
        d = rs.uniform(0.05, 0.95, size=(T, K, 2)).astype(np.float32)
        d[rs.uniform(size=(T, K)) < 0.08] = 0.0
        # shoulders missing at the start (frames stay untouched), later (box carried forward)
        d[0, 5, 0] = 0.0
        d[1, 6] = 0.0
        d[T // 2, 5] = 0.0
        d[T // 2 + 1, 6, 0] = 0.0
        out[f"in{n}"] = d.copy()
        out[f"out{n}"] = ref_dl.normalize_pose(d.copy(), body)
    out["ids"] = np.array([5, 6, 2])
    np.savez_compressed(os.path.join(HERE, "normalize_pose.npz"), **out)


class _Tape:
    """Records what the reference draws from ``random`` so tests can replay the parameters."""
    def __init__(self, seed):
        self.r = random.Random(seed)
        self.uniform_draws, self.random_draws = [], []

    def uniform(self, a, b):
        v = self.r.uniform(a, b)
        self.uniform_draws.append(v)
        return v

    def random(self):
        v = self.r.random()
        self.random_draws.append(v)
        return v


def augment_case():
    K, T = 54, 10
    pose = list(range(0, 12))
    lh = list(range(12, 33))
    rh = list(range(33, 54))
    ids = {"pose": pose, "left_hand": lh, "rigth_hand": rh}
    body = {"pose_chest_middle_up": 0, "pose_left_shoulder": 5, "pose_left_elbow": 7, "pose_left_wrist": 9,
            "pose_right_shoulder": 6, "pose_right_elbow": 8, "pose_right_wrist": 10}
    aug = ref_aug.augmentation(ids, body)
    rs = np.random.RandomState(11)
    base = rs.uniform(0.05, 0.95, size=(T, K, 2)).astype(np.float32)
    base[rs.uniform(size=(T, K)) < 0.1] = 0.0
    out = {"base": base, "pose": np.array(pose), "left_hand": np.array(lh), "right_hand": np.array(rh),
           "arm_chains": np.array(aug.ARM_IDENTIFIERS_ORDER)}
    real_uniform, real_random = ref_aug.random.uniform, ref_aug.random.random
    try:
        for n in range(3):
            tape = _Tape(100 + n)
            ref_aug.random.uniform, ref_aug.random.random = tape.uniform, tape.random
            out[f"rotate{n}"] = aug.augment_rotate(torch.from_numpy(base.copy()), (-15, 15)).numpy()
            out[f"rotate{n}_u"] = np.array(tape.uniform_draws)
            tape = _Tape(200 + n)
            ref_aug.random.uniform, ref_aug.random.random = tape.uniform, tape.random
            out[f"squeeze{n}"] = aug.augment_shear(torch.from_numpy(base.copy()), "squeeze", (-0.15, 0.15)).numpy()
            out[f"squeeze{n}_u"] = np.array(tape.uniform_draws)
            tape = _Tape(300 + n)
            ref_aug.random.uniform, ref_aug.random.random = tape.uniform, tape.random
            out[f"persp{n}"] = aug.augment_shear(torch.from_numpy(base.copy()), "perspective", (-0.15, 0.15)).numpy()
            out[f"persp{n}_u"] = np.array(tape.uniform_draws)
            out[f"persp{n}_r"] = np.array(tape.random_draws)
            tape = _Tape(400 + n)
            ref_aug.random.uniform, ref_aug.random.random = tape.uniform, tape.random
            out[f"arm{n}"] = aug.augment_arm_joint_rotate(torch.from_numpy(base.copy()), 0.5, (-15, 15)).numpy()
            out[f"arm{n}_u"] = np.array(tape.uniform_draws)
            out[f"arm{n}_r"] = np.array(tape.random_draws)
    finally:
        ref_aug.random.uniform, ref_aug.random.random = real_uniform, real_random
    np.savez_compressed(os.path.join(HERE, "augmentation.npz"), **out)


def missing_case():
    cwd = os.getcwd()
    os.chdir(REF)                      # put_missing_frames opens ./dataset_config.json
    out = {}
    try:
        n = 0
        for T in (20, 64, 131):
            for ds in ("AUTSL", "AEC", "PUCP_PSL_DGI305"):
                seed = 1000 + n
                random.seed(seed)
                np.random.seed(seed)
                video = torch.arange(T, dtype=torch.float32).view(T, 1, 1).repeat(1, 3, 2).clone()
                v2, mask = ref_dl.put_missing_frames(video, False, ds)
                v3, m3 = ref_dl.add_sos(v2, mask)
                out[f"src{n}"] = v2[:, 0, 0].numpy().astype(np.int32)
                out[f"mask{n}"] = mask.numpy()
                out[f"sos_video{n}"] = v3.numpy()
                out[f"sos_mask{n}"] = m3.numpy()
                out[f"meta{n}"] = np.array([T, seed])
                out[f"ds{n}"] = np.array(ds)
                n += 1
        # random-missing mode
        random.seed(77)
        video = (1 + torch.arange(30, dtype=torch.float32)).view(30, 1, 1).repeat(1, 2, 2).clone()
        v2, mask = ref_dl.put_missing_frames(video, True, "AUTSL")
        out["rand_video"] = v2.numpy()
        out["rand_mask"] = mask.numpy()
        out["count"] = np.array(n)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "missing_frames.npz"), **out)


def cubic_case():
    """3_test_cubic_interpolation.py:32-58 ``cubic_interpolation`` itself (the module runs main() at import, so the function is
    lifted out of its source with ast), on small sequences: AUTSL-style missing blocks, a block at each end (extrapolation),
    exact zeros in unmasked frames, a keypoint with 3 / 2 samples, an all-missing keypoint."""
    import ast
    import pandas as pd
    src = open(os.path.join(REF, "3_test_cubic_interpolation.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "cubic_interpolation"][0]
    ns = {"torch": torch, "np": np, "pd": pd}
    exec(compile(ast.Module([fn], []), "ref_cubic", "exec"), ns)
    ref_fn = ns["cubic_interpolation"]
    rs = np.random.RandomState(21)
    out = {"count": 3}
    for n, (T1, K) in enumerate([(33, 6), (65, 5), (20, 4)]):
        tt = np.arange(T1)[:, None, None]
        data = (0.5 + 0.3 * np.sin(2 * np.pi * (rs.uniform(0.5, 3, (1, K, 2)) * tt / T1 + rs.uniform(0, 1, (1, K, 2))))).astype(np.float32)
        mask = np.zeros(T1, dtype=np.float32)
        mask[5:9] = 1
        mask[T1 // 2:T1 // 2 + 6] = 1
        if n == 1:
            mask[1:4] = 1            # SOS stays, a block right behind it
            mask[T1 - 3:] = 1        # block at the end: extrapolation
        if n == 2:
            mask[0:2] = 1            # block at the very start: extrapolation to the left
        data[7, 0, 0] = 0.0          # exact zeros in unmasked frames are missing too
        data[12, 1, :] = 0.0
        keep = np.where(mask == 0)[0]
        data[:, 2, 0] = 0.0
        data[keep[[1, len(keep) // 2, -2]], 2, 0] = [0.3, 0.5, 0.2]     # three samples: parabola
        data[:, 3, 1] = 0.0
        data[keep[[2, -1]], 3, 1] = [0.4, 0.6]              # two samples: straight line
        if K > 4:
            data[:, 4, :] = 0.0                            # nothing left: zeros
        res = ref_fn(torch.from_numpy(data.copy()), torch.from_numpy(mask).unsqueeze(0))
        out[f"data{n}"], out[f"mask{n}"], out[f"out{n}"] = data, mask, res.numpy()
    np.savez_compressed(os.path.join(HERE, "cubic.npz"), **out)
    print("cubic", [out[f"out{n}"].shape for n in range(3)])


def loss_case():
    rs = np.random.RandomState(9)
    o = torch.from_numpy(rs.normal(size=(3, 17, 71, 2)).astype(np.float32))
    t = torch.from_numpy(rs.normal(size=(3, 17, 71, 2)).astype(np.float32))
    out = {"o": o.numpy(), "t": t.numpy(),
           "euclid": ref_loss.EuclideanLoss()(o, t).numpy(),
           "mse": torch.nn.MSELoss()(o, t).numpy(),
           "euclid_dist": ref_loss.EuclideanDistanceLoss()(o[0], t[0]).numpy()}
    od = o[0].clone()
    od[3, 5] = t[0, 3, 5]                      # one zero distance: torch.norm's backward gives 0 there
    od.requires_grad_(True)
    d = ref_loss.EuclideanDistanceLoss()(od, t[0])
    d.backward()
    out.update({"distance_o": od.detach().numpy(), "distance": d.detach().numpy(), "distance_grad": od.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


if __name__ == "__main__":
    get_mask_case()
    normalize_case()
    augment_case()
    missing_case()
    loss_case()
    cycle_case()
    cubic_case()
    small_full = ["learned_input_positional_encoder", "input_embedding.bias", "fc_final.weight", "fc_final.bias",
                  "transformer.encoder.layers.0.self_attn.in_proj_bias",
                  "transformer.decoder.layers.1.multihead_attn.out_proj.weight",
                  "transformer.decoder.layers.0.norm2.weight", "swiGlu_decoded.fc2.bias",
                  "filled_embedding.weight", "transformer.encoder.layers.1.linear1.bias"]
    completer_case("completer_small_k54", K=54, H=64, L=2, NH=4, T=12, B=3, seed=42, grads_full=small_full)
    completer_case("completer_small_k71", K=71, H=64, L=2, NH=4, T=12, B=3, seed=43, grads_full=small_full)
    completer_case("completer_default_k54", K=54, H=256, L=6, NH=8, T=16, B=2, seed=44,
                   grads_full=["fc_final.bias", "learned_filled_positional_encoder",
                               "transformer.encoder.layers.0.norm1.weight",
                               "transformer.decoder.layers.5.self_attn.in_proj_bias"])
    print("done")
