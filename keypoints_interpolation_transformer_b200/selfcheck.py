"""smoke(): one tiny train step of the hot path on cuda:0, checked against the CPU oracle."""
import torch


def smoke():
    from oracle import kit_oracle as ko            # the checker (allowed here by the scope rules)
    from . import model, optim, train
    Kp, H, L, NH, B, T = 71, 64, 2, 4, 4, 16
    dev = torch.device("cuda", 0)
    sd = ko.deterministic_state_dict(2 * Kp, H, L)
    m = model.KeypointCompleter(2 * Kp, H, L, NH)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    inputs, gt, mask = ko.synthetic_batch(B, T, Kp, seed=1)
    step = train.TrainStep(m, optim.FlatAdam(m, lr=1e-4), criterion="mse")
    loss = step(inputs.to(dev), gt.to(dev), mask.to(dev))
    torch.cuda.synchronize()
    params = {k: v.clone().requires_grad_(not k.endswith("pos_encoding")) for k, v in sd.items()}
    ref_loss, ref_pred = ko.train_forward_loss(params, inputs, gt, mask, NH, criterion="mse")
    ref_loss.backward()
    rel = ((step.pred.cpu() - ref_pred).norm() / ref_pred.norm()).item()
    assert rel < 2e-2, f"pred differs from the oracle: {rel}"
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    num = den = 0.0
    for n, (o, c, s) in zip(m._param_names, m._param_slices):
        g = m.flat_grads[o:o + c].view(s).cpu()
        num += float(((g - params[n].grad).double() ** 2).sum())
        den += float((params[n].grad.double() ** 2).sum())
    assert (num / den) ** 0.5 < 3e-2, f"gradients differ from the oracle: {(num / den) ** 0.5}"
    print(f"smoke ok: loss {loss.item():.6f} (oracle {ref_loss.item():.6f}), pred rel err {rel:.2e}, "
          f"grad rel err {(num / den) ** 0.5:.2e}, launches {step.last_launches}")
