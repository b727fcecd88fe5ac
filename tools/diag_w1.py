import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoints_interpolation_transformer_b200 import model, optim, train
from oracle import kit_oracle as ko
B,T,Kp,H,L,NH=32,64,71,256,6,8
sd = ko.deterministic_state_dict(2*Kp,H,L)
m = model.KeypointCompleter(2*Kp,H,L,NH); m.load_state_dict(sd); m=m.to('cuda'); m.train()
inputs, gt, mask = ko.synthetic_batch(B,T,Kp,seed=42)
params = {k: v.clone().requires_grad_(not k.endswith("pos_encoding")) for k, v in sd.items()}
ref_loss, ref_pred = ko.train_forward_loss(params, inputs, gt, mask, NH, criterion="mse")
ref_loss.backward()
step = train.TrainStep(m, optim.FlatAdam(m, lr=0.0), criterion="mse")
step.forward_backward(inputs.cuda(), gt.cuda(), mask.cuda()); torch.cuda.synchronize()
got = {n: m.flat_grads[o:o+c].view(s).cpu() for n,(o,c,s) in zip(m._param_names, m._param_slices)}
for lname in ['transformer.decoder.layers.5', 'transformer.decoder.layers.2', 'transformer.decoder.layers.1']:
    for suffix in ['.linear1.weight', '.linear1.bias', '.linear2.weight', '.linear2.bias']:
        n = lname + suffix
        g, r = got[n], params[n].grad
        print(n, 'rel', ((g-r).norm()/r.norm()).item(), 'norm ratio', (g.norm()/r.norm()).item(), 'cos', (torch.dot(g.flatten(), r.flatten())/(g.norm()*r.norm())).item())
    n = lname + '.linear1.weight'
    g, r = got[n], params[n].grad
    rowerr = (g-r).norm(dim=1); rown = r.norm(dim=1)
    idx = torch.argsort(rowerr, descending=True)[:8]
    print('  top rows by abs err:', [(int(i), round(rowerr[i].item(),6), round(rown[i].item(),6)) for i in idx])
    print('  err^2 share of top 8 rows:', (rowerr[idx]**2).sum().item()/ (rowerr**2).sum().item(), ' of top 64:', (torch.sort(rowerr**2, descending=True)[0][:64].sum()/(rowerr**2).sum()).item())
    bn = lname + '.linear1.bias'
    bg, br = got[bn], params[bn].grad
    print('  bias: top err idx', [(int(i), round((bg-br)[i].item(),7), round(br[i].item(),7)) for i in torch.argsort((bg-br).abs(), descending=True)[:8]])

print('---- rank analysis of the dec.5 linear1.weight error')
n = 'transformer.decoder.layers.5.linear1.weight'
dW = (got[n] - params[n].grad).double()
U, S, Vh = torch.linalg.svd(dW, full_matrices=False)
print('singular values', [round(float(x), 6) for x in S[:6]], 'fro', float(dW.norm()))
eng = m.engine_for(B, T, training=True)
y2 = eng.debug_read('dec5.y2').view(B * T, H).cpu().double()
for k in range(3):
    proj = (y2 @ Vh[k]).abs()
    top = torch.argsort(proj, descending=True)[:5]
    print('sv', k, 'top tokens', [(int(t) // T, int(t) % T, round(float(proj[t]), 3)) for t in top], 'median proj', float(proj.median()))
print('mask rows of those batches:')
t0 = int(torch.argmax((y2 @ Vh[0]).abs()))
b0 = t0 // T
print('b', b0, 't', t0 % T, 'y_mask', mask[b0, 1:].tolist())
s3 = eng.debug_read('dec5.s3').view(B * T, H).cpu().double()
var = s3.var(dim=1, unbiased=False)
print('s3 variance: min', float(var.min()), 'argmin', int(var.argmin()) // T, int(var.argmin()) % T, 'median', float(var.median()), 'at t0', float(var[t0]))
s2 = eng.debug_read('dec5.s2').view(B * T, H).cpu().double()
var2 = s2.var(dim=1, unbiased=False)
print('s2 variance: min', float(var2.min()), 'argmin', int(var2.argmin()) // T, int(var2.argmin()) % T, 'median', float(var2.median()), 'at t0', float(var2[t0]))
