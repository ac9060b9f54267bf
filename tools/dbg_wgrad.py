"""Debug helper: python tools/dbg_wgrad.py [dataset ngf B] -- generator parameter gradients vs autograd, per tensor."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import lsnf_b200
from lsnf_b200 import synth
from oracle import refpath
from helpers import build_nets, rel_l2, to_torch

ds = sys.argv[1] if len(sys.argv) > 1 else "svhn"
ngf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
leak = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
nz = 128 if ds == "cifar10" else 100
c = dict(dataset=ds, nz=nz, ngf=ngf, leak=leak)
args, netG, netF = build_nets(c, "cuda:0", seed=4)
img = synth.image_size(ds)
x_np, z_np, _ = synth.inputs(B, nz, 3, img, 1, seed=9)
z, x = torch.from_numpy(z_np), torch.from_numpy(x_np)
flat, pairs, loss = lsnf_b200.generator_gradients(netG, z.cuda(), x.cuda(), B)
torch.cuda.synchronize()
print("kernels ran; loss", loss.item())
gp = to_torch(synth.generator_state(ds, nz, ngf, 3, seed=4))
leaves = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
layers = refpath.generator_layers(ds, nz, ngf)
want = torch.nn.functional.mse_loss(refpath.generator_forward(leaves, z, layers, leak), x, reduction="sum") / B
want.backward()
print("oracle loss", want.item())
named = dict(netG.named_parameters())
got = {id(p): g for p, g in pairs}
for k in leaves:
    g = got[id(named[k])].cpu()
    print(k, tuple(g.shape), "rel_l2 %.3e" % rel_l2(g, leaves[k].grad), "norm", float(leaves[k].grad.norm()), float(g.norm()))
