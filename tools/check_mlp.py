"""SharedMLPMax (tensor-core path) vs a torch emulation with the same bf16 rounding points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
from maskplanner_b200.shared_mlp import shared_mlp_max, pad64
dev = torch.device("cuda", 0)
torch.manual_seed(0)

def emu(a0, K, convs, bns):
    x = a0.float()
    M = x.shape[0]
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        cout, cin = conv.weight.shape[:2]
        w = conv.weight.view(cout, cin).bfloat16().float()
        z = (x[:, :cin] @ w.t()).bfloat16().float()
        mean = z.mean(0); var = z.var(0, unbiased=False)
        s = bn.weight / torch.sqrt(var + bn.eps); t = bn.bias - mean * s
        x = F.relu(z * s + t)
        if i < len(convs) - 1:
            x = x.bfloat16().float()
    return x.view(M // K, K, -1).max(1)[0]

def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))

for (G, K, cin, mlp) in [(64, 12, 9, [16, 24, 32]), (2, 40, 35, [32, 48]), (512, 32, 3, [64, 64, 128]), (256, 64, 131, [128, 128, 256]), (4, 128, 259, [256, 512, 1024])]:
    convs = nn.ModuleList(); bns = nn.ModuleList(); c = cin
    for co in mlp:
        convs.append(nn.Conv2d(c, co, 1)); bns.append(nn.BatchNorm2d(co)); c = co
    convs.to(dev); bns.to(dev)
    for bn in bns:
        bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-0.3, 0.3)
    M = G * K
    base = torch.randn(M, cin, device=dev)
    a0 = F.pad(base, (0, pad64(cin) - cin)).bfloat16()
    wout = torch.randn(G, mlp[-1], device=dev)
    a1 = a0.clone().requires_grad_(True)
    out1 = shared_mlp_max(a1, K, convs, bns, True)
    (out1 * wout).sum().backward()
    g1 = [a1.grad.float()] + [p.grad.clone() for p in list(convs.parameters()) + list(bns.parameters())]
    for p in list(convs.parameters()) + list(bns.parameters()): p.grad = None
    a2 = a0.clone().float().requires_grad_(True)
    out2 = emu(a2, K, convs, bns)
    (out2 * wout).sum().backward()
    g2 = [a2.grad] + [(p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for p in list(convs.parameters()) + list(bns.parameters())]
    names = ["a0"] + [n for n, _ in list(convs.named_parameters())] + ["bn." + n for n, _ in bns.named_parameters()]
    print("G=%d K=%d cin=%d mlp=%s  out relL2 %.3e" % (G, K, cin, mlp, rel(out1, out2)))
    for n, x, y in zip(names, g1, g2):
        if "bias" in n and not n.startswith("bn."):
            continue
        print("   grad %-14s relL2 %.3e cos %.5f" % (n, rel(x, y), float(F.cosine_similarity(x.flatten(), y.flatten(), dim=0))))
