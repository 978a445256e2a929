import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer
from oracle import step_oracle as SO
for B in (4, 8):
    tr = Trainer("windows_v2", torch.device("cuda", 0), seed=3)
    ref = SO.Regressor(449, n_stroke_masks=22); ref.load_state_dict(tr.model.state_dict())
    tr.model.dropout.p = 0.0; ref.dropout.p = 0.0; ref.train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    batch = synthetic.make_batch(B, "windows_v2", seed0=21)
    g = torch.Generator().manual_seed(B)
    seeds = (torch.randint(0,5120,(B,),generator=g), torch.randint(0,512,(B,),generator=g))
    want = SO.train_step(ref, opt, batch, seeds)
    got = float(tr.step(tr.to_device(batch), seeds).item())
    print("B", B, "loss", got, want, abs(got-want)/abs(want))
    rp = dict(ref.named_parameters())
    for n, p in tr.model.named_parameters():
        w = rp[n].grad; gg = p.grad.cpu()
        print("  %-28s relL2 %.3e  cos %.6f  |w| %.3e maxabs %.3e" % (n, float((gg-w).norm()/(w.norm()+1e-30)), float(torch.nn.functional.cosine_similarity(gg.flatten(), w.flatten(), dim=0)), float(w.norm()), float(w.abs().max())))
