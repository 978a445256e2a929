import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, copy
from maskplanner_b200 import regressor, synthetic
from oracle import step_oracle as SO
B=2
batch = synthetic.make_batch(B, "windows_v2", seed0=0)
torch.manual_seed(0)
ref = SO.Regressor(449, n_stroke_masks=22)
opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
ref.train(); torch.manual_seed(11); SO.train_step(ref, opt, batch)
ref.eval()
ref64 = copy.deepcopy(ref).double()
mine = regressor.maskplanner_model("windows_v2"); mine.load_state_dict(ref.state_dict()); mine.cuda().eval()
cloud = batch["point_cloud"].permute(0,2,1).float()
torch.manual_seed(12); e1=torch.randint(0,5120,(B,)); e2=torch.randint(0,512,(B,))
def rel(a,b): 
    a=a.double().cpu(); b=b.double().cpu(); return float((a-b).abs().max()/b.abs().max())
with torch.no_grad():
    # layer by layer
    r1x, r1p = ref.sa1(cloud, None, seed_idx=e1)
    m1x, m1p = mine.sa1(cloud.cuda(), None, seed_idx=e1)
    d1x, d1p = ref64.sa1(cloud.double(), None, seed_idx=e1)
    print("sa1 xyz equal", torch.equal(r1x, m1x.cpu()), "points gpu-vs-cpu", rel(m1p, r1p), "cpu-vs-f64", rel(r1p, d1p), "gpu-vs-f64", rel(m1p,d1p))
    r2x, r2p = ref.sa2(r1x, r1p, seed_idx=e2)
    m2x, m2p = mine.sa2(m1x, m1p, seed_idx=e2)
    d2x, d2p = ref64.sa2(d1x, d1p, seed_idx=e2)
    print("sa2 xyz equal", torch.equal(r2x, m2x.cpu()), "points gpu-vs-cpu", rel(m2p, r2p), "cpu-vs-f64", rel(r2p, d2p), "gpu-vs-f64", rel(m2p,d2p))
    # sa2 with identical inputs
    m2p_same = mine.sa2(r1x.cuda(), r1p.cuda(), seed_idx=e2)[1]
    print("sa2 same-input gpu-vs-cpu", rel(m2p_same, r2p))
    r3 = ref.sa3(r2x, r2p)[1]; m3 = mine.sa3(m2x, m2p)[1]; d3 = ref64.sa3(d2x,d2p)[1]
    print("sa3 gpu-vs-cpu", rel(m3, r3), "cpu-vs-f64", rel(r3,d3), "gpu-vs-f64", rel(m3,d3))
    m3_same = mine.sa3(r2x.cuda(), r2p.cuda())[1]
    print("sa3 same-input gpu-vs-cpu", rel(m3_same, r3))
    a = ref(cloud,(e1,e2)); b = mine(cloud.cuda(),(e1,e2)); c = ref64(cloud.double(),(e1,e2))
    for n,x,y,z in zip(("pred","masks","scores"),a,b,c):
        print(n, "gpu-vs-cpu", rel(y,x), "cpu-vs-f64", rel(x,z), "gpu-vs-f64", rel(y,z))
    print("bn running var min", min(float(v.min()) for k,v in ref.state_dict().items() if "running_var" in k))
    for k,v in ref.state_dict().items():
        if "running_var" in k: print(k, float(v.min()), float(v.max()))
