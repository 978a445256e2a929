"""Developer check of the fused heads' mask branch against a float64 replica with intermediate gradients."""
import copy, sys, os, torch
import torch.nn.functional as Fn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskplanner_b200 import regressor, _cabi as c
from maskplanner_b200.heads import _Ctx, _tf32_round
def rel(a,b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-300))
torch.manual_seed(0)
m = regressor.maskplanner_model("windows_v2").cuda()
B=64; dev=torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(B)
k = _Ctx(B, dev, 2); lib=k.lib; Bp=k.Bp
# linear_dx accuracy for the three mask-branch shapes
for name,(Nout,Kin) in {"sm_fc3":(9878,1024),"mask_conf":(22,1024),"sm_fc2":(1024,1024)}.items():
    W = torch.randn(Nout,Kin,device=dev,generator=g)*0.03
    dYt = torch.randn(Nout,Bp,device=dev,generator=g)
    got = k.linear_dx(W,dYt)
    want = W.double().t() @ dYt.double()
    print("linear_dx",name,rel(got,want))
    got2 = k.linear_dx(W,dYt,out=got.clone())
    print("   accumulate",rel(got2,2*want))
    Xt = torch.randn(Kin,Bp,device=dev,generator=g); hi=_tf32_round(Xt)
    dW = k.linear_dw(dYt,hi,Xt-hi,Kin)
    print("   linear_dw",rel(dW,dYt.double()@Xt.double().t()))
    X = torch.randn(Bp,Kin,device=dev,generator=g); hi=_tf32_round(X)
    Yt = k.linear_fwd(W,hi,X-hi)
    print("   linear_fwd",rel(Yt,W.double()@X.double().t()))
