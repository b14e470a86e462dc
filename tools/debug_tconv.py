"""Batched vs per-graph TransformerConv backward, both against a dense torch (fp64) restatement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raindrop_b200 import functional as RF
torch.manual_seed(0)
Bn, Tn, H, F_, IN = 5, 9, 2, 5, 7
xb = torch.randn(Tn, Bn, IN).cuda()
src_e = torch.tensor([0, 1, 2, 2, 3, 0]).cuda(); tgt_e = torch.tensor([1, 2, 0, 2, 0, 0]).cuda()
eib = torch.stack([src_e, tgt_e])
P = [torch.randn(H * F_, IN).cuda() * 0.3 if i % 2 == 0 else torch.randn(H * F_).cuda() * 0.1 for i in range(8)]
Gb = torch.randn(Tn, Bn, H * F_).cuda()

def dense(x64, Ps):
    wq, bq, wk, bk, wv, bv, ws, bs = Ps
    outs = []
    for b in range(Bn):
        x = x64[:, b]
        q = (x @ wq.T + bq).view(Tn, H, F_); k = (x @ wk.T + bk).view(Tn, H, F_); v = (x @ wv.T + bv).view(Tn, H, F_)
        out = x @ ws.T + bs
        logit = (q[tgt_e] * k[src_e]).sum(-1) / F_ ** 0.5            # [E, H]
        agg = torch.zeros(Tn, H, F_, dtype=x.dtype, device=x.device)
        for node in range(Tn):
            m = (tgt_e == node)
            if m.any():
                a = torch.softmax(logit[m], 0)                       # [e, H]
                agg[node] = (a[:, :, None] * v[src_e[m]]).sum(0)
        outs.append(out + agg.reshape(Tn, H * F_))
    return torch.stack(outs, 1)

x64 = xb.double().requires_grad_(True); P64 = [p.double().requires_grad_(True) for p in P]
o = dense(x64, P64); (o * Gb.double()).sum().backward()
ref = [p.grad for p in P64]; refx = x64.grad

def nw(a, b): return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))

xa = xb.clone().requires_grad_(True); Pa = [p.clone().requires_grad_(True) for p in P]
ob, _ = RF.transformer_conv(xa.reshape(Tn * Bn, IN), eib, None, H, F_, *Pa, geom=(Tn, Bn, Bn, 1))
print("fwd batched vs dense", nw(ob.view(Tn, Bn, -1), o))
(ob.view(Tn, Bn, -1) * Gb).sum().backward()
print("batched: dx", nw(xa.grad, refx), [round(nw(p.grad, r), 7) for p, r in zip(Pa, ref)])
xl = xb.clone().requires_grad_(True); Pl = [p.clone().requires_grad_(True) for p in P]
outs = [RF.transformer_conv(xl[:, b], eib, None, H, F_, *Pl)[0] for b in range(Bn)]
ol = torch.stack(outs, 1)
print("fwd loop vs dense", nw(ol, o))
(ol * Gb).sum().backward()
print("loop   : dx", nw(xl.grad, refx), [round(nw(p.grad, r), 7) for p, r in zip(Pl, ref)])
