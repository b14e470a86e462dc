"""Structured-input probe of the grouped weight-gradient kernel (rd_linear_wgrad_group)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raindrop_b200 import lib as L

lib = L.load()
torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)


def run(dY, X):
    rows, nout = dY.shape
    kin = X.shape[1]
    items = (L.RdWgradItem * 1)()
    dW = torch.full((nout, kin), float("nan"), device="cuda"); db = torch.full((nout,), float("nan"), device="cuda")
    part = torch.full((lib.rd_linear_wgrad_partial_bytes(rows, nout, kin) // 4,), float("nan"), device="cuda")
    it = items[0]
    it.d_out, it.x, it.rows, it.out_features, it.in_features = dY.data_ptr(), X.data_ptr(), rows, nout, kin
    it.d_weight, it.d_bias, it.partial = dW.data_ptr(), db.data_ptr(), part.data_ptr()
    L.check(lib.rd_linear_wgrad_group(items, 1, L.stream_ptr()), "wgrad")
    torch.cuda.synchronize()
    return dW, db, part


for rows, nout, kin in ((256, 32, 32), (256, 128, 64), (512, 160, 152)):
    g = torch.Generator().manual_seed(0)
    dY = torch.zeros(rows, nout); X = torch.zeros(rows, kin)
    for m in range(min(nout, rows)):
        dY[m, m] = 1.0                       # dW[m, n] = X[m, n]
    X[:, :] = torch.arange(rows)[:, None] * 1000.0 + torch.arange(kin)[None, :]
    dW, db, part = run(dY.cuda(), X.cuda())
    ref = dY.double().T @ X.double()
    print("== structured rows=%d nout=%d kin=%d: max|dW| %.3f max|ref| %.3f  err %.3e  nan %d  partial nan frac %.3f"
          % (rows, nout, kin, dW.nan_to_num().abs().max().item(), ref.abs().max().item(),
             (dW.cpu().double() - ref).abs().max().item(), int(torch.isnan(dW).sum()), torch.isnan(part).float().mean().item()))
    print("dW[0:4, 0:8]\n", dW[0:4, 0:8].cpu(), "\nref\n", ref[0:4, 0:8].float())
    print("dW[33:35, 30:36]\n", dW[33:35, 30:36].cpu() if nout > 34 and kin > 35 else None)
    print("db[0:8]", db[0:8].cpu(), "ref", dY.sum(0)[0:8])
    dYr = torch.randn(rows, nout, generator=g); Xr = torch.randn(rows, kin, generator=g)
    dW, db, _ = run(dYr.cuda(), Xr.cuda())
    ref = dYr.double().T @ Xr.double()
    e = (dW.cpu().double() - ref).abs()
    print("random: err max %.3e  (ref max %.3f)  db err %.3e ; rows of dW with err>1e-3: %s"
          % (e.max().item(), ref.abs().max().item(), (db.cpu().double() - dYr.double().sum(0)).abs().max().item(),
             (e.max(1).values > 1e-3).nonzero().flatten()[:16].tolist()))
