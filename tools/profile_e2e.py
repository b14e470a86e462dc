"""cProfile of the module-API training loop (host-side overhead of the e2e path)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_dropin
from raindrop_b200.synth import model_config, make_batch
cfg = model_config("P19", dropout=0.2)
model = build_dropin(cfg, 4).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
crit = torch.nn.CrossEntropyLoss()
hb = make_batch(cfg, 128, seed=1, pin=True)
def step():
    P = hb["src"].cuda(non_blocking=True); Pt = hb["times"].cuda(non_blocking=True)
    Ps = hb["static"].cuda(non_blocking=True); y = hb["y"].cuda(non_blocking=True)
    lengths = torch.sum(Pt > 0, dim=0)
    out, _, _ = model.forward(P, Ps, Pt, lengths)
    opt.zero_grad(); loss = crit(out, y); loss.backward(); opt.step()
    return loss.item()
for _ in range(10): step()
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): step()
torch.cuda.synchronize(); print("wall per step %.3f ms" % ((time.perf_counter() - t0) * 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
