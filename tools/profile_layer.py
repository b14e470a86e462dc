"""The message-passing layer kernel alone at P19 shape, >= 1 GiB of traffic, for `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raindrop_b200 import functional as RF
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N, C = 34, 240
x = torch.randn(B * N, C, device="cuda"); W = torch.randn(C, C, device="cuda") / C ** 0.5
b = torch.zeros(C, device="cuda"); s = torch.ones(N, device="cuda")
for _ in range(3):
    RF.ObPropLayerFunction.apply(x, W, b, s, N)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    RF.ObPropLayerFunction.apply(x, W, b, s, N)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
