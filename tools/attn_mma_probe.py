"""S = Q K^T issue-time probe of the attention forward kernel under RD_ATTN_DBG modes (timing only)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for m in range(6):
        env = dict(os.environ, RD_ATTN_DBG=str(m))
        print(subprocess.run([sys.executable, __file__, str(m)], env=env, capture_output=True, text=True).stdout.strip())
    sys.exit(0)
import torch
from raindrop_b200 import lib as L
lib = L.load()
B, H, T, hd = 128, 2, 60, 76
D = H * hd
qkv = torch.randn(T, B, 3 * D, device="cuda"); ctx = torch.empty(T, B, D, device="cuda")
lengths = torch.randint(2, T + 1, (B,), device="cuda")
rng = torch.tensor([1, 2], dtype=torch.int64, device="cuda")
def run():
    L.check(lib.rd_temporal_attention_fwd(qkv.data_ptr(), lengths.data_ptr(), B, H, T, hd, 0.2, rng.data_ptr(), 16, 1, ctx.data_ptr(), L.stream_ptr()), "fwd")
for _ in range(3): run()
dbg = torch.zeros(B * H, 16, dtype=torch.int64, device="cuda")
lib.rd_debug_attention_timing(dbg.data_ptr()); run(); torch.cuda.synchronize(); lib.rd_debug_attention_timing(None)
d = dbg.cpu().double()
names = {0: "3 MMAs x 10 k-steps (product)", 1: "hi.hi only (10 MMAs)", 2: "alternating accumulators", 3: "issued twice (60 MMAs)", 4: "three accumulators round-robin", 5: "N = 128 instead of 64"}
print("mode %s %-32s lo+sync -> S issued %.2f us, -> S done %.2f us" % (sys.argv[1], names[int(sys.argv[1])], ((d[:, 4] - d[:, 3]).median()) / 1e3, ((d[:, 5] - d[:, 3]).median()) / 1e3))
