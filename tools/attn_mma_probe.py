"""S = Q K^T issue-time probe of the attention forward kernel under RD_ATTN_DBG modes (timing only)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for m in (0, 14):
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
xg = torch.randn(7680, 152, device="cuda"); Wg = torch.randn(152, 152, device="cuda"); bg = torch.zeros(152, device="cuda"); yg = torch.empty(7680, 152, device="cuda")
scg = torch.empty(lib.rd_linear_scratch_bytes(152, 152) // 4, device="cuda")
def gemm():
    L.check(lib.rd_linear_fwd(xg.data_ptr(), Wg.data_ptr(), bg.data_ptr(), 7680, 152, 152, 0, yg.data_ptr(), scg.data_ptr(), L.stream_ptr()), "lin")
gemm(); torch.cuda.synchronize()
lib.rd_debug_attention_timing(dbg.data_ptr())
if sys.argv[1] == "8":
    gemm(); gemm()
run(); torch.cuda.synchronize(); lib.rd_debug_attention_timing(None)
d = dbg.cpu().double() / 1.965      # SM cycles -> ns at 1965 MHz
names = {0: "3 MMAs x 10 k-steps (product)", 1: "hi.hi only (10 MMAs)", 2: "alternating accumulators", 3: "issued twice (60 MMAs)", 4: "three accumulators round-robin", 5: "N = 128 instead of 64", 6: "per-k-step stamps", 7: "no remainder pass / proxy fence before the MMAs", 8: "tensor-core GEMM launched right before", 9: "throw-away MMA at CTA start", 13: "stamps around the first MMAs", 14: "stamps around tcgen05.fence::after_thread_sync"}
print("mode %s %-32s lo+sync -> S issued %.2f us, -> S done %.2f us" % (sys.argv[1], names[int(sys.argv[1])], ((d[:, 4] - d[:, 3]).median()) / 1e3, ((d[:, 5] - d[:, 3]).median()) / 1e3))
if 6 <= int(sys.argv[1]) <= 8:
    for slot, n in ((13, "after k-step 0 (3 MMAs)"), (14, "after k-step 4 (15 MMAs)"), (15, "after k-step 9 (30 MMAs)"), (4, "after commit")):
        print("    %-28s +%.2f us" % (n, ((d[:, slot] - d[:, 3]).median()) / 1e3))
print("    per-CTA median %.2f us" % ((d[:, 12] - d[:, 0]).median() / 1e3))
if sys.argv[1] == "13":
    for slot, n in ((13, "before the first MMA"), (14, "after MMA 1"), (15, "after MMA 3"), (4, "after commit"), (5, "S done")):
        print("    %-28s +%.2f us" % (n, ((d[:, slot] - d[:, 3]).median()) / 1e3))
if sys.argv[1] == "14":
    for slot, n in ((13, "entering the issue block"), (14, "after the tcgen05 fence"), (4, "after commit"), (5, "S done")):
        print("    %-28s +%.2f us" % (n, ((d[:, slot] - d[:, 3]).median()) / 1e3))
names13 = ["setup done", "keep bits", "Q,K landed", "lo(Q,K)+sync", "S issued", "S done", "softmax+P stored", "V landed", "lo(V)+sync", "O issued", "O done", "ctx stored", "exit"]
print("    " + "  ".join("%s %.2f" % (n, (d[:, i] - d[:, 0]).median() / 1e3) for i, n in enumerate(names13)))
