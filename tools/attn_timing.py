"""Phase timeline of the tensor-core attention forward kernel (%globaltimer stamps per CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
lib.rd_debug_attention_timing(dbg.data_ptr())
run(); torch.cuda.synchronize()
lib.rd_debug_attention_timing(None)
d = dbg.cpu().double() / 1.965      # SM cycle counter (per-SM, not comparable across CTAs) -> ns at 1965 MHz
t0 = d[:, 0].min()
names = ["setup done", "keep bits", "Q,K landed", "lo(Q,K)+sync", "S issued", "S done", "softmax+P stored", "V landed", "lo(V)+sync", "O issued", "O done", "ctx stored", "exit"]
print("per-CTA setup -> exit: median %.2f us, max %.2f us" % ((d[:, 12] - d[:, 0]).median() / 1e3, (d[:, 12] - d[:, 0]).max() / 1e3))
for i, n in enumerate(names):
    col = d[:, i] - d[:, 0]
    print("%-18s median +%.2f us   (p10 %.2f, p90 %.2f)" % (n, col.median() / 1e3, col.quantile(0.1) / 1e3, col.quantile(0.9) / 1e3))

# ---- backward ----------------------------------------------------------------------------------------------
dctx = torch.randn(T, B, D, device="cuda"); dqkv = torch.empty(T, B, 3 * D, device="cuda")
def runb():
    L.check(lib.rd_temporal_attention_bwd(qkv.data_ptr(), dctx.data_ptr(), lengths.data_ptr(), B, H, T, hd, 0.2, rng.data_ptr(), 16, 1, dqkv.data_ptr(), L.stream_ptr()), "bwd")
for _ in range(3): runb()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [run() for _ in range(20)]; e1.record(); torch.cuda.synchronize()
print("\nforward  %.2f us per call (back to back)" % (e0.elapsed_time(e1) / 20 * 1e3))
e0.record(); [runb() for _ in range(20)]; e1.record(); torch.cuda.synchronize()
print("backward %.2f us per call (back to back)" % (e0.elapsed_time(e1) / 20 * 1e3))
dbg = torch.zeros(B * H, 16, dtype=torch.int64, device="cuda")
lib.rd_debug_attention_timing(dbg.data_ptr()); runb(); torch.cuda.synchronize(); lib.rd_debug_attention_timing(None)
d = dbg.cpu().double() / 1.965; t0 = d[:, 0].min()
names = ["setup done", "Q,K landed", "S issued", "dP issued", "S,dP done", "softmax/dS stored", "dO,K (MN) landed", "dV,dQ issued",
         "dV,dQ done; Q TMA", "dQ,dV stored", "Q (MN) landed", "dK issued", "dK stored"]
print("backward per-CTA setup -> dK stored: median %.2f us, max %.2f us" % ((d[:, 12] - d[:, 0]).median() / 1e3, (d[:, 12] - d[:, 0]).max() / 1e3))
for i, n in enumerate(names):
    col = d[:, i] - d[:, 0]
    print("%-20s median +%.2f us   (p10 %.2f, p90 %.2f)" % (n, col.median() / 1e3, col.quantile(0.1) / 1e3, col.quantile(0.9) / 1e3))

# ---- projection GEMM (rd_linear_fwd): M = 7680 tokens, K = 152, N = 152 / 456 ---------------------------------
for (K, N) in ((152, 152), (152, 456), (456, 152), (272, 152)):
    x = torch.randn(7680, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.1; b = torch.zeros(N, device="cuda")
    y = torch.empty(7680, N, device="cuda")
    sc = torch.empty(lib.rd_linear_scratch_bytes(K, N) // 4, device="cuda")
    def g():
        L.check(lib.rd_linear_fwd(x.data_ptr(), W.data_ptr(), b.data_ptr(), 7680, K, N, 0, y.data_ptr(), sc.data_ptr(), L.stream_ptr()), "lin")
    for _ in range(3): g()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [g() for _ in range(20)]; e1.record(); torch.cuda.synchronize()
    print("\nlinear M=7680 K=%d N=%d: %.2f us per call (incl. the split_weights launch)" % (K, N, e0.elapsed_time(e1) / 20 * 1e3))
    dbg = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
    lib.rd_debug_gemm_timing(dbg.data_ptr()); g(); torch.cuda.synchronize(); lib.rd_debug_gemm_timing(None)
    d = dbg.cpu().double(); d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    print("  CTAs %d, kernel span %.2f us" % (len(d), (d[:, 7].max() - t0) / 1e3))
    for i, n in enumerate(["start", "setup+sync", "first k-block ready", "all MMAs issued", "accumulator complete", "epilogue math done", "stores drained", "exit"]):
        col = d[:, i] - d[:, 0]
        print("  %-22s median +%.2f us (max %.2f)" % (n, col.median() / 1e3, col.max() / 1e3))
