import os, sys, time, datetime
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
print("rank", rank, "start", flush=True)
torch.cuda.set_device(local)
t0 = time.time()
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=60), device_id=torch.device("cuda", local))
print("rank", rank, "init done %.1fs" % (time.time() - t0), flush=True)
x = torch.ones(1 << 20, device="cuda") * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize()
print("rank", rank, "allreduce ok", x[0].item(), "%.1fs" % (time.time() - t0), flush=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): dist.all_reduce(x)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        dist.all_reduce(x)
    g.replay(); torch.cuda.synchronize()
    print("rank", rank, "graph-captured allreduce ok", flush=True)
except Exception as e:
    print("rank", rank, "graph capture failed", repr(e), flush=True)
dist.destroy_process_group()
