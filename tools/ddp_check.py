"""torchrun --nproc-per-node N tools/ddp_check.py : N-rank sample-sharded training == 1-rank training
on the concatenated batch (the gradient all-reduce is the mean of the shards).  Dropout 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from helpers import build_dropin, to_dev, rel_l2
from raindrop_b200.synth import model_config, make_batch, used_param_keys
from raindrop_b200.train import TrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = model_config("P19", dropout=0.0)
per = 32
B = per * world
batches = [make_batch(cfg, B, seed=70 + i) for i in range(3)]
for use_graph in (False, True):
    model = build_dropin(cfg, 6).train()
    ts = TrainStep(model, per, lr=1e-3, use_graph=use_graph)
    if use_graph:
        ts.load_batch(to_dev({k: (v[:, :per] if k in ("src", "times") else (v[:per] if v is not None else None)) for k, v in batches[0].items()}))
        try:
            ts.capture(warmup=0)
        except Exception as exc:  # noqa: BLE001
            print("rank %d: CUDA-graph capture with NCCL failed: %r" % (rank, exc)); sys.exit(3)
    losses = []
    for b in batches:
        sl = slice(rank * per, (rank + 1) * per)
        shard = {k: (v[:, sl] if k in ("src", "times") else (v[sl] if v is not None else None)) for k, v in b.items()}
        ts.load_batch(to_dev(shard))
        l = ts.step().clone()
        dist.all_reduce(l); losses.append(l.item() / world)
    if rank == 0:
        ref = build_dropin(cfg, 6).train()
        rs = TrainStep(ref, B, lr=1e-3, use_graph=False, distributed=False)
        ref_losses = []
        for b in batches:
            rs.load_batch(to_dev(b)); ref_losses.append(rs.step().item())
        p1, p2 = dict(ref.named_parameters()), dict(model.named_parameters())
        worst = max(rel_l2(p2[k], p1[k]) for k in used_param_keys(cfg))
        print("graph=%s world=%d losses %s vs 1-rank %s ; worst param rel_l2 %.2e" % (use_graph, world, ["%.6f" % x for x in losses], ["%.6f" % x for x in ref_losses], worst))
        assert all(abs(a - b) < 2e-4 for a, b in zip(losses, ref_losses)), (losses, ref_losses)
        assert worst < 5e-3, worst
    dist.barrier()
if rank == 0:
    print("DDP_CHECK_OK", flush=True)
torch.cuda.synchronize(); sys.stdout.flush()
os._exit(0)   # destroy_process_group() hangs on this pool once NCCL work was graph-captured
