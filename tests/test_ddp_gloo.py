"""CPU, world_size 2, gloo: the sample-sharded data-parallel host logic.
Each rank computes the gradient of ITS shard with the CPU oracle (our kernels need a GPU), exposes it
the way RaindropV2Function.backward does (views into one flat bucket), and `allreduce_gradients` must
turn it into the full-batch gradient with ONE collective; the gather fallback is exercised too."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _Shim(torch.nn.Module):
    """Duck-types what allreduce_gradients needs from Raindrop_v2: used_parameters() and _flat_grad."""

    def __init__(self, oracle, keys):
        super().__init__()
        self.oracle, self.keys, self._flat_grad = oracle, keys, None

    def used_parameters(self):
        sd = dict(self.oracle.named_parameters())
        return [sd[k] for k in self.keys]


def _get(q, procs, timeout=900.0):
    """q.get() that cannot hang: gives up when a worker died or the deadline passed (slow first `import torch` under load)"""
    import time
    t0 = time.time()
    while q.empty():
        dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
        assert not dead, "worker exited with %s" % dead
        assert time.time() - t0 < timeout, "workers produced nothing within %.0f s" % timeout
        time.sleep(0.05)
    return q.get()


def _worker(rank, world, port, alias, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.raindrop_oracle import build_oracle_model
    from raindrop_b200.synth import make_batch, model_config, synth_weights, used_param_keys
    from raindrop_b200.train import allreduce_gradients
    cfg = model_config("TINY", dropout=0.0)
    B = 8
    full = make_batch(cfg, B, seed=3)
    oracle = build_oracle_model(cfg).eval()
    synth_weights(oracle, cfg, seed=5)
    keys = used_param_keys(cfg)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    logits, _, _ = oracle.forward_dense(full["src"][:, sl], full["static"][sl], full["times"][:, sl], full["lengths"][sl])
    F.cross_entropy(logits, full["y"][sl]).backward()
    shim = _Shim(oracle, keys)
    params = shim.used_parameters()
    if alias:   # gradients are views of one flat bucket, like the CUDA backward hands them to autograd
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        shim._flat_grad = flat
    allreduce_gradients(shim)
    if rank == 0:
        # numpy, not torch tensors: tensors travel as shared-memory handles served by THIS process, which may have exited
        # before the parent unpickles them (FileNotFoundError on the handle listener)
        out_q.put({k: p.grad.detach().numpy().copy() for k, p in zip(keys, params)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("alias", [True, False])
def test_two_rank_gradients_equal_full_batch(alias):
    from oracle.raindrop_oracle import build_oracle_model
    from raindrop_b200.synth import make_batch, model_config, synth_weights, used_param_keys
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, alias, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = _get(q, procs)
    for p in procs:
        p.join(timeout=900)
        assert p.exitcode == 0
    cfg = model_config("TINY", dropout=0.0)
    full = make_batch(cfg, 8, seed=3)
    oracle = build_oracle_model(cfg).eval()
    synth_weights(oracle, cfg, seed=5)
    logits, _, _ = oracle.forward_dense(full["src"], full["static"], full["times"], full["lengths"])
    F.cross_entropy(logits, full["y"]).backward()
    ref = dict(oracle.named_parameters())
    for k in used_param_keys(cfg):
        assert torch.allclose(torch.from_numpy(got[k]), ref[k].grad, rtol=1e-4, atol=1e-7), k


def _eval_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from raindrop_b200.train import evaluate_sharded

    class Stub(torch.nn.Module):            # stands in for Raindrop_v2 on CPU: any per-sample function will do
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.arange(6.0).view(3, 2))

        def forward(self, P, Pstatic, Ptime, lengths):
            feat = torch.stack([P.sum((0, 2)), Ptime.sum(0), lengths.float()], 1)
            return feat @ self.w, torch.zeros(()), None

    g = torch.Generator().manual_seed(0)
    P, Pt = torch.randn(5, 11, 4, generator=g), torch.rand(5, 11, generator=g)
    out = evaluate_sharded(Stub(), P, None, Pt)
    if rank == 0:
        out_q.put(out.detach().numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_gathers_the_whole_set_in_order():
    """evaluate_sharded (uneven shards: 11 samples on 2 ranks) == the single-process result, same order."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = _get(q, procs)
    for p in procs:
        p.join(timeout=900)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    P, Pt = torch.randn(5, 11, 4, generator=g), torch.rand(5, 11, generator=g)
    feat = torch.stack([P.sum((0, 2)), Pt.sum(0), (Pt > 0).sum(0).float()], 1)
    got = torch.from_numpy(got)
    assert got.shape == (11, 2) and torch.allclose(got, feat @ torch.arange(6.0).view(3, 2), atol=1e-5)
