#!/usr/bin/env python
"""Benchmark of the Raindrop hot path on B200 (driver contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one training step of Raindrop_v2 (forward + CrossEntropy + backward + Adam, dropout 0.2,
code/Raindrop.py:311-324) on one batch of P19-shape synthetic data (BASELINE.json configs[1]:
B = 128 samples per GPU, 34 sensors, T_max = 60).  Weak scaling: every rank owns its own 128 samples;
the only collective is one NCCL all-reduce over the flat gradient bucket.

Printed JSON line (rank 0):
  value     samples/s, whole job, inputs resident in HBM, step = one CUDA-graph replay of TrainStep
  e2e       samples/s through the drop-in nn.Module API (model.forward -> criterion -> backward ->
            optimizer.step) with pinned HOST batches, H2D copies and the loss read-back inside the
            timed region
  roofline  the observation-propagation (message-passing) layer kernel: algorithmic bytes
            8*N*C per (sample, layer) / CUDA-event time, at a row count with >= 1 GiB of traffic
            (`rows`) and at the configuration's own batch (`at_config`)
  cpu_baseline  the CPU restatement of the reference (oracle/, same per-sample loop and per-edge
            GEMMs as code/models_rd.py:322-343) timed on this box's host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import warnings  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from raindrop_b200.synth import make_batch, model_config, synth_weights  # noqa: E402

warnings.filterwarnings("ignore")
CFG_NAME = "P19"
BATCH = 128
L2_FLUSH_BYTES = 256 << 20   # > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (rank 0)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n_gpus > 1:
        raise SystemExit("--gpus %d needs torchrun (one process per GPU)" % n_gpus)
    return world, rank, local


def build_model(cfg, device):
    from raindrop_b200.models_rd import Raindrop_v2
    torch.manual_seed(1)   # code/Raindrop.py:58
    gs = torch.ones(cfg["d_inp"], cfg["d_inp"])
    m = Raindrop_v2(cfg["d_inp"], cfg["d_model"], cfg["nhead"], cfg["nhid"], cfg["nlayers"], cfg["dropout"],
                    cfg["max_len"], cfg["d_static"], cfg["MAX"], 0.5, "mean", cfg["n_classes"], gs)
    synth_weights(m, cfg, seed=7)      # random-init weights of the named architecture, same on every rank
    return m.to(device).train()


def flush_l2(buf):
    """Write a buffer larger than L2, then read it back: the write evicts everything, the read leaves the
    cache full of CLEAN lines (otherwise the timed kernel pays for writing back ~126 MB of dirty zeros)."""
    buf.zero_()
    buf.sum()


def timed_steps(step_fn, steps, flush_buf):
    """Per-step CUDA-event timing; the L2 flush between steps sits outside the event pairs."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        flush_l2(flush_buf)
        s.record()
        step_fn()
        e.record()
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in ev]


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _round_tf32(t):
    """RN to TF32 with integer ops (bench-side data prep, so the timed launch is the GEMM kernel alone)."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def roofline_leg(cfg, device):
    """Message-passing layer kernel alone (rd_obprop_fwd on TF32-exact operands, no rounding pre-pass):
    algorithmic bytes = read x[rows,C] + write out[rows,C]."""
    from raindrop_b200 import lib as L
    lib = L.load()
    N, C = cfg["d_inp"], cfg["max_len"] * cfg["d_ob"]
    peak, how = peaks()
    out = {}
    for tag, B in (("large", 16384), ("at_config", BATCH)):
        rows = B * N
        x = _round_tf32(torch.randn(rows, C, device=device))
        W = _round_tf32(torch.randn(C, C, device=device) / C ** 0.5)
        b = torch.zeros(C, device=device)
        s = torch.ones(N, device=device)
        y = torch.empty_like(x)

        def fn():
            L.check(lib.rd_obprop_fwd(x.data_ptr(), W.data_ptr(), b.data_ptr(), s.data_ptr(), N, rows, C, y.data_ptr(),
                                      None, L.stream_ptr()), "rd_obprop_fwd")
        for _ in range(3):
            fn()
        flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=device)
        ts = timed_steps(fn, 20, flush)
        ms = sum(ts) / len(ts)
        gb = rows * C * 8 / 1e9
        # the same launch back to back (how MEASURED_PEAKS.json's copy bandwidth was taken); at `large` the
        # 1.07 GB working set cannot stay in the 126 MB L2 either way
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms_b2b = e0.elapsed_time(e1) / 10
        out[tag] = dict(rows=rows, ms=ms, achieved=gb / (ms * 1e-3), frac=gb / (ms * 1e-3) / peak,
                        tflops=2.0 * rows * C * C / (ms * 1e-3) / 1e12, ms_b2b=ms_b2b, frac_b2b=gb / (ms_b2b * 1e-3) / peak)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "obprop_tc_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    big = out["large"]
    return {"kernel": "obprop_tc_kernel (tcgen05 TF32 + TMA, one ob-prop layer)", "bound": "hbm",
            "achieved": round(big["achieved"], 1), "peak": peak, "peak_source": how, "unit": "GB/s",
            "frac": round(big["frac"], 4), "traffic": traffic, "rows": big["rows"],
            "algorithmic_bytes_per_launch": big["rows"] * C * 8, "ms_per_launch": round(big["ms"], 5),
            "tflops_tf32": round(big["tflops"], 1),
            "back_to_back": {"ms_per_launch": round(big["ms_b2b"], 5), "frac": round(big["frac_b2b"], 4),
                             "note": "no L2 flush between launches (same protocol as the measured copy peak)"},
            "at_config": {"rows": out["at_config"]["rows"], "ms_per_launch": round(out["at_config"]["ms"], 5),
                          "achieved": round(out["at_config"]["achieved"], 1), "frac": round(out["at_config"]["frac"], 4),
                          "note": "4352 rows = 34 tiles on 148 SMs, 8 MB: launch/latency bound, L2-sized"}}


def cpu_reference_leg(cfg, steps, warmup, budget_s=25.0):
    """The CPU restatement of the reference (oracle/): same per-sample Python loop with per-edge
    lin_value GEMMs and torch.nn.TransformerEncoder as code/models_rd.py:322-358, train mode
    (dropout 0.2), CrossEntropy + backward + Adam(lr=1e-4) like code/Raindrop.py:319-324."""
    from oracle.raindrop_oracle import build_oracle_model       # the checker, timed as the baseline
    model = build_oracle_model(cfg).train()
    synth_weights(model, cfg, seed=7)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    batch = make_batch(cfg, BATCH, seed=1000 * 2)
    # the path is thousands of tiny ops: more threads is not faster.  Probe a few thread counts on a
    # small forward and keep the best, so the baseline gets the host's best configuration.
    ncpu = os.cpu_count() or 1
    probe = make_batch(cfg, 8, seed=5)
    best_t, threads = None, 1
    for cand in sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(cand)
        with torch.no_grad():
            model.forward(probe["src"], probe["static"], probe["times"], probe["lengths"])
            t0 = time.perf_counter()
            model.forward(probe["src"], probe["static"], probe["times"], probe["lengths"])
            dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, threads = dt, cand
    torch.set_num_threads(threads)

    def one():
        logits, _, _ = model.forward(batch["src"], batch["static"], batch["times"], batch["lengths"])
        loss = F.cross_entropy(logits, batch["y"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(max(1, warmup)):
        one()
    ts = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    sec = sum(ts) / len(ts)
    return {"value": round(BATCH / sec, 2), "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": "%d train steps (fwd+CE+bwd+Adam, dropout 0.2) of P19 B=%d after %d warm-up, %.2f s/step; "
                      "torch %s CPU, %d threads (best of a probe over thread counts; host has %d logical CPUs)"
                      % (len(ts), BATCH, max(1, warmup), sec, torch.__version__, threads, ncpu),
            "sec_per_step": sec, "steps": len(ts)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_config(CFG_NAME, dropout=0.2)
    cb = cpu_reference_leg(cfg, steps=max(1, min(args.steps, 20)), warmup=min(args.warmup, 2), budget_s=120.0)
    line = {"impl": "reference", "metric": "samples/sec (P19-shape synthetic) training step", "value": cb["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": cb["steps"], "warmup": min(args.warmup, 2),
            "ms_per_step": round(cb["sec_per_step"] * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "P19 synthetic (batch=128, 34 sensors, T_max=60) Raindrop_v2 training step",
                       "global_batch": BATCH, "note": "reference is CPU-only here: single process, host cores"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    from raindrop_b200 import lib as L
    from raindrop_b200.train import TrainStep, allreduce_gradients
    world, rank, local = dist_setup(args.gpus)
    device = torch.device("cuda", local if world > 1 else 0)
    torch.cuda.set_device(device)
    cfg = model_config(CFG_NAME, dropout=0.2)
    lib = L.load()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=device)

    # ---- leg 1: device-resident TrainStep, one CUDA graph per step ------------------------------
    model = build_model(cfg, device)
    ts = TrainStep(model, BATCH, lr=1e-4, use_graph=True)
    host_batches = [make_batch(cfg, BATCH, seed=1000 * 2 + 17 * rank + i, pin=True) for i in range(4)]
    ts.load_batch(host_batches[0])
    n0 = lib.rd_launch_count()
    ts._enqueue()                      # eager once: counts our launches per step
    launches_per_step = int(lib.rd_launch_count() - n0)
    graph_note = "one CUDA-graph replay (TrainStep)"
    try:
        ts.capture(warmup=2)
    except Exception as exc:  # noqa: BLE001  (e.g. NCCL refusing stream capture): run the same step eagerly
        ts.use_graph, ts.graph = False, None
        graph_note = "eager TrainStep (graph capture failed: %r)" % (exc,)
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        ts.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(device.index or 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    t_wall = time.perf_counter()
    per_step = timed_steps(ts.step, args.steps, flush)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(sum(per_step), world, device)
    ms_per_step = dev_ms / args.steps
    value = world * BATCH * args.steps / (dev_ms * 1e-3)
    loss_graph = float(ts.loss.item())

    # ---- leg 2: end to end through the drop-in module API, host batches ---------------------------
    model2 = build_model(cfg, device)
    opt = torch.optim.Adam(model2.parameters(), lr=1e-4)
    crit = torch.nn.CrossEntropyLoss()
    h2d = sum(t.numel() * t.element_size() for k, t in host_batches[0].items() if t is not None)
    state = {"i": 0, "loss": 0.0}

    def e2e_step():
        hb = host_batches[state["i"] % len(host_batches)]
        state["i"] += 1
        P = hb["src"].to(device, non_blocking=True)
        Pt = hb["times"].to(device, non_blocking=True)
        Ps = hb["static"].to(device, non_blocking=True)
        y = hb["y"].to(device, non_blocking=True)
        lengths = torch.sum(Pt > 0, dim=0)                         # code/Raindrop.py:317
        outputs, _, _ = model2.forward(P, Ps, Pt, lengths)         # code/Raindrop.py:319
        opt.zero_grad()
        loss = crit(outputs, y)
        loss.backward()
        allreduce_gradients(model2)
        opt.step()
        state["loss"] = loss.item()                                # D2H read of the step's result

    for _ in range(args.warmup):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_times = timed_steps(e2e_step, args.steps, flush)
    if world > 1:
        dist.barrier()
    e2e_ms = max_over_ranks(sum(e2e_times), world, device)
    e2e_value = world * BATCH * args.steps / (e2e_ms * 1e-3)

    if rank != 0:
        _finish(world)
        return
    roof = roofline_leg(cfg, device)
    line = {
        "metric": "samples/sec (P19-shape synthetic) training step", "value": round(value, 1), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "P19 synthetic (batch=128 per GPU, 34 sensors, T_max=60) Raindrop_v2 training step: "
                               "fwd + CrossEntropy + bwd + Adam, dropout 0.2", "global_batch": world * BATCH,
                   "per_gpu_batch": BATCH,
                   "precision": "fp32 storage and accumulation; ob-prop GEMM operands rounded to TF32 (forward error 3e-4), "
                                "encoder GEMMs error-compensated 3xTF32 (fp32-level)", "parallelism": "sample-sharded dp%d, 1 NCCL all-reduce of the flat grad bucket" % world,
                   "l2": "flushed between timed steps (256 MiB write + read-back, outside the per-step CUDA-event pairs)",
                   "step": graph_note, "wall_ms_per_step_incl_flush": round(wall / args.steps * 1e3, 4)},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": round(e2e_ms / args.steps, 4),
                "path": "pinned host batch -> H2D -> models_rd.Raindrop_v2.forward -> CrossEntropyLoss -> backward -> "
                        "torch.optim.Adam.step -> loss.item()"},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "final_loss": {"graph": round(loss_graph, 5), "e2e": round(state["loss"], 5)},
    }
    if world == 1 and not args.no_cpu_baseline:
        # separate process (own thread pool, hard time limit) so a slow host cannot stall the bench
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5",
                                "--warmup", "1"], capture_output=True, text=True, timeout=240)
            ref = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "cpu leg failed or timed out: %r" % (exc,)}
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """destroy_process_group() was observed to hang on this pool after NCCL work has been captured in a
    CUDA graph; every rank is done and synchronised here, so leave without tearing the communicator down."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
