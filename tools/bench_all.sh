#!/bin/bash
# One bench.py line per BASELINE configuration on ONE GPU (per-GPU batch of the named configuration) -> gpurun_out/
for cfg in P12 PAM P19x4 LARGEx8; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg_$cfg.log 2>&1
  tail -c 300 gpurun_out/r2_bench_cfg_$cfg.log
done
