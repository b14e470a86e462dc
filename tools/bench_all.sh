#!/bin/bash
# One bench.py line per BASELINE configuration on ONE GPU (per-GPU batch of the named configuration) -> gpurun_out/r02_bench_<cfg>.json
for cfg in P19 P12 PAM P19x4 LARGEx8; do
  timeout 900 python bench.py --config $cfg --steps ${STEPS:-30} --warmup ${WARMUP:-5} > gpurun_out/r02_bench_$cfg.log 2>&1
  tail -1 gpurun_out/r02_bench_$cfg.log > gpurun_out/r02_bench_$cfg.json
  cut -c1-260 gpurun_out/r02_bench_$cfg.json
done
