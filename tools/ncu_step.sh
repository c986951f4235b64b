#!/bin/bash
# ncu launch list of one training step (cold-cache, serialised per-launch times): compare SHARES, not absolutes.
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3990 -c 1340 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log | cut -c1-300; wc -l gpurun_out/launches_step.csv
