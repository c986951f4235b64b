#!/bin/bash
# GPU side of the profile evidence (run under gpurun): per-launch ncu list of one training step, ncu --set full of
# the dominant conv launch (decoder.convtsp3.0 fprop) and of the wgrad kernel, torch.profiler step table.
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 800 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
wc -l gpurun_out/launches_step.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tma -s 6 -c 1 -o gpurun_out/prof_tsp3_fprop python tools/one_layer.py convtsp3 2 > gpurun_out/ncu_tsp3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tma -s 1 -c 1 -o gpurun_out/prof_tsp3_wgrad python tools/one_layer.py convtsp3 2 >> gpurun_out/ncu_tsp3.log 2>&1
tail -2 gpurun_out/ncu_tsp3.log
timeout 600 python tools/profile_step.py 8 > gpurun_out/profile_step.log 2>&1
tail -28 gpurun_out/profile_step.log
