#!/bin/bash
# GPU side of the profile evidence (run under gpurun, one GPU): per-launch ncu list of the training step (eager launches),
# ncu --set full of the dominant kernel's largest launch (conv_stream_kernel, decoder.convtsp3.0 fprop), of the halo
# weight-gradient kernel and of a SepConv3d fprop, torch.profiler step table.  tools/summarize_profiles.py turns the
# artefacts in gpurun_out/ into profiles/rN_*.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 650 --csv --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches_run.log 2>&1
wc -l gpurun_out/launches_step.csv
# one_layer.py runs fprop, wgrad, dgrad (one launch per temporal phase) per iteration: -s skips the first iteration
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_stream -s 6 -c 1 -f -o gpurun_out/prof_tsp3_fprop \
    python tools/one_layer.py convtsp3 2 > gpurun_out/p1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_halo -s 1 -c 1 -f -o gpurun_out/prof_tsp3_wgrad \
    python tools/one_layer.py convtsp3 2 > gpurun_out/p2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_stream -s 2 -c 1 -f -o gpurun_out/prof_b13s_fprop \
    python tools/one_layer.py base1.3.conv_s 2 > gpurun_out/p3.log 2>&1
timeout 150 python tools/profile_step.py 8 > gpurun_out/profile_step.log 2>&1
tail -28 gpurun_out/profile_step.log
