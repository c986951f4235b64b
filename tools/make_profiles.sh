#!/bin/bash
# GPU side of the profile evidence (run under gpurun, one GPU): per-launch ncu list of the training step (eager launches) with
# DRAM bytes, ncu --set full of the dominant kernel's launches that DESIGN.md quotes, torch.profiler step table.
# tools/summarize_profiles.py rN turns gpurun_out/ into profiles/rN_*.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1700 -c 480 --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-graph > gpurun_out/launches_run.log 2>&1
wc -l gpurun_out/launches_step.csv
# one_layer.py runs fprop, wgrad, dgrad (one launch per temporal phase) per iteration: -s skips the first iteration
cap() {  # name, kernel regex, skip, output
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/$4 python tools/one_layer.py $1 2 > gpurun_out/$4.log 2>&1
}
# the decoder convolutions as the step runs them: source 0 read THROUGH relu + 2x bilinear (interpolating producer warps);
# up2_bench.py --only-fused --iters 1 launches fprop twice (pack + timed) and the weight gradient twice
capup() {  # layer, kernel regex, output
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/$3 python tools/up2_bench.py $1 --iters 1 --only-fused > gpurun_out/$3.log 2>&1
}
capup convtsp3 conv_stream prof_tsp3_fprop
capup convtsp2 conv_stream prof_tsp2_fprop
capup convtsp3 conv_wgrad_halo prof_tsp3_wgrad
capup convtsp4.3 conv_stream prof_tsp43_fprop
cap base1.0.conv_s conv_stream 1 prof_stem_fprop
cap base1.3.conv_s conv_stream 2 prof_b13s_fprop
cap base1.3.conv_t conv_stream 2 prof_b13t_fprop
cap 3c.b1.conv_s conv_stream 2 prof_3cs_fprop
cap 3c.b1.conv_t conv_stream 2 prof_3ct_fprop
timeout 150 python tools/profile_step.py 8 > gpurun_out/profile_step.log 2>&1
tail -36 gpurun_out/profile_step.log
timeout 200 python tools/up2_bench.py > gpurun_out/up2_bench.txt 2>&1
cat gpurun_out/up2_bench.txt
ls -la gpurun_out/*.ncu-rep
