#!/bin/bash
# compute-sanitizer evidence: memcheck / racecheck / synccheck over the hand-rolled mbarrier / TMEM kernels on small
# shapes (one conv scenario with concat + upsample, one Mixed block, the stem, a tiny whole-model train step).
mkdir -p gpurun_out
SEL='conv_concat_relu_upsample and bf16 and 0 or mixed_block and bf16 and 3b or stem_sepconv and bf16'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
