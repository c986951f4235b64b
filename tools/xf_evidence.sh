#!/bin/bash
# Sanitizer + ncu evidence for the transformer-fusion kernels (csrc/xfmr.cu) and the ordered split-K of the audio forward GEMM.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 170 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_transformer.py -m gpu -q -x -k "block_vs_torch or dropout_kernels" \
      > gpurun_out/sanitizer_xfmr_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_xfmr_$tool.log
done
# the QKV projection of the fusion model's token block (M = 678 token rows, N = 1536, K = 512): third bgemm launch of the scenario
timeout 120 ncu --set full --clock-control none --import-source on -k regex:bgemm -s 2 -c 1 -f -o gpurun_out/prof_bgemm_qkv python tools/xf_debug.py 512 > gpurun_out/prof_bgemm_qkv.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
