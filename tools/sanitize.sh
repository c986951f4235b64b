#!/bin/bash
# compute-sanitizer evidence.  memcheck: every kernel test (hand-rolled mbarrier / TMEM / TMA conv kernels in all modes
# incl. the epilogue-statistics path, the interpolating-producer (VINET_XF_UP2) variants and the un-swizzled stem mode, BatchNorm,
# pools, upsample, cluster-launched losses, audio GEMMs), the input / output pipeline kernels and one whole-model train step on the
# tensor-core parity mode.  racecheck / synccheck: the kernel tests (the whole-model step exhausts the tools' own memory).
# Logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_preprocess.py \
    "tests/test_gpu_inference.py::test_postprocess_matches_reference_pipeline" "tests/test_gpu_parity_tc.py::test_two_term_split_is_close_but_reported_separately" \
    -m gpu -q -x > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
done
