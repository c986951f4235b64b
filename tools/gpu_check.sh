#!/bin/bash
# One GPU-box visit: diagnostics first (isolated subprocesses), then the GPU test-suite in slices so a
# sticky CUDA error in one slice cannot hide the others.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python tools/diag_tc.py > gpurun_out/diag.log 2>&1
echo "=== diag"; tail -40 gpurun_out/diag.log
PY="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu"
timeout 900 $PY tests/test_gpu_kernels.py -k "fp32 or loss" > gpurun_out/pytest_k_fp32.log 2>&1
echo "=== kernels fp32"; tail -15 gpurun_out/pytest_k_fp32.log
timeout 900 $PY tests/test_gpu_kernels.py -k "bf16" > gpurun_out/pytest_k_bf16.log 2>&1
echo "=== kernels bf16"; tail -15 gpurun_out/pytest_k_bf16.log
timeout 1200 $PY tests/test_gpu_model.py -k "fp32" > gpurun_out/pytest_m_fp32.log 2>&1
echo "=== model fp32"; tail -15 gpurun_out/pytest_m_fp32.log
timeout 1200 $PY tests/test_gpu_model.py -k "bf16" -s > gpurun_out/pytest_m_bf16.log 2>&1
echo "=== model bf16"; tail -15 gpurun_out/pytest_m_bf16.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "=== smoke"; tail -6 gpurun_out/smoke.log
