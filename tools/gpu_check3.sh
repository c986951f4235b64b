#!/bin/bash
# One GPU-box visit: TMA kernel diagnostics (isolated subprocess), GPU test-suite, quick bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python tools/diag_tma.py > gpurun_out/diag_tma.log 2>&1
echo "=== diag_tma"; tail -60 gpurun_out/diag_tma.log
PY="python -m pytest -q -p no:cacheprovider --timeout 900 -m gpu -x"
timeout 1500 $PY tests > gpurun_out/pytest_gpu.log 2>&1
echo "=== pytest gpu"; tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.json > gpurun_out/bench_quick.log 2>&1
echo "=== bench quick"; tail -5 gpurun_out/bench_quick.log
timeout 600 python tools/profile_step.py 8 > gpurun_out/profile_step.log 2>&1
echo "=== profile"; tail -32 gpurun_out/profile_step.log
