#!/bin/bash
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900 -m gpu"
timeout 1200 $PY tests/test_gpu_kernels.py > gpurun_out/pytest_k.log 2>&1
echo "=== kernels"; tail -12 gpurun_out/pytest_k.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table_r1.json > gpurun_out/bench_quick.log 2>&1
echo "=== bench quick"; tail -5 gpurun_out/bench_quick.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "=== ncu"; tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
