#!/bin/bash
# One-GPU verification of the current build: the GPU test suite, smoke(), and the default bench with / without the deferred stem BatchNorm.
mkdir -p gpurun_out
( timeout 420 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/vh_tests.log 2>&1; echo "rc=$?" >> gpurun_out/vh_tests.log ); tail -5 gpurun_out/vh_tests.log
timeout 150 python bench.py --no-cpu-baseline --no-parity > gpurun_out/vh_bench_defer.json 2> gpurun_out/vh_bench_defer.err; tail -c 400 gpurun_out/vh_bench_defer.json; echo
VINET_NO_DEFER_BN=1 timeout 150 python bench.py --no-cpu-baseline --no-parity > gpurun_out/vh_bench_nodefer.json 2> gpurun_out/vh_bench_nodefer.err; tail -c 400 gpurun_out/vh_bench_nodefer.json; echo
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/vh_smoke.log 2>&1; tail -5 gpurun_out/vh_smoke.log
