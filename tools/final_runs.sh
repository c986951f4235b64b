#!/bin/bash
# Round-end evidence on one GPU: full bench line (with CPU baseline + parity probe), the reference arm, the other BASELINE.json
# configs, then profiles and sanitizer logs.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python bench.py --kernel-table gpurun_out/final_ktable.json > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.json; echo
timeout 300 python bench.py --model avinet --no-cpu-baseline --no-parity > gpurun_out/final_avinet.json 2> gpurun_out/final_avinet.err; tail -c 200 gpurun_out/final_avinet.json; echo
for v in transformer tokens; do timeout 300 python bench.py --model avinet --av-fusion $v --no-cpu-baseline --no-parity > gpurun_out/final_avinet_$v.json 2> gpurun_out/final_avinet_$v.err; tail -c 200 gpurun_out/final_avinet_$v.json; echo; done
timeout 200 python tools/av_determinism.py > gpurun_out/final_av_determinism.txt 2>&1; tail -8 gpurun_out/final_av_determinism.txt
timeout 300 python bench.py --mode eval --batch 1 --no-cpu-baseline --no-parity > gpurun_out/final_eval1.json 2> gpurun_out/final_eval1.err; tail -c 200 gpurun_out/final_eval1.json; echo
timeout 300 python bench.py --mode eval --no-cpu-baseline --no-parity > gpurun_out/final_eval8.json 2> gpurun_out/final_eval8.err; tail -c 200 gpurun_out/final_eval8.json; echo
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -4 gpurun_out/final_smoke.log
bash tools/make_profiles.sh > gpurun_out/make_profiles.log 2>&1; tail -3 gpurun_out/make_profiles.log
bash tools/sanitize.sh 2>&1 | tail -12
bash tools/xf_evidence.sh 2>&1 | tail -8
