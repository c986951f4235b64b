#!/bin/bash
# BASELINE.json config 5 (restated per SURVEY §8d): {128x192, 224x384, 448x768} x T {16, 32, 48}, train step, on the GPUs given.
# usage: tools/run_sweep.sh <ngpus> <outfile>
N=${1:-1}; OUT=${2:-gpurun_out/r2_sweep_n$N.jsonl}
: > "$OUT"
for hw in "128 192" "224 384" "448 768"; do
  set -- $hw
  for T in 16 32 48; do
    B=8; if [ "$1" = "448" ] && [ "$T" != "16" ]; then B=4; fi
    if [ "$N" = "1" ]; then
      timeout 300 python bench.py --steps 5 --warmup 3 --clip-len $T --height $1 --width $2 --batch $B --no-cpu-baseline --no-parity 2>/dev/null | tail -1 >> "$OUT"
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 --clip-len $T --height $1 --width $2 --batch $B 2>/dev/null | tail -1 >> "$OUT"
    fi
  done
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
        r = d.get("roofline") or {}
        print("%-34s B/gpu %s n %d: %8.1f clips/s  e2e %8.1f  %.3f of sustained bf16" % (d["metric"], d["config"]["workload"].split("batch ")[1].split(" x")[0], d["n_gpus"], d["value"], d["e2e"]["value"], r.get("step_frac_of_sustained", r.get("frac", 0))))
    except Exception as e:
        print("bad line", l[:100], e)
PY
