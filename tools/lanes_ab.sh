#!/bin/bash
out=gpurun_out/r2w; mkdir -p $out
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-roofline"
for rep in 1 2 3; do
  for cfg in "2 3" "3 2" "3 1" "2 2" "4 1"; do
    set -- $cfg
    $B --in-flight $1 --chains $2 > $out/f$1c$2_$rep.json 2> $out/f$1c$2_$rep.err
    python -c "import json; d=json.load(open('$out/f$1c$2_$rep.json')); print('rep $rep lanes $1 chains $2: %.0f (e2e %.0f) %.2f ms' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
  done
done | tee $out/summary.txt
