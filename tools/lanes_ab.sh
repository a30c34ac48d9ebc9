#!/bin/bash
# lanes x chains A/B of the default bench (run-to-run noise is ~0.2 % since results are no longer kept alive)
out=${1:-gpurun_out/ab}; mkdir -p $out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-roofline"
for cfg in "2 3" "3 2" "3 1" "2 2" "4 1" "2 4" "3 3" "4 2" "2 6"; do
  set -- $cfg
  $B --in-flight $1 --chains $2 > $out/f$1c$2.json 2> $out/f$1c$2.err
  python -c "import json; d=json.load(open('$out/f$1c$2.json')); print('lanes $1 chains $2: %.0f (e2e %.0f) %.2f ms/pass' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done | tee $out/summary.txt
