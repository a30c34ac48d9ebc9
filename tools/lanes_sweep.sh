#!/bin/bash
# Lanes x chains sweep of the default benchmark (BASELINE configs[1]); prints motion-s/s per setting.
out=${1:-gpurun_out/sweep}
mkdir -p "$out"
for f in 2 3 4; do
  for c in 1 2 3 5; do
    timeout 200 python bench.py --steps 6 --warmup 2 --in-flight $f --chains $c --no-cpu-baseline --no-gpu-eager --no-roofline \
      > "$out/f${f}_c${c}.json" 2> "$out/f${f}_c${c}.err"
    python -c "import json,sys; d=json.load(open('$out/f${f}_c${c}.json')); print('lanes $f chains $c: %.0f motion-s/s (e2e %.0f), %.2f ms per pass' % (d['value'], d['e2e']['value'], d['ms_per_step']))" 2>/dev/null || echo "lanes $f chains $c: failed"
  done
done | tee "$out/summary.txt"
