#!/bin/bash
# Batch-size / batches-in-flight sweep of bench.py on one B200 (results under gpurun_out/).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for b in 128 256; do
  python bench.py --batch $b --steps 4 --no-cpu-baseline --no-roofline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
done
for f in 2 3; do
  python bench.py --in-flight $f --steps 12 --no-cpu-baseline --no-roofline > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err
done
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'))
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
