#!/bin/bash
# Do two captured step graphs on different streams overlap at all?  (small batches leave the GPU mostly idle)
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py --steps 8 --no-cpu-baseline --no-roofline "$@" > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err; }
run b8_f1 --batch 8
run b8_f2 --batch 8 --in-flight 2
run b8_f4 --batch 8 --in-flight 4
run b16_f1 --batch 16
run b16_f2 --batch 16 --in-flight 2
run b32_f1 --batch 32
run b32_f2 --batch 32 --in-flight 2
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/d_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'))
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
