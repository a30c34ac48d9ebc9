#!/bin/bash
# Linear step graphs (one chain, no side streams) with several independent batches in flight
mkdir -p gpurun_out
run() { name=$1; shift; CFB_CHAINS=$CH python bench.py --steps 12 --no-cpu-baseline --no-roofline "$@" > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err; }
CH=1 run b8_c1_f1 --batch 8
CH=1 run b8_c1_f2 --batch 8 --in-flight 2
CH=1 run b64_c1_f1
CH=1 run b64_c1_f2 --in-flight 2
CH=1 run b64_c1_f3 --in-flight 3
CH=1 run b64_c1_f4 --in-flight 4
CH=1 run b64_c1_f6 --in-flight 6
CH=2 run b64_c2_f3 --in-flight 3
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/e_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'), d['gpu_launches'])
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
