#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1
tail -5 gpurun_out/g_pytest.log
run() { name=$1; shift; CFB_CHAINS=$CH timeout 300 python bench.py --steps 12 --no-cpu-baseline --no-roofline "$@" > gpurun_out/g_$name.json 2> gpurun_out/g_$name.err; }
CH=3 run c3_f2 --in-flight 2
CH=6 run c6_f2 --in-flight 2
CH=4 run c4_f2 --in-flight 2
CH=3 run c3_f3 --in-flight 3
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/g_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'), d['gpu_launches'])
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
