#!/bin/bash
# Independent batches in flight, one host thread + stream + sampler handle per lane
mkdir -p gpurun_out
run() { name=$1; shift; CFB_CHAINS=$CH timeout 300 python bench.py --steps 12 --no-cpu-baseline --no-roofline "$@" > gpurun_out/f_$name.json 2> gpurun_out/f_$name.err; }
CH=1 run b8_c1_f2 --batch 8 --in-flight 2
CH=6 run b64_c6_f2 --in-flight 2
CH=6 run b64_c6_f3 --in-flight 3
CH=3 run b64_c3_f2 --in-flight 2
CH=3 run b64_c3_f4 --in-flight 4
CH=1 run b64_c1_f4 --in-flight 4
CH=1 run b64_c1_f6 --in-flight 6
CH=2 run b64_c2_f3 --in-flight 3
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/f_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'), d['gpu_launches'])
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
