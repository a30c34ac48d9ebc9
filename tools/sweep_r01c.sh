#!/bin/bash
# CUDA_DEVICE_MAX_CONNECTIONS sweep (hardware work queues vs the 14 streams of one captured step)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --steps 6 --no-cpu-baseline --no-roofline $EXTRA > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err; }
run conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
run conn16 CUDA_DEVICE_MAX_CONNECTIONS=16
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_ch8 CUDA_DEVICE_MAX_CONNECTIONS=32 CFB_CHAINS=8
run conn32_ch3 CUDA_DEVICE_MAX_CONNECTIONS=32 CFB_CHAINS=3
EXTRA="--in-flight 2" run conn32_f2 CUDA_DEVICE_MAX_CONNECTIONS=32
EXTRA="--batch 128" run conn32_b128 CUDA_DEVICE_MAX_CONNECTIONS=32
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/c_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'))
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
