#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_pytest.log 2>&1
tail -6 gpurun_out/i_pytest.log
timeout 600 python bench.py > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err; tail -c 300 gpurun_out/i_bench_default.err
timeout 300 python bench.py --in-flight 1 --no-cpu-baseline > gpurun_out/i_bench_f1.json 2> gpurun_out/i_bench_f1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench_ref.err
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/i_bench_*.json')):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d.get('ms_per_denoiser_step'), d.get('gpu_launches'), (d.get('roofline') or {}).get('frac'), d.get('one_batch_in_flight'))
    except Exception as e:
        print(p,'ERR',e, open(p.replace('.json','.err')).read()[-800:])
PY
