#!/bin/bash
mkdir -p gpurun_out
CFB_TC_2CTA=1 timeout 120 python tools/pair_check.py check > gpurun_out/h_check.log 2>&1; echo "check rc=$?" >> gpurun_out/h_check.log
cat gpurun_out/h_check.log | tail -14
if grep -q "check rc=0" gpurun_out/h_check.log; then
  CFB_TC_2CTA=0 timeout 200 python tools/pair_check.py time > gpurun_out/h_time0.log 2>&1
  CFB_TC_2CTA=1 timeout 200 python tools/pair_check.py time > gpurun_out/h_time1.log 2>&1
  paste gpurun_out/h_time0.log gpurun_out/h_time1.log | cut -c1-130
  for f in 1 2; do
    CFB_TC_2CTA=1 timeout 300 python bench.py --steps 12 --no-cpu-baseline --no-roofline --in-flight $f > gpurun_out/h_pair_f$f.json 2> gpurun_out/h_pair_f$f.err
    python -c "
import json;d=json.loads(open('gpurun_out/h_pair_f$f.json').read().strip().splitlines()[-1]);print('pair in-flight $f', round(d['value']), round(d['e2e']['value']), d['ms_per_denoiser_step'])" || tail -5 gpurun_out/h_pair_f$f.err
  done
fi
