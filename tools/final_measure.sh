#!/bin/bash
out=gpurun_out/final; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -rA > $out/pytest_full.log 2>&1; echo "pytest rc=$?" >> $out/pytest_full.log
grep -E "passed|failed|pytest rc" $out/pytest_full.log | tail -3
python bench.py > $out/bench_default.json 2> $out/bench_default.err
Q="--no-cpu-baseline --no-gpu-eager --no-roofline"
python bench.py --dyadic $Q > $out/bench_dyadic.json 2> $out/bench_dyadic.err
python bench.py --windows 115 --batch 1 --steps 1 --warmup 1 $Q > $out/bench_win115_b1.json 2> $out/bench_win115_b1.err
python bench.py --windows 115 --batch 64 --steps 3 --warmup 1 $Q > $out/bench_win115_b64.json 2> $out/bench_win115_b64.err
python bench.py --sweep 4096 $Q > $out/bench_sweep4096.json 2> $out/bench_sweep4096.err
python bench.py --precision fp32 --steps 6 --warmup 3 $Q > $out/bench_fp32.json 2> $out/bench_fp32.err
python bench.py --in-flight 1 --steps 10 $Q > $out/bench_lane1.json 2> $out/bench_lane1.err
for f in default dyadic win115_b1 win115_b64 sweep4096 fp32 lane1; do python -c "
import json
d=json.load(open('$out/bench_$f.json')); print('$f', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['execution'].get('batches_in_flight'), d.get('windows_per_second'), d['clocks'])" 2>&1 | tail -1; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
