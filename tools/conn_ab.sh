#!/bin/bash
# hardware work-queue count (CUDA_DEVICE_MAX_CONNECTIONS, default 8) against the step's 6-12 concurrent streams
out=${1:-gpurun_out/conn}; mkdir -p $out
B="python bench.py --steps 18 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-roofline"
run() {  # name, lanes, chains, env...
  name=$1; l=$2; c=$3; shift 3
  env "$@" $B --in-flight $l --chains $c > $out/$name.json 2> $out/$name.err
  python -c "import json; d=json.load(open('$out/$name.json')); print('$name lanes $l chains $c: %.0f (e2e %.0f) %.2f ms/pass' % (d['value'], d['e2e']['value'], d['ms_per_step']))" || tail -3 $out/$name.err
}
{
run base 3 2 X=1
run conn32 3 2 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_l4c2 4 2 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_l6c1 6 1 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_l3c3 3 3 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn4 3 2 CUDA_DEVICE_MAX_CONNECTIONS=4
run scaleq 3 2 CUDA_SCALE_LAUNCH_QUEUES=4x
run conn32_scaleq 3 2 CUDA_DEVICE_MAX_CONNECTIONS=32 CUDA_SCALE_LAUNCH_QUEUES=4x
} | tee $out/summary.txt
