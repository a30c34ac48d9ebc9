#!/bin/bash
# on the GPU box: rebuild with -DCFB_DEBUG_SKIP into the tree, run the default bench with each class dropped
out=gpurun_out/r2k; mkdir -p $out
CFB_EXTRA_NVCC="-DCFB_DEBUG_SKIP" python -c "from convofusion_b200.build import build; build(force=True)" > $out/build.log 2>&1
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-roofline"
for m in 0 1 2 4 8 16 32 64 128 256 13 15; do
  CFB_SKIP=$m $B > $out/skip$m.json 2> $out/skip$m.err
  python -c "import json; d=json.load(open('$out/skip$m.json')); print('skip $m: %.0f motion-s/s, %.2f ms/pass (2 lanes), one lane %.2f ms/pass' % (d['value'], d['ms_per_step'], d['one_batch_in_flight']['ms_per_step']))" 2>/dev/null || echo "skip $m failed"
done | tee $out/summary.txt
