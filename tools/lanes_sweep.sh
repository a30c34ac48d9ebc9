#!/bin/bash
# Reproduces profiles/r01_lanes_sweep.txt on one B200: batch size, concurrent chains per captured step (CFB_CHAINS) and
# independent batches in flight (bench.py --in-flight, SamplerPool lanes).  Results: gpurun_out/lanes_*.json.
#   gpurun --timeout 900 -- 'bash tools/lanes_sweep.sh'
mkdir -p gpurun_out
run() {  # name chains [bench args...]
  local name=$1 ch=$2; shift 2
  CFB_CHAINS=$ch timeout 300 python bench.py --steps 12 --no-cpu-baseline --no-roofline "$@" \
    > gpurun_out/lanes_$name.json 2> gpurun_out/lanes_$name.err
}
for b in 8 16 32 64 128 256; do run b${b}_c6_f1 6 --batch $b --in-flight 1; done
for c in 1 2 3 4 6; do for f in 2 3; do run b64_c${c}_f${f} $c --in-flight $f; done; done
run b8_c1_f2 1 --batch 8 --in-flight 2
# GEMM execution options that were only ever measured with one batch in flight (latency-bound), now with two lanes
CFB_TC_OCC3=1 run b64_c3_f2_occ3 3 --in-flight 2
CFB_TC_CLUSTER=21 run b64_c3_f2_mc21 3 --in-flight 2
CFB_TC_CLUSTER=42 run b64_c3_f2_mc42 3 --in-flight 2
CFB_TC_2CTA=1 CFB_TC_2CTA_MIN_ROWS=4096 run b64_c1_f1_pair4096 1 --in-flight 1
python - <<'PY'
import glob, json
print(f"{'run':<16} {'motion-s/s':>10} {'e2e':>8} {'ms/pass':>8} {'ms/den.step (1 lane)':>21} {'launches':>9}")
for p in sorted(glob.glob("gpurun_out/lanes_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(f"{p.split('lanes_')[1][:-5]:<16} {d['value']:>10.0f} {d['e2e']['value']:>8.0f} {d['ms_per_step']:>8.2f} "
              f"{d['ms_per_denoiser_step']:>21.3f} {d['gpu_launches']:>9}")
    except Exception as exc:
        print(p, "ERR", exc)
PY
