#!/bin/bash
# ncu captures of one round (run on the GPU box through gpurun): launch list of a 2-step pass, `--set full` of the
# step's kernels and of the tcgen05 per-pair attention.  The .ncu-rep files are summarised on the box
# (tools/ncu_extract.py) and removed: gpurun brings back at most 64 MiB.
out=${1:-gpurun_out/prof}; mkdir -p $out
CMD="python bench.py --steps 1 --warmup 0 --ddim-steps 2 --in-flight 1 --no-cpu-baseline --no-gpu-eager --no-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches.csv $CMD > $out/launches.out 2>&1
python tools/launch_summary.py $out/launches.csv > $out/launches.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_tma_kernel|gemm_tc_grouped|ln_rows_kernel|guidance_sched|mem_hat_kernel|softmax_shared|cross_mma|mha_mma" -s 600 -c 60 -o $out/full_step $CMD > $out/full_step.out 2>&1
python tools/ncu_extract.py $out/full_step.ncu-rep > $out/full_step.txt 2>&1
rm -f $out/full_step.ncu-rep
CFB_CROSS_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cross_tc_kernel|mem_transpose" -s 20 -c 12 -o $out/full_cross_tc $CMD > $out/full_cross_tc.out 2>&1
python tools/ncu_extract.py $out/full_cross_tc.ncu-rep > $out/full_cross_tc.txt 2>&1
rm -f $out/full_cross_tc.ncu-rep $out/launches.csv
ls -la $out
