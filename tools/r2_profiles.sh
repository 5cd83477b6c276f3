#!/bin/bash
# round 2: the ncu evidence for the bench command and the sweep record (one GPU)
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras"
timeout 200 $CMD > gpurun_out/r2_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
wc -l gpurun_out/r2_launches.csv
timeout 200 $CMD > gpurun_out/r2_plain2.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"unpool_vec_kernel|pool_patches_tma_kernel|block_forward_kernel" -s 60 -c 6 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/r2_ncu_f.log 2>&1
tail -2 gpurun_out/r2_ncu_f.log
timeout 500 python tools/sweep.py --dtype bf16 --out gpurun_out/r2_sweep_bf16.md > gpurun_out/r2_sweep_bf16.log 2>&1; tail -2 gpurun_out/r2_sweep_bf16.log | cut -c1-200
timeout 500 python tools/sweep.py --dtype f32 --out gpurun_out/r2_sweep_f32.md > gpurun_out/r2_sweep_f32.log 2>&1; tail -2 gpurun_out/r2_sweep_f32.log | cut -c1-200
SW="python tools/sweep.py --dtype bf16 --graph random --no-ref --iters 3 --points 262144:8:64,262144:32:64,262144:8:128,262144:8:512"
timeout 200 $SW > gpurun_out/r2_sweep_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_sweep.csv $SW > gpurun_out/r2_ncu_sw.log 2>&1
wc -l gpurun_out/r2_launches_sweep.csv
