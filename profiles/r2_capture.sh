#!/bin/bash
# Round 2: the ncu evidence kept under profiles/r2/ (run on a GPU box AFTER the same commands exited 0 without ncu).
#   launches.csv         : launch list of a short bench run (gpu__time_duration.sum, --clock-control none)
#   *.ncu-rep            : --set full captures of the dominant kernels (summarised by profiles/ncu_summary.py afterwards)
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2cap; mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
BENCH="python bench.py --steps 1 --warmup 3 --iters 40 --no-cpu-baseline --no-reference-gpu --sgs-iters 2 --sgs-chains 128 --target-iters 20"
GMC_BENCH_PREWARM_S=0 $BENCH > $OUT/bench_short.json 2> $OUT/bench_short.err || { echo "short bench failed"; tail -5 $OUT/bench_short.err; }
GMC_BENCH_PREWARM_S=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $BENCH > $OUT/ncu_list.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:run_kernel -s 2 -c 1 -o $OUT/run_kernel python profiles/r2_step_ab.py 256 500 40 1 > $OUT/ncu_run.log 2>&1
GMC_STEP_WIDE=1 timeout 300 $NCU -k regex:run_kernel -s 2 -c 1 -o $OUT/run_kernel_wide python profiles/r2_step_ab.py 128 500 40 1 > $OUT/ncu_run_wide.log 2>&1
timeout 300 $NCU -k regex:residual_tma -c 8 -o $OUT/residual_tma python profiles/stencil_only.py 256 500 1 > $OUT/ncu_stencil.log 2>&1
timeout 300 $NCU -k regex:sgs_run_kernel -s 1 -c 1 -o $OUT/sgs_run_kernel python profiles/sgs_bench.py 512 4 0 > $OUT/ncu_sgs.log 2>&1
ls -la $OUT
