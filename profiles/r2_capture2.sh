#!/bin/bash
# Round 2: re-capture of the kernels that changed after r2_capture.sh (lean loss-only stencil, SGS chain without batch barriers,
# 32-warp whole-grid values pass).  Run on a GPU box after the same commands exited 0 without ncu.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2cap2; mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
python profiles/stencil_only.py 256 500 1 > $OUT/stencil_plain.txt 2>&1 || exit 1
timeout 300 $NCU -k regex:residual_tma -c 8 -o $OUT/residual_tma python profiles/stencil_only.py 256 500 1 > $OUT/ncu_stencil.log 2>&1
python profiles/sgs_bench.py 512 20 0 > $OUT/sgs_plain.txt 2>&1 || exit 1
timeout 300 $NCU -k regex:sgs_run_kernel -s 1 -c 1 -o $OUT/sgs_run_kernel python profiles/sgs_bench.py 512 4 0 > $OUT/ncu_sgs.log 2>&1
python profiles/sgs_grid_bench.py 300 4 > $OUT/sgs_grid_plain.txt 2>&1
timeout 300 $NCU -k regex:sgs_grid -c 3 -o $OUT/sgs_grid python profiles/sgs_grid_bench.py 300 4 > $OUT/ncu_sgs_grid.log 2>&1
python profiles/stencil_only.py 128 2000 10 > $OUT/stencil_2000.txt 2>&1
python profiles/stencil_only.py 256 500 30 > $OUT/stencil_500.txt 2>&1
python profiles/stencil_only.py 4096 500 5 > $OUT/stencil_4096.txt 2>&1
ls -la $OUT
