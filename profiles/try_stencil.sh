#!/bin/bash
cd "$(dirname "$0")/.."
for v in "2 4 4 3" "2 4 3 3" "2 4 3 4" "2 8 2 3" "4 4 2 3"; do
  set -- $v
  make -C mcmc_gpu_b200/csrc clean >/dev/null
  make -C mcmc_gpu_b200/csrc -j4 EXTRA="-DRS_RW=$1 -DRS_WARPS=$2 -DRS_MIN_CTAS=$3 -DRS_STAGES=$4" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== RW=$1 WARPS=$2 MIN_CTAS=$3 STAGES=$4: $(grep -A2 'residual_kernelILb1ELb0ELb1E' mcmc_gpu_b200/csrc/build/residual.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
  python profiles/stencil_only.py 256 500 2>&1 | head -2
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j4 >/dev/null 2>&1
python -m pytest tests/test_gpu_residual.py -m gpu -q 2>&1 | tail -2
