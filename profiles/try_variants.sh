#!/bin/bash
# Rebuild libgmc with different step-kernel launch shapes on the GPU box and print the phase breakdown of each.
cd "$(dirname "$0")/.."
export GMC_BLOCKS=${GMC_BLOCKS:-30,40,30,40}
for v in "256 2" "128 4" "256 4" "128 8" "64 8"; do
  set -- $v
  make -C mcmc_gpu_b200/csrc clean >/dev/null
  make -C mcmc_gpu_b200/csrc -j4 EXTRA="-DGMC_STEP_THREADS=$1 -DGMC_STEP_MIN_CTAS=$2" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== threads=$1 minCTAs=$2: $(grep -A1 'run_kernel' mcmc_gpu_b200/csrc/build/step.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -B1 -A2 'run_kernel' mcmc_gpu_b200/csrc/build/step.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
  python profiles/phase_timing.py 1184 100 2>&1 | tail -11
done
