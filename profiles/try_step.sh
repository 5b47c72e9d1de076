#!/bin/bash
# A/B of step-kernel build variants on one box: usage try_step.sh "<EXTRA flags>" ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  make -C mcmc_gpu_b200/csrc clean >/dev/null
  make -C mcmc_gpu_b200/csrc -j4 EXTRA="$v" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== EXTRA='$v'"
  python profiles/phase_timing.py 256 200 > /tmp/p.txt 2>&1; grep -E "chain-steps/s|scalars|spectrum fill" /tmp/p.txt
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j4 >/dev/null 2>&1
