#!/bin/bash
# A/B of stencil build variants on one box: usage try_stencil2.sh "<EXTRA flags>" ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  make -C mcmc_gpu_b200/csrc clean >/dev/null
  make -C mcmc_gpu_b200/csrc -j4 EXTRA="$v" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== EXTRA='$v'"
  python profiles/stencil_only.py 256 500 30 > /tmp/s.txt 2>&1; sed -n 1,3p /tmp/s.txt
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j4 >/dev/null 2>&1
