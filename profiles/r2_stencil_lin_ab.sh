#!/bin/bash
# Round 2: occupancy / ring depth of the linear-form loss-only stencil.  usage: r2_stencil_lin_ab.sh "<EXTRA flags>" ...
cd "$(dirname "$0")/.."
run() {
  python profiles/stencil_only.py 256 500 30 2>&1 | sed -n '2p;6p'
  python profiles/stencil_only.py 128 2000 10 2>&1 | sed -n '2p;6p'
  python profiles/stencil_only.py 4096 500 5 2>&1 | sed -n '2p'
}
for v in "$@"; do
  make -C mcmc_gpu_b200/csrc clean >/dev/null
  make -C mcmc_gpu_b200/csrc -j8 EXTRA="$v" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== EXTRA='$v'"
  run
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j8 >/dev/null 2>&1
