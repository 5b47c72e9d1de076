#!/bin/bash
# Round 2: A/B of the stencil builds on one box.  usage: r2_stencil_ab.sh "<EXTRA flags>" ...   (env LEGACY=1 adds the round-1 kernel)
cd "$(dirname "$0")/.."
run() {
  python profiles/stencil_only.py 256 500 30 2>&1 | sed -n 1,3p
  python profiles/stencil_only.py 128 2000 10 2>&1 | sed -n 1,2p
}
if [ -n "$LEGACY" ]; then echo "=== round-1 kernel (GMC_RS_LEGACY=1)"; GMC_RS_LEGACY=1 run; fi
for v in "$@"; do
  make -C mcmc_gpu_b200/csrc clean >/dev/null
  make -C mcmc_gpu_b200/csrc -j8 EXTRA="$v" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== EXTRA='$v'"
  run
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j8 >/dev/null 2>&1
