cd "$(dirname "$0")/.."
for v in "-DR2_LS_OCTETS" "" "-DR2_LS_OCTETS" ""; do
  make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j8 EXTRA="$v" >/dev/null 2>&1 || { echo "build failed"; continue; }
  echo "=== EXTRA='$v'"
  timeout 100 python profiles/stencil_only.py 4096 500 5 2>&1 | sed -n "2,3p"
  timeout 100 python profiles/stencil_only.py 256 500 40 2>&1 | sed -n "2,3p"
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j8 >/dev/null 2>&1
