#!/bin/bash
# Round 2: A/B of the step kernel's register cap (128 = two CTAs fill the register file; 120 leaves room for one small CTA per SM)
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
for v in "" "-DGMC_STEP_MAXNREG=120"; do
  make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j8 EXTRA="$v" >/dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== EXTRA='$v'"
  for cfg in "256 500" "512 500"; do python profiles/r2_step_ab.py $cfg 1000 3 2>&1 | tail -1; done
  GMC_E2E_COUNTS=32 timeout 200 python bench.py --steps 5 --warmup 3 --no-sgs --no-targets --no-cpu-baseline --no-reference-gpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('int32 counts: value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'serial', d['e2e']['serial']['value'])"
  GMC_E2E_COUNTS=16 timeout 200 python bench.py --steps 5 --warmup 3 --no-sgs --no-targets --no-cpu-baseline --no-reference-gpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('int16 counts: value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'serial', d['e2e']['serial']['value'])"
done
make -C mcmc_gpu_b200/csrc clean >/dev/null; make -C mcmc_gpu_b200/csrc -j8 >/dev/null 2>&1
