#!/bin/bash
# Round 2: compute-sanitizer passes over small parity tests of every shared-memory kernel ("race
# detection / sanitizers").  Each tool runs under its own timeout; the summaries land in gpurun_out/r2san/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2san; mkdir -p $OUT
T1="tests/test_gpu_residual.py::test_batched_residual_and_loss tests/test_gpu_residual.py::test_signed_zero_and_special_quotients_are_bit_identical"
T2="tests/test_gpu_run.py::test_free_run_matches_oracle_on_emulated_draws[ragged_rf] tests/test_gpu_step.py::test_replay_matches_reference_trajectory[ragged_rf]"
T3="tests/test_gpu_sgs.py::test_replay_matches_oracle_trajectory[expo_raw-warp] tests/test_gpu_sgs.py::test_replay_matches_oracle_trajectory[expo_raw-cta]"
for tool in memcheck racecheck synccheck; do
  i=0
  for T in "$T1" "$T2" "$T3"; do
    i=$((i+1))
    timeout ${SAN_TIMEOUT:-200} compute-sanitizer --tool $tool --print-limit 5 --log-file $OUT/$tool.$i.log \
      python -m pytest $T -m gpu -x -q -p no:cacheprovider > $OUT/$tool.$i.pytest.txt 2>&1
    echo "$tool set $i: exit $? | $(tail -1 $OUT/$tool.$i.pytest.txt) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.$i.log | tail -1)"
  done
done 2>&1 | tee $OUT/summary.txt
