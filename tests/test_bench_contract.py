"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys, from the
unmodified reference under baseline/_ref when it is installed (kind "reference") and from the oracle port otherwise."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest

HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gstatsMCMC"))


@pytest.mark.parametrize("extra,kind", [(["--port-baseline"], "port"), ([], "reference" if HAVE_REF else "port")])
def test_reference_arm_prints_one_json_line_with_the_contract_keys(extra, kind):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-iters", "4", "--grid", "200"] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "chain-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    # both arms report the same configuration keys (bench.base_config)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(chains=256, chains_total=0, grid=200, iters=1000)
    assert set(d["config"]) == set(bench.base_config(a, 1))
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
