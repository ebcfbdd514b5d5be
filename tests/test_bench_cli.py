"""bench.py's reference arm (the CPU leg the driver launches next to ours) keeps the JSON contract: one line, the same metric /
unit / higher_is_better as our arm, `impl`, a `cpu_baseline` describing the run and an `e2e` that repeats the line's value with
zero copy bytes.  Only the gim_dkm arm runs here (one bounded oracle call, ~10 s); the gim_loftr arm shares the code path and is
timed by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_dkm_prints_one_contract_line():
    r = subprocess.run([sys.executable, "bench.py", "--workload", "dkm", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["steps"] == 1 and d["requested_steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "gim_dkm 672x896" in d["config"]["workload"]


def test_reference_arm_exits_quietly_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    for extra in ([], ["--workload", "dkm"]):
        r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra, cwd=ROOT,
                           env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
