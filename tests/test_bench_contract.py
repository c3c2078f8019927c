"""bench.py contract (driver-facing): the CPU arm (`--impl reference`) prints exactly ONE line on stdout, valid JSON with the keys the
driver reads; under a multi-rank launch only rank 0 prints.  (The GPU arm shares `emit`/`protect_stdout` and the same key set.)"""
import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "64", "--m", "256", "--d", "16", "--p", "2"]


def run(env_extra=None):
    env = dict(os.environ)
    env.update({"NK_BENCH_REF_M": "48", "NK_BENCH_REF_N": "300,900"})          # the one unmodified-reference fit, shrunk
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *ARGS], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in j, key
    assert j["impl"] == "reference" and j["unit"] == "samples/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["dtype"] == "f64" and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["value"] > 0 and j["ms_per_step"] > 0
    # the UNMODIFIED reference fit (baseline/_ref/regressors.py) is timed beside the port when it is on the box
    um = cb["unmodified_reference"]
    if (ROOT / "baseline" / "_ref" / "regressors.py").exists():
        assert um["n"] == [300, 900] and len(um["fit_s"]) == 2 and um["m"] == 48 and "T0_s" in um and "r_samples_per_s" in um
    else:
        assert um is None


def test_reference_arm_other_ranks_stay_silent():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout == ""


def test_stdout_is_protected_from_library_noise():
    code = ("import sys, os; sys.path.insert(0, %r); import bench; bench.protect_stdout(); os.write(1, b'NCCL version x\\n'); "
            "print('python-level noise'); bench.emit({'ok': 1})" % str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.stdout.splitlines() == ['{"ok": 1}']
    assert "NCCL version x" in r.stderr and "python-level noise" in r.stderr
