"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line with the agreed keys, and
non-zero ranks of a multi-rank reference launch exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "aligner_train_tokens_per_sec" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["din"] == 3584 and d["config"]["d"] == 4096


def test_reference_arm_nonzero_rank_is_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_watchdog_aborts_a_run_that_does_not_finish():
    """bench.py arms a timer before any work (a hung peer wait or collective has no timeout of its own): when it fires the process
    says so on stderr and exits with status 4 instead of sitting there until the launcher's limit. Disabled with 0."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TD_BENCH_WATCHDOG_S="0.5")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "8", "--warmup", "2"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 4 and "bench.py watchdog" in r.stderr and r.stdout.strip() == ""
