"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads.
The native arm needs a GPU and is exercised on the GPU box by the driver itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--bodies", "8192"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pair_interactions_per_s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "G pair-interactions/s" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "targets x" in cb["sample"]
    assert "workload" in d["config"] and d["config"]["n"] == 8192
    # the reference's real per-step path (Simulation::step on its own 25,000-body scene) rides in the same line
    bh = d["bh"]
    assert bh["n"] == 25000 and bh["ms_per_step"] > 0 and bh["cpu_baseline"]["kind"] in ("reference", "port")
    assert bh["e2e"]["value"] == bh["ms_per_step"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--bodies", "8192"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
