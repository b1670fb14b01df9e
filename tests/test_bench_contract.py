"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU oracle on this box and
prints ONE JSON line with the keys the driver reads; the CUDA arm's static pieces (algorithmic bytes, metric names) agree
with DESIGN.md / SURVEY.md 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--games", "2048", "--start-turn", "30"],
                          cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"].startswith("env-steps/sec") and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    # the timed sample is long enough to anchor a ratio (a 0.1 s sample moved it by +-25 % between boxes)
    assert d["config"]["start_turn"] == 30 and d["config"]["turns_per_step"] >= 1
    assert d["ms_per_step"] * d["steps"] >= 600.0 or d["config"]["turns_per_step"] * (d["steps"] + d["warmup"]) >= 250


def test_algorithmic_bytes_match_the_design():
    sys.path.insert(0, ROOT)
    from bench import algorithmic_bytes_per_env_step

    # 20x20x2p 32,009 B; 10x10x2p 8,209; 15x15x2p 18,153; 20x20x4p 61,937
    assert algorithmic_bytes_per_env_step(20, 20, 2)["total"] == 32009
    assert algorithmic_bytes_per_env_step(10, 10, 2)["total"] == 8209
    assert algorithmic_bytes_per_env_step(15, 15, 2)["total"] == 18153
    assert algorithmic_bytes_per_env_step(20, 20, 4)["total"] == 61937
