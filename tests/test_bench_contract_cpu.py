"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "2", "--batch", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # exactly one JSON line on stdout (library chatter goes to stderr)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tiles/s" and d["higher_is_better"] is True
    assert d["metric"] == "inference tiles/sec" and d["value"] > 0 and d["config"]["baseline_config"] == 2
    cb = d["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "src", "model.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "tiles" in cb["sample"]
    assert d["config"]["batch_per_gpu"] == 2 and d["config"]["parallelism"] == "dp1"
    assert d["e2e"] == {"value": d["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_reference_arm_is_the_training_config():
    """The default line is BASELINE.json's first-named number (training, configs[2]); both arms use the same config block."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.metric_name(b.CONFIGS[3][2]) == "training tiles/sec (fwd+bwd)"
    c = b.config_block(3, 16, 1)
    assert c["baseline_config"] == 3 and c["batch_per_gpu"] == 16 and c["global_batch"] == 16
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["metric"] == "training tiles/sec (fwd+bwd)" and d["config"]["baseline_config"] == 3 and d["value"] > 0
