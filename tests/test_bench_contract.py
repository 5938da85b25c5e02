"""bench.py contract on the CPU arm (no GPU needed): `--impl reference` prints ONE JSON line with the keys the
driver reads, for the metric / unit / config of the B200 arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "64",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "wgan_gp_train_iters_per_sec" and d["higher_is_better"] is True
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-3 * 1000.0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--batch", "64", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]


def test_every_baseline_config_has_a_reference_arm_with_the_same_workload_text():
    """--config 3 / 4 / 5: the reference arm prints the contract line with the metric, unit and config.workload the
    B200 arm of that configuration prints (both arms take them from bench.CONFIGS)."""
    import pytest
    sys.path.insert(0, ROOT)
    import bench
    for cid in (3, 4, 5):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(cid),
                              "--batch", "64", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                             timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads([l for l in out.stdout.splitlines() if l.strip().startswith("{")][0])
        c = bench.CONFIGS[cid]
        assert d["metric"] == c["metric"] and d["unit"] == c["unit"] and d["scaling"] == c["scaling"]
        assert d["config"]["workload"] == c["workload"] and d["config"]["baseline_config"] == cid
        assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": c["workload"]') >= 3      # reference arm, training arm, sampling arm
