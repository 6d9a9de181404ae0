"""CPU checks of bench.py's contract: the reference arm prints ONE JSON line with the agreed keys (and runs the
unmodified reference when baseline/_ref is there), and the FLOP / byte accounting matches SURVEY.md §8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "S", "--steps", "1",
                        "--warmup", "1", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    import reference_loader as RL
    assert cb["kind"] == ("reference" if RL.available() else "port")
    assert "workload" in line["config"] and "model" not in line["config"]


def test_flop_and_byte_accounting_matches_the_survey():
    import bench
    cfg, B = bench.CONFIGS["XL"]
    assert B == 1024
    f = bench.algorithmic_flops_per_image(cfg, 256, 5)
    assert abs(f / 1e9 - 163.569) < 0.01                              # XL 163.569 GFLOP / image
    assert abs(bench.algorithmic_flops_per_image(bench.CONFIGS["M"][0], 196, 5) / 1e9 - 89.133) < 0.01
    assert abs(bench.algorithmic_flops_per_image(bench.CONFIGS["S"][0], 196, 5) / 1e9 - 40.105) < 0.01
    # 177 big GEMMs per XL step; the four per block that write the residual stream move both of its planes
    M = 1024 * 261
    per_launch = bench.gemm_bytes_per_step(cfg, M) / 177
    assert 2.2e9 < per_launch < 2.5e9
