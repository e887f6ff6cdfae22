"""CPU-only checks of the measurement plumbing: synthetic inputs, the reference arm's JSON line
(bench.py --impl reference runs the oracle port on the host cores and needs no GPU), and the
profile tools on the committed evidence."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_synthetic_inputs():
    """benchmark.py:58-60 semantics ((rand > s) as integers) and the Poisson histogram variant."""
    x = bench.make_inputs(2, (64, 80), 0.9, 2, seed=3)
    assert len(x) == 2 and x[0].dtype == torch.uint8 and tuple(x[0].shape) == (2, 20, 64, 80)
    assert set(x[0].unique().tolist()) <= {0, 1}
    assert abs(x[0].float().mean().item() - 0.1) < 0.01
    assert not torch.equal(x[0], x[1])
    assert torch.equal(bench.make_inputs(2, (64, 80), 0.9, 1, seed=3)[0], x[0])       # seeded
    assert bench.make_inputs(1, (32, 32), 0.0, 1)[0].min().item() == 1                 # benchmark default: every bin active
    p = bench.make_inputs(2, (64, 80), 0.95, 1, kind="poisson")[0]
    assert p.dtype == torch.uint8 and p.max().item() <= 10
    assert abs((p > 0).float().mean().item() - 0.05) < 0.01


def test_reference_arm_line():
    """The reference arm prints one JSON line with the contract's keys; ranks other than 0 print nothing."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "gen1_b1", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "gen1_b1"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out1.returncode == 0 and not [l for l in out1.stdout.splitlines() if l.startswith("{")]


def test_committed_bench_lines_carry_the_contract():
    """Every committed bench line of this round has the keys the driver reads."""
    prof = os.path.join(ROOT, "profiles")
    names = [n for n in os.listdir(prof) if n.startswith("r01_bench_v") and n.endswith(".json")]
    assert names
    latest = max(names, key=lambda n: int(n[len("r01_bench_v"):-5]))
    d = json.load(open(os.path.join(prof, latest)))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["config"]["workload"] == "1mpx_b8" and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 20 * 384 * 640 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1


def test_launch_summary_tool_reads_committed_list():
    prof = os.path.join(ROOT, "profiles")
    names = [n for n in os.listdir(prof) if n.startswith("r01_launches_v") and n.endswith(".csv") and "eager" not in n]
    latest = max(names, key=lambda n: int(n[len("r01_launches_v"):-4]))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(prof, latest), "1"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1000:]
    assert "one forward" in out.stdout and "sast::" in out.stdout
