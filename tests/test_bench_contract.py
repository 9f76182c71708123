"""bench.py prints ONE JSON line with the keys the driver reads (metric / value / e2e / roofline / cpu_baseline ...), for both
arms.  The reference arm runs on the CPU (oracle port = the checker timed as a baseline: the one place bench.py executes
oracle/); the GPU arm is run with tiny step counts."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--prime", "4", "--no-learner"], 600)
    assert BASE <= set(d) and d["impl"] == "reference" and d["metric"] == "env agent-steps/s" and d["unit"] == "agent-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2 and d["warmup"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["loaded_product_library"] is False          # the reference arm never maps libmarl_b200.so


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "6", "--warmup", "3", "--prime", "30", "--loop-steps", "3", "--fp32-loop-steps", "1", "--host-fed-steps", "2",
              "--c2-steps", "30", "--replay-batch", "4096", "--c1-iters", "120", "--e2e-hybrid", "1"], 900)
    assert BASE | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["metric"] == "env agent-steps/s" and d["n_gpus"] == 1 and d["steps"] == 6 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["gpu_launches"] == 6 and d["value"] > 1e9 and d["dtype"] == "f64" and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["dense"]["value"] > 0 and e["wire"]["value"] > 0 and 0 < e["hybrid"]["wire_fraction"] < 1
    assert e["value"] == max(e["dense"]["value"], e["wire"]["value"], e["hybrid"]["value"]) and e["mode"] in ("dense", "wire", "hybrid")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    ln = d["learner"]
    assert ln["value"] > 0 and ln["roofline"]["bound"] == "tensor" and ln["e2e"]["value"] > 0 and ln["cpu_baseline"]["value"] > 0
    assert set(ln["segments_ms"]) == {"act", "env_step", "replay_push", "sample+learn"}
    assert d["secondary_c2"]["value"] > 0 and d["replay"]["sample"]["frac_of_hbm_peak"] > 0 and d["learner_fp32"]["value"] > 0
    assert "hw_slowdown" not in d["clocks"]["reasons"]
    c1 = d["c1_dropin"]                                   # BASELINE.json configs[0] through the drop-in classes
    assert set(c1["ms_per_call"]) == {"act", "step", "remember", "learn"} and c1["iterations"] >= 20
    assert abs(c1["ms_per_iteration"] - sum(c1["ms_per_call"].values())) < 1e-9 and 0 < c1["ms_per_iteration"] < 50
