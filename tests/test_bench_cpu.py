"""bench.py without a GPU: the reference arm (the CPU restatement on the host cores) prints the contract's JSON line, the GPU arm
refuses loudly -- there is no CPU fallback behind it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "channel256", "--gpus", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                     # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("MLUPS") and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3 and d["dtype"] == "f64" and d["config"]["workload"] == "channel256"
    assert d["value"] > 0 and abs(d["value"] - 256.0 ** 3 / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "256x256x256" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                                 # on a GPU box the arm runs; the refusal is what this test is about
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and not r.stdout.strip()


def test_clock_sampler_survives_a_box_without_nvml():
    sys.path.insert(0, ROOT)
    import bench
    c = bench.ClockSampler(0)
    c.start()
    out = c.stop()
    assert set(out) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
