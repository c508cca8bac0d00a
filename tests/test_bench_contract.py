"""bench.py's reference arm runs on CPU only, so its JSON contract can be checked here without a GPU: one line, the
driver's keys, `impl: reference`, a cpu_baseline describing the run and a zero-copy e2e object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "waveglow_infer_audio_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    # "reference" when oracle/build_ref.py has placed the reference module in oracle/_ref/, else the torch-op port
    has_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "glow.py"))
    assert cb["kind"] == ("reference" if has_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb and cb["cpu_model"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_default_precision_resolution():
    """`--precision` defaults to the fp32-path mode of the workload: f16f8 for the 256-channel WaveGlow, bf16x3 otherwise."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'args.precision = "f16f8" if (args.workload == "waveglow" and args.channels == 256 and args.config in (0, 1, 2)) else "bf16x3"' in src


import pytest


@pytest.mark.gpu
def test_gpu_arm_json_line():
    """The driver's line at N=1 (short run): every key of the contract, a live roofline and a live accuracy check."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-extra"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "waveglow_infer_audio_samples_per_sec" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["value"] > 1e7 and d["output_finite"] is True
    e = d["e2e"]
    assert e["value"] > 1e7 and e["h2d_bytes_per_step"] == 16 * (80 * 861 + 861 * 256) * 4 and e["d2h_bytes_per_step"] == 16 * 861 * 256 * 4
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["kernel"] == "k_layer_ps<2,0>" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.2 < r["frac"] < 0.6
    assert r["launches_timed"] == 2 * 96 and len(r["launch_ms_by_layer"]) == 8
    assert r["layer_share_of_step"] < 1.0                       # 96 x avg launch <= step
    assert d["gpu_launches"] == 2 * 124
    a = d["accuracy"]
    assert a["max_abs"] <= 1e-3 and a["snr_db"] >= 60.0
