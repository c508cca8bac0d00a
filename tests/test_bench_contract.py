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
