"""The reference arm of bench.py (`--impl reference`: the CPU restatement of the path timed on the host cores, the one other
place besides tests / smoke() that may execute oracle/) prints ONE JSON line in the driver's contract -- runnable without a GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "qformer_video_audio_clips_per_sec" and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["value"] - d["config"]["clips_per_gpu_per_step"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE.json configs[1]" in d["config"]["workload"] and d["vs_baseline"] is None


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mraudio_b200/ may import it (bench.py's CPU legs, tests and smoke() only)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mraudio_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if "import oracle" in src or "from oracle" in src:
                    bad.append(os.path.join(dirpath, f))
    assert bad == []
