"""The reference arm of bench.py runs without a GPU (it times the oracle's CPU restatement): check the one-line JSON
contract the driver parses -- keys, units, the cpu_baseline / e2e objects of the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RECMV_BENCH_CPU_THREADS="4", RECMV_BENCH_REF_RAYS="128")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("rays/sec at 512x512x64") and line["value"] > 0 and line["gpu_launches"] == 0
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in line
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # other ranks of a torchrun launch stay silent and exit 0
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
