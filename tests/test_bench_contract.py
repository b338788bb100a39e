"""CPU: the JSON contract of bench.py's reference arm (the GPU arm needs a device and is exercised by the
driver), and that the GPU arm refuses to run without one instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(built):
    # --scale: 960x540 frames keep the CPU suite short (the line is then marked INVALID: a debug size, not a bench number)
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports to its ranks: the arm must not inherit it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--scale", "0.125"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and "full" in d["cpu_baseline"]["sample"] and "INVALID" in d
    # the OpenMP team is set explicitly to the host cores this process may use, whatever OMP_NUM_THREADS says
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "7680x4320" in d["config"]["workload"] and d["config"]["tracks"] == 10
    assert "over 1 GPU(s)" in d["config"]["sharding"]
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_device():
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_band_mode_and_same_sharding_string_as_the_gpu_arm(built):
    # more than 32 steps: the fixed 1/16 band per step + one full frame validating the x16; --gpus N names N in `sharding`
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "4", "--steps", "33", "--warmup", "0",
                        "--scale", "0.0625"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert "1/16 band" in d["cpu_baseline"]["sample"] and "over 4 GPU(s)" in d["config"]["sharding"]
    chk = d["cpu_baseline"]["full_frame_check"]
    assert chk["full_frame_s"] > 0 and 0.3 < chk["measured_over_extrapolated"] < 3.0


def test_reference_arm_other_workloads(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["config"]["width"] == 1920 and d["value"] > 0 and d["dtype"] == "f32"
