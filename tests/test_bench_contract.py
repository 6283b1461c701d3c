"""bench.py contract, the parts that run without a GPU: the reference arm prints ONE JSON line with the agreed keys (the reference's
own kernel on the host cores, a bounded sample of the config-2 frame), and the B200 arm refuses to run without a CUDA device -- there
is no CPU fallback to fall back to."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "spheres-101k-1920x1080" and d["config"]["triangles"] == 101090
    # the same `config` keys as the B200 arm prints (the driver compares the two lines' configs); the row sampling is a key of its own
    sys.path.insert(0, str(ROOT))
    import bench
    from opencl_render_b200 import api, scenes
    cfg = scenes.CONFIGS[2]
    sc = cfg["make"]()
    m = sc.meta["camera"]
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
    assert d["config"] == bench.config_keys(cfg, sc, cam)
    assert (d["config"]["width"], d["config"]["height"], d["config"]["samples"]) == (1920, 1080, 1) and "every 4th row" in d["sampling"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "every 4th row" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_refuses_to_run_without_a_gpu():
    from opencl_render_b200 import _lib
    if _lib.load().oclr_device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and not any(l.startswith("{") for l in r.stdout.splitlines())
