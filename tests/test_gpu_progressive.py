"""GPU suite for the rows SURVEY.md section 8f-3 / 8f-4 add around the trace path: progressive accumulation over sample ranges
(raytrace_opencl.c:726-741), checkpoint / resume through the 16-bit planes, the fp32 accumulation option, the live progress
counter (raytrace.c:156-173, 566-587) and the slice cut of a launch domain.  Everything goes through the C-ABI; expected
values are the golden vectors generated from the reference build."""
import threading
import os
from pathlib import Path

import numpy as np
import pytest

from opencl_render_b200 import _lib, api
from tests import helpers

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _gold(name):
    g = np.load(GOLDEN / f"{name}.npz")
    return g["r"], g["g"], g["b"]


def _equal(img, want):
    return all(np.array_equal(img[c], want[c]) for c in range(3))


@pytest.mark.parametrize("variant", [api.KERNEL_SIMPLE, api.KERNEL_PIPE])
@pytest.mark.parametrize("name,cuts", [("soup_s4", [0, 1, 3, 4]), ("soup_mirror_glass", [0, 2, 3]), ("spheres_mirror", [0, 1, 2])])
def test_sample_ranges_accumulate_to_the_whole_job(name, cuts, variant):
    sc, cam, lists, samples = helpers.make_case(name)
    assert cuts[-1] == samples
    ds = api.DeviceScene(sc, 0)
    fr = api.DeviceFrame(ds, cam, lists)
    for a, b in zip(cuts[:-1], cuts[1:]):
        fr.render(samples, variant=variant, samples=(a, b))
    assert _equal(fr.read(), _gold(name))
    with pytest.raises(api.OclrError):
        fr.render(samples, samples=(2, samples + 1))


def test_checkpoint_and_resume_on_another_frame():
    sc, cam, lists, samples = helpers.make_case("soup_s4")
    ds = api.DeviceScene(sc, 0)
    a = api.DeviceFrame(ds, cam, lists)
    a.render(samples, samples=(0, 2))
    saved = a.read()
    a.close()
    b = api.DeviceFrame(ds, cam, lists)
    b.write(saved)
    assert _equal(b.read(), saved)
    b.render(samples, samples=(2, samples))
    assert _equal(b.read(), _gold("soup_s4"))


def test_float_accumulation_one_sample_equals_reference_planes():
    sc, cam, lists, samples = helpers.make_case("spheres")
    assert samples == 1
    ds = api.DeviceScene(sc, 0)
    fr = api.DeviceFrame(ds, cam, lists)
    fr.set_accumulation(api.ACCUMULATE_FLOAT)
    fr.render(1)
    assert _equal(fr.read(), _gold("spheres"))
    acc = fr.read_accum()
    assert np.all(acc[..., 3] == 1.0)


@pytest.mark.parametrize("name", ["soup_s4", "soup_mirror_glass"])
def test_float_accumulation_bounds_the_truncation_loss(name):
    """Reference: sum_i trunc(c_i * 65535/S); float mode: trunc(sum_i c_i * 65535/S).  Each truncation loses [0, 1) of a 16-bit
    step, so float - reference lies in [0, S) up to one step of fp32 rounding -- and the progressive float job resumes exactly."""
    sc, cam, lists, samples = helpers.make_case(name)
    ds = api.DeviceScene(sc, 0)
    fr = api.DeviceFrame(ds, cam, lists)
    fr.set_accumulation(api.ACCUMULATE_FLOAT)
    fr.render(samples)
    img = fr.read()
    want = _gold(name)
    for c in range(3):
        d = img[c].astype(np.int64) - want[c].astype(np.int64)
        sat = want[c] == 65535
        assert d[~sat].min() >= -1 and d[~sat].max() <= samples, (d.min(), d.max())
    acc = fr.read_accum()
    assert np.all(acc[..., 3] == float(samples))
    # resume the float job on a second frame from the accumulator after the first sample
    one = api.DeviceFrame(ds, cam, lists)
    one.set_accumulation(api.ACCUMULATE_FLOAT)
    one.render(samples, samples=(0, 1))
    part = one.read_accum()
    two = api.DeviceFrame(ds, cam, lists)
    two.set_accumulation(api.ACCUMULATE_FLOAT)
    two.write_accum(part)
    two.render(samples, samples=(1, samples))
    assert _equal(two.read(), img) and np.array_equal(two.read_accum(), acc)
    with pytest.raises(api.OclrError):
        two.render(samples, variant=api.KERNEL_SIMPLE)          # float mode is a wavefront-pipeline feature
    two.set_accumulation(api.ACCUMULATE_REFERENCE_16BIT)
    two.render(samples)
    assert _equal(two.read(), want)


@pytest.mark.parametrize("slices", [2, 3, 5])
def test_sliced_launch_domain_is_invisible(slices):
    try:
        for name in ("spheres", "soup_s4", "terrain_textured"):
            sc, cam, lists, samples = helpers.make_case(name)
            ds = api.DeviceScene(sc, 0)
            fr = api.DeviceFrame(ds, cam, lists)
            api.set_option("slices", 1)
            fr.render(samples)
            whole, ids, flags = fr.read(), fr.primary_ids(), fr.undefined_flags()
            api.set_option("slices", slices)
            _, launches, cnt = fr.render(samples, count=True)
            assert _equal(fr.read(), whole) and np.array_equal(fr.primary_ids(), ids) and np.array_equal(fr.undefined_flags(), flags)
            assert launches >= 3 * slices and cnt["segments"] > 0
            # the band set of one rank is cut into sub-band sets
            out = tuple(np.zeros((cam.height, cam.width), np.uint16) for _ in range(3))
            for rank in range(2):
                fr.render_bands(samples, 16, rank, 2)
                for rows in api.band_partition(cam.height, rank, 2, band_rows=16):
                    fr.read(rows=rows, out=out)
            assert _equal(out, whole)
    finally:
        api.set_option("slices", 0)
    with pytest.raises(api.OclrError):
        api.set_option("no_such_option", 1)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tracing_one_round_ahead_is_invisible(mode):
    """rt_wavefront.cuh: the spawn of a segment and the next segment's closest-hit query may run before the segment's last shadow
    ray has returned.  Every mode must reproduce the golden planes, ids and flags; modes 1 / 2 need fewer rounds on mirror chains."""
    launches = {}
    try:
        for name in ("soup_mirror_glass", "spheres_mirror", "terrain_textured", "soup_lights_sun_last", "soup_lights_omni_last", "soup_s4"):
            sc, cam, lists, samples = helpers.make_case(name)
            gold = np.load(GOLDEN / f"{name}.npz")
            ds = api.DeviceScene(sc, 0)
            fr = api.DeviceFrame(ds, cam, lists)
            api.set_option("ahead", mode)
            _, n, _ = fr.render(samples)
            fr.render(samples)                      # second render: enqueues exactly the rounds the first one needed
            launches[name] = fr.last_launches
            img, flags = fr.read(), fr.undefined_flags()
            assert helpers.compare_rgb(img, (gold["r"], gold["g"], gold["b"]), mask=(flags == 0))["diff_pixels"] == 0, name
            if samples == 1:
                assert np.array_equal(fr.primary_ids(), gold["ids"])
        api.set_option("ahead", 0)
        sc, cam, lists, samples = helpers.make_case("spheres_mirror")
        fr = api.DeviceFrame(api.DeviceScene(sc, 0), cam, lists)
        fr.render(samples)
        fr.render(samples)
        if mode:
            assert launches["spheres_mirror"] < fr.last_launches, (launches, fr.last_launches)
    finally:
        api.set_option("ahead", -1)


_SUPER_PROBE = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from opencl_render_b200 import api
from tests import helpers
out = {}
for name in ("soup", "soup_mirror_glass", "terrain_textured", "soup_axis_light", "soup_s4", "coarse_grid"):
    sc, cam, lists, samples = helpers.make_case(name)
    res = {}
    for policy in ("0", "1"):
        os.environ["OCLR_SUPER"] = policy
        ds = api.DeviceScene(sc, 0)
        fr = api.DeviceFrame(ds, cam, lists)
        _, _, cnt = fr.render(samples, count=True)
        res[policy] = (fr.read(), fr.primary_ids(), fr.undefined_flags(), cnt, len(ds.debug_read(2)))
        fr.close(); ds.close()
    two, three = res["0"], res["1"]
    gold = np.load(os.path.join("tests", "golden", name + ".npz"))
    out[name] = dict(equal=bool(all(np.array_equal(a, b) for a, b in zip(two[0], three[0])) and np.array_equal(two[1], three[1])
                                and np.array_equal(two[2], three[2])),
                     golden=int(helpers.compare_rgb(three[0], (gold["r"], gold["g"], gold["b"]), mask=(three[2] == 0))["diff_pixels"]),
                     axes=int(sc.axes_div), brick_bytes=[two[4], three[4]],
                     two={k: two[3][k] for k in ("superSteps", "superEnters", "coarseSteps", "bricksLoaded")},
                     three={k: three[3][k] for k in ("superSteps", "superEnters", "superRefines", "coarseSteps", "bricksLoaded")})
print("SUPER_PROBE " + json.dumps(out))
"""


def test_super_brick_level_of_the_walk_is_invisible():
    """rt_walk.h, three-level walk: entirely empty super-bricks (4x4x4 bricks) are crossed in one step.  The level is compiled out of
    the production kernel (it costs more than it saves at the kernel's register cap, rt_trace.cuh); build() also makes the variant
    library that has it, and this test drives that library in a process of its own: a scene packed with OCLR_SUPER=0 (no flags: the
    two-level walk) and the same scene packed with the flags give identical planes, ids and flags, equal to the golden vectors; with
    the flags the walk really takes super-brick steps and reads fewer brick records."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    lib = root / "opencl_render_b200" / "libopencl_render_b200_super.so"
    if not lib.is_file():
        pytest.skip("variant library with the super-brick level not built (python -m opencl_render_b200.build --variant super -DOCLR_SUPER_LEVEL=1)")
    env = dict(os.environ, OCLR_LIB=str(lib), OCLR_HIERARCHICAL="2")
    r = subprocess.run([sys.executable, "-c", _SUPER_PROBE], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("SUPER_PROBE ")][-1][len("SUPER_PROBE "):])
    for name, o in out.items():
        assert o["equal"] and o["golden"] == 0, (name, o)
        assert o["two"]["superSteps"] == 0 and o["two"]["superEnters"] == 0, (name, o)
        if o["axes"] >= 256:
            assert o["three"]["superSteps"] > 0 and o["three"]["coarseSteps"] < o["two"]["coarseSteps"], (name, o)
            assert o["three"]["bricksLoaded"] < o["two"]["bricksLoaded"] and o["brick_bytes"][1] > o["brick_bytes"][0], (name, o)
        elif o["axes"] < 32:
            assert o["three"]["superSteps"] == 0 and o["brick_bytes"][1] == o["brick_bytes"][0], (name, o)


_TAIL_PROBE = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from opencl_render_b200 import api
from tests import helpers
out = {}
for name in helpers.CASE_NAMES:
    sc, cam, lists, samples = helpers.make_case(name)
    gold = np.load(os.path.join("tests", "golden", name + ".npz"))
    ds = api.DeviceScene(sc, 0)
    fr = api.DeviceFrame(ds, cam, lists)
    _, launches, _ = fr.render(samples)
    img, flags = fr.read(), fr.undefined_flags()
    ids = fr.primary_ids()
    if samples != 1:
        fr.render(1)
        ids = fr.primary_ids()
    out[name] = dict(diff=int(helpers.compare_rgb(img, (gold["r"], gold["g"], gold["b"]), mask=(flags == 0))["diff_pixels"]),
                     ids=int((ids != gold["ids"]).sum()), launches=int(launches))
    fr.close(); ds.close()
print("TAIL_PROBE " + json.dumps(out))
"""


@pytest.mark.parametrize("after", ["0", "2", "off", "requeue-all", "requeue-16", "cells-0"])
def test_tail_handoff_is_invisible(after):
    """rt_tail.cuh: a launch domain may hand the rays still walking in the tail of a trace launch to wf_tail_kernel (one ray per warp,
    cooperative bursts; an experiment that is exact but not faster, hence off by default).  With the hand-off forced as early as
    possible (OCLR_HANDOFF_AFTER=0: every ray still walking at cell level when its warp sees the queue dry goes through the burst
    walker), two outer iterations later, and switched off, every golden case gives the golden planes and ids.  (Environment knobs
    are read once per process: each setting runs in a process of its own.)"""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ)
    if after == "off":
        env["OCLR_HANDOFF_MAX_PATHS"] = "0"
    elif after.startswith("requeue"):   # mode 2: the rays given up go back into a queue of walk records, a second pass of the pipe kernel takes them
        env.update(OCLR_HANDOFF_MAX_PATHS="4000000000", OCLR_HANDOFF_MODE="2", OCLR_HANDOFF_AFTER="0",
                   OCLR_HANDOFF_LANES="32" if after.endswith("all") else "16")
    else:   # mode 1: one ray per warp, bursts over brick planes ("0", "2") or over cell planes ("cells-0")
        env["OCLR_HANDOFF_MAX_PATHS"] = "4000000000"    # (off by default: not faster, rt_tail.cuh)
        env["OCLR_HANDOFF_AFTER"] = after.split("-")[-1]
        env["OCLR_TAIL_BRICKS"] = "0" if after.startswith("cells") else "1"
    r = subprocess.run([sys.executable, "-c", _TAIL_PROBE], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("TAIL_PROBE ")][-1][len("TAIL_PROBE "):])
    assert set(out) == set(helpers.CASE_NAMES)
    for name, o in out.items():
        assert o["diff"] == 0 and o["ids"] == 0, (after, name, o)


def test_ring_depth_follows_the_materials(monkeypatch):
    """runtime.cu Scene::ringSlots: only mirror / glass segments push a path's ring beyond two slots (raytrace_opencl.c:682-722), so a
    scene without reflection / transparency channels gets 2 ring slots per path instead of 12 -- same planes as with the full ring
    (OCLR_RING_SLOTS=12), about half the path state."""
    sc, cam, lists, samples = helpers.make_case("spheres")
    gold = np.load(GOLDEN / "spheres.npz")
    paths = ((cam.height + 7) // 8 * 8) * cam.width
    per_path = {}
    for slots in (None, "12"):
        if slots:
            monkeypatch.setenv("OCLR_RING_SLOTS", slots)
        ds = api.DeviceScene(sc, 0)
        fr = api.DeviceFrame(ds, cam, lists)
        assert fr.state_bytes == 0
        fr.render(samples)
        assert helpers.compare_rgb(fr.read(), (gold["r"], gold["g"], gold["b"]))["diff_pixels"] == 0
        per_path[slots] = fr.state_bytes / paths
        fr.close()
        ds.close()
    assert per_path[None] < 520 < 980 < per_path["12"] < 1000, per_path
    monkeypatch.delenv("OCLR_RING_SLOTS")
    sc, cam, lists, samples = helpers.make_case("spheres_mirror")       # mirrors: the full ring, whatever the knob says
    fr = api.DeviceFrame(api.DeviceScene(sc, 0), cam, lists)
    fr.render(samples)
    assert fr.state_bytes / (((cam.height + 7) // 8 * 8) * cam.width) > 980


def test_progress_counter_counts_pixel_samples():
    sc, cam, lists, _ = helpers.make_case("spheres")
    samples = 24
    ds = api.DeviceScene(sc, 0)
    fr = api.DeviceFrame(ds, cam, lists)
    total = cam.width * cam.height * samples
    seen = []
    stop = threading.Event()

    def poll():
        while not stop.is_set():
            seen.append(fr.progress())

    t = threading.Thread(target=poll)
    t.start()
    try:
        fr.render(samples)
    finally:
        stop.set()
        t.join()
    assert fr.progress() == (total, total)
    done = [d for d, tot in seen if tot == total]
    assert done == sorted(done) and all(0 <= d <= total for d in done)
    fr.render(samples, rows=(8, 24), samples=(3, 5))
    assert fr.progress() == (16 * cam.width * 2, 16 * cam.width * 2)


def test_raytrace_all_reports_live_progress():
    lib = _lib.load()
    sc, cam, lists, _ = helpers.make_case("spheres")
    seen = []
    stop = threading.Event()

    def poll():
        while not stop.is_set():
            seen.append((float(lib.GetProgress()), float(lib.oclr_estimated_seconds_left())))

    lib.SetProgress(0.0)
    t = threading.Thread(target=poll)
    t.start()
    try:
        api.raytrace_all(1, cam, lists, 320, sc)          # long enough (a few hundred ms) for the poller to catch it in flight
    finally:
        stop.set()
        t.join()
    p = [a for a, _ in seen]
    assert all(0.0 <= a <= 0.9991 for a in p)
    assert abs(lib.GetProgress() - 0.999) < 1e-4            # raytrace.c:580: the call itself never reports 1.0 (render.cpp:1397 does)
    assert any(0.0 < a < 0.999 for a in p), "no intermediate progress value was observed"
    assert all(e >= -1.0 for _, e in seen)
