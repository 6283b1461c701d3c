#!/usr/bin/env python3
"""Generates tests/golden/*.npz and builders.json from the REFERENCE ITSELF (oracle/_ref = /root/reference compiled
unmodified by oracle/build_ref.py).  Run in the authoring container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Per case: the reference's RGB planes (RaytraceAll computationType 0 semantics), the primary-hit triangle ids obtained
from the unmodified kernel through the ID-material scene variant, and a digest of the inputs so
generator drift is detected.  builders.json holds digests of the reference builders' outputs."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
from opencl_render_b200 import api, scenes  # noqa: E402
import ref  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_builders import CASES as BUILDER_CASES  # noqa: E402


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scene_digest(sc, cam, lists):
    return digest(sc.vertex, sc.tri_idx, sc.tri_mat, sc.tri_uv, sc.tri_normal, sc.mat_size, sc.mat_start, sc.textures, sc.light_type,
                  sc.light_pos, sc.light_dir, sc.light_colour, sc.light_radius, sc.light_half, sc.box_min, sc.grid_start, sc.grid_list,
                  lists.start, lists.end, lists.list, cam.eye, cam.eye_to_top_left, cam.left_to_right, cam.top_to_bottom,
                  np.float32(cam.pixel_size_inv))


def main():
    ref.load()
    for name in helpers.CASE_NAMES:
        sc, cam, lists, samples = helpers.make_case(name)
        # the lists themselves must be the reference builders' (256 only; coarse grids are this repo's generalisation)
        if sc.axes_div == 256:
            rs, re_, rl = ref.camera_lists(cam, sc)
            assert np.array_equal(rs, lists.start) and np.array_equal(re_, lists.end) and np.array_equal(rl, lists.list)
            box, gs, gl = ref.scene_grid(sc)
            assert np.array_equal(box, sc.box_min) and np.array_equal(gs, sc.grid_start) and np.array_equal(gl, sc.grid_list)
        r, g, b = ref.raytrace_all(cam, lists, sc, samples)            # the reference's own single-thread entry point
        rt = ref.render(cam, lists, sc, samples, threads=4)            # threaded driver must agree bit for bit
        assert all(np.array_equal(x, y) for x, y in zip((r, g, b), rt))
        idsc = scenes.id_material_variant(sc)
        ir, ig, ib = ref.raytrace_all(cam, lists, idsc, 1)
        ids = scenes.decode_id_planes(ir, ig, ib)
        np.savez_compressed(HERE / f"{name}.npz", r=r, g=g, b=b, ids=ids, samples=np.int32(samples),
                            inputs=np.frombuffer(bytes.fromhex(scene_digest(sc, cam, lists)), np.uint8))
        print(name, r.shape, "hits", int((ids != 0xFFFFFFFF).sum()), "nonzero", int((r > 0).sum()))
    out = {}
    for name, (make, w, h) in BUILDER_CASES.items():
        sc = make()
        m = sc.meta["camera"]
        cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], w, h)
        sc.normalise()
        rs, re_, rl = ref.camera_lists(cam, sc)
        box, gs, gl = ref.scene_grid(sc)
        out[name] = dict(camera=digest(cam.eye, cam.eye_to_top_left, cam.left_to_right, cam.top_to_bottom, np.float32(cam.pixel_size_inv)),
                         camera_lists=digest(rs, re_, rl), scene_grid=digest(box, gs, gl))
    (HERE / "builders.json").write_text(json.dumps(out, indent=1))
    print("builders.json written")


if __name__ == "__main__":
    main()
