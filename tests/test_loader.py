"""Scene loader (opencl_render_b200/loader.py): OBJ/MTL -> the plugin's scene-array conventions (source/render.cpp:707-1308).
CPU tests: conventions one by one on hand-written files, and a save/load round trip that reproduces the arrays of the synthetic
scenes bit for bit -- so a loaded scene renders to the golden planes (GPU test at the end)."""
from pathlib import Path

import numpy as np
import pytest

from opencl_render_b200 import api, loader, scenes
from tests import helpers

GOLDEN = Path(__file__).resolve().parent / "golden"


def _write(tmp_path, obj, mtl=None):
    (tmp_path / "s.obj").write_text(obj)
    if mtl is not None:
        (tmp_path / "s.mtl").write_text(mtl)
    return tmp_path / "s.obj"


def test_quad_split_default_uv_and_camera_facing_flat_normals(tmp_path):
    p = _write(tmp_path, "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 2\nf 1 2 3 4\nf -5 -4 -1\n")
    sc = loader.load_obj(p, eye=(0.5, 0.5, -3.0))
    assert sc.triangle_count == 3
    # quad (a,b,c,d) -> (a,b,c), (a,c,d)   render.cpp:733-736, 778-781
    assert sc.tri_idx[:, :3].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4]]
    # no vt -> (0,0), (0,1), (1,1)   render.cpp:946-951
    assert np.array_equal(sc.tri_uv[0], np.array([[0, 0], [0, 1], [1, 1]], np.float32))
    # no vn -> flat normal, flipped to face the eye: render.cpp:757-772
    assert np.allclose(sc.tri_normal[0, :, :3], [[0, 0, -1]] * 3) and np.allclose(sc.tri_normal[1, :, :3], [[0, 0, -1]] * 3)
    other = loader.load_obj(p, eye=(0.5, 0.5, +3.0))
    assert np.allclose(other.tri_normal[0, :, :3], [[0, 0, 1]] * 3)
    # one default material: white colour, the other channels 1x1 black (the plugin's fallback, render.cpp:1237-1246)
    assert sc.material_count == 1 and sc.mat_size.tolist() == [[1, 1]] * 5
    assert sc.textures[:5, :3].tolist() == [[255, 255, 255], [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]]
    assert sc.mat_start.tolist() == [0, 1, 2, 3, 4, 5]
    lean = loader.load_obj(p, eye=(0, 0, -3), reference_fallbacks=False)
    assert lean.mat_size.tolist() == [[1, 1], [0, 0], [0, 0], [0, 0], [0, 0]] and lean.mat_start.tolist() == [0, 1, 1, 1, 1, 1]
    # default light: one distant sun with the plugin's constants (render.cpp:961, 976-980)
    assert sc.light_count == 1 and sc.light_type[0] == api.LIGHT_DISTANT and abs(sc.light_radius[0] - 0.52) < 1e-7
    assert np.isinf(sc.light_half[0]) and abs(np.linalg.norm(sc.light_dir[0, :3]) - 1) < 1e-6


def test_materials_channels_bitmaps_and_corner_attributes(tmp_path):
    from PIL import Image
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3) * 9
    Image.fromarray(img, "RGB").save(tmp_path / "wood.png")
    mtl = ("newmtl wood\nmap_Kd wood.png\nKe 0.5 0 0\n"
           "newmtl glass\nKd 0.2 0.4 1.0\nTf 0.5 0.5 0.5\nillum 5\n"
           "newmtl chrome\nKd 1 1 1\nKs 0.8 0.8 0.8\nillum 3\nd 0.25\nmap_bump -bm 2 wood.png\n")
    obj = ("mtllib s.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvt 0.25 0.5\nvt 1 0\nvt 0 1\nvn 0 0 -2\n"
           "usemtl glass\nf 1/1/1 2/2/1 3/3/1\nusemtl wood\nf 2/2/1 4/1/1 3/3/1\nusemtl chrome\nf 1//1 2//1 4//1\nusemtl glass\nf 1 2 4\n")
    sc = loader.load_obj(_write(tmp_path, obj, mtl), eye=(0, 0, -5))
    assert sc.meta["materials"] == ["wood", "glass", "chrome"]            # ids = order of the material list (the MTL file)
    assert sc.tri_mat.tolist() == [1, 0, 2, 1]
    assert np.array_equal(sc.tri_uv[0], np.array([[0.25, 0.5], [1, 0], [0, 1]], np.float32))
    assert np.array_equal(sc.tri_uv[2], np.array([[0, 0], [0, 1], [1, 1]], np.float32))       # no vt on that face
    assert np.array_equal(sc.tri_normal[0, :, :3], np.array([[0, 0, -1]] * 3, np.float32))   # vn normalised (render.cpp:744-749)
    size = sc.mat_size.reshape(3, 5, 2)
    start = sc.mat_start
    tex = lambda m, ch: sc.textures[start[5 * m + ch]: start[5 * m + ch] + size[m, ch, 0] * size[m, ch, 1], :3]
    # glass: Kd -> floor(0.5 + 255 c) (render.cpp:1268), Tf grey, illum 5 without Ks -> reflection 0.2 (render.cpp:1219-1227)
    assert tex(1, api.CH_COLOR).tolist() == [[51, 102, 255]] and tex(1, api.CH_TRANSPARENCY).tolist() == [[128, 128, 128]]
    assert tex(1, api.CH_REFLECTION).tolist() == [[51, 51, 51]] and tex(1, api.CH_BUMP).tolist() == [[0, 0, 0]]
    # wood: bitmap rows from the top, x fastest (render.cpp:1174-1184); Ke -> luminance
    assert size[0, api.CH_COLOR].tolist() == [3, 2] and np.array_equal(tex(0, api.CH_COLOR), img.reshape(-1, 3))
    assert tex(0, api.CH_LUMINANCE).tolist() == [[128, 0, 0]]
    # chrome: Ks with illum 3 -> reflection, d -> transparency 1 - d, bump bitmap (options before the file name are skipped)
    assert tex(2, api.CH_REFLECTION).tolist() == [[204, 204, 204]] and tex(2, api.CH_TRANSPARENCY).tolist() == [[191, 191, 191]]
    assert size[2, api.CH_BUMP].tolist() == [3, 2]
    assert start[-1] == sc.textures.shape[0]


def test_bad_files_fail_loudly(tmp_path):
    with pytest.raises(ValueError):
        loader.load_obj(_write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n"))
    with pytest.raises(ValueError):
        loader.load_obj(_write(tmp_path, "v 0 0 0\n"))


ROUND_TRIP = ["soup", "spheres_mirror", "terrain_textured"]


@pytest.mark.parametrize("name", ROUND_TRIP)
def test_save_load_round_trip_reproduces_the_arrays(tmp_path, name):
    sc, cam, lists, samples = helpers.make_case(name)
    loader.save_obj(sc, tmp_path / "scene.obj")
    lights = [dict(type=int(sc.light_type[i]), pos=sc.light_pos[i, :3], dir=sc.light_dir[i, :3], colour=sc.light_colour[i, :3],
                   radius=float(sc.light_radius[i]), half=float(sc.light_half[i])) for i in range(sc.light_count)]
    back = loader.load_obj(tmp_path / "scene.obj", eye=sc.meta["camera"]["eye"], lights=lights, reference_fallbacks=False,
                           normalise_normals=False)
    for f in ("vertex", "tri_idx", "tri_mat", "tri_uv", "tri_normal", "mat_size", "mat_start", "textures", "light_type", "light_pos",
              "light_colour", "light_radius", "light_half"):
        assert np.array_equal(getattr(back, f), getattr(sc, f)), f
    assert np.allclose(back.light_dir, sc.light_dir / np.maximum(np.linalg.norm(sc.light_dir[:, :3], axis=1, keepdims=True), 1e-30), atol=1e-6)


@pytest.mark.gpu
def test_loaded_scene_renders_to_the_golden_planes(tmp_path):
    """OBJ -> loader -> device builders (grid + camera lists in HBM) -> trace -> PNG: the whole harness path."""
    sc, cam, lists, samples = helpers.make_case("spheres_mirror")
    loader.save_obj(sc, tmp_path / "scene.obj")
    light = dict(type=int(sc.light_type[0]), pos=sc.light_pos[0, :3], dir=sc.light_dir[0, :3], colour=sc.light_colour[0, :3],
                 radius=float(sc.light_radius[0]), half=float(sc.light_half[0]))
    back = loader.load_obj(tmp_path / "scene.obj", eye=sc.meta["camera"]["eye"], lights=[light], reference_fallbacks=False,
                           normalise_normals=False)
    if not np.array_equal(back.light_dir, sc.light_dir):
        pytest.skip("light direction of this case is not unit length")
    ds = api.DeviceScene(back, 0)
    fr = api.DeviceFrame(ds, cam)
    fr.render(samples)
    img = fr.read()
    gold = np.load(GOLDEN / "spheres_mirror.npz")
    assert all(np.array_equal(img[c], gold[k]) for c, k in enumerate("rgb"))
    api.write_image(tmp_path / "out.png", img)
    assert (tmp_path / "out.png").stat().st_size > 100


@pytest.mark.gpu
def test_render_harness_cli(tmp_path):
    from opencl_render_b200 import render
    out = tmp_path / "img.ppm"
    ck = tmp_path / "job.npz"
    args = ["config1", "--size", "160x120", "--samples", "6", "--pass-samples", "4", "--out", str(out), "--checkpoint", str(ck), "--quiet"]
    assert render.main(args) == 0
    first = out.read_bytes()
    sc = scenes.CONFIGS[1]["make"]()
    m = sc.meta["camera"]
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], 160, 120)
    fr = api.DeviceFrame(api.DeviceScene(sc, 0), cam)
    fr.render(6)
    want = fr.read()
    px = np.frombuffer(first[len(b"P6\n160 120\n65535\n"):], ">u2").reshape(120, 160, 3)
    assert all(np.array_equal(px[..., c], want[c]) for c in range(3))
    # the finished checkpoint resumes to the same picture without rendering anything
    assert int(np.load(ck)["done"]) == 6 and render.main(args) == 0 and out.read_bytes() == first


def _ply_text(verts, faces, normals=None, uvs=None):
    props = ["property float x", "property float y", "property float z"]
    if normals is not None:
        props += ["property float nx", "property float ny", "property float nz"]
    if uvs is not None:
        props += ["property float s", "property float t"]
    head = ["ply", "format ascii 1.0", "comment test", f"element vertex {len(verts)}", *props, f"element face {len(faces)}",
            "property list uchar int vertex_indices", "end_header"]
    rows = []
    for i, p in enumerate(verts):
        r = list(p) + (list(normals[i]) if normals is not None else []) + (list(uvs[i]) if uvs is not None else [])
        rows.append(" ".join(f"{float(x):.9g}" for x in r))
    for f in faces:
        rows.append(" ".join(str(x) for x in [len(f), *f]))
    return "\n".join(head + rows) + "\n"


def test_ply_ascii_and_binary_agree_with_the_obj_loader(tmp_path):
    import struct
    verts = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0.5, 0.5, 1.5)]
    faces = [[0, 1, 2, 3], [0, 1, 4], [1, 2, 4]]
    (tmp_path / "a.ply").write_text(_ply_text(verts, faces))
    ply = loader.load_scene(tmp_path / "a.ply", eye=(0.3, 0.2, -4.0))
    (tmp_path / "a.obj").write_text("".join(f"v {x} {y} {z}\n" for x, y, z in verts) + "".join("f " + " ".join(str(i + 1) for i in f) + "\n" for f in faces))
    obj = loader.load_scene(tmp_path / "a.obj", eye=(0.3, 0.2, -4.0))
    assert ply.triangle_count == 4 and ply.tri_idx[:, :3].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [1, 2, 4]]      # quad split as render.cpp:733-781
    for f in ("vertex", "tri_idx", "tri_mat", "tri_uv", "tri_normal", "mat_size", "mat_start", "textures", "light_type", "light_dir"):
        assert np.array_equal(getattr(ply, f), getattr(obj, f)), f
    # the same file as binary_little_endian / binary_big_endian with normals and uvs per vertex
    normals = [(0, 0, -2), (0, 0, -1), (0, 0, -1), (0, 0, -1), (0, 3, 0)]
    uvs = [(0, 0), (1, 0), (1, 1), (0, 1), (0.5, 0.25)]
    ascii_scene = None
    for fmt, order in (("ascii", None), ("binary_little_endian", "<"), ("binary_big_endian", ">")):
        p = tmp_path / f"{fmt}.ply"
        if order is None:
            p.write_text(_ply_text(verts, faces, normals, uvs))
        else:
            head = "\n".join(["ply", f"format {fmt} 1.0", f"element vertex {len(verts)}", "property float x", "property float y",
                              "property float z", "property float nx", "property float ny", "property float nz", "property float s",
                              "property float t", f"element face {len(faces)}", "property list uchar uint vertex_indices", "end_header"]) + "\n"
            blob = b"".join(struct.pack(order + "8f", *verts[i], *normals[i], *uvs[i]) for i in range(len(verts)))
            blob += b"".join(struct.pack(order + "B" + str(len(f)) + "I", len(f), *f) for f in faces)
            p.write_bytes(head.encode() + blob)
        sc = loader.load_ply(p, eye=(0, 0, -4), material={api.CH_COLOR: (10, 20, 30), api.CH_REFLECTION: (128, 128, 128)})
        if ascii_scene is None:
            ascii_scene = sc
            assert np.array_equal(sc.tri_normal[0, 0, :3], np.array([0, 0, -1], np.float32))           # normalised (render.cpp:744-749)
            assert np.array_equal(sc.tri_uv[2], np.array([[0, 0], [1, 0], [0.5, 0.25]], np.float32))
            assert sc.textures[:2, :3].tolist() == [[10, 20, 30], [128, 128, 128]] and sc.material_count == 1
        for f in ("vertex", "tri_idx", "tri_uv", "tri_normal", "textures", "mat_size"):
            assert np.array_equal(getattr(sc, f), getattr(ascii_scene, f)), (fmt, f)
    with pytest.raises(ValueError):
        (tmp_path / "bad.ply").write_text(_ply_text(verts, [[0, 1, 9]]))
        loader.load_ply(tmp_path / "bad.ply")
    with pytest.raises(ValueError):
        (tmp_path / "nope.ply").write_text("solid\n")
        loader.load_ply(tmp_path / "nope.ply")


def test_render_harness_arguments_without_a_gpu(capsys):
    """The harness's device list is the dialog's combo (index 0 = the reference's CPU entry, a label only); picking it is refused."""
    from opencl_render_b200 import _lib, render
    assert render.main(["--list-devices"]) == 0
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "0: Local CPU single thread"
    for bad in (["config1", "--device", "0"], []):
        with pytest.raises(SystemExit):
            render.main(bad)
    if _lib.load().oclr_device_count() == 0:
        with pytest.raises(SystemExit):
            render.main(["config1", "--device", "1"])      # no CUDA device enumerated: there is nothing to pick


def test_texture_tag_projections():
    """project_uv restates ShdProjectPoint (render.cpp:495-673): spot values per projection, derived from the reference's formulas."""
    P = loader.project_uv
    n = (0.0, 0.0, 1.0)
    # spherical: u = angle about y / 2pi (mirrored for z < 0), v = -(0.5 + atan(y / sqrt(x^2+z^2)) / pi)
    assert P((1, 0, 0), n, "spherical") == pytest.approx((0.0, -0.5))
    assert P((0, 0, 1), n, "spherical") == pytest.approx((0.25, -0.5))
    assert P((0, 0, -1), n, "spherical") == pytest.approx((0.75, -0.5))
    assert P((1, 1, 0), n, "spherical") == pytest.approx((0.0, -0.75))
    assert P((0, 2, 0), n, "spherical") == pytest.approx((0.0, -0.5)) and P((0, -2, 0), n, "spherical") == pytest.approx((0.0, 0.5))
    assert P((0, 0, 1), n, "spherical", ox=0.5, lenx=0.5) == pytest.approx((1.5, -0.5))        # (0.25 - 0.5 + 1) / 0.5
    # cylindrical / flat: v = -(y/2 + oy) / leny
    assert P((0, 3, 1), n, "cylindrical", oy=0.5, leny=2.0) == pytest.approx((0.25, -1.0))
    assert P((0, 3, 0), n, "cylindrical") == pytest.approx((0.0, -1.5))
    assert P((2, 3, 9), n, "flat", ox=0.25) == pytest.approx((0.75, -1.5))
    # cubic: dominant normal axis picks the plane, sign picks the mirror
    assert P((2, 4, 6), (1, 0, 0), "cubic") == pytest.approx((3.0, -2.0)) and P((2, 4, 6), (-1, 0, 0), "cubic") == pytest.approx((-3.0, -2.0))
    assert P((2, 4, 6), (0, 1, 0), "cubic") == pytest.approx((1.0, -3.0)) and P((2, 4, 6), (0, -1, 0), "cubic") == pytest.approx((1.0, 3.0))
    assert P((2, 4, 6), (0, 0, 1), "cubic") == pytest.approx((-1.0, -2.0)) and P((2, 4, 6), (0, 0, -1), "cubic") == pytest.approx((1.0, -2.0))
    assert P((2, 4, 6), (1, 1, 1), "cubic") == pytest.approx((-1.0, -2.0))        # ties fall through to z (render.cpp:594-606)
    # shrink wrap: the pole maps to the centre, the equator to a circle of radius 1/4
    assert P((0, 5, 0), n, "shrinkwrap") == pytest.approx((0.5, 0.5)) and P((1, 0, 0), n, "shrinkwrap") == pytest.approx((0.75, 0.5))
    assert P((1, 2, 3), n, "volume") == (1.0, 2.0)
    with pytest.raises(ValueError):
        P((0, 0, 0), n, "frontal")


def test_loader_applies_the_projection_to_faces_without_uvs(tmp_path):
    p = _write(tmp_path, "v 0 0 0\nv 2 0 0\nv 2 2 0\nvt 0.1 0.2\nf 1 2 3\nf 1/1 2/1 3/1\n")
    sc = loader.load_obj(p, eye=(0, 0, -5), uv_projection=("flat", dict(ox=0.0, oy=0.0, lenx=2.0, leny=1.0)))
    assert np.allclose(sc.tri_uv[0], [[0, 0], [0.5, 0], [0.5, -1.0]])          # u = x/2/lenx, v = -y/2
    assert np.allclose(sc.tri_uv[1], [[0.1, 0.2]] * 3)                          # faces with vt keep their own
