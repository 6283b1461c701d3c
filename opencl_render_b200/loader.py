"""Scene loader: Wavefront OBJ/MTL (+ bitmaps) and PLY -> the scene arrays RaytraceAll takes.

Stands in for the Cinema4D scene extraction of the plugin (source/render.cpp:707-1308), which needs the C4D SDK, so that real
assets can be rendered through the same boundary.  It follows the plugin's conventions wherever they are observable:

* polygons: a triangle stays (a, b, c); a quad becomes (a, b, c) and (a, c, d) (render.cpp:733-736, 778-781); polygons with more
  corners (OBJ allows them, C4D does not) are fanned the same way from their first corner;
* normals: per-corner normals when the file has them, normalised in double and rounded to float (render.cpp:744-755); otherwise
  one flat normal per triangle, cross(b - a, c - a) scaled by +-1/length so that it FACES THE CAMERA (render.cpp:757-772) -- which
  is why the eye position is an input of the loader;
* UVs: the file's `vt` per corner; else, on request, one of the plugin's texture-tag projections (spherical, cylindrical, flat, cubic,
  shrink wrap; `project_uv` = ShdProjectPoint, render.cpp:495-673, applied as at render.cpp:920-944); else (0,0), (0,1), (1,1)
  (render.cpp:946-951);
* materials: five channel images per material in one 4-byte-per-texel atlas, channel order COLOR, REFLECTION, TRANSPARENCY, BUMP,
  LUMINANCE (raytrace_opencl.h:14-22), a bitmap copied row by row from the top (render.cpp:1165-1186), a plain colour stored as a
  1x1 image `floor(0.5 + 255 c)` (render.cpp:1256-1275), an enabled reflection channel without bitmap = 0.2 (render.cpp:1219-1227),
  and -- the plugin's quirk -- every ABSENT non-colour channel becomes a 1x1 BLACK image, bump included (render.cpp:1237-1246);
  `reference_fallbacks=False` leaves them absent (size 0) instead, which skips the kernel's bump path;
* lights: OBJ has none, so they are given by the caller; the defaults are the plugin's (radius 0.52 degrees = the sun's angular
  size, half-attenuation distance infinity; render.cpp:961, 976), directions are normalised (render.cpp:977-980).

MTL keys: Kd / map_Kd (colour), refl / map_refl or Ks with illum >= 3 (reflection), Tf / map_Tf, or d / Tr as a grey level
(transparency), map_bump / bump (bump), Ke / map_Ke (luminance).  `save_obj` writes a scene back in the same dialect (round-trip
tested).  Host code only; bitmaps are decoded with Pillow.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

from .api import (CH_BUMP, CH_COLOR, CH_LUMINANCE, CH_REFLECTION, CH_TRANSPARENCY, LIGHT_DISTANT, MATERIAL_CHANNEL_COUNT, HostScene)
from .scenes import MaterialAtlas, _lights

SUN_ANGLE_DEGREES = 0.52      # render.cpp:961
_CHANNEL_NAMES = {CH_COLOR: "color", CH_REFLECTION: "reflection", CH_TRANSPARENCY: "transparency", CH_BUMP: "bump", CH_LUMINANCE: "luminance"}


PROJECTIONS = ("spherical", "cylindrical", "flat", "cubic", "shrinkwrap", "volume")


def project_uv(p, n, projection: str = "spherical", ox: float = 0.0, oy: float = 0.0, lenx: float = 1.0, leny: float = 1.0):
    """The texture-tag projections the plugin falls back to for polygons without UVW data (ShdProjectPoint, render.cpp:495-673, used at
    render.cpp:920-944): point p (object space, double), polygon normal n -> (u, v).  `ox, oy, lenx, leny` = the tag's offset / length."""
    x, y, z = (float(c) for c in p[:3])
    lxi = 1.0 / lenx if lenx != 0.0 else 0.0
    lyi = 1.0 / leny if leny != 0.0 else 0.0
    pi, pi2 = math.pi, 2.0 * math.pi
    sq = math.sqrt(x * x + z * z)

    def around():     # angle of (x, z) about the y axis as a fraction of a turn  (:528-529, :552-553, :566-567)
        u = math.acos(max(-1.0, min(1.0, x / sq))) / pi2
        return 1.0 - u if z < 0.0 else u

    def wrap(u):      # :530-535
        u -= ox
        if lenx > 0.0 and u < 0.0:
            u += 1.0
        elif lenx < 0.0 and u > 0.0:
            u -= 1.0
        return u * lxi

    if projection == "volume":                                   # P_VOLUMESHADER :514-518
        return x, y
    if projection == "spherical":                                # :519-541
        if sq == 0.0:
            u, v = 0.0, (0.5 if y > 0.0 else -0.5)
        else:
            u, v = wrap(around()), 0.5 + math.atan(y / sq) / pi
        return u, -(v - oy) * lyi
    if projection == "shrinkwrap":                               # :542-561
        if sq == 0.0:
            u, v = 0.0, (0.0 if y > 0.0 else 1.0)
        else:
            u, v = around(), 0.5 - math.atan(y / sq) / pi
        sn, cs = math.sin(u * pi2), math.cos(u * pi2)
        return (0.5 + 0.5 * cs * v - ox) * lxi, (0.5 + 0.5 * sn * v - oy) * lyi
    if projection == "cylindrical":                              # :562-580
        u = 0.0 if sq == 0.0 else wrap(around())
        return u, -(y * 0.5 + oy) * lyi
    if projection == "flat":                                     # P_FLAT / P_SPATIAL :581-587
        return (x * 0.5 - ox) * lxi, -(y * 0.5 + oy) * lyi
    if projection == "cubic":                                    # :588-639
        nx, ny, nz = (float(c) for c in n[:3])
        if abs(nx) > abs(ny):
            axis = 0 if abs(nx) > abs(nz) else 2
        else:
            axis = 1 if abs(ny) > abs(nz) else 2
        if axis == 0:
            return ((-z if nx < 0.0 else z) * 0.5 - ox) * lxi, -(y * 0.5 + oy) * lyi
        if axis == 1:
            return (x * 0.5 - ox) * lxi, ((z if ny < 0.0 else -z) * 0.5 - oy) * lyi
        return ((x if nz < 0.0 else -x) * 0.5 - ox) * lxi, -(y * 0.5 + oy) * lyi
    raise ValueError(f"unknown projection {projection!r} (one of {PROJECTIONS})")


def _byte(c: float) -> int:
    """render.cpp:1268: (cl_uchar)floor(0.5f + c * 255.f)."""
    return int(math.floor(0.5 + float(np.float32(c)) * 255.0)) & 0xFF


def _load_bitmap(path: Path) -> np.ndarray:
    from PIL import Image     # Pillow is the only decoder in the image; fail loudly without it
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def parse_mtl(path: Path) -> dict:
    """-> {material name: {channel id: (r,g,b) bytes tuple | uint8 [h,w,3] array}}; only channels the file enables."""
    mats: dict = {}
    cur = None
    raw: dict = {}
    base = Path(path).parent

    def finish(name, r):
        if name is None:
            return
        ch = {}
        if "map_kd" in r:
            ch[CH_COLOR] = _load_bitmap(base / r["map_kd"])
        elif "kd" in r:
            ch[CH_COLOR] = tuple(_byte(c) for c in r["kd"])
        if "map_refl" in r:
            ch[CH_REFLECTION] = _load_bitmap(base / r["map_refl"])
        elif "refl" in r:
            ch[CH_REFLECTION] = tuple(_byte(c) for c in r["refl"])
        elif r.get("illum", 0) >= 3:
            ks = r.get("ks")
            ch[CH_REFLECTION] = tuple(_byte(c) for c in ks) if ks else (_byte(0.2),) * 3      # render.cpp:1219-1227
        if "map_tf" in r:
            ch[CH_TRANSPARENCY] = _load_bitmap(base / r["map_tf"])
        elif "tf" in r:
            ch[CH_TRANSPARENCY] = tuple(_byte(c) for c in r["tf"])
        elif "tr" in r and r["tr"] > 0:
            ch[CH_TRANSPARENCY] = (_byte(r["tr"]),) * 3
        elif "d" in r and r["d"] < 1:
            ch[CH_TRANSPARENCY] = (_byte(1.0 - r["d"]),) * 3
        if "map_bump" in r:
            ch[CH_BUMP] = _load_bitmap(base / r["map_bump"])
        if "map_ke" in r:
            ch[CH_LUMINANCE] = _load_bitmap(base / r["map_ke"])
        elif "ke" in r:
            ch[CH_LUMINANCE] = tuple(_byte(c) for c in r["ke"])
        mats[name] = ch

    for line in Path(path).read_text().splitlines():
        t = line.split("#", 1)[0].split()
        if not t:
            continue
        key = t[0].lower()
        if key == "newmtl":
            finish(cur, raw)
            cur, raw = " ".join(t[1:]), {}
        elif key in ("kd", "ks", "ke", "tf", "refl") and len(t) >= 4:
            raw[key] = tuple(float(x) for x in t[1:4])
        elif key in ("d", "tr") and len(t) >= 2:
            raw[key] = float(t[1])
        elif key == "illum" and len(t) >= 2:
            raw["illum"] = int(float(t[1]))
        elif key in ("map_kd", "map_refl", "map_tf", "map_ke", "map_bump", "bump") and len(t) >= 2:
            raw["map_bump" if key == "bump" else key] = t[-1]     # options (-bm ...) precede the file name
    finish(cur, raw)
    return mats


def _atlas_from(materials: list[dict], reference_fallbacks: bool) -> MaterialAtlas:
    atlas = MaterialAtlas()
    for ch in materials:
        args = {}
        for cid, name in _CHANNEL_NAMES.items():
            img = ch.get(cid)
            if img is None and cid == CH_COLOR:
                img = (255, 255, 255)                       # render.cpp:1249-1276: colour defaults to white
            if img is None and reference_fallbacks:
                img = (0, 0, 0)                             # render.cpp:1237-1246: absent non-colour channel = 1x1 black
            args[name] = img
        atlas.add(**args)
    return atlas


def load_obj(path, eye=(0.0, 0.0, 0.0), lights=None, reference_fallbacks: bool = True, normalise_normals: bool = True,
             name: str | None = None, uv_projection=None) -> HostScene:
    """Reads `path` (and the MTL files it names) into a HostScene.  `eye`: camera position (flat normals face it).
    `lights`: list of dict(type, pos, dir, colour, radius, half) -- see scenes._lights; default one distant sun.
    `uv_projection`: for faces without `vt`, a projection name (PROJECTIONS) or (name, dict(ox, oy, lenx, leny)) -- the plugin's
    texture-tag fallback (project_uv); None keeps the default triangle (0,0), (0,1), (1,1) it uses without a texture tag."""
    path = Path(path)
    v, vt, vn = [], [], []
    corners = []           # per triangle: 3 x (vi, ti, ni) with -1 = absent
    tri_mat_name = []
    mtl: dict = {}
    cur_mat = None
    for line in path.read_text().splitlines():
        t = line.split("#", 1)[0].split()
        if not t:
            continue
        key = t[0]
        if key == "v":
            v.append([float(x) for x in t[1:4]])
        elif key == "vt":
            vt.append([float(t[1]), float(t[2]) if len(t) > 2 else 0.0])
        elif key == "vn":
            vn.append([float(x) for x in t[1:4]])
        elif key == "mtllib":
            for m in t[1:]:
                if (path.parent / m).is_file():
                    mtl.update(parse_mtl(path.parent / m))
        elif key == "usemtl":
            cur_mat = " ".join(t[1:])
        elif key == "f":
            poly = []
            for c in t[1:]:
                parts = (c.split("/") + ["", ""])[:3]

                def idx(s, n):
                    if not s:
                        return -1
                    k = int(s)
                    k = k - 1 if k > 0 else n + k
                    if not 0 <= k < n:
                        raise ValueError(f"{path}: index {s} out of range in face '{line.strip()}'")
                    return k
                poly.append((idx(parts[0], len(v)), idx(parts[1], len(vt)), idx(parts[2], len(vn))))
            if len(poly) < 3:
                raise ValueError(f"{path}: face with fewer than 3 corners")
            for k in range(1, len(poly) - 1):              # (a,b,c), (a,c,d), ...   render.cpp:733-736, 778-781
                corners.append((poly[0], poly[k], poly[k + 1]))
                tri_mat_name.append(cur_mat)
    return _assemble(path, v, vt, vn, corners, tri_mat_name, mtl, eye, lights, reference_fallbacks, normalise_normals, name, uv_projection)


def _assemble(path, v, vt, vn, corners, tri_mat_name, mtl, eye, lights, reference_fallbacks, normalise_normals, name,
              uv_projection=None) -> HostScene:
    """Parsed geometry (positions, UVs, normals, per-triangle corner index triples (vertex, uv, normal; -1 = absent) and material names)
    -> HostScene with the plugin's conventions (module docstring)."""
    if not corners:
        raise ValueError(f"{path}: no faces")
    V = np.zeros((len(v), 4), np.float32)
    V[:, :3] = np.asarray(v, np.float64).astype(np.float32)
    n_tri = len(corners)
    c = np.asarray(corners, np.int64)                      # [N,3,3]
    idx = np.zeros((n_tri, 4), np.int32)
    idx[:, :3] = c[:, :, 0]

    # UVs: file values, else the plugin's default triangle (render.cpp:946-951)
    uv = np.tile(np.array([[0, 0], [0, 1], [1, 1]], np.float32), (n_tri, 1, 1))
    has_uv = (c[:, :, 1] >= 0).all(axis=1)
    if has_uv.any():
        vt_a = np.asarray(vt, np.float64).astype(np.float32)
        uv[has_uv] = vt_a[c[has_uv][:, :, 1]]
    if uv_projection is not None and (~has_uv).any():      # render.cpp:920-944: project the (untransformed) corners, n = (p1-p0) x (p2-p0)
        kind, kw = (uv_projection, {}) if isinstance(uv_projection, str) else (uv_projection[0], dict(uv_projection[1]))
        vd = np.asarray(v, np.float64)
        for t in np.nonzero(~has_uv)[0]:
            p0, p1, p2 = (vd[c[t, k, 0]] for k in range(3))
            nn = np.cross(p1 - p0, p2 - p0)
            for k, pk in enumerate((p0, p1, p2)):
                uv[t, k] = project_uv(pk, nn, kind, **kw)

    # normals
    nrm = np.zeros((n_tri, 3, 4), np.float32)
    has_n = (c[:, :, 2] >= 0).all(axis=1)
    if has_n.any():
        vn_a = np.asarray(vn, np.float64)
        n = vn_a[c[has_n][:, :, 2]]
        if normalise_normals:                              # Vector::Normalize in double, then (cl_float): render.cpp:744-755
            ln = np.sqrt((n * n).sum(axis=2, keepdims=True))
            n = n / np.where(ln > 0, ln, 1.0)
        nrm[has_n, :, :3] = n.astype(np.float32)
    if (~has_n).any():                                     # flat normal facing the camera: render.cpp:757-772 (fp32)
        f = np.float32
        sel = np.nonzero(~has_n)[0]
        a, b, cc = V[idx[sel, 0], :3], V[idx[sel, 1], :3], V[idx[sel, 2], :3]
        ab, ac = (b - a).astype(f), (cc - a).astype(f)
        tn = np.stack([ab[:, 1] * ac[:, 2] - ab[:, 2] * ac[:, 1], ab[:, 2] * ac[:, 0] - ab[:, 0] * ac[:, 2],
                       ab[:, 0] * ac[:, 1] - ab[:, 1] * ac[:, 0]], axis=1).astype(f)
        with np.errstate(divide="ignore", invalid="ignore"):
            len_inv = (f(1.0) / np.sqrt((tn * tn).sum(axis=1).astype(np.float64)).astype(f)).astype(f)
        to_a = (a - np.asarray(eye, f)[None, :3]).astype(f)        # vector(cameraEye, a) = a - eye
        away = (to_a * tn).sum(axis=1) >= 0
        len_inv = np.where(away, -len_inv, len_inv).astype(f)
        nrm[sel, :, :3] = (tn * len_inv[:, None])[:, None, :]

    # material ids = position in the material list of the document (render.cpp:1086-1100 numbers the document's materials, used or
    # not): here the order of the MTL files; names a face uses that no MTL defines -- and faces without usemtl -- follow, white
    order: list = list(mtl.keys())
    for m in tri_mat_name:
        if m not in order:
            order.append(m)
    mats = [mtl.get(m, {}) if m is not None else {} for m in order]
    atlas = _atlas_from(mats, reference_fallbacks)
    lut = {m: i for i, m in enumerate(order)}
    tri_mat = np.array([lut[m] for m in tri_mat_name], np.int32)
    size, start, tex = atlas.arrays()

    if lights is None:
        lights = [dict(type=LIGHT_DISTANT, dir=(0.3, -1.0, 0.2), colour=(1, 1, 1))]
    fixed = []
    for e in lights:
        e = dict(e)
        e.setdefault("radius", SUN_ANGLE_DEGREES)
        d = np.asarray(e.get("dir", (0, 0, 1)), np.float64)
        ln = math.sqrt(float((d * d).sum()))
        e["dir"] = tuple(d / ln) if ln > 0 else tuple(d)   # render.cpp:977-980
        fixed.append(e)
    lt, pos, dr, col, rad, half = _lights(fixed)
    return HostScene(vertex=V, tri_idx=idx, tri_mat=tri_mat, tri_uv=uv, tri_normal=nrm, mat_size=size, mat_start=start, textures=tex,
                     light_type=lt, light_pos=pos, light_dir=dr, light_colour=col, light_radius=rad, light_half=half,
                     name=name or path.stem, meta=dict(materials=[m or "(default)" for m in order])).normalise()


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def load_ply(path, eye=(0.0, 0.0, 0.0), lights=None, material: dict | None = None, reference_fallbacks: bool = True,
             normalise_normals: bool = True, name: str | None = None, uv_projection=None) -> HostScene:
    """Reads a PLY file (ascii, binary_little_endian or binary_big_endian) into a HostScene with the same conventions as load_obj.
    Used per vertex: x y z, nx ny nz (optional), s t or u v (optional); per face: vertex_indices / vertex_index (triangles, quads and
    larger polygons are fanned as render.cpp:733-781 does).  PLY has no materials: the one material of the scene is `material`
    ({channel id: (r, g, b) bytes | uint8 [h,w,3] image}, default: the plugin's white)."""
    path = Path(path)
    data = path.read_bytes()
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        raise ValueError(f"{path}: not a PLY file")
    body = data[data.index(b"\n", end) + 1:]
    fmt = None
    elements = []          # (name, count, [(kind, ...)])
    for line in data[:end].decode("ascii", "replace").splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append((t[1], int(t[2]), []))
        elif t[0] == "property" and elements:
            if t[1] == "list":
                elements[-1][2].append(("list", _PLY_TYPES[t[2]], _PLY_TYPES[t[3]], t[4]))
            else:
                elements[-1][2].append(("scalar", _PLY_TYPES[t[1]], t[2]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"{path}: unknown PLY format {fmt!r}")
    order = "<" if fmt == "binary_little_endian" else ">"
    vertex_cols: dict = {}
    faces: list = []
    tokens = body.split() if fmt == "ascii" else None
    tpos = 0
    off = 0
    for ename, count, props in elements:
        scalar_only = all(p[0] == "scalar" for p in props)
        if fmt != "ascii" and scalar_only:
            dt = np.dtype([(p[2], order + p[1]) for p in props])
            arr = np.frombuffer(body, dt, count, off)
            off += dt.itemsize * count
            cols = {p[2]: arr[p[2]] for p in props}
            rows = None
        else:
            cols, rows = {p[-1]: [] for p in props}, None
            for _ in range(count):
                for pr in props:
                    if pr[0] == "scalar":
                        if fmt == "ascii":
                            val = float(tokens[tpos]); tpos += 1
                        else:
                            dt = np.dtype(order + pr[1]); val = np.frombuffer(body, dt, 1, off)[0]; off += dt.itemsize
                        cols[pr[2]].append(val)
                    else:
                        if fmt == "ascii":
                            n = int(tokens[tpos]); tpos += 1
                            vals = [int(float(x)) for x in tokens[tpos:tpos + n]]; tpos += n
                        else:
                            dc, di = np.dtype(order + pr[1]), np.dtype(order + pr[2])
                            n = int(np.frombuffer(body, dc, 1, off)[0]); off += dc.itemsize
                            vals = np.frombuffer(body, di, n, off).astype(np.int64).tolist(); off += di.itemsize * n
                        cols[pr[3]].append(vals)
        if ename == "vertex":
            vertex_cols = {k: np.asarray(val, np.float64) for k, val in cols.items()}
        elif ename == "face":
            key = "vertex_indices" if "vertex_indices" in cols else ("vertex_index" if "vertex_index" in cols else None)
            if key is None:
                raise ValueError(f"{path}: face element without vertex_indices")
            faces = cols[key]
    if not {"x", "y", "z"} <= set(vertex_cols):
        raise ValueError(f"{path}: vertex element without x y z")
    nv = len(vertex_cols["x"])
    v = np.stack([vertex_cols["x"], vertex_cols["y"], vertex_cols["z"]], axis=1).tolist()
    has_n = {"nx", "ny", "nz"} <= set(vertex_cols)
    vn = np.stack([vertex_cols["nx"], vertex_cols["ny"], vertex_cols["nz"]], axis=1).tolist() if has_n else []
    ukey = ("s", "t") if {"s", "t"} <= set(vertex_cols) else (("u", "v") if {"u", "v"} <= set(vertex_cols) else None)
    vt = np.stack([vertex_cols[ukey[0]], vertex_cols[ukey[1]]], axis=1).tolist() if ukey else []
    corners, names = [], []
    for poly in faces:
        if len(poly) < 3:
            raise ValueError(f"{path}: face with fewer than 3 corners")
        if min(poly) < 0 or max(poly) >= nv:
            raise ValueError(f"{path}: vertex index out of range in a face")
        pc = [(int(i), int(i) if ukey else -1, int(i) if has_n else -1) for i in poly]
        for k in range(1, len(pc) - 1):
            corners.append((pc[0], pc[k], pc[k + 1]))
            names.append("ply")
    mtl = {"ply": dict(material or {})}
    return _assemble(path, v, vt, vn, corners, names, mtl, eye, lights, reference_fallbacks, normalise_normals, name, uv_projection)


def load_scene(path, **kw) -> HostScene:
    """OBJ or PLY by file extension."""
    return load_ply(path, **kw) if str(path).lower().endswith(".ply") else load_obj(path, **kw)


def save_obj(scene: HostScene, path, bitmap_format: str = "png") -> Path:
    """Writes `scene` as OBJ + MTL (+ one bitmap per non-1x1 channel) in the dialect load_obj reads; floats with 9 significant
    digits so that a load of the result reproduces the arrays bit for bit."""
    path = Path(path)
    scene.normalise()
    g = lambda x: f"{float(x):.9g}"
    lines = [f"mtllib {path.stem}.mtl"]
    for p in scene.vertex[:, :3]:
        lines.append("v " + " ".join(g(x) for x in p))
    for t in range(scene.triangle_count):
        for k in range(3):
            lines.append("vt " + " ".join(g(x) for x in scene.tri_uv[t, k]))
    for t in range(scene.triangle_count):
        for k in range(3):
            lines.append("vn " + " ".join(g(x) for x in scene.tri_normal[t, k, :3]))
    last = None
    for t in range(scene.triangle_count):
        m = int(scene.tri_mat[t])
        if m != last:
            lines.append(f"usemtl m{m}")
            last = m
        lines.append("f " + " ".join(f"{int(scene.tri_idx[t, k]) + 1}/{3 * t + k + 1}/{3 * t + k + 1}" for k in range(3)))
    path.write_text("\n".join(lines) + "\n")
    keys = {CH_COLOR: ("Kd", "map_Kd"), CH_REFLECTION: ("refl", "map_refl"), CH_TRANSPARENCY: ("Tf", "map_Tf"), CH_BUMP: (None, "map_bump"),
            CH_LUMINANCE: ("Ke", "map_Ke")}
    out = []
    for m in range(scene.material_count):
        out.append(f"newmtl m{m}")
        for ch in range(MATERIAL_CHANNEL_COUNT):
            w, h = (int(x) for x in scene.mat_size[MATERIAL_CHANNEL_COUNT * m + ch])
            if w == 0:
                continue
            s = int(scene.mat_start[MATERIAL_CHANNEL_COUNT * m + ch])
            img = scene.textures[s:s + w * h, :3].reshape(h, w, 3)
            plain, bitmap = keys[ch]
            if w == 1 and h == 1 and plain:
                # c with floor(0.5 + 255 c) == byte
                out.append(f"{plain} " + " ".join(g(np.float32(int(b)) / np.float32(255.0)) for b in img[0, 0]))
            else:
                from PIL import Image
                fn = f"{path.stem}_m{m}_{_CHANNEL_NAMES[ch]}.{bitmap_format}"
                Image.fromarray(img, "RGB").save(path.parent / fn)
                out.append(f"{bitmap} {fn}")
    (path.parent / f"{path.stem}.mtl").write_text("\n".join(out) + "\n")
    return path
