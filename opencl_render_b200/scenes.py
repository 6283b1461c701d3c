"""Synthetic scenes for the five BASELINE.json configs in the reference's array conventions.

The Cinema4D scene extraction (source/render.cpp:707-1308) is out of scope; these generators stand in for it and
produce exactly what `parseAndRender` would hand to RaytraceAll: vertices (cl_float3), triangles (cl_int3), per-corner
normals and UVs, materials as five channel images each in one uchar3 atlas (channel order COLOR, REFLECTION,
TRANSPARENCY, BUMP, LUMINANCE; size.x == 0 = absent; render.cpp:1136-1306), and lights.  All geometry stays in front of
every camera (trianglelist.cpp:546 TODO: triangles behind the eye break the camera lists).

Every generator takes size parameters so tests can run small instances of the same scene family.
"""
from __future__ import annotations

import math

import numpy as np

from .api import (CH_BUMP, CH_COLOR, CH_LUMINANCE, CH_REFLECTION, CH_TRANSPARENCY, LIGHT_DISTANT, LIGHT_SPOT,
                  MATERIAL_CHANNEL_COUNT, HostScene)

INF = np.float32(np.inf)

# dimensions of the reference's res/tex/*.jpg bitmaps (12 054 446 texels in total); config 4 uses procedural images of
# the same sizes because the JPEGs themselves are reference assets and are not copied into this repository.
RES_TEX_SIZES = [(1024, 1024), (512, 512), (730, 229), (756, 512), (730, 229), (930, 2000), (2600, 1800), (2000, 756),
                 (550, 552), (1926, 794), (295, 466)]


class MaterialAtlas:
    """Builds materialImageSize / materialImageStart / textures like render.cpp:1136-1306."""

    def __init__(self):
        self.sizes = []
        self.starts = []
        self.texels = []
        self.cursor = 0

    def add(self, color=None, reflection=None, transparency=None, bump=None, luminance=None) -> int:
        """Each channel: None (absent), an (r,g,b) tuple (1x1 image) or a uint8 [h,w,3] array.  Returns material id."""
        chans = {CH_COLOR: color, CH_REFLECTION: reflection, CH_TRANSPARENCY: transparency, CH_BUMP: bump, CH_LUMINANCE: luminance}
        for ch in range(MATERIAL_CHANNEL_COUNT):
            img = chans[ch]
            self.starts.append(self.cursor)
            if img is None:
                self.sizes.append((0, 0))
                continue
            a = np.asarray(img, dtype=np.uint8)
            if a.ndim == 1:
                a = a.reshape(1, 1, 3)
            h, w = a.shape[:2]
            t = np.zeros((h * w, 4), dtype=np.uint8)
            t[:, :3] = a.reshape(-1, 3)
            self.texels.append(t)
            self.sizes.append((w, h))
            self.cursor += h * w
        return len(self.sizes) // MATERIAL_CHANNEL_COUNT - 1

    def arrays(self):
        size = np.array(self.sizes, dtype=np.uint32).reshape(-1, 2)
        start = np.array(self.starts + [self.cursor], dtype=np.int32)
        tex = np.concatenate(self.texels, axis=0) if self.texels else np.zeros((1, 4), dtype=np.uint8)
        return size, start, tex


def _lights(entries):
    """entries: list of dict(type,pos,dir,colour,radius,half)."""
    n = len(entries)
    lt = np.zeros(n, np.int32)
    pos = np.zeros((n, 4), np.float32)
    dr = np.zeros((n, 4), np.float32)
    col = np.zeros((n, 4), np.float32)
    rad = np.zeros(n, np.float32)
    half = np.zeros(n, np.float32)
    for i, e in enumerate(entries):
        lt[i] = e["type"]
        pos[i, :3] = e.get("pos", (0, 0, 0))
        dr[i, :3] = e.get("dir", (0, -1, 0))
        col[i, :3] = e.get("colour", (1, 1, 1))
        rad[i] = e.get("radius", 0.0)
        half[i] = e.get("half", INF)
    return lt, pos, dr, col, rad, half


def _scene(name, vertex, tri_idx, tri_mat, tri_uv, tri_normal, atlas, lights, meta=None) -> HostScene:
    v4 = np.zeros((vertex.shape[0], 4), np.float32)
    v4[:, :3] = vertex
    i4 = np.zeros((tri_idx.shape[0], 4), np.int32)
    i4[:, :3] = tri_idx
    n4 = np.zeros((tri_idx.shape[0], 3, 4), np.float32)
    n4[:, :, :3] = tri_normal
    size, start, tex = atlas.arrays()
    lt, pos, dr, col, rad, half = _lights(lights)
    return HostScene(vertex=v4, tri_idx=i4, tri_mat=np.asarray(tri_mat, np.int32), tri_uv=np.asarray(tri_uv, np.float32),
                     tri_normal=n4, mat_size=size, mat_start=start, textures=tex, light_type=lt, light_pos=pos, light_dir=dr,
                     light_colour=col, light_radius=rad, light_half=half, name=name, meta=meta or {}).normalise()


def _flat_normals(vertex, tri_idx):
    a, b, c = vertex[tri_idx[:, 0]], vertex[tri_idx[:, 1]], vertex[tri_idx[:, 2]]
    n = np.cross(b - a, c - a)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = n / np.where(ln > 0, ln, 1)
    return np.repeat(n[:, None, :], 3, axis=1).astype(np.float32)


# ---- config 1: random triangle soup -------------------------------------------------------------------------------
def soup(n_tri: int = 1000, seed: int = 1234, light_radius: float = 0.0, reflective: bool = False,
         transparent: bool = False) -> HostScene:
    rng = np.random.Generator(np.random.PCG64(seed))
    centre = np.stack([rng.uniform(-2, 2, n_tri), rng.uniform(-0.5, 2.5, n_tri), rng.uniform(-2, 2, n_tri)], axis=1)
    off = rng.uniform(-0.3, 0.3, (n_tri, 3, 3))
    vertex = (centre[:, None, :] + off).reshape(-1, 3).astype(np.float32)
    tri_idx = np.arange(3 * n_tri, dtype=np.int32).reshape(n_tri, 3)
    atlas = MaterialAtlas()
    m0 = atlas.add(color=(230, 140, 60), reflection=(0, 0, 0), transparency=(0, 0, 0), luminance=(0, 0, 0))
    mats = np.full(n_tri, m0, np.int32)
    if reflective:
        m1 = atlas.add(color=(200, 200, 220), reflection=(160, 160, 160), transparency=(0, 0, 0), luminance=(0, 0, 0))
        mats[1::3] = m1
    if transparent:
        m2 = atlas.add(color=(120, 220, 160), reflection=(0, 0, 0), transparency=(140, 150, 130), luminance=(10, 0, 0))
        mats[2::3] = m2
    uv = np.tile(np.array([[0, 0], [1, 0], [0, 1]], np.float32), (n_tri, 1, 1))
    lights = [dict(type=LIGHT_SPOT, pos=(3.0, 8.0, -5.0), colour=(1, 1, 1), radius=light_radius)]
    return _scene(f"soup{n_tri}", vertex, tri_idx, mats, uv, _flat_normals(vertex, tri_idx), atlas, lights,
                  meta=dict(camera=dict(eye=(0, 4.4, -8), look_at=(0, 0, 0), up=(0, 1, 0), fov=0.9)))


# ---- config 2: grid of UV spheres on a floor ----------------------------------------------------------------------------
def _uv_sphere(stacks: int, slices: int):
    """Unit sphere: 2 poles + (stacks-1) rings of `slices` vertices; 2*slices + (stacks-2)*slices*2 triangles."""
    verts = [(0.0, 1.0, 0.0)]
    for i in range(1, stacks):
        th = math.pi * i / stacks
        for j in range(slices):
            ph = 2 * math.pi * j / slices
            verts.append((math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)))
    verts.append((0.0, -1.0, 0.0))
    verts = np.array(verts, np.float64)
    tris = []
    ring = lambda i, j: 1 + (i - 1) * slices + (j % slices)
    for j in range(slices):
        tris.append((0, ring(1, j + 1), ring(1, j)))
    for i in range(1, stacks - 1):
        for j in range(slices):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j + 1), ring(i + 1, j)
            tris.append((a, b, c))
            tris.append((a, c, d))
    south = len(verts) - 1
    for j in range(slices):
        tris.append((south, ring(stacks - 1, j), ring(stacks - 1, j + 1)))
    return verts, np.array(tris, np.int32)


def sphere_grid(grid: int = 6, stacks: int = 27, slices: int = 54, pitch: float = 2.5, radius: float = 1.0,
                reflection: int = 0, light_radius: float = 0.0) -> HostScene:
    sv, st = _uv_sphere(stacks, slices)
    vs, ts, ns = [], [], []
    base = 0
    half = (grid - 1) * pitch / 2
    for gz in range(grid):
        for gx in range(grid):
            centre = np.array([gx * pitch - half, radius, gz * pitch - half])
            vs.append(sv * radius + centre)
            ts.append(st + base)
            ns.append(sv)
            base += sv.shape[0]
    ext = half + 2 * pitch
    floor_v = np.array([[-ext, 0, -ext], [ext, 0, -ext], [ext, 0, ext], [-ext, 0, ext]], np.float64)
    vs.append(floor_v)
    ts.append(np.array([[0, 2, 1], [0, 3, 2]], np.int32) + base)
    ns.append(np.tile(np.array([[0, 1, 0]], np.float64), (4, 1)))
    vertex = np.concatenate(vs).astype(np.float32)
    tri_idx = np.concatenate(ts).astype(np.int32)
    vnorm = np.concatenate(ns).astype(np.float32)
    tri_normal = vnorm[tri_idx]
    n_tri = tri_idx.shape[0]
    atlas = MaterialAtlas()
    refl = (reflection,) * 3
    m_sphere = atlas.add(color=(200, 60, 50), reflection=refl, transparency=(0, 0, 0), luminance=(0, 0, 0))
    m_floor = atlas.add(color=(180, 180, 170), reflection=(0, 0, 0), transparency=(0, 0, 0), luminance=(0, 0, 0))
    mats = np.full(n_tri, m_sphere, np.int32)
    mats[-2:] = m_floor
    uv = np.tile(np.array([[0, 0], [1, 0], [0, 1]], np.float32), (n_tri, 1, 1))
    lights = [dict(type=LIGHT_SPOT, pos=(6.0, 14.0, -8.0), colour=(1, 1, 1), radius=light_radius)]
    return _scene(f"spheres{grid}x{grid}", vertex, tri_idx, mats, uv, tri_normal, atlas, lights,
                  meta=dict(camera=dict(eye=(0, 9.0, -16.0), look_at=(0, 0.5, 0), up=(0, 1, 0), fov=0.9)))


# ---- config 3/4/5: procedural terrain ---------------------------------------------------------------------------------------
def _value_noise(nx: int, nz: int, seed: int, octaves: int = 4, base_cells: int = 8) -> np.ndarray:
    """4-octave value noise on an (nz+1) x (nx+1) lattice, values in [0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xs = np.linspace(0, 1, nx + 1)
    zs = np.linspace(0, 1, nz + 1)
    out = np.zeros((nz + 1, nx + 1))
    amp, total = 1.0, 0.0
    for o in range(octaves):
        cells = base_cells * (2 ** o)
        lat = rng.random((cells + 2, cells + 2))
        fx, fz = xs * cells, zs * cells
        ix, iz = np.minimum(fx.astype(int), cells), np.minimum(fz.astype(int), cells)
        tx, tz = fx - ix, fz - iz
        tx, tz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        a = lat[np.ix_(iz, ix)]
        b = lat[np.ix_(iz, ix + 1)]
        c = lat[np.ix_(iz + 1, ix)]
        d = lat[np.ix_(iz + 1, ix + 1)]
        out += amp * ((a * (1 - tx)[None, :] + b * tx[None, :]) * (1 - tz)[:, None] + (c * (1 - tx)[None, :] + d * tx[None, :]) * tz[:, None])
        total += amp
        amp *= 0.5
    return out / total


def _procedural_texture(w: int, h: int, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    y, x = np.mgrid[0:h, 0:w]
    base = rng.integers(40, 216, 3)
    f = rng.uniform(2, 9, 4)
    img = np.zeros((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = base[c] + 38 * np.sin(2 * np.pi * (f[0] * x / w + 0.3 * c)) * np.cos(2 * np.pi * f[1] * y / h) \
            + 22 * np.sin(2 * np.pi * (f[2] * (x + y) / (w + h)))
    img += rng.integers(-12, 13, (h, w, 1))
    return np.clip(img, 0, 255).astype(np.uint8)


def terrain(quads: int = 708, seed: int = 7, extent: float = 40.0, height: float = 6.0, textured: bool = False,
            tile_quads: int = 32, mirror_spheres: int = 0) -> HostScene:
    """Heightfield of quads x quads cells (2 triangles each), finite-difference normals, 1 distant light (0.52 deg)."""
    n = quads
    hgt = (_value_noise(n, n, seed) * height).astype(np.float32)
    xs = np.linspace(-extent / 2, extent / 2, n + 1, dtype=np.float32)
    zs = np.linspace(-extent / 2, extent / 2, n + 1, dtype=np.float32)
    gx, gz = np.meshgrid(xs, zs)
    vertex = np.stack([gx, hgt, gz], axis=-1).reshape(-1, 3).astype(np.float32)
    step = extent / n
    dhdx = np.gradient(hgt.astype(np.float64), step, axis=1)
    dhdz = np.gradient(hgt.astype(np.float64), step, axis=0)
    vn = np.stack([-dhdx, np.ones_like(dhdx), -dhdz], axis=-1)
    vn /= np.linalg.norm(vn, axis=-1, keepdims=True)
    vnorm = vn.reshape(-1, 3).astype(np.float32)
    iz, ix = np.mgrid[0:n, 0:n]
    v00 = (iz * (n + 1) + ix).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    tri_idx = np.empty((2 * n * n, 3), np.int32)
    tri_idx[0::2] = np.stack([v00, v01, v11], axis=1)   # counter-clockwise seen from +y
    tri_idx[1::2] = np.stack([v00, v11, v10], axis=1)
    # UVs tiled per tile_quads x tile_quads quads
    u = (ix.reshape(-1) % tile_quads) / tile_quads
    v = (iz.reshape(-1) % tile_quads) / tile_quads
    du = 1.0 / tile_quads
    uv = np.empty((2 * n * n, 3, 2), np.float32)
    uv[0::2, 0] = np.stack([u, v], 1)
    uv[0::2, 1] = np.stack([u, v + du], 1)
    uv[0::2, 2] = np.stack([u + du, v + du], 1)
    uv[1::2, 0] = np.stack([u, v], 1)
    uv[1::2, 1] = np.stack([u + du, v + du], 1)
    uv[1::2, 2] = np.stack([u + du, v], 1)
    atlas = MaterialAtlas()
    n_tri = tri_idx.shape[0]
    if textured:
        mids = []
        for k, (w, h) in enumerate(RES_TEX_SIZES):
            img = _procedural_texture(w, h, k)
            bump = img if k == 3 else None   # one image doubles as BUMP
            mids.append(atlas.add(color=img, reflection=(0, 0, 0), transparency=(0, 0, 0), bump=bump, luminance=(0, 0, 0)))
        tile = (iz.reshape(-1) // tile_quads) * ((n + tile_quads - 1) // tile_quads) + (ix.reshape(-1) // tile_quads)
        quad_mat = np.array(mids, np.int32)[(tile * 7 + 3) % len(mids)]
        mats = np.repeat(quad_mat, 2).astype(np.int32)
    else:
        m0 = atlas.add(color=(120, 160, 90), reflection=(0, 0, 0), transparency=(0, 0, 0), luminance=(0, 0, 0))
        mats = np.full(n_tri, m0, np.int32)
    tri_normal = vnorm[tri_idx]
    if mirror_spheres:
        sv, st = _uv_sphere(16, 32)
        m_mir = atlas.add(color=(230, 230, 235), reflection=(128, 128, 128), transparency=(0, 0, 0), luminance=(0, 0, 0))
        vs, ts, ns, ms, us = [vertex], [tri_idx], [tri_normal], [mats], [uv]
        base = vertex.shape[0]
        for k in range(mirror_spheres):
            ang = 2 * math.pi * k / mirror_spheres
            cx, cz = 0.22 * extent * math.cos(ang), 0.22 * extent * math.sin(ang)
            r = 0.035 * extent
            cy = float(hgt.max()) + 1.5 * r
            vs.append((sv * r + np.array([cx, cy, cz])).astype(np.float32))
            ts.append(st + base)
            ns.append(sv.astype(np.float32)[st])
            ms.append(np.full(st.shape[0], m_mir, np.int32))
            us.append(np.tile(np.array([[0, 0], [1, 0], [0, 1]], np.float32), (st.shape[0], 1, 1)))
            base += sv.shape[0]
        vertex, tri_idx, tri_normal, mats, uv = np.concatenate(vs), np.concatenate(ts), np.concatenate(ns), np.concatenate(ms), np.concatenate(us)
    lights = [dict(type=LIGHT_DISTANT, dir=(-0.45, -0.8, 0.35), colour=(1, 1, 1), radius=0.52)]
    top = float(hgt.max())
    return _scene(f"terrain{quads}", vertex, tri_idx, mats, uv, tri_normal, atlas, lights,
                  meta=dict(camera=dict(eye=(0.0, top + 0.55 * extent, -0.95 * extent), look_at=(0, 0.3 * height, 0), up=(0, 1, 0), fov=0.75),
                            top=top, extent=extent))


def sweep_cameras(scene: HostScene, frames: int = 64):
    """Config 5: eye positions on a circle around the terrain, all geometry in front of every camera."""
    ext, top = scene.meta["extent"], scene.meta["top"]
    r = 0.95 * ext
    out = []
    for k in range(frames):
        a = 2 * math.pi * k / frames
        out.append(dict(eye=(r * math.sin(a), top + 0.55 * ext, -r * math.cos(a)), look_at=(0, 0.3 * top, 0), up=(0, 1, 0), fov=0.75))
    return out


# ---- the five BASELINE.json configs ----------------------------------------------------------------------------------------------
CONFIGS = {
    1: dict(name="soup-1k-512x512", make=lambda: soup(1000), width=512, height=512, samples=1),
    2: dict(name="spheres-101k-1920x1080", make=lambda: sphere_grid(6, 27, 54), width=1920, height=1080, samples=1),
    3: dict(name="terrain-1M-3840x2160", make=lambda: terrain(708), width=3840, height=2160, samples=1),
    4: dict(name="terrain-10M-textured-7680x4320", make=lambda: terrain(2237, textured=True), width=7680, height=4320, samples=1),
    5: dict(name="terrain-1M-mirrors-sweep64-1920x1080", make=lambda: terrain(708, mirror_spheres=6), width=1920, height=1080,
            samples=1, frames=64),
}


def id_material_variant(scene: HostScene) -> HostScene:
    """ID-material scene: material i for triangle i, LUMINANCE 1x1 = (i+1) as 24-bit little-endian RGB, no
    lights -> the UNMODIFIED kernel's RGB output encodes the primary-hit triangle id.  N < 2^24."""
    n = scene.triangle_count
    assert n < (1 << 24) - 1
    size = np.zeros((n * MATERIAL_CHANNEL_COUNT, 2), np.uint32)
    size[CH_LUMINANCE::MATERIAL_CHANNEL_COUNT] = 1
    start = np.zeros(n * MATERIAL_CHANNEL_COUNT + 1, np.int32)
    start[CH_LUMINANCE:-1:MATERIAL_CHANNEL_COUNT] = np.arange(n)
    start[-1] = n
    code = np.arange(1, n + 1, dtype=np.uint32)
    tex = np.zeros((n, 4), np.uint8)
    tex[:, 0] = code & 0xFF
    tex[:, 1] = (code >> 8) & 0xFF
    tex[:, 2] = (code >> 16) & 0xFF
    z = np.zeros
    return HostScene(vertex=scene.vertex, tri_idx=scene.tri_idx, tri_mat=np.arange(n, dtype=np.int32), tri_uv=scene.tri_uv,
                     tri_normal=scene.tri_normal, mat_size=size, mat_start=start, textures=tex, light_type=z(0, np.int32),
                     light_pos=z((0, 4), np.float32), light_dir=z((0, 4), np.float32), light_colour=z((0, 4), np.float32),
                     light_radius=z(0, np.float32), light_half=z(0, np.float32), axes_div=scene.axes_div, box_min=scene.box_min,
                     grid_start=scene.grid_start, grid_list=scene.grid_list, name=scene.name + "-idmat", meta=scene.meta).normalise()


def decode_id_planes(r: np.ndarray, g: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Inverse of the ID-material encoding: out = (int)(byte/255*65535) per channel -> byte = round(out*255/65535)."""
    dec = lambda p: np.rint(p.astype(np.float64) * 255.0 / 65535.0).astype(np.uint32)
    code = dec(r) | (dec(g) << 8) | (dec(b) << 16)
    return np.where(code == 0, np.uint32(0xFFFFFFFF), code - 1).astype(np.uint32)
