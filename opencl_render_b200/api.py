"""Host-side mirror of the reference's render interface (source/opencl/raytrace.h:46-106, called from
source/render.cpp:1311-1352), on top of the C-ABI of libopencl_render_b200.so.

Names and argument meaning follow the reference: `raytrace_all` is `RaytraceAll`, `set_camera` is `SetCamera`
(render.cpp:461-491), `camera_triangle_list` / `scene_triangle_list` are `CameraTriangleList::New` /
`SceneTriangleList::New` (source/util/trianglelist.cpp:520-626, 655-737).  Arrays use the reference's element layouts
(cl_float3 = 4 floats, cl_int3 = 4 ints, cl_uchar3 = 4 bytes).  Every compute call goes to the CUDA library; there is
no Python implementation behind any of them.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

AXES_DIVISION = 256          # SceneTriangleList::AXES_DIVISION, trianglelist.h:44
MATERIAL_CHANNEL_COUNT = 5   # raytrace_opencl.h:14-22
CH_COLOR, CH_REFLECTION, CH_TRANSPARENCY, CH_BUMP, CH_LUMINANCE = range(5)
LIGHT_OMNI, LIGHT_SPOT, LIGHT_SPOTRECT, LIGHT_DISTANT, LIGHT_PARALLEL, LIGHT_PARSPOT, LIGHT_PARSPOTRECT, LIGHT_TUBE, \
    LIGHT_AREA, LIGHT_PHOTOMETRIC = range(10)
KERNEL_SIMPLE, KERNEL_PIPE, KERNEL_DEFAULT = 0, 2, -1
ACCUMULATE_REFERENCE_16BIT, ACCUMULATE_FLOAT = 0, 1
NO_TRIANGLE = 0xFFFFFFFF


class OclrError(RuntimeError):
    pass


def _c(a, dtype, shape_tail=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape_tail is not None and tuple(a.shape[1:]) != tuple(shape_tail):
        raise ValueError(f"expected trailing shape {shape_tail}, got {a.shape}")
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


@dataclass
class HostScene:
    """The scene arrays the plugin hands to RaytraceAll (everything except camera, camera lists and outputs)."""
    vertex: np.ndarray               # float32 [V,4]
    tri_idx: np.ndarray              # int32   [N,4]
    tri_mat: np.ndarray              # int32   [N]
    tri_uv: np.ndarray               # float32 [N,3,2]
    tri_normal: np.ndarray           # float32 [N,3,4]
    mat_size: np.ndarray             # uint32  [5*M,2]
    mat_start: np.ndarray            # int32   [5*M+1]
    textures: np.ndarray             # uint8   [T,4]
    light_type: np.ndarray           # int32   [L]
    light_pos: np.ndarray            # float32 [L,4]
    light_dir: np.ndarray            # float32 [L,4]
    light_colour: np.ndarray         # float32 [L,4]
    light_radius: np.ndarray         # float32 [L]
    light_half: np.ndarray           # float32 [L]
    axes_div: int = AXES_DIVISION
    box_min: np.ndarray | None = None     # float32 [axes_div+1,4]
    grid_start: np.ndarray | None = None  # uint32 [axes_div^3+1]
    grid_list: np.ndarray | None = None   # uint32
    name: str = ""
    meta: dict = field(default_factory=dict)

    def normalise(self) -> "HostScene":
        self.vertex = _c(self.vertex, np.float32, (4,))
        self.tri_idx = _c(self.tri_idx, np.int32, (4,))
        self.tri_mat = _c(self.tri_mat, np.int32)
        self.tri_uv = _c(self.tri_uv, np.float32, (3, 2))
        self.tri_normal = _c(self.tri_normal, np.float32, (3, 4))
        self.mat_size = _c(self.mat_size, np.uint32, (2,))
        self.mat_start = _c(self.mat_start, np.int32)
        self.textures = _c(self.textures, np.uint8, (4,))
        self.light_type = _c(self.light_type, np.int32)
        self.light_pos = _c(self.light_pos, np.float32, (4,))
        self.light_dir = _c(self.light_dir, np.float32, (4,))
        self.light_colour = _c(self.light_colour, np.float32, (4,))
        self.light_radius = _c(self.light_radius, np.float32)
        self.light_half = _c(self.light_half, np.float32)
        if self.box_min is not None:
            self.box_min = _c(self.box_min, np.float32, (4,))
            self.grid_start = _c(self.grid_start, np.uint32)
            self.grid_list = _c(self.grid_list, np.uint32)
        return self

    @property
    def triangle_count(self) -> int:
        return int(self.tri_idx.shape[0])

    @property
    def vertex_count(self) -> int:
        return int(self.vertex.shape[0])

    @property
    def material_count(self) -> int:
        return int(self.mat_size.shape[0] // MATERIAL_CHANNEL_COUNT)

    @property
    def light_count(self) -> int:
        return int(self.light_type.shape[0])

    def desc(self, allow_missing_grid: bool = False) -> _lib.SceneDesc:
        if self.box_min is None and not allow_missing_grid:
            raise OclrError("scene has no grid: call scene_triangle_list(scene) first")
        d = _lib.SceneDesc()
        d.vertexCount = self.vertex_count
        d.vertex = _ptr(self.vertex)
        d.triangleCount = self.triangle_count
        d.triangleVertexIndex = _ptr(self.tri_idx)
        d.triangleMaterialId = _ptr(self.tri_mat)
        d.triangleUv = _ptr(self.tri_uv)
        d.triangleNormal = _ptr(self.tri_normal)
        d.axesDivCount = self.axes_div
        if self.box_min is not None:
            d.sceneBoxMin = _ptr(self.box_min)
            d.scenePixelTriangleListStart = _ptr(self.grid_start)
            d.scenePixelTriangleList = _ptr(self.grid_list)
        d.materialCount = self.material_count
        d.materialImageSize = _ptr(self.mat_size)
        d.materialImageStart = _ptr(self.mat_start)
        d.texturesSize = int(self.textures.shape[0])
        d.textures = _ptr(self.textures)
        d.lightCount = self.light_count
        d.lightType = _ptr(self.light_type)
        d.lightPosition = _ptr(self.light_pos)
        d.lightDirection = _ptr(self.light_dir)
        d.lightColour = _ptr(self.light_colour)
        d.lightRadius = _ptr(self.light_radius)
        d.lightHalfAttenuationDistance = _ptr(self.light_half)
        return d


@dataclass
class CameraSetup:
    """Output of SetCamera (render.cpp:461-491) + the image size."""
    width: int
    height: int
    eye: np.ndarray
    eye_to_top_left: np.ndarray
    left_to_right: np.ndarray
    top_to_bottom: np.ndarray
    pixel_size_inv: float

    def c(self) -> _lib.Camera:
        k = _lib.Camera()
        k.width, k.height = self.width, self.height
        for i in range(3):
            k.eye[i] = float(self.eye[i])
            k.eyeToTopLeft[i] = float(self.eye_to_top_left[i])
            k.leftToRight[i] = float(self.left_to_right[i])
            k.topToBottom[i] = float(self.top_to_bottom[i])
        k.pixelSizeInv = float(self.pixel_size_inv)
        return k


@dataclass
class CameraLists:
    """CameraTriangleList: per-pixel candidate triangles (Start/End have W*H entries, no +1)."""
    start: np.ndarray
    end: np.ndarray
    list: np.ndarray


def set_camera(position, look_at, up, fov: float, width: int, height: int) -> CameraSetup:
    """SetCamera (render.cpp:461-491): fov is the horizontal field of view in radians."""
    lib = _lib.load()
    out = _lib.Camera()
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v[:3]])
    lib.oclr_set_camera(C.byref(out), f3(position), f3(look_at), f3(up), C.c_float(fov), width, height)
    g = lambda a: np.array([a[0], a[1], a[2], 0.0], dtype=np.float32)
    return CameraSetup(width, height, g(out.eye), g(out.eyeToTopLeft), g(out.leftToRight), g(out.topToBottom),
                       float(np.float32(out.pixelSizeInv)))


def camera_triangle_list(camera: CameraSetup, scene: HostScene) -> CameraLists:
    """CameraTriangleList::New (trianglelist.cpp:520-626)."""
    lib = _lib.load()
    scene.normalise()
    out = _lib.CameraLists()
    cam = camera.c()
    if not lib.oclr_build_camera_lists(C.byref(cam), scene.vertex_count, _ptr(scene.vertex), scene.triangle_count,
                                       _ptr(scene.tri_idx), C.byref(out)):
        raise OclrError(_lib.last_error())
    try:
        p = int(out.pixelCount)
        start = np.ctypeslib.as_array(out.start, shape=(p,)).copy()
        end = np.ctypeslib.as_array(out.end, shape=(p,)).copy()
        lst = np.ctypeslib.as_array(out.list, shape=(max(int(out.listSize), 1),))[:int(out.listSize)].copy()
    finally:
        lib.oclr_free_camera_lists(C.byref(out))
    return CameraLists(start, end, lst)


def scene_triangle_list(scene: HostScene, axes_div: int = AXES_DIVISION, device: int | None = None) -> HostScene:
    """SceneTriangleList::New (trianglelist.cpp:655-737): fills scene.box_min / grid_start / grid_list.  `device` = a CUDA device
    index: the builder runs on the GPU (same output, entry for entry); None: the host builder."""
    lib = _lib.load()
    scene.normalise()
    out = _lib.SceneGrid()
    if device is None:
        ok = lib.oclr_build_scene_grid(axes_div, scene.vertex_count, _ptr(scene.vertex), scene.triangle_count, _ptr(scene.tri_idx), C.byref(out))
    else:
        ok = lib.oclr_build_scene_grid_device(device, axes_div, scene.vertex_count, _ptr(scene.vertex), scene.triangle_count,
                                              _ptr(scene.tri_idx), C.byref(out))
    if not ok:
        raise OclrError(_lib.last_error())
    try:
        n = int(out.axesDivCount)
        scene.axes_div = n
        scene.box_min = np.ctypeslib.as_array(C.cast(out.boxMin, C.POINTER(C.c_float)), shape=(n + 1, 4)).copy()
        scene.grid_start = np.ctypeslib.as_array(out.start, shape=(n ** 3 + 1,)).copy()
        scene.grid_list = np.ctypeslib.as_array(out.list, shape=(max(int(out.listSize), 1),))[:int(out.listSize)].copy()
    finally:
        lib.oclr_free_scene_grid(C.byref(out))
    return scene


def computation_types() -> list[str]:
    """InitOpenCL + GetComputationTypeCount/Name (raytrace.c:78-153): index 0 is a label only."""
    lib = _lib.load()
    lib.InitOpenCL()
    names = []
    for i in range(lib.GetComputationTypeCount()):
        buf = C.create_string_buffer(256)
        names.append(buf.value.decode() if lib.GetComputationTypeName(i, 255, buf) else "")
    return names


def raytrace_all(computation_type: int, camera: CameraSetup, lists: CameraLists, sample_count: int, scene: HostScene,
                 out=None):
    """RaytraceAll (raytrace.h:58-106) with HOST arrays in and out: upload, repack, trace, read back.
    Returns (red, green, blue) uint16 [H,W].  Raises OclrError where the reference returns CL_FALSE."""
    lib = _lib.load()
    scene.normalise()
    w, h = camera.width, camera.height
    if out is None:
        out = tuple(np.empty((h, w), dtype=np.uint16) for _ in range(3))
    r, g, b = out
    dim = (C.c_uint32 * 2)(w, h)
    f = lambda v: np.ascontiguousarray(v, dtype=np.float32).ctypes.data_as(_lib.c_float_p)
    start = _c(lists.start, np.uint32)
    end = _c(lists.end, np.uint32)
    lst = _c(lists.list, np.uint32)
    d = scene.desc()
    ok = lib.oclr_raytrace_all_p(computation_type, dim, f(camera.eye), f(camera.eye_to_top_left), f(camera.left_to_right),
                                 f(camera.top_to_bottom), C.c_float(camera.pixel_size_inv), C.byref(d), _ptr(start), _ptr(end),
                                 _ptr(lst), int(lst.size), sample_count, _ptr(r), _ptr(g), _ptr(b))
    if not ok:
        raise OclrError(_lib.last_error())
    return r, g, b


class DeviceScene:
    """Resident scene (extension): uploaded and repacked once, rendered many times."""

    def __init__(self, scene: HostScene, device: int = 0, axes_div: int | None = None):
        """A scene without a grid (scene_triangle_list not called) is accepted: SceneTriangleList::New then runs on the device
        during the upload and the grid never exists on the host (`axes_div` = its axesDivCount, default 256)."""
        self._lib = _lib.load()
        scene.normalise()
        if scene.box_min is None:
            scene.axes_div = axes_div or AXES_DIVISION
        d = scene.desc(allow_missing_grid=True)
        self.handle = self._lib.oclr_scene_create(device, C.byref(d))
        if not self.handle:
            raise OclrError(_lib.last_error())
        self.device = device
        self.host = scene

    @property
    def device_bytes(self) -> int:
        return int(self._lib.oclr_scene_device_bytes(self.handle))

    def debug_read(self, which: int) -> np.ndarray:
        """Packed device array as bytes (0 triGeo, 1 triShade, 2 bricks, 3 cellRange, 4 planes, 5 cellList)."""
        n = int(self._lib.oclr_scene_debug_read(self.handle, which, None, 0))
        out = np.zeros(max(n, 1), np.uint8)
        self._lib.oclr_scene_debug_read(self.handle, which, _ptr(out), n)
        return out[:n]

    def close(self):
        if getattr(self, "handle", None):
            self._lib.oclr_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()


class DeviceFrame:
    """One camera on a resident scene: camera lists + three 16-bit planes in HBM."""

    def __init__(self, scene: DeviceScene, camera: CameraSetup, lists: CameraLists | None = None):
        """`lists` = CameraTriangleList::New output built on the host, or None: the lists are built on the device from the
        resident scene (entry-for-entry the same; see camera_lists())."""
        self._lib = _lib.load()
        self.scene = scene
        self.camera = camera
        cam = camera.c()
        if lists is None:
            self.handle = self._lib.oclr_frame_create_device_lists(scene.handle, C.byref(cam))
        else:
            start = _c(lists.start, np.uint32)
            end = _c(lists.end, np.uint32)
            lst = _c(lists.list, np.uint32)
            self.handle = self._lib.oclr_frame_create(scene.handle, C.byref(cam), _ptr(start), _ptr(end), _ptr(lst), int(lst.size))
        if not self.handle:
            raise OclrError(_lib.last_error())

    @property
    def state_bytes(self) -> int:
        """Wavefront path state held in HBM (0 before the first render)."""
        return int(self._lib.oclr_frame_state_bytes(self.handle))

    def camera_lists(self) -> "CameraLists":
        """Copies the frame's camera lists (uploaded or device-built) back to the host."""
        P = self.camera.width * self.camera.height
        n = int(self._lib.oclr_frame_camera_list_size(self.handle))
        start, end, lst = np.empty(P, np.uint32), np.empty(P, np.uint32), np.empty(max(n, 1), np.uint32)
        if not self._lib.oclr_frame_read_camera_lists(self.handle, _ptr(start), _ptr(end), _ptr(lst)):
            raise OclrError(_lib.last_error())
        return CameraLists(start, end, lst[:n])

    def render(self, sample_count: int = 1, rows=None, variant: int = KERNEL_DEFAULT, count: bool = False, stream: int = 0,
               sync: bool = True, samples=None):
        """Trace rows [rows[0], rows[1]) (default: all).  `samples` = (begin, end): only that range of the sample_count-sample
        job (progressive rendering: sample 0 overwrites the planes, later samples add).  Returns (device_ms, launches,
        counters dict | None)."""
        r0, r1 = rows if rows is not None else (0, self.camera.height)
        s0, s1 = samples if samples is not None else (0, sample_count)
        stats = _lib.RenderStats()
        ok = self._lib.oclr_frame_render_samples(self.handle, sample_count, s0, s1, r0, r1, variant, 1 if count else 0,
                                                 C.c_void_p(stream), C.byref(stats) if (sync or count) else None)
        if not ok:
            raise OclrError(_lib.last_error())
        self.last_trace_ms, self.last_trace_launches = float(stats.traceMs), int(stats.traceLaunches)
        return float(stats.deviceMs), int(stats.launches), (stats.counters.as_dict() if count else None)

    def render_bands(self, sample_count: int, band_rows: int, rank: int, world: int, variant: int = KERNEL_DEFAULT,
                     count: bool = False, stream: int = 0, sync: bool = True, samples=None):
        """Trace the rows y with (y // band_rows) % world == rank in one launch."""
        s0, s1 = samples if samples is not None else (0, sample_count)
        stats = _lib.RenderStats()
        ok = self._lib.oclr_frame_render_bands_samples(self.handle, sample_count, s0, s1, band_rows, rank, world, variant,
                                                       1 if count else 0, C.c_void_p(stream),
                                                       C.byref(stats) if (sync or count) else None)
        if not ok:
            raise OclrError(_lib.last_error())
        self.last_trace_ms, self.last_trace_launches = float(stats.traceMs), int(stats.traceLaunches)
        return float(stats.deviceMs), int(stats.launches), (stats.counters.as_dict() if count else None)

    def read(self, rows=None, out=None, stream: int = 0):
        h, w = self.camera.height, self.camera.width
        r0, r1 = rows if rows is not None else (0, h)
        if out is None:
            out = tuple(np.zeros((h, w), dtype=np.uint16) for _ in range(3))
        if not self._lib.oclr_frame_read(self.handle, r0, r1, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), C.c_void_p(stream)):
            raise OclrError(_lib.last_error())
        return out

    def write(self, planes, rows=None, stream: int = 0):
        """Uploads full-frame uint16 planes (rows [rows[0], rows[1]) of them): restores a checkpoint taken with read()."""
        r0, r1 = rows if rows is not None else (0, self.camera.height)
        p = [_c(a, np.uint16) for a in planes]
        if not self._lib.oclr_frame_write(self.handle, r0, r1, _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), C.c_void_p(stream)):
            raise OclrError(_lib.last_error())

    def set_accumulation(self, mode: int):
        """ACCUMULATE_REFERENCE_16BIT (default, raytrace_opencl.c:726-741) or ACCUMULATE_FLOAT (fp32 sums, no per-sample truncation)."""
        if not self._lib.oclr_frame_set_accumulation(self.handle, mode):
            raise OclrError(_lib.last_error())

    def read_accum(self) -> np.ndarray:
        """float32 [H,W,4]: (sum r, sum g, sum b, samples) of the float accumulator."""
        out = np.empty((self.camera.height, self.camera.width, 4), np.float32)
        if not self._lib.oclr_frame_read_accum(self.handle, _ptr(out)):
            raise OclrError(_lib.last_error())
        return out

    def write_accum(self, acc: np.ndarray):
        a = _c(acc, np.float32)
        if not self._lib.oclr_frame_write_accum(self.handle, _ptr(a)):
            raise OclrError(_lib.last_error())

    def progress(self):
        """(done, total) pixel-samples of the render call in flight (thread-safe)."""
        d, t = C.c_ulonglong(), C.c_ulonglong()
        if not self._lib.oclr_frame_progress(self.handle, C.byref(d), C.byref(t)):
            raise OclrError(_lib.last_error())
        return int(d.value), int(t.value)

    def primary_ids(self) -> np.ndarray:
        ids = np.empty((self.camera.height, self.camera.width), dtype=np.uint32)
        if not self._lib.oclr_frame_read_primary_ids(self.handle, _ptr(ids)):
            raise OclrError(_lib.last_error())
        return ids

    def undefined_flags(self) -> np.ndarray:
        """uint8 [H,W]: 1 where the reference's own result is undefined (uninitialised read, see oclr_abi.h)."""
        flags = np.empty((self.camera.height, self.camera.width), dtype=np.uint8)
        if not self._lib.oclr_frame_read_flags(self.handle, _ptr(flags)):
            raise OclrError(_lib.last_error())
        return flags

    @property
    def last_launches(self) -> int:
        return int(self._lib.oclr_frame_last_launches(self.handle))

    def push_rows(self, band_rows: int, rank: int, world: int, peer_planes, stream: int = 0):
        """Stores this rank's rows into the full-frame planes of every GPU in `peer_planes` (device addresses, one per rank)."""
        arr = (C.c_void_p * world)(*[int(p) for p in peer_planes])
        if not self._lib.oclr_frame_push_rows(self.handle, band_rows, rank, world, arr, C.c_void_p(stream)):
            raise OclrError(_lib.last_error())

    def device_planes(self):
        r, g, b = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._lib.oclr_frame_device_planes(self.handle, C.byref(r), C.byref(g), C.byref(b))
        return r.value, g.value, b.value

    def close(self):
        if getattr(self, "handle", None):
            self._lib.oclr_frame_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()


def write_image(path: str, planes, fmt: str | None = None, bmp_reference_cast: bool = False):
    """Writes uint16 [H,W] planes (r, g, b): .bmp (24 bit, value/256 -- or the reference's low-byte cast, writebmp.cpp:136-141),
    .ppm (16 bit) or .png (16 bit)."""
    lib = _lib.load()
    r, g, b = (_c(a, np.uint16) for a in planes)
    h, w = r.shape
    fmt = (fmt or str(path).rsplit(".", 1)[-1]).lower()
    p = str(path).encode()
    if fmt == "bmp":
        ok = lib.oclr_write_bmp(p, w, h, _ptr(r), _ptr(g), _ptr(b), 1 if bmp_reference_cast else 0)
    elif fmt == "ppm":
        ok = lib.oclr_write_ppm16(p, w, h, _ptr(r), _ptr(g), _ptr(b))
    elif fmt == "png":
        ok = lib.oclr_write_png16(p, w, h, _ptr(r), _ptr(g), _ptr(b))
    else:
        raise ValueError(f"unknown image format {fmt!r}")
    if not ok:
        raise OclrError(f"could not write {path}")


def set_option(name: str, value: int):
    if not _lib.load().oclr_set_option(name.encode(), int(value)):
        raise OclrError(_lib.last_error())


def band_partition(height: int, rank: int, world: int, band_rows: int = 128) -> list[tuple[int, int]]:
    """Rows owned by `rank`: bands of 128 rows (the reference's tile height, raytrace.c:507) dealt round-robin."""
    lib = _lib.load()
    n = lib.oclr_band_partition(height, band_rows, rank, world, None, 0)
    rows = (C.c_uint32 * (2 * max(n, 1)))()
    lib.oclr_band_partition(height, band_rows, rank, world, rows, n)
    return [(int(rows[2 * i]), int(rows[2 * i + 1])) for i in range(n)]
