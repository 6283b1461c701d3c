"""Multi-GPU host layer (one process per GPU, torch.distributed over NCCL/NVLink for the plumbing).

The path shards by pixels: the image is cut into row bands dealt round-robin to the ranks, the scene
is replicated per GPU, and the only exchange step is assembling the three 16-bit planes -- one NCCL all-gather of each
rank's compact rows.  No collective sits on the trace path itself."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


def owned_rows(height: int, rank: int, world: int, band_rows: int) -> np.ndarray:
    """Frame rows y with (y // band_rows) % world == rank, ascending (mirrors map_row() in csrc/rt_types.h)."""
    y = np.arange(height)
    return y[(y // band_rows) % world == rank]


class BandPartition:
    def __init__(self, height: int, width: int, rank: int, world: int, band_rows: int = 16):
        self.height, self.width, self.rank, self.world, self.band_rows = height, width, rank, world, band_rows
        self.rows = owned_rows(height, rank, world, band_rows)
        self.owned_rows = int(self.rows.size)
        self.max_owned = max(int(owned_rows(height, r, world, band_rows).size) for r in range(world))

    def render(self, frame: api.DeviceFrame, samples: int, variant: int, stream: int) -> int:
        """Async launch on `stream`; returns the number of kernels launched."""
        if self.owned_rows == 0:
            return 0
        frame.render_bands(samples, self.band_rows, self.rank, self.world, variant=variant, stream=stream, sync=False)
        return frame.last_launches

    def render_timed(self, frame, samples, variant) -> float:
        return frame.render_bands(samples, self.band_rows, self.rank, self.world, variant=variant)[0]

    def render_counted(self, frame, samples, variant):
        return frame.render_bands(samples, self.band_rows, self.rank, self.world, variant=variant, count=True)


def torch_u8(t):
    return t.uint8    # byte view: every backend (NCCL, gloo) moves uint8


class _DevPtr:
    """Exposes a raw device address to torch through __cuda_array_interface__ (int16 view of the ushort planes)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i2", "data": (int(ptr), False), "version": 3}


class PlaneGather:
    """Assembles the full frame on every rank: compact owned rows -> all_gather_into_tensor -> scatter to frame rows."""

    def __init__(self, frame, part: BandPartition, device, planes=None):
        import torch
        self.torch = torch
        self.part = part
        h, w = part.height, part.width
        if planes is None:
            r, g, b = frame.device_planes()
            assert g == r + 2 * h * w and b == g + 2 * h * w
            planes = torch.as_tensor(_DevPtr(r, (3, h, w)), device=device)          # the frame's own planes, no copy
        self.planes = planes
        self.mine = torch.as_tensor(part.rows, device=device, dtype=torch.long)
        self.send = torch.zeros((3, part.max_owned, w), dtype=torch.int16, device=device)
        self.recv = torch.empty((part.world, 3, part.max_owned, w), dtype=torch.int16, device=device)
        self.full = torch.zeros((3, h, w), dtype=torch.int16, device=device)
        self.all_rows = [torch.as_tensor(owned_rows(h, r_, part.world, part.band_rows), device=device, dtype=torch.long)
                         for r_ in range(part.world)]
        self.launches = 0

    def run(self):
        import torch.distributed as dist
        t = self.torch
        n = self.part.owned_rows
        if n:
            t.index_select(self.planes, 1, self.mine, out=self.send[:, :n, :])
        dist.all_gather_into_tensor(self.recv.view(torch_u8(t)).view(-1), self.send.view(torch_u8(t)).view(-1))                           # NCCL over NVLink: the only exchange step
        for r_, rows in enumerate(self.all_rows):
            if rows.numel():
                self.full.index_copy_(1, rows, self.recv[r_, :, :rows.numel(), :])
        return self.full


class PlanePush:
    """Assembles the full frame on every rank WITHOUT a collective call: each rank stores its finished rows straight into the
    full-frame planes of all GPUs of the box over NVLink peer memory (one kernel, oclr_frame_push_rows) and one cross-GPU barrier
    on the stream orders the readers behind the writers.  The destination planes live in torch symmetric memory (peer-mapped on
    every rank), double-buffered: while a rank still reads frame k from one buffer, its peers may already push frame k + 1 into the
    other; the barrier of frame k + 1 is behind every read of frame k that was enqueued on the same stream."""

    def __init__(self, frame, part: BandPartition, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.frame, self.part = torch, frame, part
        h, w = part.height, part.width
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty((3, h, w), dtype=torch.int16, device=device)
            self.handles.append(symm_mem.rendezvous(t, dist.group.WORLD))
            self.bufs.append(t)
        for hdl in self.handles:
            if len(hdl.buffer_ptrs) != part.world:
                raise RuntimeError("symmetric memory rendezvous returned the wrong number of peers")
        self.step = 0
        self.launches = 2       # push kernel + barrier kernel

    def run(self):
        i = self.step & 1
        self.step += 1
        hdl = self.handles[i]
        self.frame.push_rows(self.part.band_rows, self.part.rank, self.part.world, hdl.buffer_ptrs,
                             stream=self.torch.cuda.current_stream().cuda_stream)
        hdl.barrier(channel=0)
        return self.bufs[i]


def plane_exchange(frame, part: BandPartition, device):
    """The frame-assembly step for world > 1: peer stores over NVLink (PlanePush); the NCCL all-gather (PlaneGather) only where
    peer-mapped memory cannot be set up (said on stderr).  Returns (object with .run() / .launches, name)."""
    import sys
    try:
        return PlanePush(frame, part, device), "peer stores over NVLink (oclr_frame_push_rows) + symmetric-memory barrier"
    except Exception as e:      # no P2P mapping between these devices / symmetric memory unavailable in this torch build
        print(f"[opencl_render_b200] peer-memory frame assembly unavailable ({type(e).__name__}: {e}); using the NCCL all-gather",
              file=sys.stderr, flush=True)
        return PlaneGather(frame, part, device), "NCCL all-gather of compact rows + scatter"


_SCENE_ARRAYS = ("vertex", "tri_idx", "tri_mat", "tri_uv", "tri_normal", "mat_size", "mat_start", "textures", "light_type", "light_pos",
                 "light_dir", "light_colour", "light_radius", "light_half", "box_min", "grid_start", "grid_list")


class EndToEnd:
    """The drop-in call with HOST buffers, made by ONE process: RaytraceAll on one GPU (world == 1) or on the first `world` GPUs of
    the box (computation type deviceCount + 1 with the "devices" option: every GPU pulls 1/N of the scene over PCIe, fans it out over
    NVLink, traces its row bands and copies them back).  Upload + repack + trace + read back are all inside the call."""

    def __init__(self, scene: api.HostScene, cam: api.CameraSetup, lists: api.CameraLists, world: int, device: int, n_devices: int = 1,
                 pinned: bool = True):
        import copy
        self.cam, self.world, self.device, self.n_devices = cam, world, device, n_devices
        self._keep = []
        self.scene = copy.copy(scene)
        conv = self._pin if pinned else (lambda a: np.array(a, copy=True))
        for name in _SCENE_ARRAYS:
            setattr(self.scene, name, conv(getattr(scene, name)))
        self.lists = api.CameraLists(conv(lists.start), conv(lists.end), conv(lists.list))
        self.out = tuple(conv(np.zeros((cam.height, cam.width), np.uint16)) for _ in range(3))
        self.h2d_bytes = int(sum(getattr(self.scene, n).nbytes for n in _SCENE_ARRAYS) + self.lists.start.nbytes + self.lists.end.nbytes +
                             self.lists.list.nbytes)
        self.d2h_bytes = int(6 * cam.height * cam.width)
        if world > 1:
            api.set_option("devices", world)
            self.computation_type = n_devices + 1
            self.call = f"RaytraceAll(all devices) on {world} GPUs from one process (C-ABI, host buffers; shared upload over NVLink)"
        else:
            self.computation_type = 1 + device
            self.call = "RaytraceAll (C-ABI, host buffers)"
        self._source = (scene, lists)

    def _pin(self, a):
        import torch
        a = np.ascontiguousarray(a)
        if a.size == 0:
            return a
        t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        self._keep.append(t)
        return v

    def pageable_copy(self) -> "EndToEnd":
        """The same call fed from plain (pageable) arrays, as the plugin allocates them (render.cpp:1086-1134)."""
        return EndToEnd(self._source[0], self.cam, self._source[1], self.world, self.device, self.n_devices, pinned=False)

    def step(self, samples: int):
        api.raytrace_all(self.computation_type, self.cam, self.lists, samples, self.scene, out=self.out)
