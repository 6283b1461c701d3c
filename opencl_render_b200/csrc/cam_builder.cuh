// cam_builder.cuh -- CameraTriangleList::New on the device (source/util/trianglelist.cpp:520-626 with GetCameraPosition
// :74-90 and FillRectangle :131-217).  The per-pixel camera lists are per-FRAME inputs of the kernel: built on
// the host they cost 0.15 s (config 2) to 0.45 s (config 3) per frame against a few milliseconds of tracing.
//
// Same fp32 decisions as the host restatement (builders.cpp; compiled -fmad=false), different machinery:
//   cam_project_kernel   one thread per triangle: the three GetCameraPosition projections, the clamped bounding box of
//                        FillRectangle, and the triangle's slot budget = bbox pixels + 1;
//   (cub exclusive scan) slot offsets -- every (triangle, bbox pixel) pair owns one slot of a key pool, so the raster kernels
//                        need no atomics for their output and no second pass;
//   cam_raster_*_kernel  the reference's per-pixel tests (12-term edge-crossing test, else the corner-inside test); a hit writes
//                        key = pixel * N + triangle -- the reference's own sort key -- a miss writes the sentinel P * N.
//                        Small boxes: one thread per triangle; large boxes: one block per triangle;
//   (cub radix sort)     keys ascending = pixels ascending, triangle ids ascending per pixel: the order the reference gets
//                        from its quicksort (keys are unique, so any correct sort gives the same array);
//   cam_*compress*       the reference's storage compression (a pixel whose list equals its left, else its upper, neighbour's
//                        shares that copy), sequential there, here: equality flags -> scan of the kept lengths -> pointer
//                        jumping to the surviving copy.
// Output is entry-for-entry the host builder's (tests/test_gpu_parity.py::test_device_camera_lists_equal_host_builder).
#pragma once
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "rt_core.h"

namespace oclr {

struct CamProjD {
    f3 eye, tl, lr, tb, screenN;
    float tlDotN, psiSq;
    uint32_t W, H;
};

struct P2 {
    float x, y;
};

// float -> cl_uint as the reference's x86-64 build performs it: cvttss2si to 64 bits (out of range / NaN give
// 0x8000000000000000), low half kept.
__device__ __forceinline__ uint32_t to_u32_x86(float f) {
    if (!(f >= -9223372036854775808.f && f < 9223372036854775808.f)) return 0u;
    return (uint32_t)(long long)f;
}
__device__ __forceinline__ uint32_t to_u32_x86(double f) {
    if (!(f >= -9223372036854775808. && f < 9223372036854775808.)) return 0u;
    return (uint32_t)(long long)f;
}

// trianglelist.cpp:74-90
__device__ __forceinline__ P2 cam_project(const CamProjD& c, float vx, float vy, float vz) {
    const f3 e = mk3(vx - c.eye.x, vy - c.eye.y, vz - c.eye.z);
    const float s = c.tlDotN / dot3(e, c.screenN);
    const f3 q = mk3(s * e.x - c.tl.x, s * e.y - c.tl.y, s * e.z - c.tl.z);
    P2 p;
    p.x = dot3(c.lr, q) * c.psiSq;
    p.y = dot3(c.tb, q) * c.psiSq;
    return p;
}

struct TriScreen {   // 48 bytes per triangle
    float ax, ay, bx, by, cx, cy;
    uint32_t x0, y0, x1, y1;   // clamped bounding box (empty when x1 < x0 or y1 < y0)
    uint32_t aPixel;           // pixel of vertex a when it is on screen (:157-160), else 0xFFFFFFFF
    uint32_t pad;
};

__device__ __forceinline__ uint64_t tri_slots(const TriScreen& t) {
    const uint64_t w = t.x0 <= t.x1 ? (uint64_t)(t.x1 - t.x0) + 1u : 0u, h = t.y0 <= t.y1 ? (uint64_t)(t.y1 - t.y0) + 1u : 0u;
    return w * h + 1u;
}

enum { kCamSmallBox = 256 };   // boxes up to this many pixels are rastered by one thread

__global__ void __launch_bounds__(256) cam_project_kernel(CamProjD c, uint32_t triangleCount, const float4* __restrict__ triShade,
                                                          TriScreen* __restrict__ screen, uint64_t* __restrict__ slots,
                                                          uint32_t* __restrict__ largeList, uint32_t* __restrict__ largeCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= triangleCount) return;
    const float4 A = __ldg(triShade + 8 * (size_t)i), B = __ldg(triShade + 8 * (size_t)i + 1), C = __ldg(triShade + 8 * (size_t)i + 2);
    const P2 a = cam_project(c, A.x, A.y, A.z), b = cam_project(c, B.x, B.y, B.z), cc = cam_project(c, C.x, C.y, C.z);
    const float wm = (float)(c.W - 1), hm = (float)(c.H - 1);
    TriScreen t;
    t.ax = a.x; t.ay = a.y; t.bx = b.x; t.by = b.y; t.cx = cc.x; t.cy = cc.y;
    t.x0 = to_u32_x86(fmaxf(0.f, fminf(fminf(a.x, b.x), fminf(cc.x, wm))));
    t.y0 = to_u32_x86(fmaxf(0.f, fminf(fminf(a.y, b.y), fminf(cc.y, hm))));
    t.x1 = to_u32_x86(fminf(wm, fmaxf(fmaxf(a.x, b.x), fmaxf(cc.x, 0.f))));
    t.y1 = to_u32_x86(fminf(hm, fmaxf(fmaxf(a.y, b.y), fmaxf(cc.y, 0.f))));
    const uint32_t ax = to_u32_x86(floor((double)a.x)), ay = to_u32_x86(floor((double)a.y));
    t.aPixel = (0.f <= a.x && a.x < (float)c.W && 0.f <= a.y && a.y < (float)c.H) ? ax + ay * c.W : 0xFFFFFFFFu;
    t.pad = 0;
    screen[i] = t;
    const uint64_t n = tri_slots(t);
    slots[i] = n;
    if (n - 1u > (uint64_t)kCamSmallBox) largeList[atomicAdd(largeCount, 1u)] = i;
}

// The per-pixel decision of FillRectangle (:163-213) for pixel (x, y) != the pixel of vertex a.
struct Edge {
    float px, py, qx, qy, sx, sy;  // sx = dx/dy, sy = 1/sx  (a division by zero fails the tests by design, :143)
};
__device__ __forceinline__ Edge make_edge(float px, float py, float qx, float qy) {
    Edge e;
    e.px = px; e.py = py; e.qx = qx; e.qy = qy;
    e.sx = (qx - px) / (qy - py);
    e.sy = 1.f / e.sx;
    return e;
}
__device__ __forceinline__ bool edge_touches(const Edge& e, uint32_t x, uint32_t y) {
    const float i0 = e.px + ((float)y - e.py) * e.sx;
    const float i1 = e.py + ((float)x - e.px) * e.sy;
    const float i2 = i0 + e.sx;
    const float i3 = i1 + e.sy;
    return ((0.f <= (e.px - i0) * (i0 - e.qx)) & (x == to_u32_x86(i0))) | ((0.f <= (e.px - i2) * (i2 - e.qx)) & (x == to_u32_x86(i2))) |
           ((0.f <= (e.py - i1) * (i1 - e.qy)) & (y == to_u32_x86(i1))) | ((0.f <= (e.py - i3) * (i3 - e.qy)) & (y == to_u32_x86(i3)));
}

struct TriRaster {
    Edge eab, ebc, eca;
    float abx, aby, bcx, bcy, cax, cay;
    float ax, ay, bx, by, cx, cy;
};
__device__ __forceinline__ TriRaster make_raster(const TriScreen& t) {
    TriRaster r;
    r.eab = make_edge(t.ax, t.ay, t.bx, t.by);
    r.ebc = make_edge(t.bx, t.by, t.cx, t.cy);
    r.eca = make_edge(t.cx, t.cy, t.ax, t.ay);
    r.abx = t.bx - t.ax; r.aby = t.by - t.ay;
    r.bcx = t.cx - t.bx; r.bcy = t.cy - t.by;
    r.cax = t.ax - t.cx; r.cay = t.ay - t.cy;
    r.ax = t.ax; r.ay = t.ay; r.bx = t.bx; r.by = t.by; r.cx = t.cx; r.cy = t.cy;
    return r;
}
__device__ __forceinline__ bool pixel_taken(const TriRaster& r, uint32_t x, uint32_t y) {
    bool take = edge_touches(r.eab, x, y) | edge_touches(r.ebc, x, y) | edge_touches(r.eca, x, y);
    if (!take) {  // pixel corner inside the triangle (:196-211)
        const float axx = (float)x - r.ax, axy = (float)y - r.ay;
        const float bxx = (float)x - r.bx, bxy = (float)y - r.by;
        const float cxx = (float)x - r.cx, cxy = (float)y - r.cy;
        const float k0 = r.abx * axy - r.aby * axx;
        const float k1 = r.bcx * bxy - r.bcy * bxx;
        const float k2 = r.cax * cxy - r.cay * cxx;
        take = (0 <= k0 * k1) & (0 <= k1 * k2);
    }
    return take;
}

// Slot layout of triangle i: [offset] = the vertex-a pixel, [offset + 1 + j] = bbox pixel j (x-major like the reference's loops).
__device__ __forceinline__ void raster_slot(const TriRaster& r, const TriScreen& t, uint32_t tri, uint64_t N, uint64_t sentinel,
                                            uint32_t W, uint32_t j, uint64_t* __restrict__ keys, uint32_t* __restrict__ pixelCount) {
    const uint32_t bh = t.y1 - t.y0 + 1u;
    const uint32_t x = t.x0 + j / bh, y = t.y0 + j % bh;
    const uint32_t pix = x + y * W;
    uint64_t key = sentinel;
    if (pix != t.aPixel && pixel_taken(r, x, y)) {
        key = (uint64_t)pix * N + tri;
        atomicAdd(pixelCount + pix, 1u);
    }
    keys[j] = key;
}

__global__ void __launch_bounds__(128) cam_raster_small_kernel(uint32_t triangleCount, uint32_t W, uint64_t sentinel,
                                                               const TriScreen* __restrict__ screen, const uint64_t* __restrict__ offset,
                                                               uint64_t* __restrict__ keys, uint32_t* __restrict__ pixelCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= triangleCount) return;
    const TriScreen t = screen[i];
    const uint64_t slots = tri_slots(t);
    uint64_t* out = keys + offset[i];
    if (t.aPixel != 0xFFFFFFFFu) {
        out[0] = (uint64_t)t.aPixel * triangleCount + i;
        atomicAdd(pixelCount + t.aPixel, 1u);
    } else {
        out[0] = sentinel;
    }
    if (slots - 1u == 0u || slots - 1u > (uint64_t)kCamSmallBox) return;   // large boxes: cam_raster_large_kernel
    const TriRaster r = make_raster(t);
    for (uint32_t j = 0; j < (uint32_t)(slots - 1u); ++j) raster_slot(r, t, i, triangleCount, sentinel, W, j, out + 1, pixelCount);
}

__global__ void __launch_bounds__(256) cam_raster_large_kernel(uint32_t triangleCount, uint32_t W, uint64_t sentinel,
                                                               const TriScreen* __restrict__ screen, const uint64_t* __restrict__ offset,
                                                               const uint32_t* __restrict__ largeList, uint64_t* __restrict__ keys,
                                                               uint32_t* __restrict__ pixelCount) {
    const uint32_t i = largeList[blockIdx.x];
    const TriScreen t = screen[i];
    const uint64_t slots = tri_slots(t) - 1u;
    const TriRaster r = make_raster(t);
    uint64_t* out = keys + offset[i] + 1;
    for (uint64_t j = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; j < slots; j += (uint64_t)gridDim.y * blockDim.x)
        raster_slot(r, t, i, triangleCount, sentinel, W, (uint32_t)j, out, pixelCount);
}

// sorted keys -> triangle ids (the valid keys come first; `real` = their count = startInc[P])
__global__ void __launch_bounds__(256) cam_split_kernel(const uint64_t* __restrict__ keys, uint64_t real, uint64_t N, uint32_t* __restrict__ list) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < real) list[k] = (uint32_t)(keys[k] % N);
}

// ---- storage compression (:580-613) ----------------------------------------------------------------------------------------------
// parent[p] = p when pixel p keeps its own copy, else the neighbour whose copy it shares; keptLen[p] = its length when kept.
__global__ void __launch_bounds__(256) cam_equal_kernel(uint32_t W, uint32_t P, const uint32_t* __restrict__ startInc,
                                                        const uint32_t* __restrict__ list, uint32_t* __restrict__ parent,
                                                        uint32_t* __restrict__ keptLen) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint32_t s = startInc[p], len = startInc[p + 1] - s;
    const uint32_t x = p % W, y = p / W;
    uint32_t par = p;
    auto same = [&](uint32_t q) {
        const uint32_t qs = startInc[q];
        if (startInc[q + 1] - qs != len) return false;
        for (uint32_t k = 0; k < len; ++k)
            if (list[qs + k] != list[s + k]) return false;
        return true;
    };
    if (0 < x && same(p - 1))
        par = p - 1;
    else if (0 < y && same(p - W))
        par = p - W;
    parent[p] = par;
    keptLen[p] = par == p ? len : 0u;
}

// Pointer doubling towards the pixel that keeps the copy (parents point left / up, so chains are at most W + H long).
__global__ void __launch_bounds__(256) cam_jump_kernel(uint32_t P, uint32_t* __restrict__ parent, uint32_t* __restrict__ changed) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint32_t a = parent[p], b = parent[a];
    if (a != b) {
        parent[p] = b;
        *changed = 1u;
    }
}

__global__ void __launch_bounds__(256) cam_finish_kernel(uint32_t P, const uint32_t* __restrict__ startInc, const uint32_t* __restrict__ parent,
                                                         const uint32_t* __restrict__ keptStart, const uint32_t* __restrict__ list,
                                                         uint32_t* __restrict__ outStart, uint32_t* __restrict__ outEnd,
                                                         uint32_t* __restrict__ outList) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint32_t root = parent[p];
    const uint32_t len = startInc[root + 1] - startInc[root];
    const uint32_t ns = keptStart[root];
    outStart[p] = ns;
    outEnd[p] = ns + len;
    if (root == p) {
        const uint32_t s = startInc[p];
        for (uint32_t k = 0; k < len; ++k) outList[ns + k] = list[s + k];
    }
}

}  // namespace oclr
