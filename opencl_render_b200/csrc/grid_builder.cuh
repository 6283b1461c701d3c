// grid_builder.cuh -- SceneTriangleList::New on the device (source/util/trianglelist.cpp:655-737 with FillCube :452-503 and
// BoxIntersectsTriangle / Cull :381-449).  The reference spends a 2 MB memset per triangle here (10 s at 101 k triangles, ~17 min
// at 10 M); the host restatement (builders.cpp) takes 0.3-1.4 s on 16 cores; this is the same result from the GPU.
//
//   planes   per axis: radix-sort the vertex coordinates, plane i = midpoint of the sorted values at index i*(V-1)/n and its
//            predecessor (:660-678, unsigned 32-bit index arithmetic like the reference);
//   cells    the reference flood-fills from the cell of vertex a over face neighbours whose box clips the triangle to a
//            non-empty polygon.  Here every triangle gets the block of cells whose closed boxes overlap its closed bounding
//            box (only those can pass the clip test) and walks the SAME flood fill inside it -- small blocks by one thread with
//            an explicit queue, large ones (a floor quad covers 65 k cells) by one CTA with parallel frontier sweeps -- so the set
//            is the reference's connected component, not merely "all cells that pass";
//   lists    every (triangle, block cell) owns one slot of a key pool: key = cell * N + triangle (the reference's sort key) when
//            the fill reached the cell, a sentinel otherwise; cub radix sort; per-cell histogram + scan = the CSR starts.
// The clip arithmetic is rt_clip.h, shared with the host builder.  Output is entry-for-entry the host builder's
// (tests/test_gpu_parity.py::test_device_scene_grid_equals_host_builder).
#pragma once
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "rt_clip.h"

namespace oclr {

struct TriCells {   // per triangle: block of candidate cells and the fill's seed
    int x0, y0, z0, dx, dy, dz;
    uint32_t seed;   // index of the seed cell inside the block
    uint32_t pad;
};

enum { kGridSmallBlock = 512 };   // blocks up to this many cells are filled by one thread

__global__ void __launch_bounds__(256) grid_coord_kernel(uint32_t V, const float4* __restrict__ vertex, int axis, float* __restrict__ out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float4 p = vertex[v];
    out[v] = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
}

__global__ void grid_planes_kernel(uint32_t V, const float* __restrict__ sorted, int n, float* __restrict__ planes /* n + 1 */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t idx = ((uint32_t)i * (V - 1u)) / (uint32_t)n;
    planes[i] = (0u < idx && idx < V) ? (sorted[idx] + sorted[idx - 1]) / 2.f : sorted[idx];
}

// lowest cell c in [0, n-1] whose upper plane p[c+1] >= v (n-1 when there is none)
__device__ __forceinline__ int first_cell_reaching(const float* __restrict__ p, int n, float v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p[mid + 1] >= v) hi = mid; else lo = mid + 1;
    }
    return lo;
}
// highest cell c in [0, n-1] whose lower plane p[c] <= v (0 when there is none)
__device__ __forceinline__ int last_cell_reached(const float* __restrict__ p, int n, float v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) grid_range_kernel(uint32_t N, const float4* __restrict__ vertex, const int4* __restrict__ triIdx, int n,
                                                         const float* __restrict__ planes, TriCells* __restrict__ cells,
                                                         uint64_t* __restrict__ slots, uint32_t* __restrict__ largeList,
                                                         uint32_t* __restrict__ largeCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* px = planes;
    const float* py = planes + (n + 1);
    const float* pz = planes + 2 * (n + 1);
    const int4 vi = triIdx[i];
    const float4 A = vertex[vi.x], B = vertex[vi.y], C = vertex[vi.z];
    int sx, sy, sz;
    box_address(n, px, py, pz, mk3(A.x, A.y, A.z), sx, sy, sz);   // FillCube starts in the cell of vertex a (:455)
    TriCells t;
    int x0 = first_cell_reaching(px, n, fminf(A.x, fminf(B.x, C.x))), x1 = last_cell_reached(px, n, fmaxf(A.x, fmaxf(B.x, C.x)));
    int y0 = first_cell_reaching(py, n, fminf(A.y, fminf(B.y, C.y))), y1 = last_cell_reached(py, n, fmaxf(A.y, fmaxf(B.y, C.y)));
    int z0 = first_cell_reaching(pz, n, fminf(A.z, fminf(B.z, C.z))), z1 = last_cell_reached(pz, n, fmaxf(A.z, fmaxf(B.z, C.z)));
    x0 = min(x0, sx); x1 = max(x1, sx);
    y0 = min(y0, sy); y1 = max(y1, sy);
    z0 = min(z0, sz); z1 = max(z1, sz);
    t.x0 = x0; t.y0 = y0; t.z0 = z0;
    t.dx = x1 - x0 + 1; t.dy = y1 - y0 + 1; t.dz = z1 - z0 + 1;
    t.seed = (uint32_t)((sx - x0) + t.dx * ((sy - y0) + t.dy * (sz - z0)));
    t.pad = 0;
    cells[i] = t;
    const uint64_t s = (uint64_t)t.dx * t.dy * t.dz;
    slots[i] = s;
    if (s > (uint64_t)kGridSmallBlock) largeList[atomicAdd(largeCount, 1u)] = i;
}

// Clip test of block cell `local` of triangle (a, b, c).
__device__ __forceinline__ bool cell_hits(const TriCells& t, uint32_t local, int n, const float* __restrict__ planes, f3 a, f3 b, f3 c) {
    const int cx = t.x0 + (int)(local % (uint32_t)t.dx), cy = t.y0 + (int)((local / (uint32_t)t.dx) % (uint32_t)t.dy),
              cz = t.z0 + (int)(local / ((uint32_t)t.dx * (uint32_t)t.dy));
    const float lo[3] = {planes[cx], planes[(n + 1) + cy], planes[2 * (n + 1) + cz]};
    const float hi[3] = {planes[cx + 1], planes[(n + 1) + cy + 1], planes[2 * (n + 1) + cz + 1]};
    return box_hits_triangle(lo, hi, a, b, c);
}

__device__ __forceinline__ uint32_t cell_global(const TriCells& t, uint32_t local, int n) {
    const uint32_t cx = (uint32_t)t.x0 + local % (uint32_t)t.dx, cy = (uint32_t)t.y0 + (local / (uint32_t)t.dx) % (uint32_t)t.dy,
                   cz = (uint32_t)t.z0 + local / ((uint32_t)t.dx * (uint32_t)t.dy);
    return cx + (uint32_t)n * cy + (uint32_t)n * (uint32_t)n * cz;
}

// state: 0 not looked at, 1 in the fill, 2 rejected by the clip test
// Small blocks: one thread runs the reference's flood fill with an explicit queue (the clip test is only evaluated for face
// neighbours of cells already in the fill, like FillCube).
__global__ void __launch_bounds__(128) grid_fill_small_kernel(uint32_t N, const float4* __restrict__ vertex, const int4* __restrict__ triIdx, int n,
                                                              const float* __restrict__ planes, const TriCells* __restrict__ cells,
                                                              const uint64_t* __restrict__ offset, uint8_t* __restrict__ state,
                                                              uint16_t* __restrict__ queue, uint64_t sentinel, uint64_t* __restrict__ keys,
                                                              uint32_t* __restrict__ cellCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const TriCells t = cells[i];
    const uint32_t total = (uint32_t)t.dx * (uint32_t)t.dy * (uint32_t)t.dz;
    if ((uint64_t)t.dx * t.dy * t.dz > (uint64_t)kGridSmallBlock) return;   // grid_fill_large_kernel
    const uint64_t off = offset[i];
    uint8_t* st = state + off;
    uint16_t* q = queue + off;
    const int4 vi = triIdx[i];
    const float4 A = vertex[vi.x], B = vertex[vi.y], C = vertex[vi.z];
    const f3 a = mk3(A.x, A.y, A.z), b = mk3(B.x, B.y, B.z), c = mk3(C.x, C.y, C.z);
    for (uint32_t k = 0; k < total; ++k) st[k] = 0;
    uint32_t head = 0, tail = 0;
    st[t.seed] = 1;
    q[tail++] = (uint16_t)t.seed;
    const int stride[3] = {1, t.dx, t.dx * t.dy}, dim[3] = {t.dx, t.dy, t.dz};
    while (head < tail) {
        const uint32_t cur = q[head++];
        const int cc[3] = {(int)(cur % (uint32_t)t.dx), (int)((cur / (uint32_t)t.dx) % (uint32_t)t.dy), (int)(cur / ((uint32_t)t.dx * (uint32_t)t.dy))};
        for (int k = 0; k < 3; ++k)
            for (int d = -1; d <= 1; d += 2) {
                const int nc = cc[k] + d;
                if (nc < 0 || nc >= dim[k]) continue;   // outside the block no cell can pass the clip test
                const uint32_t nb = (uint32_t)((int)cur + d * stride[k]);
                if (st[nb] != 0) continue;
                const bool hit = cell_hits(t, nb, n, planes, a, b, c);
                st[nb] = hit ? 1 : 2;
                if (hit) q[tail++] = (uint16_t)nb;
            }
    }
    for (uint32_t k = 0; k < total; ++k) {
        uint64_t key = sentinel;
        if (st[k] == 1) {
            const uint32_t cell = cell_global(t, k, n);
            key = (uint64_t)cell * N + i;
            atomicAdd(cellCount + cell, 1u);
        }
        keys[off + k] = key;
    }
}

// Large blocks: one CTA per triangle.  The clip test is evaluated for every block cell in parallel (state 3 = passes, not yet
// reached), then the fill grows from the seed by frontier sweeps until a sweep adds nothing.
__global__ void __launch_bounds__(1024) grid_fill_large_kernel(uint32_t N, const float4* __restrict__ vertex, const int4* __restrict__ triIdx, int n,
                                                              const float* __restrict__ planes, const TriCells* __restrict__ cells,
                                                              const uint64_t* __restrict__ offset, const uint32_t* __restrict__ largeList,
                                                              uint8_t* __restrict__ state, uint64_t sentinel, uint64_t* __restrict__ keys,
                                                              uint32_t* __restrict__ cellCount) {
    __shared__ int changed;
    const uint32_t i = largeList[blockIdx.x];
    const TriCells t = cells[i];
    const uint64_t total = (uint64_t)t.dx * t.dy * t.dz;
    const uint64_t off = offset[i];
    uint8_t* st = state + off;
    const int4 vi = triIdx[i];
    const float4 A = vertex[vi.x], B = vertex[vi.y], C = vertex[vi.z];
    const f3 a = mk3(A.x, A.y, A.z), b = mk3(B.x, B.y, B.z), c = mk3(C.x, C.y, C.z);
    for (uint64_t k = threadIdx.x; k < total; k += blockDim.x) st[k] = (k == t.seed) ? 1 : (cell_hits(t, (uint32_t)k, n, planes, a, b, c) ? 3 : 2);
    __syncthreads();
    const int64_t sx = 1, sy = t.dx, sz = (int64_t)t.dx * t.dy;
    // Every thread owns a contiguous run of cells and sweeps it forwards, then backwards (raster-scan style): a frontier crosses
    // a whole run in one sweep instead of advancing one cell per sweep.  A cell promoted in a sweep may already serve its
    // neighbours in the same sweep: the fixed point -- the connected component of the seed -- is the same.
    const uint64_t chunk = (total + blockDim.x - 1) / blockDim.x;
    const uint64_t kBegin = (uint64_t)threadIdx.x * chunk, kEnd = kBegin + chunk < total ? kBegin + chunk : total;
    auto grow = [&](uint64_t k) {
        if (st[k] != 3) return;
        const int cx = (int)(k % (uint64_t)t.dx), cy = (int)((k / (uint64_t)t.dx) % (uint64_t)t.dy), cz = (int)(k / (uint64_t)sz);
        const bool reach = (cx > 0 && st[k - sx] == 1) || (cx + 1 < t.dx && st[k + sx] == 1) || (cy > 0 && st[k - sy] == 1) ||
                           (cy + 1 < t.dy && st[k + sy] == 1) || (cz > 0 && st[k - sz] == 1) || (cz + 1 < t.dz && st[k + sz] == 1);
        if (reach) {
            st[k] = 1;
            changed = 1;
        }
    };
    for (;;) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        for (uint64_t k = kBegin; k < kEnd; ++k) grow(k);
        __syncthreads();
        for (uint64_t k = kEnd; k > kBegin; --k) grow(k - 1);
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    for (uint64_t k = threadIdx.x; k < total; k += blockDim.x) {
        uint64_t key = sentinel;
        if (st[k] == 1) {
            const uint32_t cell = cell_global(t, (uint32_t)k, n);
            key = (uint64_t)cell * N + i;
            atomicAdd(cellCount + cell, 1u);
        }
        keys[off + k] = key;
    }
}

__global__ void __launch_bounds__(256) grid_split_kernel(const uint64_t* __restrict__ keys, uint64_t real, uint64_t N, uint32_t* __restrict__ list) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < real) list[k] = (uint32_t)(keys[k] % N);
}

__global__ void grid_boxmin_kernel(int n, const float* __restrict__ planes, float4* __restrict__ boxMin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    boxMin[i] = make_float4(planes[i], planes[(n + 1) + i], planes[2 * (n + 1) + i], 0.f);
}

}  // namespace oclr
