// rt_kernels.cuh -- CUDA kernels of the raytrace path (sm_100a; no tensor cores: the path is scattered 128-bit
// fetches + fp32 scalar math).
#pragma once
#include <cuda_runtime.h>

#include "rt_core.h"

namespace oclr {

// Copies the 3*(n+1) split planes into shared memory; every GetBoxAddress / DDA step reads them (divergent
// indices -> shared memory, not constant memory).
__device__ __forceinline__ void load_planes(float* sh, const SceneView& S) {
    const int total = 3 * (S.n + 1);
    for (int i = threadIdx.x; i < total; i += blockDim.x) sh[i] = __ldg(S.planes + i);
    __syncthreads();
}

__device__ __forceinline__ void flush_counters(const Counters& c, Counters* g) {
    atomicAdd(&g->segments, c.segments);
    atomicAdd(&g->primCandidates, c.primCandidates);
    atomicAdd(&g->gridRays, c.gridRays);
    atomicAdd(&g->cells, c.cells);
    atomicAdd(&g->cellsNonEmpty, c.cellsNonEmpty);
    atomicAdd(&g->gridCandidates, c.gridCandidates);
    atomicAdd(&g->shadedHits, c.shadedHits);
    atomicAdd(&g->occluderLookups, c.occluderLookups);
    atomicAdd(&g->bricksLoaded, c.bricksLoaded);
    atomicAdd(&g->emptyBrickCells, c.emptyBrickCells);
    if (c.walkWarpIters) atomicAdd(&g->walkWarpIters, c.walkWarpIters);
    if (c.walkLaneIters) atomicAdd(&g->walkLaneIters, c.walkLaneIters);
    if (c.testWarpIters) atomicAdd(&g->testWarpIters, c.testWarpIters);
    if (c.testLaneIters) atomicAdd(&g->testLaneIters, c.testLaneIters);
    if (c.mailboxSkips) atomicAdd(&g->mailboxSkips, c.mailboxSkips);
    if (c.coarseSteps) atomicAdd(&g->coarseSteps, c.coarseSteps);
    if (c.coarseEnters) atomicAdd(&g->coarseEnters, c.coarseEnters);
    if (c.switchWarpIters) atomicAdd(&g->switchWarpIters, c.switchWarpIters);
    if (c.switchLaneIters) atomicAdd(&g->switchLaneIters, c.switchLaneIters);
    if (c.walkIdleLanes) atomicAdd(&g->walkIdleLanes, c.walkIdleLanes);
    if (c.walkParkedLanes) atomicAdd(&g->walkParkedLanes, c.walkParkedLanes);
    if (c.walkFinishedLanes) atomicAdd(&g->walkFinishedLanes, c.walkFinishedLanes);
    if (c.walkLowIters) atomicAdd(&g->walkLowIters, c.walkLowIters);
    if (c.walkExhaustedIters) atomicAdd(&g->walkExhaustedIters, c.walkExhaustedIters);
    if (c.splitAttempts) atomicAdd(&g->splitAttempts, c.splitAttempts);
    if (c.splitsDone) atomicAdd(&g->splitsDone, c.splitsDone);
    if (c.splitParts) atomicAdd(&g->splitParts, c.splitParts);
    if (c.splitCancelled) atomicAdd(&g->splitCancelled, c.splitCancelled);
    if (c.superSteps) atomicAdd(&g->superSteps, c.superSteps);
    if (c.superEnters) atomicAdd(&g->superEnters, c.superEnters);
    if (c.superRefines) atomicAdd(&g->superRefines, c.superRefines);
}

// ---- kernel A: one thread per pixel, serial control flow (the straightforward restatement) --------------------
// Block = 128 threads = 16x8 pixels; each warp covers an 8x4 pixel tile so primary and shadow rays of a warp are
// spatially coherent.  All samples of a pixel are accumulated by its thread in order (raytrace.c:615-652).
template <bool COUNT>
__global__ void __launch_bounds__(128) raytrace_simple_kernel(SceneView S, FrameView F, Counters* gcnt) {
    extern __shared__ float shPlanes[];
    if (F.camBad != nullptr && *F.camBad != 0u) return;   // caller's camera lists failed the range check
    load_planes(shPlanes, S);
    const float* px = shPlanes;
    const float* py = shPlanes + (S.n + 1);
    const float* pz = shPlanes + 2 * (S.n + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = map_row(F, blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3));
    if (x >= F.cam.width || y >= F.cam.height) return;
    const uint32_t pixel = y * F.cam.width + x;

    Counters cnt = {};
    const float scale = 65535.f / (float)F.sampleCount;
    // planes start from zero (raytrace.c:476-486); a later sample range of the same job adds to what is there
    uint16_t r = F.sampleBegin ? F.outR[pixel] : (uint16_t)0, g = F.sampleBegin ? F.outG[pixel] : (uint16_t)0,
             b = F.sampleBegin ? F.outB[pixel] : (uint16_t)0;
    bool undef = false;
    for (uint32_t s = F.sampleBegin; s < F.sampleEnd; ++s) {
        uint32_t pid;
        const f3 c = trace_sample<COUNT>(S, F, px, py, pz, pixel, s, &pid, undef, &cnt);
        if (s == 0 && F.idOut) F.idOut[pixel] = pid;
        r = accumulate16(r, c.x, scale);
        g = accumulate16(g, c.y, scale);
        b = accumulate16(b, c.z, scale);
    }
    F.outR[pixel] = r;
    F.outG[pixel] = g;
    F.outB[pixel] = b;
    if (F.flagOut) F.flagOut[pixel] = undef ? 1 : 0;
    if (COUNT) flush_counters(cnt, gcnt);
}

}  // namespace oclr
