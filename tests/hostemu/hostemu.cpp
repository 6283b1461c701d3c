// tests/hostemu/hostemu.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the device arithmetic of the product (opencl_render_b200/csrc/rt_core.h, the same header the CUDA kernels
// instantiate) and the scene packer for the HOST, so that the packed layout + restated control flow can be compared
// bit-for-bit with the reference on a machine without a GPU.  It is built into tests/_build/ by tests/conftest.py and is
// never part of, nor reachable from, libopencl_render_b200.so (which has no CPU compute path).
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/oclr_abi.h"
#include "../../opencl_render_b200/csrc/rt_core.h"
#include "../../opencl_render_b200/csrc/runtime.h"

using namespace oclr;

extern "C" int hostemu_render(const oclr_scene_desc* d, const oclr_camera* cam, const uint32_t* camStart, const uint32_t* camEnd,
                              const uint32_t* camList, uint32_t sampleCount, uint32_t rowBegin, uint32_t rowEnd, uint16_t* outR,
                              uint16_t* outG, uint16_t* outB, uint32_t* ids, uint8_t* flags, oclr_counters* counters, int threads) {
    HostScene h;
    h.vertexCount = d->vertexCount; h.vertex = (const float4*)d->vertex;
    h.triangleCount = d->triangleCount; h.triIdx = (const int32_t*)d->triangleVertexIndex; h.triMat = d->triangleMaterialId;
    h.triUv = (const float*)d->triangleUv; h.triNormal = (const float4*)d->triangleNormal;
    h.axesDivCount = d->axesDivCount; h.boxMin = (const float4*)d->sceneBoxMin; h.gridStart = d->scenePixelTriangleListStart;
    h.gridList = d->scenePixelTriangleList; h.materialCount = d->materialCount; h.matSize = (const uint2*)d->materialImageSize;
    h.matStart = d->materialImageStart; h.texturesSize = d->texturesSize; h.textures = (const uchar4*)d->textures;
    h.lightCount = d->lightCount; h.lightType = d->lightType; h.lightPos = (const float4*)d->lightPosition;
    h.lightDir = (const float4*)d->lightDirection; h.lightColour = (const float4*)d->lightColour; h.lightRadius = d->lightRadius;
    h.lightHalf = d->lightHalfAttenuationDistance;
    std::string err;
    if (!validate_scene(h, err)) { fprintf(stderr, "hostemu: %s\n", err.c_str()); return 0; }
    std::vector<float4> geo(4 * (size_t)h.triangleCount + 1), shade(8 * (size_t)h.triangleCount + 1);
    pack_triangles(h, geo.data(), shade.data(), threads);
    PackedGrid grid;
    if (!pack_grid(h, grid, err)) { fprintf(stderr, "hostemu: %s\n", err.c_str()); return 0; }
    std::vector<Light> lights;
    pack_lights(h, lights);
    SceneView S;
    S.triGeo = geo.data(); S.triShade = shade.data(); S.bricks = grid.bricks.data(); S.cellRange = grid.cellRange.data();
    S.cellList = grid.cellList.data(); S.faceMask = grid.faceMask.data(); S.planes = grid.planes.data(); S.matSize = h.matSize; S.matStart = h.matStart;
    S.textures = h.textures; S.lights = lights.data(); S.triangleCount = h.triangleCount; S.materialCount = h.materialCount;
    S.lightCount = h.lightCount; S.n = grid.n; S.nb = grid.nb;
    FrameView F;
    memcpy(&F.cam, cam, sizeof(Camera));
    F.camStart = camStart; F.camEnd = camEnd; F.camList = camList; F.sampleCount = sampleCount; F.rowBegin = rowBegin; F.rowEnd = rowEnd; F.bandRows = 0; F.bandRank = 0; F.bandWorld = 1; F.ownedRows = rowEnd - rowBegin;
    F.outR = outR; F.outG = outG; F.outB = outB; F.idOut = ids; F.flagOut = flags;
    const float* px = S.planes; const float* py = px + (S.n + 1); const float* pz = py + (S.n + 1);
    if (rowEnd > cam->height) rowEnd = cam->height;
    std::atomic<uint32_t> nextRow(rowBegin);
    std::vector<Counters> cnts(threads > 0 ? threads : 1);
    // walk mode: unset = the reference's cell walk, 1 = two-level (brick-skipping) walk, 2 = the trace kernel's packed form (rt_walk.h)
    // >= 3 = the packed walk cut into parts of that many cells (exact random access into the walk, rt_walk.h pwalk_jump)
    const int hierEnv = getenv("HOSTEMU_HIERARCHICAL") ? atoi(getenv("HOSTEMU_HIERARCHICAL")) : 0;
    const int hier = getenv("HOSTEMU_HIERARCHICAL") ? (hierEnv >= 2 ? hierEnv : 1) : 0;
    auto worker = [&](int tid) {
        Counters& cnt = cnts[tid];
        memset(&cnt, 0, sizeof(cnt));
        const float scale = 65535.f / (float)sampleCount;
        for (;;) {
            const uint32_t y = nextRow.fetch_add(1);
            if (y >= rowEnd) break;
            for (uint32_t x = 0; x < cam->width; ++x) {
                const uint32_t pixel = y * cam->width + x;
                uint16_t r = 0, g = 0, b = 0;
                bool undef = false;
                for (uint32_t s = 0; s < sampleCount; ++s) {
                    uint32_t pid;
                    const f3 c = trace_sample<true>(S, F, px, py, pz, pixel, s, &pid, undef, &cnt, hier);
                    if (s == 0 && ids) ids[pixel] = pid;
                    r = accumulate16(r, c.x, scale); g = accumulate16(g, c.y, scale); b = accumulate16(b, c.z, scale);
                }
                outR[pixel] = r; outG[pixel] = g; outB[pixel] = b;
                if (flags) flags[pixel] = undef ? 1 : 0;
            }
        }
    };
    if (threads <= 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
        for (auto& t : pool) t.join();
    }
    if (counters) {
        Counters tot; memset(&tot, 0, sizeof(tot));
        unsigned long long* a = (unsigned long long*)&tot;
        for (auto& c : cnts) { const unsigned long long* b = (const unsigned long long*)&c; for (size_t i = 0; i < sizeof(Counters) / 8; ++i) a[i] += b[i]; }
        memcpy(counters, &tot, sizeof(tot));
    }
    return 1;
}

// Host packer output for comparison with the device packers (pack_kernels.cuh).  Buffers sized by the caller:
// geo 64 B/tri, shade 128 B/tri, bricks 16 B/brick, ranges 8 B per non-empty cell (count returned), planes 3*(n+1) floats.
extern "C" long hostemu_pack(const oclr_scene_desc* d, void* geo, void* shade, void* bricks, void* ranges, size_t rangesCap, void* planes,
                             void* faceMask) {
    HostScene h;
    h.vertexCount = d->vertexCount; h.vertex = (const float4*)d->vertex;
    h.triangleCount = d->triangleCount; h.triIdx = (const int32_t*)d->triangleVertexIndex; h.triMat = d->triangleMaterialId;
    h.triUv = (const float*)d->triangleUv; h.triNormal = (const float4*)d->triangleNormal;
    h.axesDivCount = d->axesDivCount; h.boxMin = (const float4*)d->sceneBoxMin; h.gridStart = d->scenePixelTriangleListStart;
    h.gridList = d->scenePixelTriangleList; h.materialCount = d->materialCount; h.matSize = (const uint2*)d->materialImageSize;
    h.matStart = d->materialImageStart; h.texturesSize = d->texturesSize; h.textures = (const uchar4*)d->textures;
    std::string err;
    pack_triangles(h, (float4*)geo, (float4*)shade, 2);
    PackedGrid grid;
    if (!pack_grid(h, grid, err)) return -1;
    memcpy(bricks, grid.bricks.data(), sizeof(uint4) * grid.bricks.size());
    if (grid.cellRange.size() * sizeof(uint2) > rangesCap) return -2;
    memcpy(ranges, grid.cellRange.data(), sizeof(uint2) * grid.cellRange.size());
    memcpy(planes, grid.planes.data(), sizeof(float) * grid.planes.size());
    if (faceMask) memcpy(faceMask, grid.faceMask.data(), sizeof(uint32_t) * 6 * grid.cellRange.size());
    return (long)grid.cellRange.size();
}
