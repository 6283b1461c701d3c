// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// C calling-convention doors onto the *unmodified* reference objects that oracle/build_ref.py compiles
// from /root/reference:
//   * CameraTriangleList::New / SceneTriangleList::New (source/util/trianglelist.cpp:520-626, 655-737)
//     are C++ static factories returning objects; ctypes needs flat functions and caller-owned copies.
//   * ref_raytrace_threads(): drives the reference's exported kernel-as-C-function `Raytrace`
//     (source/opencl/raytrace_opencl.c:406-742) over disjoint interleaved rows from N host threads, with
//     private nextPixelId / sampleId cells -- the stand-in for "the reference kernel on all host cores
//     through a CPU OpenCL runtime" (PoCL is not installed).  Output is
//     bit-identical to RaytraceAll(0, ...) because the C-path seed depends only on (pixel, sample)
//     (raytrace_opencl.c:478-481 with raytrace.c:612-616).
#include <thread>
#include <vector>
#include <atomic>
#include <cstring>
#include <cstddef>

#include "trianglelist.h"   // pulls raytrace.h, which redefines int/float/... as macros: undo that here
#undef bool
#undef double
#undef float
#undef float2
#undef float3
#undef int
#undef int2
#undef int3
#undef uchar3
#undef uint
#undef uint2
#undef ulong
#undef ushort
#undef max
#undef min

extern "C" void Raytrace(cl_uint* nextPixelId, cl_uint2* dim, cl_float3* eye, cl_float3* eyeToTopLeft,
                         cl_float3* leftToRight, cl_float3* topToBottom, cl_float* pixelSizeInv,
                         cl_uint* camStart, cl_uint* camEnd, cl_uint* camList, cl_uint* sampleId,
                         cl_uint* sampleCount, cl_float3* vertex, cl_uint* triangleCount, cl_int3* triIdx,
                         cl_int* triMat, cl_float2* triUv, cl_float3* triNormal, cl_int* axesDivCount,
                         cl_float3* boxMin, cl_uint* gridStart, cl_uint* gridList, cl_uint2* matSize,
                         cl_int* matStart, cl_uchar3* textures, cl_uint* lightCount, cl_int* lightType,
                         cl_float3* lightPos, cl_float3* lightDir, cl_float3* lightColour,
                         cl_float* lightRadius, cl_float* lightHalf, cl_ushort* outR, cl_ushort* outG,
                         cl_ushort* outB);

// ---- event counters of the instrumented build (oracle/build_ref.py, libref_raytrace_counted.so) -------------
// The instrumented kernel copy bumps thread-local cells; every worker adds its cells to the totals when it is done.
#ifdef REF_COUNTED
#include <mutex>
extern "C" { __thread unsigned long long ref_cnt[16]; }
static unsigned long long g_refTotals[16];
static std::mutex g_refMutex;
static void ref_counters_flush() {
    std::lock_guard<std::mutex> lock(g_refMutex);
    for (int i = 0; i < 16; ++i) { g_refTotals[i] += ref_cnt[i]; ref_cnt[i] = 0; }
}
extern "C" void ref_counters_reset() {
    std::lock_guard<std::mutex> lock(g_refMutex);
    for (int i = 0; i < 16; ++i) g_refTotals[i] = 0;
}
extern "C" void ref_counters_read(unsigned long long* out) {
    ref_counters_flush();   // (the calling thread's own cells: RaytraceAll(0) runs in the caller)
    std::lock_guard<std::mutex> lock(g_refMutex);
    for (int i = 0; i < 16; ++i) out[i] = g_refTotals[i];
}
#else
static void ref_counters_flush() {}
#endif

extern "C" {

// ---- camera list ---------------------------------------------------------------------------------
void* ref_camera_list_new(cl_uint w, cl_uint h, const float* eye, const float* eyeToTopLeft,
                          const float* leftToRight, const float* topToBottom, float pixelSizeInv,
                          cl_uint vertexCount, cl_uint triangleCount, cl_float3* vertex, cl_int3* triIdx) {
    cl_uint2 dim; dim.s[0] = w; dim.s[1] = h;
    cl_float3 e, tl, lr, tb;
    for (int i = 0; i < 3; ++i) { e.s[i] = eye[i]; tl.s[i] = eyeToTopLeft[i]; lr.s[i] = leftToRight[i]; tb.s[i] = topToBottom[i]; }
    e.s[3] = tl.s[3] = lr.s[3] = tb.s[3] = 0.f;
    return CameraTriangleList::New(dim, e, tl, lr, tb, pixelSizeInv, vertexCount, triangleCount, vertex, triIdx);
}
ptrdiff_t ref_camera_list_size(void* p) { return ((CameraTriangleList*)p)->GetTriangleListSize(); }
void ref_camera_list_copy(void* p, cl_uint pixelCount, cl_uint* start, cl_uint* end, cl_uint* list) {
    CameraTriangleList* c = (CameraTriangleList*)p;
    memcpy(start, c->GetTriangleListStart(), sizeof(cl_uint) * pixelCount);
    memcpy(end, c->GetTriangleListEnd(), sizeof(cl_uint) * pixelCount);
    memcpy(list, c->GetTriangleList(), sizeof(cl_uint) * (size_t)c->GetTriangleListSize());
}
void ref_camera_list_free(void* p) { delete (CameraTriangleList*)p; }

// ---- scene grid ----------------------------------------------------------------------------------
int ref_scene_axes_division() { return (int)SceneTriangleList::AXES_DIVISION; }
void* ref_scene_list_new(cl_uint vertexCount, cl_uint triangleCount, cl_float3* vertex, cl_int3* triIdx) {
    return SceneTriangleList::New(vertexCount, triangleCount, vertex, triIdx);
}
cl_uint ref_scene_list_size(void* p) {
    const int n = SceneTriangleList::AXES_DIVISION;
    return ((SceneTriangleList*)p)->GetTriangleListStart()[n * n * n];
}
void ref_scene_list_copy(void* p, cl_float3* boxMin, cl_uint* start, cl_uint* list) {
    SceneTriangleList* s = (SceneTriangleList*)p;
    const int n = SceneTriangleList::AXES_DIVISION;
    memcpy(boxMin, s->GetBoxMin(), sizeof(cl_float3) * (n + 1));
    memcpy(start, s->GetTriangleListStart(), sizeof(cl_uint) * ((size_t)n * n * n + 1));
    memcpy(list, s->GetTriangleListStart()[n * n * n] ? s->GetTriangleList() : list,
           sizeof(cl_uint) * (size_t)s->GetTriangleListStart()[n * n * n]);
}
void ref_scene_list_free(void* p) { delete (SceneTriangleList*)p; }

// ---- the reference kernel on N host threads --------------------------------------------------------
// Rows [rowBegin, rowEnd) only (so a bounded sample of a big frame can be timed); rows are dealt to
// threads dynamically.  Output planes are full-frame sized; untouched rows keep their contents.
// rowStep > 1: only rows rowBegin, rowBegin + rowStep, ... (a sample spread evenly over the frame).
void ref_raytrace_threads_step(int nThreads, cl_uint rowBegin, cl_uint rowEnd, cl_uint rowStep,
                          cl_uint w, cl_uint h, const float* eye, const float* eyeToTopLeft,
                          const float* leftToRight, const float* topToBottom, float pixelSizeInv,
                          cl_uint* camStart, cl_uint* camEnd, cl_uint* camList, cl_uint sampleCount,
                          cl_float3* vertex, cl_uint triangleCount, cl_int3* triIdx, cl_int* triMat,
                          cl_float2* triUv, cl_float3* triNormal, cl_int axesDivCount, cl_float3* boxMin,
                          cl_uint* gridStart, cl_uint* gridList, cl_uint2* matSize, cl_int* matStart,
                          cl_uchar3* textures, cl_uint lightCount, cl_int* lightType, cl_float3* lightPos,
                          cl_float3* lightDir, cl_float3* lightColour, cl_float* lightRadius,
                          cl_float* lightHalf, cl_ushort* outR, cl_ushort* outG, cl_ushort* outB) {
    cl_uint2 dim; dim.s[0] = w; dim.s[1] = h;
    cl_float3 e, tl, lr, tb;
    for (int i = 0; i < 3; ++i) { e.s[i] = eye[i]; tl.s[i] = eyeToTopLeft[i]; lr.s[i] = leftToRight[i]; tb.s[i] = topToBottom[i]; }
    e.s[3] = tl.s[3] = lr.s[3] = tb.s[3] = 0.f;
    if (rowEnd > h) rowEnd = h;
    std::atomic<cl_uint> nextRow(rowBegin);
    auto worker = [&]() {
        cl_uint2 d = dim; cl_float3 le = e, ltl = tl, llr = lr, ltb = tb; cl_float psi = pixelSizeInv;
        cl_uint sc = sampleCount, tc = triangleCount, lc = lightCount; cl_int adc = axesDivCount;
        for (;;) {
            cl_uint row = nextRow.fetch_add(rowStep ? rowStep : 1);
            if (row >= rowEnd) break;
            for (cl_uint x = 0; x < w; ++x) {
                cl_uint pixel = row * w + x;
                cl_uint sampleId = 0;
                while (sampleId < sc) {   // Raytrace() pre-increments *sampleId (raytrace_opencl.c:474)
                    Raytrace(&pixel, &d, &le, &ltl, &llr, &ltb, &psi, camStart, camEnd, camList, &sampleId, &sc,
                             vertex, &tc, triIdx, triMat, triUv, triNormal, &adc, boxMin, gridStart, gridList,
                             matSize, matStart, textures, &lc, lightType, lightPos, lightDir, lightColour,
                             lightRadius, lightHalf, outR, outG, outB);
                }
            }
        }
        ref_counters_flush();
    };
    if (nThreads <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
}

void ref_raytrace_threads(int nThreads, cl_uint rowBegin, cl_uint rowEnd,
                          cl_uint w, cl_uint h, const float* eye, const float* eyeToTopLeft,
                          const float* leftToRight, const float* topToBottom, float pixelSizeInv,
                          cl_uint* camStart, cl_uint* camEnd, cl_uint* camList, cl_uint sampleCount,
                          cl_float3* vertex, cl_uint triangleCount, cl_int3* triIdx, cl_int* triMat,
                          cl_float2* triUv, cl_float3* triNormal, cl_int axesDivCount, cl_float3* boxMin,
                          cl_uint* gridStart, cl_uint* gridList, cl_uint2* matSize, cl_int* matStart,
                          cl_uchar3* textures, cl_uint lightCount, cl_int* lightType, cl_float3* lightPos,
                          cl_float3* lightDir, cl_float3* lightColour, cl_float* lightRadius,
                          cl_float* lightHalf, cl_ushort* outR, cl_ushort* outG, cl_ushort* outB) {
    ref_raytrace_threads_step(nThreads, rowBegin, rowEnd, 1, w, h, eye, eyeToTopLeft, leftToRight, topToBottom, pixelSizeInv, camStart, camEnd,
                              camList, sampleCount, vertex, triangleCount, triIdx, triMat, triUv, triNormal, axesDivCount, boxMin, gridStart,
                              gridList, matSize, matStart, textures, lightCount, lightType, lightPos, lightDir, lightColour, lightRadius,
                              lightHalf, outR, outG, outB);
}

// ---- the reference's own RaytraceAll through a pointer-only door (ctypes cannot pass the vector unions by value) ------
cl_uint ref_raytrace_all(cl_uint computationType, cl_uint w, cl_uint h, const float* eye, const float* eyeToTopLeft,
                         const float* leftToRight, const float* topToBottom, float pixelSizeInv,
                         cl_uint* camStart, cl_uint* camEnd, cl_uint* camList, cl_uint sampleCount,
                         cl_float3* vertex, cl_uint triangleCount, cl_int3* triIdx, cl_int* triMat,
                         cl_float2* triUv, cl_float3* triNormal, cl_int axesDivCount, cl_float3* boxMin,
                         cl_uint* gridStart, cl_uint* gridList, cl_uint2* matSize, cl_int* matStart,
                         cl_uchar3* textures, cl_uint lightCount, cl_int* lightType, cl_float3* lightPos,
                         cl_float3* lightDir, cl_float3* lightColour, cl_float* lightRadius,
                         cl_float* lightHalf, cl_ushort* outR, cl_ushort* outG, cl_ushort* outB) {
    cl_uint2 dim; dim.s[0] = w; dim.s[1] = h;
    cl_float3 e, tl, lr, tb;
    for (int i = 0; i < 3; ++i) { e.s[i] = eye[i]; tl.s[i] = eyeToTopLeft[i]; lr.s[i] = leftToRight[i]; tb.s[i] = topToBottom[i]; }
    e.s[3] = tl.s[3] = lr.s[3] = tb.s[3] = 0.f;
    // vertexCount / materialCount / texturesSize / camera list size are only used by the OpenCL branch for buffer sizes
    return RaytraceAll(computationType, dim, e, tl, lr, tb, pixelSizeInv, camStart, camEnd, camList, 0, sampleCount, 0, vertex,
                       triangleCount, triIdx, triMat, triUv, triNormal, axesDivCount, boxMin, gridStart, gridList, 0, matSize,
                       matStart, 0, textures, lightCount, lightType, lightPos, lightDir, lightColour, lightRadius, lightHalf,
                       outR, outG, outB);
}

}  // extern "C"
