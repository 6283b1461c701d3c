// builders.cpp -- host builders for the two acceleration structures the kernel consumes, and the camera set-up.
//
// Restates (not copies) the reference's
//   SetCamera                    source/render.cpp:461-491
//   CameraTriangleList::New      source/util/trianglelist.cpp:520-626  (+ GetCameraPosition :74-90, FillRectangle :131-217)
//   SceneTriangleList::New       source/util/trianglelist.cpp:655-737  (+ FillCube :452-503, BoxIntersectsTriangle/Cull :381-449)
// with the same fp32 decisions (compiled -ffp-contract=off) so the lists come out entry-for-entry identical, but
// with different machinery: triangles are processed in parallel chunks, (bin, triangle) pairs are bucketed by a
// counting sort instead of a 2 GB key array + quicksort (keys are unique, so any correct sort gives the same
// order), and the per-triangle 2 MB bitset memset (trianglelist.cpp:457) is replaced by clearing touched bits.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/oclr_abi.h"
#include "rt_clip.h"
#include "rt_core.h"

namespace oclr {

struct BinRef {
    uint32_t bin, tri;
};

static int worker_count(size_t items) {
    unsigned hw = std::thread::hardware_concurrency();
    int t = hw ? (int)hw : 1;
    if (items < 4096) t = 1;
    if (t > 64) t = 64;
    return t;
}

template <class Fn>
static void parallel_chunks(size_t items, int threads, Fn fn) {
    if (threads <= 1) {
        fn(0, (size_t)0, items);
        return;
    }
    std::vector<std::thread> pool;
    const size_t chunk = (items + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const size_t b = (size_t)t * chunk, e = std::min(items, b + chunk);
        if (b < e) pool.emplace_back(fn, t, b, e);
    }
    for (auto& th : pool) th.join();
}

// Chunks hold ascending triangle ids and chunk t precedes chunk t+1, so scattering chunk after chunk keeps every
// bin's entries in ascending triangle order -- the order the reference obtains by sorting bin*N+tri keys.
static void csr_from_chunks(size_t bins, const std::vector<std::vector<BinRef>>& chunks, uint32_t* start /*bins+1*/,
                            std::vector<uint32_t>& list) {
    memset(start, 0, sizeof(uint32_t) * (bins + 1));
    size_t total = 0;
    for (const auto& c : chunks) {
        total += c.size();
        for (const BinRef& r : c) ++start[r.bin + 1];
    }
    for (size_t i = 1; i <= bins; ++i) start[i] += start[i - 1];
    list.resize(total);
    std::vector<uint32_t> cursor(start, start + bins);
    for (const auto& c : chunks)
        for (const BinRef& r : c) list[cursor[r.bin]++] = r.tri;
}

// float -> cl_uint as the reference's x86-64 build performs it (cvttss2si to 64 bits, low half kept)
static inline uint32_t to_u32(float f) { return (uint32_t)(int64_t)f; }
static inline uint32_t to_u32(double f) { return (uint32_t)(int64_t)f; }

// ---- SetCamera: render.cpp:461-491 ---------------------------------------------------------------------------------
void set_camera(oclr_camera* out, const float position[3], const float object[3], const float up3[3], float fov, uint32_t w,
                uint32_t h) {
    const f3 camV = mk3(object[0] - position[0], object[1] - position[1], object[2] - position[2]);
    const f3 up = mk3(up3);
    const f3 r2l = cross3(up, camV);
    // render.cpp is C++: `tan(fov / 2.f)` and `sqrt(dot(...))` on float arguments resolve to the float overloads (tanf / sqrtf), not to
    // the double functions the C kernel path uses -- pinned against the reference's own lines compiled by g++ (oracle/_ref, ref_set_camera)
    const float midToLeft = sqrt_c(dot3(camV, camV)) * tanf(fov / 2.f);
    const float midToTop = midToLeft * (float)h / (float)w;
    const float r2lLen = sqrt_c(dot3(r2l, r2l));
    const float upLen = sqrt_c(dot3(up, up));
    const f3 r2lU = mk3(r2l.x / r2lLen, r2l.y / r2lLen, r2l.z / r2lLen);
    const f3 upU = mk3(up.x / upLen, up.y / upLen, up.z / upLen);
    const float psi = ((float)w) / (2.f * midToLeft);
    memset(out, 0, sizeof(*out));
    out->width = w;
    out->height = h;
    out->pixelSizeInv = psi;
    out->eye[0] = position[0];
    out->eye[1] = position[1];
    out->eye[2] = position[2];
    out->eyeToTopLeft[0] = camV.x - midToLeft * r2lU.x + midToTop * upU.x;
    out->eyeToTopLeft[1] = camV.y - midToLeft * r2lU.y + midToTop * upU.y;
    out->eyeToTopLeft[2] = camV.z - midToLeft * r2lU.z + midToTop * upU.z;
    out->leftToRight[0] = r2lU.x / psi;
    out->leftToRight[1] = r2lU.y / psi;
    out->leftToRight[2] = r2lU.z / psi;
    out->topToBottom[0] = -upU.x / psi;
    out->topToBottom[1] = -upU.y / psi;
    out->topToBottom[2] = -upU.z / psi;
}

// ---- camera lists ------------------------------------------------------------------------------------------------------
struct P2 {
    float x, y;
};

struct CamProj {
    f3 eye, tl, lr, tb, screenN;
    float tlDotN, psiSq;
};

// trianglelist.cpp:74-90
static inline P2 project(const CamProj& c, const float4& v) {
    const f3 e = mk3(v.x - c.eye.x, v.y - c.eye.y, v.z - c.eye.z);
    const float s = c.tlDotN / dot3(e, c.screenN);
    const f3 q = mk3(s * e.x - c.tl.x, s * e.y - c.tl.y, s * e.z - c.tl.z);
    P2 p;
    p.x = dot3(c.lr, q) * c.psiSq;
    p.y = dot3(c.tb, q) * c.psiSq;
    return p;
}

// One triangle edge p->q against the pixel (x,y): the four crossing tests of trianglelist.cpp:182-185
struct Edge {
    float px, py, qx, qy, sx, sy;  // sx = dx/dy, sy = 1/sx  (division by zero fails the tests by design, :143)
};
static inline Edge make_edge(P2 p, P2 q) {
    Edge e;
    e.px = p.x; e.py = p.y; e.qx = q.x; e.qy = q.y;
    e.sx = (q.x - p.x) / (q.y - p.y);
    e.sy = 1.f / e.sx;
    return e;
}
static inline bool edge_touches(const Edge& e, uint32_t x, uint32_t y) {
    const float i0 = e.px + ((float)y - e.py) * e.sx;
    const float i1 = e.py + ((float)x - e.px) * e.sy;
    const float i2 = i0 + e.sx;
    const float i3 = i1 + e.sy;
    return ((0.f <= (e.px - i0) * (i0 - e.qx)) & (x == to_u32(i0))) | ((0.f <= (e.px - i2) * (i2 - e.qx)) & (x == to_u32(i2))) |
           ((0.f <= (e.py - i1) * (i1 - e.qy)) & (y == to_u32(i1))) | ((0.f <= (e.py - i3) * (i3 - e.qy)) & (y == to_u32(i3)));
}

// trianglelist.cpp:131-217
static void raster_triangle(uint32_t W, uint32_t H, P2 a, P2 b, P2 c, uint32_t tri, std::vector<BinRef>& out) {
    const P2 ab = {b.x - a.x, b.y - a.y}, bc = {c.x - b.x, c.y - b.y}, ca = {a.x - c.x, a.y - c.y};
    const Edge eab = make_edge(a, b), ebc = make_edge(b, c), eca = make_edge(c, a);
    const float wm = (float)(W - 1), hm = (float)(H - 1);
    const uint32_t x0 = to_u32(fmaxf(0.f, fminf(fminf(a.x, b.x), fminf(c.x, wm))));
    const uint32_t y0 = to_u32(fmaxf(0.f, fminf(fminf(a.y, b.y), fminf(c.y, hm))));
    const uint32_t x1 = to_u32(fminf(wm, fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, 0.f))));
    const uint32_t y1 = to_u32(fminf(hm, fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, 0.f))));
    const uint32_t ax = to_u32(floor((double)a.x)), ay = to_u32(floor((double)a.y));
    if (0.f <= a.x && a.x < (float)W && 0.f <= a.y && a.y < (float)H) out.push_back({ax + ay * W, tri});
    for (uint32_t x = x0; x <= x1; ++x) {
        for (uint32_t y = y0; y <= y1; ++y) {
            if (x == ax && y == ay) continue;
            bool take = edge_touches(eab, x, y) | edge_touches(ebc, x, y) | edge_touches(eca, x, y);
            if (!take) {  // pixel corner inside the triangle (:196-211)
                const float axx = (float)x - a.x, axy = (float)y - a.y;
                const float bxx = (float)x - b.x, bxy = (float)y - b.y;
                const float cxx = (float)x - c.x, cxy = (float)y - c.y;
                const float k0 = ab.x * axy - ab.y * axx;
                const float k1 = bc.x * bxy - bc.y * bxx;
                const float k2 = ca.x * cxy - ca.y * cxx;
                take = (0 <= k0 * k1) & (0 <= k1 * k2);
            }
            if (take) out.push_back({x + y * W, tri});
        }
    }
}

bool build_camera_lists(const oclr_camera* cam, uint32_t vertexCount, const float4* vertex, uint32_t triangleCount,
                        const int32_t* triIdx, oclr_camera_lists* out) {
    const uint32_t W = cam->width, H = cam->height;
    const size_t P = (size_t)W * H;
    CamProj c;
    c.eye = mk3(cam->eye);
    c.tl = mk3(cam->eyeToTopLeft);
    c.lr = mk3(cam->leftToRight);
    c.tb = mk3(cam->topToBottom);
    c.screenN = cross3(c.lr, c.tb);
    c.tlDotN = dot3(c.tl, c.screenN);
    c.psiSq = cam->pixelSizeInv * cam->pixelSizeInv;
    const int threads = worker_count(triangleCount);
    std::vector<std::vector<BinRef>> chunks(threads);
    parallel_chunks(triangleCount, threads, [&](int t, size_t b, size_t e) {
        std::vector<BinRef>& v = chunks[t];
        v.reserve((e - b) * 4);
        for (size_t i = b; i < e; ++i) {
            const int32_t* vi = triIdx + 4 * i;
            raster_triangle(W, H, project(c, vertex[vi[0]]), project(c, vertex[vi[1]]), project(c, vertex[vi[2]]), (uint32_t)i, v);
        }
    });
    (void)vertexCount;
    std::vector<uint32_t> startInc(P + 1), list;
    csr_from_chunks(P, chunks, startInc.data(), list);
    chunks.clear();
    chunks.shrink_to_fit();

    // Storage compression against the left / upper neighbour (:580-613), sequential by construction.
    uint32_t* start = (uint32_t*)malloc(sizeof(uint32_t) * (P ? P : 1));
    uint32_t* end = (uint32_t*)malloc(sizeof(uint32_t) * (P ? P : 1));
    if (!start || !end) {
        free(start);
        free(end);
        return false;
    }
    uint32_t saved = 0;
    for (size_t p = 0; p < P; ++p) {
        const uint32_t s0 = startInc[p], len = startInc[p + 1] - s0;
        const uint32_t s = s0 - saved;
        if (saved && len) memmove(&list[s], &list[s0], sizeof(uint32_t) * len);
        start[p] = s;
        end[p] = s + len;
        const uint32_t x = (uint32_t)(p % W), y = (uint32_t)(p / W);
        bool merged = false;
        if (0 < x) {
            const size_t q = p - 1;
            if (len == end[q] - start[q] && 0 == memcmp(&list[start[q]], &list[s], sizeof(uint32_t) * len)) {
                saved += len;
                start[p] = start[q];
                end[p] = end[q];
                merged = true;
            }
        }
        if (0 < y && !merged) {
            const size_t q = p - W;
            if (len == end[q] - start[q] && 0 == memcmp(&list[start[q]], &list[s], sizeof(uint32_t) * len)) {
                saved += len;
                start[p] = start[q];
                end[p] = end[q];
            }
        }
    }
    const size_t kept = list.size() - saved;
    uint32_t* outList = (uint32_t*)malloc(sizeof(uint32_t) * (kept ? kept : 1));
    if (!outList) {
        free(start);
        free(end);
        return false;
    }
    if (kept) memcpy(outList, list.data(), sizeof(uint32_t) * kept);
    out->start = start;
    out->end = end;
    out->list = outList;
    out->listSize = kept;
    out->pixelCount = (uint32_t)P;
    return true;
}

// ---- scene grid ----------------------------------------------------------------------------------------------------------
// BoxIntersectsTriangle / Cull: rt_clip.h (shared with the device builder)

bool build_scene_grid(int32_t n, uint32_t vertexCount, const float4* vertex, uint32_t triangleCount, const int32_t* triIdx,
                      oclr_scene_grid* out) {
    if (n < 1 || (n & (n - 1)) || n > 1024) return false;
    const size_t cells = (size_t)n * n * n;
    std::vector<float> planes(3 * (size_t)(n + 1), 0.f);
    // Split planes at vertex quantiles (:660-678): plane i = midpoint of the sorted values at index i*(V-1)/n and
    // its predecessor (unsigned 32-bit index arithmetic, as in the reference).
    if (0 < vertexCount) {
        std::vector<float> val(vertexCount);
        for (int w = 0; w < 3; ++w) {
            for (uint32_t v = 0; v < vertexCount; ++v) val[v] = w == 0 ? vertex[v].x : (w == 1 ? vertex[v].y : vertex[v].z);
            std::sort(val.begin(), val.end());
            for (int i = 0; i <= n; ++i) {
                const uint32_t idx = ((uint32_t)i * (vertexCount - 1u)) / (uint32_t)n;
                planes[(size_t)w * (n + 1) + i] = (0 < idx && idx < vertexCount) ? (val[idx] + val[idx - 1]) / 2.f : val[idx];
            }
        }
    }
    const float* px = planes.data();
    const float* py = px + (n + 1);
    const float* pz = py + (n + 1);
    const float* pl[3] = {px, py, pz};

    const int threads = worker_count(triangleCount);
    std::vector<std::vector<BinRef>> chunks(threads);
    parallel_chunks(triangleCount, threads, [&](int t, size_t b, size_t e) {
        std::vector<BinRef>& refs = chunks[t];
        std::vector<uint8_t> seen((cells + 7) / 8, 0);
        std::vector<uint32_t> queue;
        for (size_t i = b; i < e; ++i) {
            const int32_t* vi = triIdx + 4 * i;
            const float4 &A = vertex[vi[0]], &B = vertex[vi[1]], &C = vertex[vi[2]];
            int c0[3];
            box_address(n, px, py, pz, mk3(A.x, A.y, A.z), c0[0], c0[1], c0[2]);
            queue.clear();
            uint32_t id = (uint32_t)c0[0] + (uint32_t)n * c0[1] + (uint32_t)n * n * c0[2];
            seen[id >> 3] |= (uint8_t)(1u << (id & 7));
            queue.push_back(id);
            for (size_t head = 0; head < queue.size(); ++head) {  // flood fill over face neighbours (:460-497)
                id = queue[head];
                int cc[3] = {(int)(id % n), (int)((id / n) % n), (int)(id / ((uint32_t)n * n))};
                float lo[3], hi[3];
                for (int k = 0; k < 3; ++k) {
                    lo[k] = pl[k][cc[k]];
                    hi[k] = pl[k][cc[k] + 1];
                }
                for (int k = 0; k < 3; ++k) {
                    for (int d = -1; d <= 1; d += 2) {
                        const int nc = cc[k] + d;
                        if (nc < 0 || n <= nc) continue;
                        int q[3] = {cc[0], cc[1], cc[2]};
                        q[k] = nc;
                        const uint32_t nid = (uint32_t)q[0] + (uint32_t)n * q[1] + (uint32_t)n * n * q[2];
                        if (seen[nid >> 3] & (1u << (nid & 7))) continue;
                        const float slo = lo[k], shi = hi[k];
                        lo[k] = pl[k][nc];
                        hi[k] = pl[k][nc + 1];
                        if (box_hits_triangle(lo, hi, mk3(A.x, A.y, A.z), mk3(B.x, B.y, B.z), mk3(C.x, C.y, C.z))) {
                            seen[nid >> 3] |= (uint8_t)(1u << (nid & 7));
                            queue.push_back(nid);
                        }
                        lo[k] = slo;
                        hi[k] = shi;
                    }
                }
            }
            for (uint32_t cid : queue) {
                refs.push_back({cid, (uint32_t)i});
                seen[cid >> 3] = 0;  // every set bit of this byte belongs to this triangle's fill
            }
        }
    });

    uint32_t* start = (uint32_t*)malloc(sizeof(uint32_t) * (cells + 1));
    cl_float3* boxMin = (cl_float3*)malloc(sizeof(cl_float3) * (n + 1));
    if (!start || !boxMin) {
        free(start);
        free(boxMin);
        return false;
    }
    std::vector<uint32_t> list;
    csr_from_chunks(cells, chunks, start, list);
    uint32_t* outList = (uint32_t*)malloc(sizeof(uint32_t) * (list.empty() ? 1 : list.size()));
    if (!outList) {
        free(start);
        free(boxMin);
        return false;
    }
    if (!list.empty()) memcpy(outList, list.data(), sizeof(uint32_t) * list.size());
    for (int i = 0; i <= n; ++i) {
        boxMin[i].s[0] = px[i];
        boxMin[i].s[1] = py[i];
        boxMin[i].s[2] = pz[i];
        boxMin[i].s[3] = 0.f;
    }
    out->axesDivCount = n;
    out->boxMin = boxMin;
    out->start = start;
    out->list = outList;
    out->listSize = list.size();
    return true;
}

}  // namespace oclr
