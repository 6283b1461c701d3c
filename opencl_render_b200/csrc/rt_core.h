// rt_core.h -- the arithmetic of the raytrace path, shared by every CUDA kernel of this library.
//
// Everything here is `__host__ __device__` so that tests can compile the very same expressions with g++
// (-ffp-contract=off) and compare them bit-for-bit with the reference on a machine without a GPU; the product
// library only ever instantiates it inside CUDA kernels (compiled with -fmad=false, IEEE div/sqrt).
//
// Bit-exactness contract (SURVEY.md section 7 "Hard parts"): same fp32 operation order as the reference, no FMA
// contraction, dot = ((a0*b0 + a1*b1) + a2*b2) (source/opencl/raytrace.c:18-20), the libm calls the C path
// evaluates in double stay in double (raytrace_opencl.c:22, 27, 251-253, 594, 631).
#pragma once
#include <math.h>
#include <stdint.h>

#include "rt_types.h"

namespace oclr {

struct f3 {
    float x, y, z;
};

#if defined(__CUDA_ARCH__)
#define OCLR_LDG(p) __ldg(p)
#define OCLR_POPCLL(v) __popcll(v)
#define OCLR_INF __int_as_float(0x7f800000)
#else
#define OCLR_LDG(p) (*(p))
#define OCLR_POPCLL(v) __builtin_popcountll(v)
#define OCLR_INF __builtin_inff()
#endif

OCLR_HD f3 mk3(float x, float y, float z) {
    f3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
OCLR_HD f3 mk3(const float* p) { return mk3(p[0], p[1], p[2]); }
// raytrace.c:18-20
OCLR_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// raytrace.c:21-27
OCLR_HD f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// (float)sqrt((double)x): double sqrt then narrowing == correctly rounded single sqrt
OCLR_HD float sqrt_c(float v) { return sqrtf(v); }

// ---- RNG: raytrace_opencl.c:1-23 -----------------------------------------------------------------------------
OCLR_HD uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }
OCLR_HD uint64_t xorshift64star(uint64_t v) {
    v ^= v >> 12;
    v ^= v << 25;
    v ^= v >> 27;
    return v * 2685821657736338717ull;
}
OCLR_HD_OUT(1) float rand_f(uint64_t& s, float lo, float hi) {
    s ^= xorshift64star((rotl64(s, 55) ^ rotl64(s, 3)) * 0xc23f3c0ad9da6357ull);
    s ^= xorshift64star((rotl64(s, 35) ^ rotl64(s, 3)) ^ 0xce84d6af03c16b89ull);
    s ^= xorshift64star((rotl64(s, 63) ^ rotl64(s, 35)) * 0xf097ef8bbe03ddccull);
    s ^= xorshift64star((rotl64(s, 41) ^ rotl64(s, 12)) ^ 0x48302294fbfe30bfull);
    s ^= xorshift64star((rotl64(s, 1) ^ rotl64(s, 62)) * 0x79e7425e3f4f147dull);
    s ^= xorshift64star((rotl64(s, 42) ^ rotl64(s, 29)) ^ 0x14d1d30856e5be9aull);
    s ^= xorshift64star((rotl64(s, 47) ^ rotl64(s, 45)) * 0x24289d47a66617c3ull);
    s ^= xorshift64star((rotl64(s, 39) ^ rotl64(s, 6)) ^ 0x5576fb2f80a05d14ull);
    // (double)state / (double)0xffffffffffffffff: the divisor rounds to 2^64, so the quotient is an exact scaling
    return lo + (hi - lo) * (float)((double)s * (1.0 / 18446744073709551616.0));
}

// raytrace_opencl.c:30-45.  The number of draws is state-visible (rejection loop).
OCLR_HD_OUT(2) f3 sphere_point(uint64_t& s, float radius) {
    f3 p;
    float len;
    do {
        p.x = rand_f(s, -1.f, 1.f);
        p.y = rand_f(s, -1.f, 1.f);
        p.z = rand_f(s, -1.f, 1.f);
        len = sqrt_c(dot3(p, p));
    } while (len <= 0.f);
    float tmp = sqrt_c(rand_f(s, 0.f, 1.f)) * radius / len;
    return mk3(tmp * p.x, tmp * p.y, tmp * p.z);
}

// raytrace_opencl.c:25-28 (double modf twice)
OCLR_HD float positive_modf(float v) {
    double ip;
    return (float)modf(modf((double)v, &ip) + 1., &ip);
}

// raytrace_opencl.c:83-101
OCLR_HD float point_to_line_sq(f3 o, f3 d, f3 p) {
    f3 od = mk3(d.x - o.x, d.y - o.y, d.z - o.z);
    float odSq = dot3(od, od);
    f3 op = mk3(p.x - o.x, p.y - o.y, p.z - o.z);
    float k = dot3(op, od) / odSq;
    f3 proj = mk3(o.x + k * od.x, o.y + k * od.y, o.z + k * od.z);
    f3 dv = mk3(proj.x - p.x, proj.y - p.y, proj.z - p.z);
    return dot3(dv, dv);
}

// raytrace_opencl.c:103-122.  `size.x - 1` is unsigned arithmetic in the reference.  In two halves: the wrapped texture coordinate
// depends on the triangle's UVs and the hit's barycentrics only -- one value for all the channels fetched at a hit (the reference
// recomputes it, two double-precision modf each, for every channel) --, the fetch on the channel's image.
OCLR_HD void table_uv(float u0, float v0, float u1, float v1, float u2, float v2, float abL, float acL, float& pu, float& pv) {
    pu = positive_modf(u0 + (u1 - u0) * abL + (u2 - u0) * acL);
    pv = positive_modf(v0 + (v1 - v0) * abL + (v2 - v0) * acL);
}
OCLR_HD f3 table_fetch(const uchar4* table, uint2 size, float pu, float pv) {
    float lx = pu * (float)(size.x - 1u);
    float ly = pv * (float)(size.y - 1u);
    int ix = (int)floorf(lx);
    int iy = (int)floorf(ly);
    int idx = (int)((uint32_t)ix + (uint32_t)iy * size.x);
    uchar4 t = OCLR_LDG(table + idx);
    return mk3((float)t.x / 255.f, (float)t.y / 255.f, (float)t.z / 255.f);
}
OCLR_HD_OUT(3) f3 table_value(const uchar4* table, uint2 size, float u0, float v0, float u1, float v1, float u2, float v2,
                       float abL, float acL) {
    float pu, pv;
    table_uv(u0, v0, u1, v1, u2, v2, abL, acL, pu, pv);
    return table_fetch(table, size, pu, pv);
}

// ---- ray / triangle: raytrace_opencl.c:124-172 against the packed triGeo record --------------------------------
// Stage 1 reads 32 B (q0,q1), stage 2 another 32 B (q2,q3) only when the plane distance is inside (minD, maxD).
OCLR_HD bool tri_test(const float4* g, f3 o, f3 r, float minD, float maxD, float& t, float& abL, float& acL) {
    const float4 q0 = OCLR_LDG(g + 0);
    const float4 q1 = OCLR_LDG(g + 1);
    const f3 n = mk3(q0.x, q0.y, q0.z);
    const f3 a = mk3(q0.w, q1.x, q1.y);
    const f3 ao = mk3(o.x - a.x, o.y - a.y, o.z - a.z);
    t = -dot3(n, ao) / dot3(n, r);
    if (minD < t && t < maxD) {
        const float4 q2 = OCLR_LDG(g + 2);
        const float4 q3 = OCLR_LDG(g + 3);
        const float abab = q1.z, abac = q1.w, acac = q2.w, D = q3.w;
        const f3 ab = mk3(q2.x, q2.y, q2.z);
        const f3 ac = mk3(q3.x, q3.y, q3.z);
        const f3 proj = mk3(o.x + t * r.x, o.y + t * r.y, o.z + t * r.z);
        const f3 ap = mk3(proj.x - a.x, proj.y - a.y, proj.z - a.z);
        const float apab = dot3(ap, ab);
        const float apac = dot3(ap, ac);
        abL = (abac * apac - acac * apab) * D;
        acL = (abac * apab - abab * apac) * D;
        return (0.f <= abL && 0.f <= acL && abL + acL <= 1.f);
    }
    return false;
}

// ---- grid addressing: raytrace_opencl.c:174-193 (binary search, strict <) ---------------------------------------
OCLR_HD void box_address(int n, const float* px, const float* py, const float* pz, f3 p, int& cx, int& cy, int& cz) {
    cx = 0;
    cy = 0;
    cz = 0;
    while (1 < n) {
        n /= 2;
        if (px[cx + n] < p.x) cx += n;
        if (py[cy + n] < p.y) cy += n;
        if (pz[cz + n] < p.z) cz += n;
    }
}

// raytrace_opencl.c:265-322.  The caller ignores the return value, so partial moves are kept (:354, :360).
OCLR_HD void bind_in_cube(f3& p, f3 r, f3 lo, f3 hi) {
    float t;
    if (p.x < lo.x) {
        if (r.x <= 0) return;
        t = (lo.x - p.x) / r.x;
        p.x += t * r.x; p.y += t * r.y; p.z += t * r.z;
    }
    if (hi.x < p.x) {
        if (0 <= r.x) return;
        t = (hi.x - p.x) / r.x;
        p.x += t * r.x; p.y += t * r.y; p.z += t * r.z;
    }
    if (p.y < lo.y) {
        if (r.y <= 0) return;
        t = (lo.y - p.y) / r.y;
        p.x += t * r.x; p.y += t * r.y; p.z += t * r.z;
    }
    if (hi.y < p.y) {
        if (0 <= r.y) return;
        t = (hi.y - p.y) / r.y;
        p.x += t * r.x; p.y += t * r.y; p.z += t * r.z;
    }
    if (p.z < lo.z) {
        if (r.z <= 0) return;
        t = (lo.z - p.z) / r.z;
        p.x += t * r.x; p.y += t * r.y; p.z += t * r.z;
    }
    if (hi.z < p.z) {
        if (0 <= r.z) return;
        t = (hi.z - p.z) / r.z;
        p.x += t * r.x; p.y += t * r.y; p.z += t * r.z;
    }
}

// ---- grid walk state: raytrace_opencl.c:324-401 -------------------------------------------------------------------
// The reference recomputes all three plane distances every cell (:383-385); each depends only on its own axis'
// cell index, so only the axis that stepped is re-divided here -- identical values, one IEEE division per step.
struct GridWalk {
    f3 o, r;
    float minD, maxD;
    uint32_t excl;
    int cx, cy, cz;
    int ex, ey, ez;
    float tx, ty, tz;
    int curBrick;
    uint64_t mask;
    uint32_t rankBase;
    // Two-level walk (see walk_enter_coarse): shift == 0 -> (cx,cy,cz) are cells and (tx,ty,tz) the next cell-plane
    // crossings; shift == 2 -> (cx,cy,cz) are 4x4x4-brick coordinates and (tx,ty,tz) the next BRICK-plane crossings.
    int shift;
    int endBrick;   // linear id of the brick holding the end cell (never skipped), or -1
    bool coarseOk;  // all three direction components non-zero (no NaN crossings possible)
};

OCLR_HD void walk_begin(GridWalk& w, const SceneView& S, const float* px, const float* py, const float* pz, f3 o, f3 r,
                        float minD, float maxD, uint32_t excl) {
    const int n = S.n;
    w.o = o;
    w.r = r;
    w.minD = minD;
    w.maxD = maxD;
    w.excl = excl;
    const f3 lo = mk3(px[0], py[0], pz[0]);
    const f3 hi = mk3(px[n], py[n], pz[n]);
    f3 start = mk3(o.x + minD * r.x, o.y + minD * r.y, o.z + minD * r.z);
    bind_in_cube(start, r, lo, hi);
    box_address(n, px, py, pz, start, w.cx, w.cy, w.cz);
    w.ex = w.ey = w.ez = -1;
    if (maxD < OCLR_INF) {
        f3 end = mk3(o.x + maxD * r.x, o.y + maxD * r.y, o.z + maxD * r.z);
        bind_in_cube(end, r, lo, hi);
        box_address(n, px, py, pz, end, w.ex, w.ey, w.ez);
    }
    w.tx = (px[w.cx + (0 <= r.x)] - o.x) / r.x;
    w.ty = (py[w.cy + (0 <= r.y)] - o.y) / r.y;
    w.tz = (pz[w.cz + (0 <= r.z)] - o.z) / r.z;
    w.curBrick = -1;
    w.mask = 0;
    w.rankBase = 0;
    w.shift = 0;
    w.endBrick = (w.ex < 0) ? -1 : (w.ex >> 2) + S.nb * ((w.ey >> 2) + S.nb * (w.ez >> 2));
    w.coarseOk = (n >= 4) & (r.x != 0.f) & (r.y != 0.f) & (r.z != 0.f);
}

// Loads the 16-byte brick record of the current position (cell or brick level) when the walk entered a new brick.
template <bool COUNT>
OCLR_HD void walk_load_brick(GridWalk& w, const SceneView& S, Counters* cnt) {
    const int sh = 2 - w.shift;
    const int b = (w.cx >> sh) + S.nb * ((w.cy >> sh) + S.nb * (w.cz >> sh));
    if (b != w.curBrick) {
        const uint4 br = OCLR_LDG(S.bricks + b);
        w.mask = (uint64_t)br.x | ((uint64_t)br.y << 32);
        w.rankBase = br.z;
        w.curBrick = b;
        if (COUNT) cnt->bricksLoaded++;
    }
}

// Occupancy of the current cell; loads the brick record when the walk entered a new brick.
// Returns true and the list range when the cell has triangles.
template <bool COUNT>
OCLR_HD bool walk_cell(GridWalk& w, const SceneView& S, uint2& range, Counters* cnt) {
    walk_load_brick<COUNT>(w, S, cnt);
    const int bit = (w.cx & 3) | ((w.cy & 3) << 2) | ((w.cz & 3) << 4);
    if (COUNT) cnt->cells++;
    if (COUNT && w.mask == 0ull) cnt->emptyBrickCells++;
    if ((w.mask >> bit) & 1ull) {
        const uint32_t rank = w.rankBase + (uint32_t)OCLR_POPCLL(w.mask & ((1ull << bit) - 1ull));
        range = OCLR_LDG(S.cellRange + rank);
        if (COUNT) cnt->cellsNonEmpty++;
        return true;
    }
    return false;
}

// Advance to the next cell (:383-398).  Returns false when the walk left the grid.
// Branch-free form of the reference's three-way if/else: the axis is selected first (same comparisons, same tie rule:
// x only if strictly smallest, else y if strictly smaller than z, else z -- NaNs fall through to z), then ONE step, one
// shared-memory plane fetch and one IEEE division run for whichever axis was chosen, so the lanes of a warp stay converged.
// At brick level (w.shift == 2) the very same code steps 4 planes at a time.  `axis`/`tEvent` report the crossing taken.
OCLR_HD bool walk_step_ex(GridWalk& w, int n, const float* px, const float* py, const float* pz, int& axis, float& tEvent) {
    const bool xmin = (w.tx < w.ty) & (w.tx < w.tz);
    const bool ymin = (!xmin) & (w.ty < w.tz);
    axis = xmin ? 0 : (ymin ? 1 : 2);
    tEvent = xmin ? w.tx : (ymin ? w.ty : w.tz);
    int c = xmin ? w.cx : (ymin ? w.cy : w.cz);
    const float rr = xmin ? w.r.x : (ymin ? w.r.y : w.r.z);
    const float oo = xmin ? w.o.x : (ymin ? w.o.y : w.o.z);
    const float* p = xmin ? px : (ymin ? py : pz);
    const int up = (0 <= rr);
    c += up ? 1 : -1;
    if (c < 0 || (n >> w.shift) <= c) return false;
    const float t = (p[(c + up) << w.shift] - oo) / rr;
    w.cx = xmin ? c : w.cx;
    w.tx = xmin ? t : w.tx;
    w.cy = ymin ? c : w.cy;
    w.ty = ymin ? t : w.ty;
    w.cz = (xmin | ymin) ? w.cz : c;
    w.tz = (xmin | ymin) ? w.tz : t;
    return true;
}
OCLR_HD bool walk_step(GridWalk& w, int n, const float* px, const float* py, const float* pz) {
    int axis;
    float tEvent;
    return walk_step_ex(w, n, px, py, pz, axis, tEvent);
}

// ---- two-level walk: exact skipping of empty 4x4x4 bricks ----------------------------------------------------------------
// The reference's DDA is a 3-way merge of the per-axis plane-crossing sequences t_a(k) = (plane_a[k] - o_a) / r_a, each
// non-decreasing along the ray, merged by "smallest first, ties to the later axis" (:387-398).  The crossings of every 4th
// plane (brick faces) are a sub-sequence of each, so merging only those -- the same comparisons on the same fp32
// quotients -- visits the bricks in exactly the order the cell walk passes through them.  While the bricks are empty
// nothing else is needed; on entering a brick that has triangles (or holds the end cell) the exact cell-level state is
// rebuilt: along the entry axis it is the brick's first cell; along each other axis it is the number of that axis' cell
// crossings inside the brick that the merge would have taken before the entry crossing E ("t < E" for an axis that wins
// ties against the entry axis' predecessor rule, "t <= E" otherwise), found with two probes of the three interior planes.
// Requires all direction components non-zero (then no crossing is NaN and the merge order is a total preorder).

// From a cell inside an empty brick: switch to brick coordinates; the heads become the brick-exit crossings.
OCLR_HD void walk_enter_coarse(GridWalk& w, const float* px, const float* py, const float* pz) {
    w.cx >>= 2;
    w.cy >>= 2;
    w.cz >>= 2;
    w.tx = (px[(w.cx + (0 <= w.r.x)) << 2] - w.o.x) / w.r.x;
    w.ty = (py[(w.cy + (0 <= w.r.y)) << 2] - w.o.y) / w.r.y;
    w.tz = (pz[(w.cz + (0 <= w.r.z)) << 2] - w.o.z) / w.r.z;
    w.shift = 2;
}

// One non-entry axis of the refinement: brick coordinate b, current brick-exit crossing tExit (not yet taken).
// `strict`: this axis' crossings precede E only when strictly smaller (axis index below the entry axis).
OCLR_HD void refine_axis(int b, float tExit, float o, float r, const float* p, float E, bool strict, int& cell, float& tNext) {
    const int up = (0 <= r);
    const int base = b << 2;
    // cells in travel order u_j = up ? base + j : base + 3 - j; crossing j leaves u_j through plane u_j + up
    const int p1 = up ? base + 2 : base + 2;  // plane left by u_1: up -> base+1+1, down -> base+2+0
    const float t1 = (p[p1] - o) / r;
    const bool pre1 = strict ? (t1 < E) : (t1 <= E);
    const int p2 = pre1 ? (up ? base + 3 : base + 1) : (up ? base + 1 : base + 3);  // crossing 2 or crossing 0
    const float t2 = (p[p2] - o) / r;
    const bool pre2 = strict ? (t2 < E) : (t2 <= E);
    int j;
    if (pre1) {
        j = pre2 ? 3 : 2;
        tNext = pre2 ? tExit : t2;
    } else {
        j = pre2 ? 1 : 0;
        tNext = pre2 ? t1 : t2;
    }
    cell = up ? base + j : base + 3 - j;
}

// After a brick-level step along `axis` (crossing value E) into a brick that must be walked cell by cell.
OCLR_HD void walk_refine(GridWalk& w, int axis, float E, const float* px, const float* py, const float* pz) {
    int cx, cy, cz;
    float tx, ty, tz;
    if (axis == 0) {
        const int up = (0 <= w.r.x);
        cx = up ? (w.cx << 2) : (w.cx << 2) + 3;
        tx = (px[cx + up] - w.o.x) / w.r.x;
    } else {
        refine_axis(w.cx, w.tx, w.o.x, w.r.x, px, E, true, cx, tx);  // x precedes y/z crossings only when strictly smaller
    }
    if (axis == 1) {
        const int up = (0 <= w.r.y);
        cy = up ? (w.cy << 2) : (w.cy << 2) + 3;
        ty = (py[cy + up] - w.o.y) / w.r.y;
    } else {
        refine_axis(w.cy, w.ty, w.o.y, w.r.y, py, E, axis == 2, cy, ty);  // y: "<= E" against x, "< E" against z
    }
    if (axis == 2) {
        const int up = (0 <= w.r.z);
        cz = up ? (w.cz << 2) : (w.cz << 2) + 3;
        tz = (pz[cz + up] - w.o.z) / w.r.z;
    } else {
        refine_axis(w.cz, w.tz, w.o.z, w.r.z, pz, E, false, cz, tz);  // z wins ties against x and y
    }
    w.cx = cx; w.cy = cy; w.cz = cz;
    w.tx = tx; w.ty = ty; w.tz = tz;
    w.shift = 0;
}

// Whole traversal for one ray, serial form (one thread walks one ray).  `outT` is reset to maxD in every cell
// (:366); the walk stops at the first cell that produced any hit (:380), at the end cell (:381) or off the grid.
template <bool COUNT>
OCLR_HD uint32_t grid_trace(const SceneView& S, const float* px, const float* py, const float* pz, f3 o, f3 r,
                            float minD, float maxD, uint32_t excl, float& outT, float& outAB, float& outAC,
                            Counters* cnt, bool hierarchical = false) {
    GridWalk w;
    walk_begin(w, S, px, py, pz, o, r, minD, maxD, excl);
    if (COUNT) cnt->gridRays++;
    uint32_t closest = kNoTriangle;
    for (;;) {
        uint2 range;
        outT = maxD;
        const bool occupied = walk_cell<COUNT>(w, S, range, cnt);
        if (occupied) {
            for (uint32_t i = range.x; i < range.y; ++i) {
                const uint32_t tri = OCLR_LDG(S.cellList + i);
                if (tri != excl) {
                    float t, ab, ac;
                    if (COUNT) cnt->gridCandidates++;
                    if (tri_test(S.triGeo + 4 * (size_t)tri, o, r, minD, outT, t, ab, ac)) {
                        closest = tri;
                        outT = t;
                        outAB = ab;
                        outAC = ac;
                    }
                }
            }
        }
        if (closest != kNoTriangle || (w.cx == w.ex && w.cy == w.ey && w.cz == w.ez)) break;
        if (hierarchical && w.coarseOk && w.mask == 0ull && w.curBrick != w.endBrick) {
            // the whole brick is empty: cross it, and any empty bricks behind it, at brick granularity
            walk_enter_coarse(w, px, py, pz);
            bool inside = true;
            for (;;) {
                int axis;
                float E;
                if (!walk_step_ex(w, S.n, px, py, pz, axis, E)) {
                    inside = false;
                    break;
                }
                walk_load_brick<COUNT>(w, S, cnt);
                if (w.mask != 0ull || w.curBrick == w.endBrick) {
                    walk_refine(w, axis, E, px, py, pz);
                    break;
                }
            }
            if (!inside) break;
            continue;
        }
        if (!walk_step(w, S.n, px, py, pz)) break;
    }
    return closest;
}

// ---- shading pieces ---------------------------------------------------------------------------------------------
struct TriShade {
    f3 a, b, c, nA, nB, nC;
    float u0, v0, u1, v1, u2, v2;
    int mat;
};

OCLR_HD TriShade load_shade(const SceneView& S, uint32_t tri) {
    const float4* p = S.triShade + 8 * (size_t)tri;
    const float4 s0 = OCLR_LDG(p + 0), s1 = OCLR_LDG(p + 1), s2 = OCLR_LDG(p + 2), s3 = OCLR_LDG(p + 3);
    const float4 s4 = OCLR_LDG(p + 4), s5 = OCLR_LDG(p + 5), s6 = OCLR_LDG(p + 6), s7 = OCLR_LDG(p + 7);
    TriShade t;
    t.a = mk3(s0.x, s0.y, s0.z);
#if defined(__CUDA_ARCH__)
    t.mat = __float_as_int(s0.w);
#else
    union { float f; int i; } cv;
    cv.f = s0.w;
    t.mat = cv.i;
#endif
    t.b = mk3(s1.x, s1.y, s1.z);
    t.c = mk3(s2.x, s2.y, s2.z);
    t.nA = mk3(s3.x, s3.y, s3.z);
    t.nB = mk3(s4.x, s4.y, s4.z);
    t.nC = mk3(s5.x, s5.y, s5.z);
    t.u0 = s6.x; t.v0 = s6.y; t.u1 = s6.z; t.v1 = s6.w;
    t.u2 = s7.x; t.v2 = s7.y;
    return t;
}

// Material id + UVs only (transparent-occluder lookups in the shadow loop, :613-619)
OCLR_HD void load_mat_uv(const SceneView& S, uint32_t tri, int& mat, float& u0, float& v0, float& u1, float& v1,
                         float& u2, float& v2) {
    const float4* p = S.triShade + 8 * (size_t)tri;
    const float4 s0 = OCLR_LDG(p + 0), s6 = OCLR_LDG(p + 6), s7 = OCLR_LDG(p + 7);
#if defined(__CUDA_ARCH__)
    mat = __float_as_int(s0.w);
#else
    union { float f; int i; } cv;
    cv.f = s0.w;
    mat = cv.i;
#endif
    u0 = s6.x; v0 = s6.y; u1 = s6.z; v1 = s6.w;
    u2 = s7.x; v2 = s7.y;
}

OCLR_HD bool channel_present(const SceneView& S, int mat, int channel, uint2& size) {
    if (mat < 0 || (uint32_t)mat >= S.materialCount) return false;  // reference reads OOB for mat < 0 (:226 vs :231); guarded
    size = OCLR_LDG(S.matSize + kMaterialChannels * mat + channel);
    return 0u < size.x;
}

OCLR_HD f3 channel_value(const SceneView& S, int mat, int channel, uint2 size, const TriShade& ts, float abL, float acL) {
    const int start = OCLR_LDG(S.matStart + kMaterialChannels * mat + channel);
    return table_value(S.textures + start, size, ts.u0, ts.v0, ts.u1, ts.v1, ts.u2, ts.v2, abL, acL);
}

// The four material channels read at every shaded hit (:541-561; absent channels leave their argument untouched), with the
// texture coordinate computed once.
OCLR_HD void shade_channels(const SceneView& S, const TriShade& ts, float abL, float acL, f3& tex, f3& transp, f3& refl, f3& lum) {
    float pu, pv;
    table_uv(ts.u0, ts.v0, ts.u1, ts.v1, ts.u2, ts.v2, abL, acL, pu, pv);
    uint2 sz;
    if (channel_present(S, ts.mat, kChColor, sz)) tex = table_fetch(S.textures + OCLR_LDG(S.matStart + kMaterialChannels * ts.mat + kChColor), sz, pu, pv);
    if (channel_present(S, ts.mat, kChTransparency, sz))
        transp = table_fetch(S.textures + OCLR_LDG(S.matStart + kMaterialChannels * ts.mat + kChTransparency), sz, pu, pv);
    if (channel_present(S, ts.mat, kChReflection, sz))
        refl = table_fetch(S.textures + OCLR_LDG(S.matStart + kMaterialChannels * ts.mat + kChReflection), sz, pu, pv);
    if (channel_present(S, ts.mat, kChLuminance, sz))
        lum = table_fetch(S.textures + OCLR_LDG(S.matStart + kMaterialChannels * ts.mat + kChLuminance), sz, pu, pv);
}

// raytrace_opencl.c:124-172 on raw vertices (bump-mapping helper rays only, :244, :249): result flag ignored there,
// abL/acL keep their previous value when the plane distance is outside (0, inf).
OCLR_HD bool tri_bary_raw(f3 o, f3 r, f3 a, f3 b, f3 c, float& abL, float& acL) {
    const f3 ab = mk3(b.x - a.x, b.y - a.y, b.z - a.z);
    const f3 ac = mk3(c.x - a.x, c.y - a.y, c.z - a.z);
    const f3 ao = mk3(o.x - a.x, o.y - a.y, o.z - a.z);
    const f3 n = cross3(ac, ab);
    const float t = -dot3(n, ao) / dot3(n, r);
    if (0.f < t && t < OCLR_INF) {
        const float abab = dot3(ab, ab), abac = dot3(ab, ac), acac = dot3(ac, ac);
        const float D = 1.f / (abac * abac - abab * acac);
        const f3 proj = mk3(o.x + t * r.x, o.y + t * r.y, o.z + t * r.z);
        const f3 ap = mk3(proj.x - a.x, proj.y - a.y, proj.z - a.z);
        const float apab = dot3(ap, ab), apac = dot3(ap, ac);
        abL = (abac * apac - acac * apab) * D;
        acL = (abac * apab - abab * apac) * D;
        return true;
    }
    return false;
}

// raytrace_opencl.c:195-263
// `undefined` is set when the reference would read uninitialised variables: the first bump helper ray (:244) does not
// meet the triangle's plane at a positive distance (happens for bounce rays, whose direction is shorter than a pixel
// step), so abL/acL of :245 are stack garbage there.  This implementation then uses the hit's own abL/acL.
OCLR_HD f3 triangle_normal(const SceneView& S, const Camera& cam, const TriShade& ts, f3 loc, f3 rayO, f3 rayV, float abL,
                           float acL, bool& undefined) {
    const float dab = sqrt_c(point_to_line_sq(ts.a, ts.b, loc));
    const float dbc = sqrt_c(point_to_line_sq(ts.b, ts.c, loc));
    const float dca = sqrt_c(point_to_line_sq(ts.c, ts.a, loc));
    const float inv = 1.f / (dab + dbc + dca);
    f3 nrm = mk3((dab * ts.nC.x + dbc * ts.nA.x + dca * ts.nB.x) * inv, (dab * ts.nC.y + dbc * ts.nA.y + dca * ts.nB.y) * inv,
                 (dab * ts.nC.z + dbc * ts.nA.z + dca * ts.nB.z) * inv);
    uint2 bumpSize;
    if (channel_present(S, ts.mat, kChBump, bumpSize)) {
        const f3 tb = mk3(cam.topToBottom), lr = mk3(cam.leftToRight);
        const f3 h = channel_value(S, ts.mat, kChBump, bumpSize, ts, abL, acL);
        float bL = abL, cL = acL;  // uninitialised in the reference when the first helper ray misses the plane
        if (!tri_bary_raw(rayO, mk3(rayV.x + tb.x, rayV.y + tb.y, rayV.z + tb.z), ts.a, ts.b, ts.c, bL, cL)) undefined = true;
        const f3 hS = channel_value(S, ts.mat, kChBump, bumpSize, ts, bL, cL);
        tri_bary_raw(rayO, mk3(rayV.x + lr.x, rayV.y + lr.y, rayV.z + lr.z), ts.a, ts.b, ts.c, bL, cL);
        const f3 hE = channel_value(S, ts.mat, kChBump, bumpSize, ts, bL, cL);
        const float kPi = 3.14159265f;  // raytrace.h:33 (float macro in the C path)
        const float aE = (hE.x - h.x) * kPi / 2.f;
        const float aS = (hS.x - h.x) * kPi / 2.f;
        const float xPart = (float)sin((double)aE);
        const float yPart = (float)sin((double)aS);
        const float nPart = (float)cos((double)aE) * (float)cos((double)aS);
        nrm.x = nPart * nrm.x / cam.pixelSizeInv + xPart * lr.x + yPart * tb.x;
        nrm.y = nPart * nrm.y / cam.pixelSizeInv + xPart * lr.y + yPart * tb.y;
        nrm.z = nPart * nrm.z / cam.pixelSizeInv + xPart * lr.z + yPart * tb.z;
        const float li = 1.f / sqrt_c(dot3(nrm, nrm));
        nrm.x *= li;
        nrm.y *= li;
        nrm.z *= li;
    }
    return nrm;
}

// (int)f as the x86-64 C path evaluates it (cvttss2si: out-of-range and NaN give INT_MIN), raytrace_opencl.c:729
OCLR_HD int float_to_int_x86(float f) {
    if (f >= -2147483648.f && f < 2147483648.f) return (int)f;
    return (int)0x80000000;
}

// 16-bit accumulate with the reference's per-sample truncation and clamp (:726-741)
OCLR_HD uint16_t accumulate16(uint16_t prev, float c, float scale) {
    int v = (int)prev + float_to_int_x86(c * scale);
    if (v < 0) v = 0;
    if (0xFFFF < v) v = 0xFFFF;
    return (uint16_t)v;
}

// One light's sample point and shadow-ray interval (:564-607).  Returns false for a light type that is not in the
// reference's switch (leaves a zero vector and an empty interval, like the omni case :585-588).
struct LightRay {
    f3 dir;
    float minLen, maxLen;
};
OCLR_HD void light_ray(const Light& L, f3 loc, uint64_t& rng, LightRay& lr) {
    lr.dir = mk3(0.f, 0.f, 0.f);
    lr.minLen = 0.f;
    lr.maxLen = 0.f;
    switch (L.type) {
        case 1: case 2: case 7: case 8: case 9: {  // SPOT, SPOTRECT, TUBE, AREA, PHOTOMETRIC
            const f3 rl = sphere_point(rng, L.radius);
            lr.dir = mk3(rl.x + L.pos[0] - loc.x, rl.y + L.pos[1] - loc.y, rl.z + L.pos[2] - loc.z);
            lr.maxLen = sqrt_c(dot3(lr.dir, lr.dir));
            const float inv = 1.f / lr.maxLen;
            lr.dir.x *= inv;
            lr.dir.y *= inv;
            lr.dir.z *= inv;
        } break;
        case 3: case 4: case 5: case 6: {  // DISTANT, PARALLEL, PARSPOT, PARSPOTRECT
            f3 d = sphere_point(rng, L.distantRadius);
            d.x -= L.dir[0];
            d.y -= L.dir[1];
            d.z -= L.dir[2];
            const float inv = 1.f / sqrt_c(dot3(d, d));
            lr.dir = mk3(d.x * inv, d.y * inv, d.z * inv);
            lr.maxLen = OCLR_INF;
        } break;
        default:  // OMNI (0) and unknown types contribute nothing but still take the ambient path
            break;
    }
}

// :629-635
OCLR_HD_OUT(2) void light_accumulate(const Light& L, f3 nrm, const LightRay& lr, f3 att, f3 face[2]) {
    const float d = dot3(nrm, lr.dir);
    const float effect = fabsf(d);
    const int idx = (int)(0.f <= d);
    const float x = lr.maxLen / L.halfDistance;
    const float a = (x == 0.f) ? 1.f : (float)pow(0.5, (double)x);  // pow(0.5f, 0) == 1 exactly
    const float e = effect * (a == a ? a : 1.f);
    face[idx].x += (1.f - face[idx].x) * att.x * e * L.colour[0];
    face[idx].y += (1.f - face[idx].y) * att.y * e * L.colour[1];
    face[idx].z += (1.f - face[idx].z) * att.z * e * L.colour[2];
}

// rt_walk.h: the walk in the trace kernel's packed formulation (walkMode 2 of trace_sample exercises it on the host)
template <bool COUNT>
OCLR_HD uint32_t grid_trace_packed(const SceneView& S, const float* planes, f3 o, f3 r, float minD, float maxD, uint32_t excl,
                                   float& outT, float& outAB, float& outAC, Counters* cnt);
template <bool COUNT>
OCLR_HD uint32_t grid_trace_split_mid(const SceneView& S, const float* planes, f3 o, f3 r, float minD, float maxD, uint32_t excl, float& outT,
                                      float& outAB, float& outAC, Counters* cnt, int walkFirst, int maxParts, int minPartCells);

template <bool COUNT>
OCLR_HD uint32_t grid_trace_split(const SceneView& S, const float* planes, f3 o, f3 r, float minD, float maxD, uint32_t excl, float& outT,
                                  float& outAB, float& outAC, Counters* cnt, int partCells);

template <bool COUNT>
OCLR_HD uint32_t grid_trace_coop(const SceneView& S, const float* planes, f3 o, f3 r, float minD, float maxD, uint32_t excl, float& outT,
                                 float& outAB, float& outAC, Counters* cnt);

template <bool COUNT>
OCLR_HD uint32_t grid_trace_coop_bricks(const SceneView& S, const float* planes, f3 o, f3 r, float minD, float maxD, uint32_t excl, float& outT,
                                        float& outAB, float& outAC, Counters* cnt);

template <bool COUNT>
OCLR_HD uint32_t grid_trace_mode(int walkMode, const SceneView& S, const float* px, const float* py, const float* pz, f3 o, f3 r,
                                 float minD, float maxD, uint32_t excl, float& outT, float& outAB, float& outAC, Counters* cnt) {
    if (walkMode == 2) return grid_trace_packed<COUNT>(S, px, o, r, minD, maxD, excl, outT, outAB, outAC, cnt);  // px = base of all planes
    if (walkMode >= 2000)   // walkMode = 2000 + 100 * (cells walked before the cut) + parts: the trace kernel's run-time split (rt_walk.h)
        return grid_trace_split_mid<COUNT>(S, px, o, r, minD, maxD, excl, outT, outAB, outAC, cnt, (walkMode - 2000) / 100, (walkMode - 2000) % 100, 3);
    if (walkMode == 1000) return grid_trace_coop<COUNT>(S, px, o, r, minD, maxD, excl, outT, outAB, outAC, cnt);   // walked by cooperative bursts
#if !defined(__CUDA_ARCH__)   // (host form of rt_tail.cuh's kernel: test infrastructure, kept out of the per-pixel kernel's stack frame)
    if (walkMode == 1001) return grid_trace_coop_bricks<COUNT>(S, px, o, r, minD, maxD, excl, outT, outAB, outAC, cnt);   // ... over brick planes
#endif
    if (walkMode >= 3) return grid_trace_split<COUNT>(S, px, o, r, minD, maxD, excl, outT, outAB, outAC, cnt, walkMode);  // parts of `walkMode` cells
    return grid_trace<COUNT>(S, px, py, pz, o, r, minD, maxD, excl, outT, outAB, outAC, cnt, walkMode == 1);
}

// ---- one pixel-sample, serial form: raytrace_opencl.c:452-725 --------------------------------------------------------
// Ring storage is supplied by the caller (registers/local memory on the simple kernel, shared memory on the
// persistent kernel) through RingT: fields indexed by slot.
struct RingLocal {
    int maxB[kRingSize];
    uint32_t excl[kRingSize];
    f3 o[kRingSize], v[kRingSize], mul[kRingSize];
    bool cam[kRingSize];
    float minD[kRingSize], maxD[kRingSize];
};

template <bool COUNT>
OCLR_HD f3 trace_sample(const SceneView& S, const FrameView& F, const float* px, const float* py, const float* pz,
                        uint32_t pixel, uint32_t sampleIdx, uint32_t* primaryId, bool& undefinedRef, Counters* cnt,
                        int walkMode = 0 /* 0 cell walk, 1 two-level walk, 2 packed two-level walk (rt_walk.h), >= 3 packed walk cut into parts of that many cells */) {
    const Camera& cam = F.cam;
    uint64_t rng = (uint64_t)pixel * (uint64_t)F.sampleCount + (uint64_t)(sampleIdx + 1u);
    const float fx = (float)(pixel % cam.width);
    const float fy = (float)(pixel / cam.width);
    RingLocal ring;
    f3 colour = mk3(0.f, 0.f, 0.f);
    int begin = 0, end = 1;
    float tmp;
    ring.maxB[0] = kMaxBounces;
    ring.excl[0] = kNoTriangle;
    ring.o[0] = mk3(cam.eye);
    ring.v[0] = mk3(cam.eyeToTopLeft);
    tmp = fx + rand_f(rng, 0.f, 1.f);
    ring.v[0].x += cam.leftToRight[0] * tmp;
    ring.v[0].y += cam.leftToRight[1] * tmp;
    ring.v[0].z += cam.leftToRight[2] * tmp;
    tmp = fy + rand_f(rng, 0.f, 1.f);
    ring.v[0].x += cam.topToBottom[0] * tmp;
    ring.v[0].y += cam.topToBottom[1] * tmp;
    ring.v[0].z += cam.topToBottom[2] * tmp;
    ring.mul[0] = mk3(1.f, 1.f, 1.f);
    ring.cam[0] = true;
    ring.minD[0] = 0.f;
    ring.maxD[0] = OCLR_INF;
    bool first = true;
    for (; begin != end; begin = (begin + 1) % kRingSize) {
        const f3 ro = ring.o[begin], rv = ring.v[begin], rm = ring.mul[begin];
        float hitT = ring.maxD[begin];
        uint32_t hit = kNoTriangle;
        float hitAB = 0.f, hitAC = 0.f;
        if (COUNT) cnt->segments++;
        if (ring.cam[begin]) {
            const uint32_t e = OCLR_LDG(F.camEnd + pixel);
            for (uint32_t i = OCLR_LDG(F.camStart + pixel); i < e; ++i) {
                const uint32_t tri = OCLR_LDG(F.camList + i);
                if (ring.excl[begin] != tri) {
                    float t, ab, ac;
                    if (COUNT) cnt->primCandidates++;
                    if (tri_test(S.triGeo + 4 * (size_t)tri, ro, rv, ring.minD[begin], hitT, t, ab, ac)) {
                        hitT = t;
                        hit = tri;
                        hitAB = ab;
                        hitAC = ac;
                    }
                }
            }
        } else {
            hit = grid_trace_mode<COUNT>(walkMode, S, px, py, pz, ro, rv, ring.minD[begin], ring.maxD[begin], ring.excl[begin], hitT,
                                         hitAB, hitAC, cnt);
        }
        if (first) {
            if (primaryId) *primaryId = hit;
            first = false;
        }
        if (hit == kNoTriangle) continue;
        if (COUNT) cnt->shadedHits++;

        const TriShade ts = load_shade(S, hit);
        const int m = ts.mat;
        f3 tex = mk3(0.f, 0.f, 0.f), transp = tex, refl = tex, lum = tex;
        f3 face[2] = {mk3(0.1f, 0.1f, 0.1f), mk3(0.1f, 0.1f, 0.1f)};
        const f3 loc = mk3(ro.x + hitT * rv.x, ro.y + hitT * rv.y, ro.z + hitT * rv.z);
        const f3 nrm = triangle_normal(S, cam, ts, loc, ro, rv, hitAB, hitAC, undefinedRef);
        shade_channels(S, ts, hitAB, hitAC, tex, transp, refl, lum);
        for (uint32_t j = 0; j < S.lightCount; ++j) {
            const Light& L = S.lights[j];
            LightRay lr;
            f3 att = mk3(1.f, 1.f, 1.f);
            light_ray(L, loc, rng, lr);
            if (lr.minLen < lr.maxLen) {
                for (;;) {
                    float t, ab = 0.f, ac = 0.f;
                    const uint32_t occ = grid_trace_mode<COUNT>(walkMode, S, px, py, pz, loc, lr.dir, lr.minLen, lr.maxLen, hit, t, ab, ac, cnt);
                    if (occ == kNoTriangle) break;
                    int om;
                    float u0, v0, u1, v1, u2, v2;
                    load_mat_uv(S, occ, om, u0, v0, u1, v1, u2, v2);
                    f3 tr = mk3(0.f, 0.f, 0.f);
                    uint2 sz;
                    if (channel_present(S, om, kChTransparency, sz)) {
                        if (COUNT) cnt->occluderLookups++;
                        const int start = OCLR_LDG(S.matStart + kMaterialChannels * om + kChTransparency);
                        tr = table_value(S.textures + start, sz, u0, v0, u1, v1, u2, v2, ab, ac);
                    }
                    att.x *= tr.x;
                    att.y *= tr.y;
                    att.z *= tr.z;
                    if (!(0.f < att.x && 0.f < att.y && 0.f < att.z)) break;
                    lr.minLen = t;
                }
            }
            light_accumulate(L, nrm, lr, att, face);
        }
        colour.x += (1.f - colour.x) * lum.x * rm.x;
        colour.y += (1.f - colour.y) * lum.y * rm.y;
        colour.z += (1.f - colour.z) * lum.z * rm.z;
        const int front = (int)(dot3(nrm, rv) <= 0.f);
        const f3 light = face[front];
        colour.x += (1.f - colour.x) * rm.x * (1.f - transp.x) * tex.x * light.x;
        colour.y += (1.f - colour.y) * rm.y * (1.f - transp.y) * tex.y * light.y;
        colour.z += (1.f - colour.z) * rm.z * (1.f - transp.z) * tex.z * light.z;

        if (ring.maxB[begin] <= 0) continue;
        // max() is the reference's ternary macro (raytrace.h:30), not fmaxf
        float total = (refl.x + transp.x) > (refl.y + transp.y) ? (refl.x + transp.x) : (refl.y + transp.y);
        total = total > (refl.z + transp.z) ? total : (refl.z + transp.z);
        f3 diffuse = mk3(0.f, 0.f, 0.f);
        if (total < 1.f) diffuse = mk3(1.f - total, 1.f - total, 1.f - total);
        f3 mul = mk3(rm.x * tex.x * diffuse.x, rm.y * tex.y * diffuse.y, rm.z * tex.z * diffuse.z);
        if (3.f / 256.f <= mul.x + mul.y + mul.z) {  // diffuse bounce (:667-683)
            ring.maxB[end] = 0;
            ring.excl[end] = hit;
            ring.o[end] = loc;
            f3 d = sphere_point(rng, 1.f);
            if (front != (int)(0 <= dot3(d, nrm))) d = mk3(-d.x, -d.y, -d.z);
            ring.v[end] = d;
            ring.mul[end] = mul;
            ring.cam[end] = false;
            ring.minD[end] = 0.f;
            ring.maxD[end] = OCLR_INF;
            end = (end + 1) % kRingSize;
            if ((end + 1) % kRingSize == begin) continue;
        }
        mul = mk3(rm.x * tex.x * refl.x, rm.y * tex.y * refl.y, rm.z * tex.z * refl.z);
        if (3.f / 256.f <= mul.x + mul.y + mul.z) {  // mirror (:690-705)
            ring.maxB[end] = ring.maxB[begin] - 1;
            ring.excl[end] = hit;
            ring.o[end] = loc;
            tmp = -2.f * dot3(nrm, rv);
            ring.v[end] = mk3(rv.x + tmp * nrm.x, rv.y + tmp * nrm.y, rv.z + tmp * nrm.z);
            ring.mul[end] = mul;
            ring.cam[end] = false;
            ring.minD[end] = 0.f;
            ring.maxD[end] = OCLR_INF;
            end = (end + 1) % kRingSize;
            if ((end + 1) % kRingSize == begin) continue;
        }
        mul = mk3(rm.x * tex.x * transp.x, rm.y * tex.y * transp.y, rm.z * tex.z * transp.z);
        if (3.f / 256.f <= mul.x + mul.y + mul.z) {  // glass (:711-722)
            ring.maxB[end] = ring.maxB[begin] - 1;
            ring.excl[end] = hit;
            ring.o[end] = ro;
            ring.v[end] = rv;
            ring.mul[end] = mul;
            ring.cam[end] = ring.cam[begin];
            ring.minD[end] = hitT;
            ring.maxD[end] = OCLR_INF;
            end = (end + 1) % kRingSize;
            if ((end + 1) % kRingSize == begin) continue;
        }
    }
    return colour;
}

}  // namespace oclr

#include "rt_walk.h"
