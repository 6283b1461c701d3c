// rt_walk.h -- the grid walk of raytrace_opencl.c:324-401 in the form the production trace kernel runs it: packed cell
// coordinates, one select-based step that serves both levels of the two-level walk, incremental brick addressing.
//
// Same contract as rt_core.h: `__host__ __device__`, the reference's fp32 operation order, IEEE division, no FMA
// contraction; tests/hostemu compiles grid_trace_packed() for the host and compares it with the reference walk cell by cell
// (tests/test_hostemu_parity.py).  The exactness argument of the two-level walk is the one written above
// walk_enter_coarse() in rt_core.h; this file only changes the bookkeeping:
//
//   cpk     cell coordinates (level 0) or 4x4x4-brick coordinates (level 1), 10 bits per axis: x | y << 10 | z << 20
//           (axesDivCount <= 1024; the plugin passes 256, render.cpp:1334);
//   brick   linear id of the brick the walk is in, updated by +-nb^axis only when a step crosses a brick face, so the
//           16-byte brick record is fetched exactly once per brick;
//   step    axis = argmin of the three next-crossing values with the reference's tie rule (x only if strictly smallest,
//           else y if strictly smaller than z, else z; NaNs fall through to z), then ONE plane fetch + ONE division.
#pragma once
#include "rt_core.h"

namespace oclr {

enum { kPkBits = 10, kPkMask = 1023, kPkNone = 0xFFFFFFFFu };
enum { kFaceNone = 7 };

OCLR_HD uint32_t pk_make(int x, int y, int z) { return (uint32_t)x | ((uint32_t)y << kPkBits) | ((uint32_t)z << (2 * kPkBits)); }
OCLR_HD int pk_get(uint32_t pk, int axis) { return (int)((pk >> (axis * kPkBits)) & kPkMask); }

struct PackedWalk {
    f3 o, r;
    float tx, ty, tz;    // next crossing per axis at the current level
    uint32_t cpk;        // current cell (level 0) / brick (level 1)
    uint32_t epk;        // end cell, or kPkNone
    int brick;           // linear id of the current brick
    int endBrick;        // brick of the end cell (walked cell by cell, never skipped), or -1
    uint32_t maskLo, maskHi, rankBase;   // record of `brick`
    int level;           // 0: cells, 1: bricks
    bool coarseOk;       // all direction components non-zero and n >= 4
};

// Ray -> initial walk state: raytrace_opencl.c:350-362 (BindInCube on start and end, GetBoxAddress) + the first three crossing values.
OCLR_HD void pwalk_setup(PackedWalk& w, int n, int nb, const float* px, const float* py, const float* pz, f3 o, f3 r, float minD,
                         float maxD) {
    w.o = o;
    w.r = r;
    const f3 lo = mk3(px[0], py[0], pz[0]);
    const f3 hi = mk3(px[n], py[n], pz[n]);
    f3 start = mk3(o.x + minD * r.x, o.y + minD * r.y, o.z + minD * r.z);
    bind_in_cube(start, r, lo, hi);
    int cx, cy, cz;
    box_address(n, px, py, pz, start, cx, cy, cz);
    w.cpk = pk_make(cx, cy, cz);
    w.epk = kPkNone;
    w.endBrick = -1;
    if (maxD < OCLR_INF) {
        f3 end = mk3(o.x + maxD * r.x, o.y + maxD * r.y, o.z + maxD * r.z);
        bind_in_cube(end, r, lo, hi);
        int ex, ey, ez;
        box_address(n, px, py, pz, end, ex, ey, ez);
        w.epk = pk_make(ex, ey, ez);
        w.endBrick = (ex >> 2) + nb * ((ey >> 2) + nb * (ez >> 2));
    }
    w.tx = (px[cx + (0 <= r.x)] - o.x) / r.x;
    w.ty = (py[cy + (0 <= r.y)] - o.y) / r.y;
    w.tz = (pz[cz + (0 <= r.z)] - o.z) / r.z;
    w.brick = (cx >> 2) + nb * ((cy >> 2) + nb * (cz >> 2));
    w.level = 0;
    w.coarseOk = (n >= 4) & (r.x != 0.f) & (r.y != 0.f) & (r.z != 0.f);
    w.maskLo = w.maskHi = w.rankBase = 0;
}

OCLR_HD void pwalk_load_brick(PackedWalk& w, const uint4* bricks) {
    const uint4 br = OCLR_LDG(bricks + w.brick);
    w.maskLo = br.x;
    w.maskHi = br.y;
    w.rankBase = br.z;
}

// Bit of the current cell inside its brick's occupancy mask: (x & 3) | (y & 3) << 2 | (z & 3) << 4.
OCLR_HD int pwalk_bit(uint32_t cpk) {
    const uint32_t t = cpk & 0x00300C03u;
    return (int)((t | (t >> 8) | (t >> 16)) & 63u);
}
OCLR_HD bool pwalk_occupied(const PackedWalk& w, int bit) {
    const uint32_t half = (bit & 32) ? w.maskHi : w.maskLo;
    return ((half >> (bit & 31)) & 1u) != 0u;
}
// Index of the current (occupied) cell among the non-empty cells of the scene, brick-major.
OCLR_HD uint32_t pwalk_rank(const PackedWalk& w, int bit) {
    const uint64_t m = (uint64_t)w.maskLo | ((uint64_t)w.maskHi << 32);
    return w.rankBase + (uint32_t)OCLR_POPCLL(m & ((1ull << bit) - 1ull));
}

// One step at the current level (:383-398).  Returns false when the walk left the grid.  `axis`, `up` and `tEvent`
// describe the crossing taken; `crossed` is set when the step entered another brick (always at level 1).
OCLR_HD bool pwalk_step(PackedWalk& w, int n, int nbShift, const float* planes, int& axis, int& up, float& tEvent, bool& crossed) {
    const bool xmin = (w.tx < w.ty) & (w.tx < w.tz);
    const bool ymin = (!xmin) & (w.ty < w.tz);
    axis = xmin ? 0 : (ymin ? 1 : 2);
    tEvent = xmin ? w.tx : (ymin ? w.ty : w.tz);
    const float rr = xmin ? w.r.x : (ymin ? w.r.y : w.r.z);
    const float oo = xmin ? w.o.x : (ymin ? w.o.y : w.o.z);
    up = (0 <= rr) ? 1 : 0;
    const int sh = axis * kPkBits;
    const int c = (int)((w.cpk >> sh) & kPkMask);
    const int dir = up ? 1 : -1;
    const int cn = c + dir;
    const int lsh = 2 * w.level;
    crossed = false;
    if ((uint32_t)cn >= (uint32_t)(n >> lsh)) return false;
    const float t = (planes[axis * (n + 1) + ((cn + up) << lsh)] - oo) / rr;
    w.cpk += (uint32_t)dir << sh;   // two's complement: -1 << sh subtracts one from the axis' field
    w.tx = xmin ? t : w.tx;
    w.ty = ymin ? t : w.ty;
    w.tz = (xmin | ymin) ? w.tz : t;
    crossed = (w.level != 0) | (((c ^ cn) & ~3) != 0);
    if (crossed) w.brick += dir * (1 << (axis * nbShift));
    return true;
}

// Level 0 -> level 1 inside an empty brick: coordinates become brick coordinates, the heads become the brick-exit crossings.
OCLR_HD void pwalk_enter_coarse(PackedWalk& w, int n, const float* planes) {
    w.cpk = (w.cpk >> 2) & 0x0FF3FCFFu;
    const float* px = planes;
    const float* py = planes + (n + 1);
    const float* pz = planes + 2 * (n + 1);
    w.tx = (px[(pk_get(w.cpk, 0) + (0 <= w.r.x)) << 2] - w.o.x) / w.r.x;
    w.ty = (py[(pk_get(w.cpk, 1) + (0 <= w.r.y)) << 2] - w.o.y) / w.r.y;
    w.tz = (pz[(pk_get(w.cpk, 2) + (0 <= w.r.z)) << 2] - w.o.z) / w.r.z;
    w.level = 1;
}

// Level 1 -> level 0 after the brick-level step along `axis` (crossing value E) entered a brick that has to be walked cell
// by cell: the exact cell state the cell-level walk would have on entering this brick (rt_core.h: walk_refine).
OCLR_HD void pwalk_refine(PackedWalk& w, int n, const float* planes, int axis, float E) {
    const float* px = planes;
    const float* py = planes + (n + 1);
    const float* pz = planes + 2 * (n + 1);
    const int bx = pk_get(w.cpk, 0), by = pk_get(w.cpk, 1), bz = pk_get(w.cpk, 2);
    int cx, cy, cz;
    float tx, ty, tz;
    if (axis == 0) {
        const int up = (0 <= w.r.x);
        cx = up ? (bx << 2) : (bx << 2) + 3;
        tx = (px[cx + up] - w.o.x) / w.r.x;
    } else {
        refine_axis(bx, w.tx, w.o.x, w.r.x, px, E, true, cx, tx);
    }
    if (axis == 1) {
        const int up = (0 <= w.r.y);
        cy = up ? (by << 2) : (by << 2) + 3;
        ty = (py[cy + up] - w.o.y) / w.r.y;
    } else {
        refine_axis(by, w.ty, w.o.y, w.r.y, py, E, axis == 2, cy, ty);
    }
    if (axis == 2) {
        const int up = (0 <= w.r.z);
        cz = up ? (bz << 2) : (bz << 2) + 3;
        tz = (pz[cz + up] - w.o.z) / w.r.z;
    } else {
        refine_axis(bz, w.tz, w.o.z, w.r.z, pz, E, false, cz, tz);
    }
    w.cpk = pk_make(cx, cy, cz);
    w.tx = tx;
    w.ty = ty;
    w.tz = tz;
    w.level = 0;
}

// Position of the k-th still-untested list entry at or after `k` (rt_wavefront.cuh next_candidate, serial form without a mailbox).
OCLR_HD uint32_t pwalk_next_entry(uint32_t begin, uint32_t end, uint32_t faceBits, uint32_t k) {
    const uint32_t rel = k - begin;
    if (rel < 32u) {
        const uint32_t mm = faceBits >> rel;
        if (mm == 0u) return begin + 32u < end ? begin + 32u : end;
        uint32_t s = 0;
        while (((mm >> s) & 1u) == 0u) ++s;
        k += s;
    }
    return k < end ? k : end;
}

// Whole traversal for one ray in the trace kernel's formulation, serial (test infrastructure for the host; the kernel
// runs the same functions with the work of 32 rays interleaved).  Face masks skip entries shared with the cell just left;
// a small direct-mapped mailbox skips triangles this ray already tested -- both exact (rt_wavefront.cuh).
template <bool COUNT>
OCLR_HD uint32_t grid_trace_packed(const SceneView& S, const float* planes, f3 o, f3 r, float minD, float maxD, uint32_t excl,
                                   float& outT, float& outAB, float& outAC, Counters* cnt) {
    const int n = S.n;
    int nbShift = 0;
    while ((1 << nbShift) < S.nb) ++nbShift;
    PackedWalk w;
    pwalk_setup(w, n, S.nb, planes, planes + (n + 1), planes + 2 * (n + 1), o, r, minD, maxD);
    pwalk_load_brick(w, S.bricks);
    if (COUNT) {
        cnt->gridRays++;
        cnt->bricksLoaded++;
    }
    uint32_t mailbox[16];
    for (int k = 0; k < 16; ++k) mailbox[k] = kNoTriangle;
    int face = kFaceNone, lastAxis = 0;
    float lastE = 0.f;
    outT = maxD;
    for (;;) {
        if (w.level == 0) {
            const int bit = pwalk_bit(w.cpk);
            if (COUNT) {
                cnt->cells++;
                if ((w.maskLo | w.maskHi) == 0u) cnt->emptyBrickCells++;
            }
            if (pwalk_occupied(w, bit)) {
                const uint32_t rank = pwalk_rank(w, bit);
                const uint2 range = OCLR_LDG(S.cellRange + rank);
                const uint32_t fm = face != kFaceNone ? OCLR_LDG(S.faceMask + 6 * (size_t)rank + face) : 0xFFFFFFFFu;
                if (COUNT) cnt->cellsNonEmpty++;
                uint32_t closest = kNoTriangle;
                outT = maxD;
                for (uint32_t k = pwalk_next_entry(range.x, range.y, fm, range.x); k < range.y; k = pwalk_next_entry(range.x, range.y, fm, k + 1)) {
                    const uint32_t tri = OCLR_LDG(S.cellList + k);
                    if (tri == excl) continue;
                    if (mailbox[tri & 15u] == tri) {
                        if (COUNT) cnt->mailboxSkips++;
                        continue;
                    }
                    mailbox[tri & 15u] = tri;
                    float t, ab, ac;
                    if (COUNT) cnt->gridCandidates++;
                    if (tri_test(S.triGeo + 4 * (size_t)tri, o, r, minD, outT, t, ab, ac)) {
                        closest = tri;
                        outT = t;
                        outAB = ab;
                        outAC = ac;
                    }
                }
                if (closest != kNoTriangle) return closest;
            }
            if (w.cpk == w.epk) break;
            if (((w.maskLo | w.maskHi) == 0u) & w.coarseOk & (w.brick != w.endBrick)) {
                pwalk_enter_coarse(w, n, planes);
                if (COUNT) cnt->coarseEnters++;
                continue;
            }
        } else if (((w.maskLo | w.maskHi) != 0u) | (w.brick == w.endBrick)) {
            pwalk_refine(w, n, planes, lastAxis, lastE);
            face = kFaceNone;
            continue;
        }
        int up;
        bool crossed;
        if (COUNT && w.level) cnt->coarseSteps++;
        if (!pwalk_step(w, n, nbShift, planes, lastAxis, up, lastE, crossed)) break;
        face = w.level ? (int)kFaceNone : lastAxis * 2 + up;
        if (crossed) {
            pwalk_load_brick(w, S.bricks);
            if (COUNT) cnt->bricksLoaded++;
        }
    }
    outT = maxD;
    return kNoTriangle;
}

}  // namespace oclr
