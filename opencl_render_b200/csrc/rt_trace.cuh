// rt_trace.cuh -- production trace stage of the wavefront pipeline (rt_wavefront.cuh has the logic kernel and the first
// trace kernel, kept for A/B): a setup kernel + a persistent trace kernel built on the packed walk of rt_walk.h.
//
// What ncu said about the first trace kernel (profiles/r01_wf_trace_cfg2_*): issue-bound (66-69 % of issue slots), 14 of 32
// lanes active per instruction.  Per-line attribution: the common path of one walk iteration was ~125 instructions at 22 lanes,
// and on top of it every iteration paid ~65 instructions for "open the cell" work at ~6 lanes, ~90 for level switches of the
// two-level walk at 1.4 lanes, and every refill ran BindInCube/GetBoxAddress with a handful of lanes.  Hence:
//
//   wf_setup_kernel   one thread per queued ray, all lanes busy: BindInCube + GetBoxAddress + first crossings
//                     (raytrace_opencl.c:350-362), written as 64-byte records in QUEUE order, so the trace kernel's refill is
//                     four coalesced 128-bit loads per lane instead of a gather plus ~400 divergent instructions.
//   wf_trace2_kernel  WALK phase: one select-based step on packed coordinates (rt_walk.h), an occupied cell costs one
//                     shared-memory store (its rank + entry face go to the lane's pending queue; range, face mask and
//                     candidate scan happen when the cell is popped in the TEST phase, with the lanes that test);
//                     level switches (enter-coarse / refine) are parked and run batched once several lanes need one.
//
// Results are bit-identical to the other kernels by construction (same rt_core.h / rt_walk.h arithmetic; tests assert it).
#pragma once
#include "rt_wavefront.cuh"

namespace oclr {

// Walk records written by wf_setup_kernel, indexed by queue slot.
struct WalkRecords {
    float4* o;    // (o.xyz, minD)
    float4* d;    // (r.xyz, maxD)
    float4* s0;   // (tx, ty, tz, as_float(cpk))
    uint4* s1;    // (epk, excl, path, coarseOk)
};

__global__ void __launch_bounds__(256) wf_setup_kernel(SceneView S, WfState w, WalkRecords rec) {
    extern __shared__ float shPlanes[];
    load_planes(shPlanes, S);
    const int n = S.n;
    const uint32_t count = *w.queueCount;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const uint32_t path = w.queue[idx];
        const float4 ro = w.rayO[path], rd = w.rayD[path];
        PackedWalk g;
        pwalk_setup(g, n, S.nb, shPlanes, shPlanes + (n + 1), shPlanes + 2 * (n + 1), mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), ro.w, rd.w);
        rec.o[idx] = ro;
        rec.d[idx] = rd;
        rec.s0[idx] = make_float4(g.tx, g.ty, g.tz, __uint_as_float(g.cpk));
        rec.s1[idx] = make_uint4(g.epk, w.rayExcl[path], path, g.coarseOk ? 1u : 0u);
    }
}

enum { kWsNone = 0, kWsRun = 1, kWsRefine = 2, kWsEnter = 3, kWsFinished = 4 };

__device__ __forceinline__ void prefetch_l1(const void* p) {
#ifndef OCLR_NO_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

#ifndef OCLR_PENDING2_DEPTH
#define OCLR_PENDING2_DEPTH 8
#endif

template <bool COUNT>
__global__ void __launch_bounds__(128, OCLR_TRACE_MIN_CTAS) wf_trace2_kernel(SceneView S, WfState w, WalkRecords rec, TraceTuning tune,
                                                                            Counters* gcnt) {
    enum { DEPTH = OCLR_PENDING2_DEPTH };
    extern __shared__ float shPlanes[];
    __shared__ uint32_t mailbox[kMailboxSlots][128];   // per-lane direct-mapped cache of tested triangle ids (exact, see rt_wavefront.cuh)
    __shared__ uint32_t pending[DEPTH][128];           // per-lane FIFO of occupied cells found by the walker: rank | entry face << 29
    load_planes(shPlanes, S);
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint32_t count = *w.queueCount;
    const int n = S.n;
    const int nbShift = 31 - __clz(S.nb);

    Counters cnt = {};
    PackedWalk g;
    g.level = 0;
    float minD = 0.f, maxD = 0.f;
    uint32_t excl = kNoTriangle, path = 0;
    int ws = kWsNone;
    int qHead = 0, qCount = 0;
    bool testing = false, exhausted = false;
    uint32_t i = 0, iEnd = 0, nextTri = 0, curMask = 0, curBegin = 0;
    uint32_t best = kNoTriangle;
    float bestT = 0.f, bestAB = 0.f, bestAC = 0.f;
    int face = kFaceNone, lastAxis = 0;
    float lastE = 0.f;

    for (;;) {
        // ---- refill idle lanes: one atomic per warp, records read in queue order ----
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, ws == kWsNone);
        if (idle != 0u && !exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= tune.refillMin)) {
            const int nIdle = __popc(idle);
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(w.queueCursor, (uint32_t)nIdle);
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base + (uint32_t)nIdle >= count) exhausted = true;
            if (ws == kWsNone) {
                const uint32_t idx = base + (uint32_t)__popc(idle & ltMask);
                if (idx < count) {
                    const float4 ro = rec.o[idx], rd = rec.d[idx], s0 = rec.s0[idx];
                    const uint4 s1 = rec.s1[idx];
                    g.o = mk3(ro.x, ro.y, ro.z);
                    g.r = mk3(rd.x, rd.y, rd.z);
                    minD = ro.w;
                    maxD = rd.w;
                    g.tx = s0.x;
                    g.ty = s0.y;
                    g.tz = s0.z;
                    g.cpk = __float_as_uint(s0.w);
                    g.epk = s1.x;
                    excl = s1.y;
                    path = s1.z;
                    g.coarseOk = s1.w != 0u;
                    g.level = 0;
                    g.brick = ((pk_get(g.cpk, 0) >> 2) + (((pk_get(g.cpk, 1) >> 2) + ((pk_get(g.cpk, 2) >> 2) << nbShift)) << nbShift));
                    g.endBrick = g.epk == kPkNone
                                     ? -1
                                     : ((pk_get(g.epk, 0) >> 2) + (((pk_get(g.epk, 1) >> 2) + ((pk_get(g.epk, 2) >> 2) << nbShift)) << nbShift));
                    pwalk_load_brick(g, S.bricks);
                    best = kNoTriangle;
                    ws = kWsRun;
                    qHead = qCount = 0;
                    testing = false;
                    face = kFaceNone;
#pragma unroll
                    for (int k = 0; k < kMailboxSlots; ++k) mailbox[k][threadIdx.x] = kNoTriangle;
                    if (COUNT) {
                        cnt.gridRays++;
                        cnt.bricksLoaded++;
                    }
                }
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, ws != kWsNone) == 0u) break;

        // ---- pick ONE kind of work for this iteration: the one most lanes are ready for (weighted) -------------------------------
        // walk-ready: walking and room in the pending queue; test-ready: a cell being tested or queued; switch: parked for a
        // level switch.  Every live lane is in at least one set, so some score is > 0 and the warp always makes progress.
        const bool rdyWalk = (ws == kWsRun) & (qCount < DEPTH);
        const bool rdyTest = testing | (qCount > 0);
        const bool rdySwitch = (ws == kWsRefine) | (ws == kWsEnter);
        const int sWalk = __popc(__ballot_sync(0xFFFFFFFFu, rdyWalk)) * tune.wWalk;
        const int sTest = __popc(__ballot_sync(0xFFFFFFFFu, rdyTest)) * tune.wTest;
        const int sSwitch = __popc(__ballot_sync(0xFFFFFFFFu, rdySwitch)) * tune.wSwitch;

        if (sWalk >= sTest && sWalk >= sSwitch && sWalk > 0) {
            // ---- WALK: one step of every walk-ready lane ----------------------------------------------------------------------
            if (COUNT) {
                if (lane == 0) cnt.walkWarpIters++;
                if (rdyWalk) cnt.walkLaneIters++;
            }
            if (rdyWalk) {
                const bool coarse = g.level != 0;
                const bool brickEmpty = (g.maskLo | g.maskHi) == 0u;
                const bool inEnd = g.brick == g.endBrick;
                const int bit = pwalk_bit(g.cpk);
                const bool occupied = (!coarse) & pwalk_occupied(g, bit);
                if (COUNT && !coarse) {
                    cnt.cells++;
                    if (brickEmpty) cnt.emptyBrickCells++;
                }
                if (occupied) {
                    int slot = qHead + qCount;
                    slot = slot >= DEPTH ? slot - DEPTH : slot;
                    const uint32_t rank = pwalk_rank(g, bit);
                    pending[slot][threadIdx.x] = rank | ((uint32_t)face << 29);
                    ++qCount;
                    prefetch_l1(S.cellRange + rank);
                    if (COUNT) cnt.cellsNonEmpty++;
                }
                const bool atEnd = (!coarse) & (g.cpk == g.epk);
                const bool needRefine = coarse & ((!brickEmpty) | inEnd);
                const bool needEnter = (!coarse) & brickEmpty & g.coarseOk & (!inEnd) & (tune.hierarchical != 0);
                if (atEnd) {
                    ws = kWsFinished;
                } else if (needRefine | needEnter) {
                    ws = needRefine ? kWsRefine : kWsEnter;
                } else {
                    int up;
                    bool crossed;
                    if (COUNT && coarse) cnt.coarseSteps++;
                    if (!pwalk_step(g, n, nbShift, shPlanes, lastAxis, up, lastE, crossed)) {
                        ws = kWsFinished;
                    } else {
                        face = coarse ? (int)kFaceNone : lastAxis * 2 + up;
                        if (crossed) {
                            pwalk_load_brick(g, S.bricks);
                            if (COUNT) cnt.bricksLoaded++;
                        }
                    }
                }
            }
        } else if (sSwitch > sTest) {
            // ---- SWITCH: parked level switches of the two-level walk, run together -----------------------------------------------
            if (COUNT) {
                if (lane == 0) cnt.switchWarpIters++;
                if (rdySwitch) cnt.switchLaneIters++;
            }
            if (ws == kWsEnter) {
                pwalk_enter_coarse(g, n, shPlanes);
                ws = kWsRun;
                if (COUNT) cnt.coarseEnters++;
            } else if (ws == kWsRefine) {
                pwalk_refine(g, n, shPlanes, lastAxis, lastE);
                face = kFaceNone;
                ws = kWsRun;
            }
        } else {
            // ---- TEST: pop cells in walk order, one real ray/triangle test per lane -----------------------------------------------
            if (COUNT) {
                if (lane == 0) cnt.testWarpIters++;
                if (rdyTest) cnt.testLaneIters++;
            }
            if (rdyTest) {
                if (!testing) {  // open the next queued cell
                    const uint32_t e = pending[qHead][threadIdx.x];
                    qHead = qHead + 1 >= DEPTH ? 0 : qHead + 1;
                    --qCount;
                    const uint32_t rank = e & 0x1FFFFFFFu;
                    const uint32_t f = e >> 29;
                    const uint2 range = __ldg(S.cellRange + rank);
                    // entries shared with the cell the walk came from were examined there (exact, rt_types.h faceMask)
                    curMask = f != (uint32_t)kFaceNone ? __ldg(S.faceMask + 6 * (size_t)rank + f) : 0xFFFFFFFFu;
                    i = curBegin = range.x;
                    iEnd = range.y;
                    bestT = maxD;  // *outRayMult = maxDistance at every cell (:366)
                    testing = true;
                    next_candidate<COUNT>(S.cellList, curBegin, iEnd, curMask, excl, mailbox, i, nextTri, cnt);
                    if (i != iEnd) prefetch_l1(S.triGeo + 4 * (size_t)nextTri);
                }
                if (i != iEnd) {
                    const uint32_t tri = nextTri;
                    float t, ab, ac;
                    if (COUNT) cnt.gridCandidates++;
                    mailbox[tri & (kMailboxSlots - 1)][threadIdx.x] = tri;
                    ++i;
                    next_candidate<COUNT>(S.cellList, curBegin, iEnd, curMask, excl, mailbox, i, nextTri, cnt);
                    if (i != iEnd) prefetch_l1(S.triGeo + 4 * (size_t)nextTri);
                    if (tri_test(S.triGeo + 4 * (size_t)tri, g.o, g.r, minD, bestT, t, ab, ac)) {
                        best = tri;
                        bestT = t;
                        bestAB = ab;
                        bestAC = ac;
                    }
                }
                if (i == iEnd) {  // cell done
                    testing = false;
                    if (best != kNoTriangle) {  // first cell with any hit wins (:380); whatever the walker found beyond it is dropped
                        w.hit[path] = make_float4(__uint_as_float(best), bestT, bestAB, bestAC);
                        ws = kWsNone;
                        qCount = 0;
                    }
                }
            }
        }

        // walk over, nothing left to test, no hit: miss
        if ((ws == kWsFinished) & (!testing) & (qCount == 0)) {
            w.hit[path] = make_float4(__uint_as_float(kNoTriangle), maxD, 0.f, 0.f);
            ws = kWsNone;
        }
    }
    if (COUNT) flush_counters(cnt, gcnt);
}

}  // namespace oclr
