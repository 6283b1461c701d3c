// rt_clip.h -- BoxIntersectsTriangle / Cull of the scene-grid builder (source/util/trianglelist.cpp:381-449), shared by the
// host builder (builders.cpp, g++ -ffp-contract=off) and the device builder (grid_builder.cuh, nvcc -fmad=false) so that both
// make the same fp32 decisions.
#pragma once
#include "rt_core.h"

namespace oclr {

// Clip `poly` against the half-space on one side of `limit` along `dim` (:381-430).  Crossing edges get an interpolated
// vertex; original vertices strictly outside are dropped; vertices on the plane stay.
OCLR_HD bool clip_axis(bool keepBelow, float limit, int dim, int& count, float poly[16][3]) {
    bool fresh[16];
    for (int i = 0; i < 16; ++i) fresh[i] = false;
    for (int i = 0; i < count; ++i) {
        const int nx = (i + 1) % count;
        const float di = limit - poly[i][dim], dn = limit - poly[nx][dim];
        if (di * dn < 0.f) {
            const float ex = poly[nx][0] - poly[i][0], ey = poly[nx][1] - poly[i][1], ez = poly[nx][2] - poly[i][2];
            const float ed = dim == 0 ? ex : (dim == 1 ? ey : ez);
            const float k = di / ed;
            const float bx = poly[i][0], by = poly[i][1], bz = poly[i][2];
            const int at = i + 1;
            for (int j = count++; at < j; --j) {
                poly[j][0] = poly[j - 1][0];
                poly[j][1] = poly[j - 1][1];
                poly[j][2] = poly[j - 1][2];
            }
            poly[at][0] = bx + k * ex;
            poly[at][1] = by + k * ey;
            poly[at][2] = bz + k * ez;
            fresh[at] = true;
            i = at;
        }
    }
    for (int i = 0; i < count; ++i) {
        const bool outside = keepBelow ? (limit < poly[i][dim]) : (poly[i][dim] < limit);
        if (!fresh[i] && outside) {
            for (int j = i + 1; j < count; ++j) {
                poly[j - 1][0] = poly[j][0];
                poly[j - 1][1] = poly[j][1];
                poly[j - 1][2] = poly[j][2];
                fresh[j - 1] = fresh[j];
            }
            --count;
            --i;
        }
    }
    return 0 < count;
}

// :433-449
OCLR_HD bool box_hits_triangle(const float lo[3], const float hi[3], f3 a, f3 b, f3 c) {
    float poly[16][3];
    poly[0][0] = a.x; poly[0][1] = a.y; poly[0][2] = a.z;
    poly[1][0] = b.x; poly[1][1] = b.y; poly[1][2] = b.z;
    poly[2][0] = c.x; poly[2][1] = c.y; poly[2][2] = c.z;
    int count = 3;
    return clip_axis(false, lo[0], 0, count, poly) && clip_axis(false, lo[1], 1, count, poly) && clip_axis(false, lo[2], 2, count, poly) &&
           clip_axis(true, hi[0], 0, count, poly) && clip_axis(true, hi[1], 1, count, poly) && clip_axis(true, hi[2], 2, count, poly);
}

}  // namespace oclr
