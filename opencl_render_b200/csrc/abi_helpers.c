/* abi_helpers.c -- the small host-side helpers the reference declares in source/opencl/raytrace.h:37-44 and its
 * callers use (render.cpp:422,758-760,1197; trianglelist.cpp:54-58,122-124,278,455).  Exported with the
 * reference's names and by-value vector arguments so a plugin built against raytrace.h links unchanged.
 * Compiled as C with -ffp-contract=off: fp32, reference operation order (raytrace.c:18-45,
 * raytrace_opencl.c:83-101,124-193). */
#include <math.h>

#include "../../include/oclr_abi.h"

cl_float dot(cl_float3 a, cl_float3 b) { return a.s[0] * b.s[0] + a.s[1] * b.s[1] + a.s[2] * b.s[2]; }

cl_float3 cross(cl_float3 a, cl_float3 b) {
    cl_float3 c;
    c.s[0] = a.s[1] * b.s[2] - a.s[2] * b.s[1];
    c.s[1] = a.s[2] * b.s[0] - a.s[0] * b.s[2];
    c.s[2] = a.s[0] * b.s[1] - a.s[1] * b.s[0];
    c.s[3] = 0.f;
    return c;
}

cl_float3 normalize(cl_float3 v) {
    const cl_float len = (cl_float)sqrt(dot(v, v));
    cl_float3 r;
    r.s[0] = v.s[0] / len;
    r.s[1] = v.s[1] / len;
    r.s[2] = v.s[2] / len;
    r.s[3] = 0.f;
    return r;
}

cl_float3 vector(cl_float3 a, cl_float3 b) {
    cl_float3 r;
    r.s[0] = b.s[0] - a.s[0];
    r.s[1] = b.s[1] - a.s[1];
    r.s[2] = b.s[2] - a.s[2];
    r.s[3] = 0.f;
    return r;
}

cl_float bindf(cl_float value, cl_float a, cl_float b) {
    const cl_float lo = value > a ? value : a;
    return lo < b ? lo : b;
}

cl_float GetPointToLineSqLen(cl_float3 origin, cl_float3 destination, cl_float3 point) {
    const cl_float3 od = vector(origin, destination);
    const cl_float3 op = vector(origin, point);
    const cl_float k = dot(op, od) / dot(od, od);
    cl_float3 d;
    d.s[0] = (origin.s[0] + k * od.s[0]) - point.s[0];
    d.s[1] = (origin.s[1] + k * od.s[1]) - point.s[1];
    d.s[2] = (origin.s[2] + k * od.s[2]) - point.s[2];
    d.s[3] = 0.f;
    return dot(d, d);
}

cl_bool RayIntersectsTriangle(cl_float3 origin, cl_float3 ray, cl_float minDistance, cl_float maxDistance, cl_float3 a,
                              cl_float3 b, cl_float3 c, cl_float* outRayMult, cl_float* outABL, cl_float* outACL) {
    const cl_float3 ab = vector(a, b), ac = vector(a, c), ao = vector(a, origin);
    const cl_float3 n = cross(ac, ab);
    cl_bool hit = CL_FALSE;
    *outRayMult = -dot(n, ao) / dot(n, ray);
    if (minDistance < *outRayMult && *outRayMult < maxDistance) {
        const cl_float abab = dot(ab, ab), abac = dot(ab, ac), acac = dot(ac, ac);
        const cl_float D = 1.f / (abac * abac - abab * acac);
        cl_float3 ap;
        cl_float apab, apac;
        ap.s[0] = (origin.s[0] + *outRayMult * ray.s[0]) - a.s[0];
        ap.s[1] = (origin.s[1] + *outRayMult * ray.s[1]) - a.s[1];
        ap.s[2] = (origin.s[2] + *outRayMult * ray.s[2]) - a.s[2];
        ap.s[3] = 0.f;
        apab = dot(ap, ab);
        apac = dot(ap, ac);
        *outABL = (abac * apac - acac * apab) * D;
        *outACL = (abac * apab - abab * apac) * D;
        hit = (0 <= *outABL && 0 <= *outACL && *outABL + *outACL <= 1.f);
    }
    return hit;
}

cl_int3 GetBoxAddress(cl_int axesDivCount, cl_float3* boxMin, cl_float3 position) {
    cl_int3 lo;
    lo.s[0] = lo.s[1] = lo.s[2] = lo.s[3] = 0;
    while (1 < axesDivCount) {
        int k;
        axesDivCount /= 2;
        for (k = 0; k < 3; ++k)
            if (boxMin[lo.s[k] + axesDivCount].s[k] < position.s[k]) lo.s[k] += axesDivCount;
    }
    return lo;
}
