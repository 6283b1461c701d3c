/* raytrace_all_min.c -- the drop-in boundary from plain C, end to end: what source/render.cpp:1311-1386 does after scene extraction.
 *
 *   gcc -O1 -std=gnu11 -Iinclude examples/raytrace_all_min.c -Lopencl_render_b200 -lopencl_render_b200 \
 *       -Wl,-rpath,$PWD/opencl_render_b200 -lm -o /tmp/raytrace_all_min && /tmp/raytrace_all_min out.bmp
 *
 * A floor quad and a tilted triangle above it, one material each, one spot light; SetCamera -> CameraTriangleList::New ->
 * SceneTriangleList::New (the library's builders) -> RaytraceAll (by-value OpenCL vector types, exactly the reference's signature)
 * -> writebmp3s layout.  Exit codes: 0 rendered, 3 no CUDA device (the library has no CPU path), 1 anything else. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oclr_abi.h"

static cl_float3 f3(float x, float y, float z) {
    cl_float3 v;
    v.s[0] = x; v.s[1] = y; v.s[2] = z; v.s[3] = 0.f;
    return v;
}

int main(int argc, char** argv) {
    enum { W = 320, H = 240, S = 4 };
    /* geometry: vertices, triangles (a,b,c), per-corner normals and uvs, material ids */
    cl_float3 vertex[7] = {f3(-4, 0, -4), f3(4, 0, -4), f3(4, 0, 4), f3(-4, 0, 4), f3(-1, 0.5f, 0), f3(1, 0.7f, 0.3f), f3(0, 2.2f, -0.2f)};
    cl_int3 tri[3];
    const int idx[3][3] = {{0, 1, 2}, {0, 2, 3}, {4, 5, 6}};   /* the quad split the plugin uses: (a,b,c), (a,c,d) */
    cl_int mat[3] = {0, 0, 1};
    cl_float2 uv[9];
    cl_float3 normal[9];
    for (int t = 0; t < 3; ++t) {
        for (int k = 0; k < 3; ++k) tri[t].s[k] = idx[t][k];
        tri[t].s[3] = 0;
        const cl_float3 n = normalize(cross(vector(vertex[idx[t][0]], vertex[idx[t][1]]), vector(vertex[idx[t][0]], vertex[idx[t][2]])));
        const float flip = n.s[1] < 0.f ? -1.f : 1.f;     /* face up / towards the camera */
        for (int k = 0; k < 3; ++k) {
            normal[3 * t + k] = f3(flip * n.s[0], flip * n.s[1], flip * n.s[2]);
            uv[3 * t + k].s[0] = (float)(k == 2);
            uv[3 * t + k].s[1] = (float)(k >= 1);
        }
    }
    /* two materials x five channels (colour, reflection, transparency, bump, luminance): 1x1 images in one atlas, bump absent */
    cl_uint2 matSize[10];
    cl_int matStart[11];
    cl_uchar3 atlas[8];
    const unsigned char texel[8][3] = {{200, 190, 170}, {40, 40, 40}, {0, 0, 0}, {0, 0, 0}, {220, 60, 40}, {0, 0, 0}, {0, 0, 0}, {12, 0, 0}};
    int cursor = 0;
    for (int m = 0; m < 2; ++m)
        for (int ch = 0; ch < 5; ++ch) {
            const int present = ch != OCLR_MATERIAL_CHANNEL_BUMP;
            matSize[5 * m + ch].s[0] = matSize[5 * m + ch].s[1] = present ? 1u : 0u;
            matStart[5 * m + ch] = cursor;
            if (present) {
                memcpy(atlas[cursor].s, texel[cursor], 3);
                atlas[cursor].s[3] = 0;
                ++cursor;
            }
        }
    matStart[10] = cursor;
    cl_int lightType[1] = {OCLR_LIGHT_TYPE_SPOT};
    cl_float3 lightPos[1] = {f3(3, 6, -4)}, lightDir[1] = {f3(0, -1, 0)}, lightColour[1] = {f3(1, 1, 1)};
    cl_float lightRadius[1] = {0.3f}, lightHalf[1] = {INFINITY};

    /* SetCamera + the two acceleration lists (render.cpp:1055, 1311-1312) */
    oclr_camera cam;
    const cl_float eye[3] = {0.f, 3.f, -7.f}, at[3] = {0.f, 0.8f, 0.f}, up[3] = {0.f, 1.f, 0.f};
    oclr_set_camera(&cam, eye, at, up, 0.8f, W, H);
    oclr_camera_lists lists;
    oclr_scene_grid grid;
    if (!oclr_build_camera_lists(&cam, 7, vertex, 3, tri, &lists) || !oclr_build_scene_grid(256, 7, vertex, 3, tri, &grid)) {
        fprintf(stderr, "builders failed: %s\n", oclr_last_error());
        return 1;
    }
    InitOpenCL();
    if (GetComputationTypeCount() < 2) {   /* index 0 is the reference's CPU entry: a label only in this library */
        fprintf(stderr, "no CUDA device: %s has no CPU path\n", oclr_version());
        oclr_free_camera_lists(&lists);
        oclr_free_scene_grid(&grid);
        return 3;
    }
    cl_ushort* plane = (cl_ushort*)calloc((size_t)3 * W * H, sizeof(cl_ushort));
    cl_uint2 dim;
    dim.s[0] = W; dim.s[1] = H;
    const cl_bool ok = RaytraceAll(1, dim, f3(cam.eye[0], cam.eye[1], cam.eye[2]), f3(cam.eyeToTopLeft[0], cam.eyeToTopLeft[1], cam.eyeToTopLeft[2]),
                                   f3(cam.leftToRight[0], cam.leftToRight[1], cam.leftToRight[2]),
                                   f3(cam.topToBottom[0], cam.topToBottom[1], cam.topToBottom[2]), cam.pixelSizeInv, lists.start, lists.end,
                                   lists.list, (ptrdiff_t)lists.listSize, S, 7, vertex, 3, tri, mat, uv, normal, grid.axesDivCount, grid.boxMin,
                                   grid.start, grid.list, 2, matSize, matStart, (cl_uint)cursor, atlas, 1, lightType, lightPos, lightDir,
                                   lightColour, lightRadius, lightHalf, plane, plane + W * H, plane + 2 * W * H);
    SetProgress(1.f);                      /* render.cpp:1397 */
    int rc = 1;
    if (ok) {
        size_t lit = 0;
        for (size_t i = 0; i < (size_t)W * H; ++i) lit += plane[i] != 0;
        const char* path = argc > 1 ? argv[1] : "img.bmp";
        rc = oclr_write_bmp(path, W, H, plane, plane + W * H, plane + 2 * W * H, 0) ? 0 : 1;
        printf("%s: %dx%d, %d samples, %zu lit pixels, progress %.3f\n", path, W, H, S, lit, GetProgress());
        if (lit == 0) rc = 1;
    } else {
        fprintf(stderr, "RaytraceAll failed: %s\n", oclr_last_error());
    }
    free(plane);
    oclr_free_camera_lists(&lists);
    oclr_free_scene_grid(&grid);
    return rc;
}
