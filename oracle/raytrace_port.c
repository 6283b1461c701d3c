/* oracle/raytrace_port.c -- TEST INFRASTRUCTURE ONLY: a plain-C CPU restatement of the reference's raytrace path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or call this file; the product
 * library (opencl_render_b200/) never links or executes it and has no CPU path of its own.
 *
 * PARITY PIN: the reference ships no golden vectors, so this port is pinned against the reference
 * ITSELF: oracle/build_ref.py compiles the unmodified /root/reference sources into oracle/_ref/, tests/test_oracle.py
 * checks this port bit-for-bit against that build on seeded scenes, and tests/golden/ holds outputs generated from the
 * reference build (tests/golden/make_golden.py) which this port must also reproduce where /root/reference is absent.
 *
 * What it follows (all in /root/reference/source/opencl/):
 *   port_rand            raytrace_opencl.c:1-23        rotl64 / xorshift64star / randF
 *   port_ball_sample     raytrace_opencl.c:30-45       GetSpherePoint
 *   port_line_dist2      raytrace_opencl.c:83-101      GetPointToLineSqLen
 *   port_texel           raytrace_opencl.c:25-28,103-122  positive_modf / Get2dTableValue3
 *   port_hit_triangle    raytrace_opencl.c:124-172     RayIntersectsTriangle
 *   port_locate          raytrace_opencl.c:174-193     GetBoxAddress
 *   port_clamp_to_grid   raytrace_opencl.c:265-322     BindInCube
 *   port_walk_grid       raytrace_opencl.c:324-401     RayIntersectsTriangles
 *   port_surface_normal  raytrace_opencl.c:195-263     GetTriangleNormal
 *   port_pixel_sample    raytrace_opencl.c:406-742     Raytrace (C-path seed rule :478-481 with raytrace.c:612-616)
 *   port_render_rows     raytrace.c:604-655            RaytraceAll, computationType 0
 * Arithmetic: fp32 in the reference's operation order, compiled -O2 -ffp-contract=off; dot is ((x+y)+z)
 * (raytrace.c:18-20); sqrt/sin/cos/pow/modf/floor go through double exactly where the C path does.
 * Data layout: the reference's own arrays (cl_float3 = 4 floats, cl_int3 = 4 ints, cl_uchar3 = 4 bytes).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

typedef struct {
    uint32_t width, height;
    v3 eye, top_left, step_right, step_down;
    float pixel_size_inv;
    const uint32_t *cam_start, *cam_end, *cam_list;
    uint32_t samples;
    const float* vertex;        /* 4 floats each */
    const int32_t* tri_index;   /* 4 ints each */
    const int32_t* tri_material;
    const float* tri_uv;        /* 6 floats each */
    const float* tri_normal;    /* 12 floats each */
    int32_t divisions;
    const float* planes;        /* sceneBoxMin: 4 floats per plane index */
    const uint32_t *cell_start, *cell_list;
    const uint32_t* mat_size;   /* 2 uints per (material, channel) */
    const int32_t* mat_start;
    const uint8_t* texels;      /* 4 bytes each */
    uint32_t light_count;
    const int32_t* light_type;
    const float *light_pos, *light_dir, *light_colour, *light_radius, *light_half;
    uint16_t *out_r, *out_g, *out_b;
    uint32_t* primary_id;       /* optional */
} port_job;

enum { CH_COLOR = 0, CH_REFLECTION, CH_TRANSPARENCY, CH_BUMP, CH_LUMINANCE, CH_COUNT };
#define NONE 0xFFFFFFFFu
#define RING 12

static v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static v3 load3(const float* p) { return V(p[0], p[1], p[2]); }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 along(v3 o, float t, v3 d) { return V(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z); }

static uint64_t rol(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }
static uint64_t star(uint64_t v) { v ^= v >> 12; v ^= v << 25; v ^= v >> 27; return v * 2685821657736338717ull; }

float port_rand(uint64_t* s, float lo, float hi) {
    static const struct { int a, b, mul; uint64_t k; } round[8] = {
        {55, 3, 1, 0xc23f3c0ad9da6357ull}, {35, 3, 0, 0xce84d6af03c16b89ull}, {63, 35, 1, 0xf097ef8bbe03ddccull},
        {41, 12, 0, 0x48302294fbfe30bfull}, {1, 62, 1, 0x79e7425e3f4f147dull}, {42, 29, 0, 0x14d1d30856e5be9aull},
        {47, 45, 1, 0x24289d47a66617c3ull}, {39, 6, 0, 0x5576fb2f80a05d14ull}};
    int i;
    for (i = 0; i < 8; ++i) {
        const uint64_t m = rol(*s, round[i].a) ^ rol(*s, round[i].b);
        *s ^= star(round[i].mul ? m * round[i].k : m ^ round[i].k);
    }
    return lo + (hi - lo) * (float)((double)*s / (double)0xffffffffffffffffull);
}

v3 port_ball_sample(uint64_t* s, float radius) {
    v3 p;
    float len, k;
    do {
        p.x = port_rand(s, -1, 1);
        p.y = port_rand(s, -1, 1);
        p.z = port_rand(s, -1, 1);
        len = (float)sqrt(dot(p, p));
    } while (len <= 0.f);
    k = (float)sqrt(port_rand(s, 0, 1)) * radius / len;
    return V(k * p.x, k * p.y, k * p.z);
}

static float port_line_dist2(v3 from, v3 to, v3 p) {
    const v3 d = sub(to, from);
    const float dd = dot(d, d);
    const float k = dot(sub(p, from), d) / dd;
    return dot(sub(along(from, k, d), p), sub(along(from, k, d), p));
}

static float wrap01(float v) {
    double ip;
    return (float)modf(modf((double)v, &ip) + 1., &ip);
}

static v3 port_texel(const uint8_t* image, const uint32_t size[2], const float* uv, float bl, float cl) {
    const float u = wrap01(uv[0] + (uv[2] - uv[0]) * bl + (uv[4] - uv[0]) * cl);
    const float v = wrap01(uv[1] + (uv[3] - uv[1]) * bl + (uv[5] - uv[1]) * cl);
    const int ix = (int)floor(u * (float)(size[0] - 1u));
    const int iy = (int)floor(v * (float)(size[1] - 1u));
    const uint8_t* t = image + 4 * (size_t)(int)((uint32_t)ix + (uint32_t)iy * size[0]);
    return V(t[0] / 255.f, t[1] / 255.f, t[2] / 255.f);
}

int port_hit_triangle(v3 o, v3 d, float lo, float hi, v3 a, v3 b, v3 c, float* t, float* bl, float* cl) {
    const v3 ab = sub(b, a), ac = sub(c, a), ao = sub(o, a);
    const v3 n = cross(ac, ab);
    *t = -dot(n, ao) / dot(n, d);
    if (lo < *t && *t < hi) {
        const float bb = dot(ab, ab), bc = dot(ab, ac), cc = dot(ac, ac);
        const float inv = 1.f / (bc * bc - bb * cc);
        const v3 ap = sub(along(o, *t, d), a);
        const float pb = dot(ap, ab), pc = dot(ap, ac);
        *bl = (bc * pc - cc * pb) * inv;
        *cl = (bc * pb - bb * pc) * inv;
        return 0 <= *bl && 0 <= *cl && *bl + *cl <= 1.f;
    }
    return 0;
}

static void port_locate(const port_job* j, v3 p, int cell[3]) {
    int n = j->divisions;
    cell[0] = cell[1] = cell[2] = 0;
    while (1 < n) {
        n /= 2;
        if (j->planes[4 * (cell[0] + n) + 0] < p.x) cell[0] += n;
        if (j->planes[4 * (cell[1] + n) + 1] < p.y) cell[1] += n;
        if (j->planes[4 * (cell[2] + n) + 2] < p.z) cell[2] += n;
    }
}

/* Pull a point lying outside the grid box back onto it along the ray, one face at a time; gives up (keeping the moves
 * made so far) as soon as the ray points away from a face it is beyond. */
static void port_clamp_to_grid(const port_job* j, v3* p, v3 d) {
    const float* lo = j->planes;
    const float* hi = j->planes + 4 * j->divisions;
    float* pc = &p->x;
    const float* dc = &d.x;
    int k;
    for (k = 0; k < 3; ++k) {
        if (pc[k] < lo[k]) {
            if (dc[k] <= 0) return;
            *p = along(*p, (lo[k] - pc[k]) / dc[k], d);
        }
        if (hi[k] < pc[k]) {
            if (0 <= dc[k]) return;
            *p = along(*p, (hi[k] - pc[k]) / dc[k], d);
        }
    }
}

static v3 tri_vertex(const port_job* j, uint32_t tri, int corner) {
    return load3(j->vertex + 4 * (size_t)j->tri_index[4 * (size_t)tri + corner]);
}

uint32_t port_walk_grid(const port_job* j, v3 o, v3 d, float lo, float hi, uint32_t skip, float* t, float* bl, float* cl) {
    const int n = j->divisions;
    const int fwd[3] = {0 <= d.x, 0 <= d.y, 0 <= d.z};
    int cell[3], last[3] = {-1, -1, -1};
    uint32_t found = NONE;
    v3 p = along(o, lo, d);
    port_clamp_to_grid(j, &p, d);
    port_locate(j, p, cell);
    if (hi < INFINITY) {
        p = along(o, hi, d);
        port_clamp_to_grid(j, &p, d);
        port_locate(j, p, last);
    }
    for (;;) {
        const uint32_t id = (uint32_t)(cell[0] + n * cell[1] + n * n * cell[2]);
        uint32_t i;
        float tx, ty, tz;
        *t = hi;
        for (i = j->cell_start[id]; i < j->cell_start[id + 1]; ++i) {
            const uint32_t tri = j->cell_list[i];
            float tt, b, c;
            if (tri != skip && port_hit_triangle(o, d, lo, *t, tri_vertex(j, tri, 0), tri_vertex(j, tri, 1), tri_vertex(j, tri, 2), &tt, &b, &c)) {
                found = tri;
                *t = tt;
                *bl = b;
                *cl = c;
            }
        }
        if (found != NONE || (cell[0] == last[0] && cell[1] == last[1] && cell[2] == last[2])) break;
        tx = (j->planes[4 * (cell[0] + fwd[0]) + 0] - o.x) / d.x;
        ty = (j->planes[4 * (cell[1] + fwd[1]) + 1] - o.y) / d.y;
        tz = (j->planes[4 * (cell[2] + fwd[2]) + 2] - o.z) / d.z;
        {
            const int axis = ((tx < ty) & (tx < tz)) ? 0 : (ty < tz ? 1 : 2);
            cell[axis] += fwd[axis] ? 1 : -1;
            if (cell[axis] < 0 || n <= cell[axis]) break;
        }
    }
    return found;
}

static int channel(const port_job* j, int mat, int ch, const uint32_t** size, const uint8_t** image) {
    if (mat < 0) return 0;
    *size = j->mat_size + 2 * (size_t)(CH_COUNT * mat + ch);
    if (!(0 < (*size)[0])) return 0;
    *image = j->texels + 4 * (size_t)j->mat_start[CH_COUNT * mat + ch];
    return 1;
}

static v3 port_surface_normal(const port_job* j, v3 at, v3 o, v3 d, uint32_t tri, float bl, float cl) {
    const v3 a = tri_vertex(j, tri, 0), b = tri_vertex(j, tri, 1), c = tri_vertex(j, tri, 2);
    const float* nn = j->tri_normal + 12 * (size_t)tri;
    const float wab = (float)sqrt(port_line_dist2(a, b, at));
    const float wbc = (float)sqrt(port_line_dist2(b, c, at));
    const float wca = (float)sqrt(port_line_dist2(c, a, at));
    const float inv = 1.f / (wab + wbc + wca);
    v3 n = V((wab * nn[8] + wbc * nn[0] + wca * nn[4]) * inv, (wab * nn[9] + wbc * nn[1] + wca * nn[5]) * inv,
             (wab * nn[10] + wbc * nn[2] + wca * nn[6]) * inv);
    const uint32_t* size;
    const uint8_t* image;
    if (channel(j, j->tri_material[tri], CH_BUMP, &size, &image)) {
        const float* uv = j->tri_uv + 6 * (size_t)tri;
        const float pi = 3.14159265f;
        float tt, b2 = bl, c2 = cl; /* uninitialised in the reference: see the note in rt_core.h triangle_normal */
        const v3 h0 = port_texel(image, size, uv, bl, cl);
        v3 hs, he;
        float ax, ay, sx, sy, cc, li;
        port_hit_triangle(o, V(d.x + j->step_down.x, d.y + j->step_down.y, d.z + j->step_down.z), 0.f, INFINITY, a, b, c, &tt, &b2, &c2);
        hs = port_texel(image, size, uv, b2, c2);
        port_hit_triangle(o, V(d.x + j->step_right.x, d.y + j->step_right.y, d.z + j->step_right.z), 0.f, INFINITY, a, b, c, &tt, &b2, &c2);
        he = port_texel(image, size, uv, b2, c2);
        ax = (he.x - h0.x) * pi / 2.f;
        ay = (hs.x - h0.x) * pi / 2.f;
        sx = (float)sin(ax);
        sy = (float)sin(ay);
        cc = (float)cos(ax) * (float)cos(ay);
        n.x = cc * n.x / j->pixel_size_inv + sx * j->step_right.x + sy * j->step_down.x;
        n.y = cc * n.y / j->pixel_size_inv + sx * j->step_right.y + sy * j->step_down.y;
        n.z = cc * n.z / j->pixel_size_inv + sx * j->step_right.z + sy * j->step_down.z;
        li = 1.f / (float)sqrt(dot(n, n));
        n = V(n.x * li, n.y * li, n.z * li);
    }
    return n;
}

typedef struct {
    int depth;
    uint32_t skip;
    v3 o, d, weight;
    int from_camera;
    float lo, hi;
} segment;

static int x86_trunc(float f) { return (f >= -2147483648.f && f < 2147483648.f) ? (int)f : (int)0x80000000; }

static uint16_t add16(uint16_t prev, float value, float scale) {
    int v = (int)prev + x86_trunc(value * scale);
    return (uint16_t)(v < 0 ? 0 : (v > 0xFFFF ? 0xFFFF : v));
}

/* One pixel-sample.  sample_id is 1-based (the C path pre-increments it, raytrace_opencl.c:474). */
void port_pixel_sample(const port_job* j, uint32_t pixel, uint32_t sample_id) {
    uint64_t rng = (uint64_t)pixel * (uint64_t)j->samples + (uint64_t)sample_id;
    segment ring[RING];
    int head = 0, tail = 1, first = 1;
    v3 colour = V(0, 0, 0);
    float k;
    ring[0].depth = 12;
    ring[0].skip = NONE;
    ring[0].o = j->eye;
    k = (float)(pixel % j->width) + port_rand(&rng, 0.f, 1.f);
    ring[0].d = V(j->top_left.x + j->step_right.x * k, j->top_left.y + j->step_right.y * k, j->top_left.z + j->step_right.z * k);
    k = (float)(pixel / j->width) + port_rand(&rng, 0.f, 1.f);
    ring[0].d = V(ring[0].d.x + j->step_down.x * k, ring[0].d.y + j->step_down.y * k, ring[0].d.z + j->step_down.z * k);
    ring[0].weight = V(1, 1, 1);
    ring[0].from_camera = 1;
    ring[0].lo = 0.f;
    ring[0].hi = INFINITY;
    for (; head != tail; head = (head + 1) % RING) {
        const segment s = ring[head];
        float t = s.hi, bl = 0.f, cl = 0.f;
        uint32_t tri = NONE;
        if (s.from_camera) {
            uint32_t i;
            for (i = j->cam_start[pixel]; i < j->cam_end[pixel]; ++i) {
                const uint32_t cand = j->cam_list[i];
                float tt, b, c;
                if (cand != s.skip &&
                    port_hit_triangle(s.o, s.d, s.lo, t, tri_vertex(j, cand, 0), tri_vertex(j, cand, 1), tri_vertex(j, cand, 2), &tt, &b, &c)) {
                    t = tt;
                    tri = cand;
                    bl = b;
                    cl = c;
                }
            }
        } else {
            tri = port_walk_grid(j, s.o, s.d, s.lo, s.hi, s.skip, &t, &bl, &cl);
        }
        if (first) {
            if (j->primary_id && sample_id == 1) j->primary_id[pixel] = tri;
            first = 0;
        }
        if (tri == NONE) continue;
        {
            const int mat = j->tri_material[tri];
            const float* uv = j->tri_uv + 6 * (size_t)tri;
            const v3 at = along(s.o, t, s.d);
            const v3 n = port_surface_normal(j, at, s.o, s.d, tri, bl, cl);
            v3 albedo = V(0, 0, 0), glass = albedo, mirror = albedo, glow = albedo;
            v3 lit[2] = {{0.1f, 0.1f, 0.1f}, {0.1f, 0.1f, 0.1f}};
            const uint32_t* size;
            const uint8_t* image;
            uint32_t l;
            int facing;
            v3 shade, w;
            float busy;
            if (channel(j, mat, CH_COLOR, &size, &image)) albedo = port_texel(image, size, uv, bl, cl);
            if (channel(j, mat, CH_TRANSPARENCY, &size, &image)) glass = port_texel(image, size, uv, bl, cl);
            if (channel(j, mat, CH_REFLECTION, &size, &image)) mirror = port_texel(image, size, uv, bl, cl);
            if (channel(j, mat, CH_LUMINANCE, &size, &image)) glow = port_texel(image, size, uv, bl, cl);
            for (l = 0; l < j->light_count; ++l) {
                v3 to = V(0, 0, 0), through = V(1, 1, 1);
                float near_ = 0.f, far_ = 0.f;
                const int type = j->light_type[l];
                if (type == 1 || type == 2 || type == 7 || type == 8 || type == 9) {
                    const v3 jitter = port_ball_sample(&rng, j->light_radius[l]);
                    float inv;
                    to = V(jitter.x + j->light_pos[4 * l] - at.x, jitter.y + j->light_pos[4 * l + 1] - at.y, jitter.z + j->light_pos[4 * l + 2] - at.z);
                    far_ = (float)sqrt(dot(to, to));
                    inv = 1.f / far_;
                    to = V(to.x * inv, to.y * inv, to.z * inv);
                } else if (type >= 3 && type <= 6) {
                    const v3 dir = load3(j->light_dir + 4 * l);
                    const float spread = (float)(sin((j->light_radius[l] / 2.f) * 3.14159265f / 180.f) * sqrt(dot(dir, dir)));
                    float inv;
                    to = port_ball_sample(&rng, spread);
                    to = sub(to, dir);
                    inv = 1.f / (float)sqrt(dot(to, to));
                    to = V(to.x * inv, to.y * inv, to.z * inv);
                    far_ = INFINITY;
                }
                if (near_ < far_) {
                    for (;;) {
                        float tt, b, c;
                        v3 pass = V(0, 0, 0);
                        const uint32_t blocker = port_walk_grid(j, at, to, near_, far_, tri, &tt, &b, &c);
                        if (blocker == NONE) break;
                        if (channel(j, j->tri_material[blocker], CH_TRANSPARENCY, &size, &image))
                            pass = port_texel(image, size, j->tri_uv + 6 * (size_t)blocker, b, c);
                        through = V(through.x * pass.x, through.y * pass.y, through.z * pass.z);
                        if (!(0.f < through.x && 0.f < through.y && 0.f < through.z)) break;
                        near_ = tt;
                    }
                }
                {
                    const float cosine = dot(n, to);
                    const int side = 0.f <= cosine;
                    const float fall = (float)pow(0.5f, far_ / j->light_half[l]);
                    const float e = (float)fabs(cosine) * (fall == fall ? fall : 1.f);
                    lit[side].x += (1.f - lit[side].x) * through.x * e * j->light_colour[4 * l];
                    lit[side].y += (1.f - lit[side].y) * through.y * e * j->light_colour[4 * l + 1];
                    lit[side].z += (1.f - lit[side].z) * through.z * e * j->light_colour[4 * l + 2];
                }
            }
            colour.x += (1.f - colour.x) * glow.x * s.weight.x;
            colour.y += (1.f - colour.y) * glow.y * s.weight.y;
            colour.z += (1.f - colour.z) * glow.z * s.weight.z;
            facing = dot(n, s.d) <= 0.f;
            shade = lit[facing];
            colour.x += (1.f - colour.x) * s.weight.x * (1.f - glass.x) * albedo.x * shade.x;
            colour.y += (1.f - colour.y) * s.weight.y * (1.f - glass.y) * albedo.y * shade.y;
            colour.z += (1.f - colour.z) * s.weight.z * (1.f - glass.z) * albedo.z * shade.z;
            if (s.depth <= 0) continue;

            busy = (mirror.x + glass.x) > (mirror.y + glass.y) ? (mirror.x + glass.x) : (mirror.y + glass.y);
            busy = busy > (mirror.z + glass.z) ? busy : (mirror.z + glass.z);
            busy = busy < 1.f ? 1.f - busy : 0.f;
            w = V(s.weight.x * albedo.x * busy, s.weight.y * albedo.y * busy, s.weight.z * albedo.z * busy);
            if (3.f / 256.f <= w.x + w.y + w.z) { /* diffuse bounce: depth 0, random direction on the side the ray came from */
                segment* q = &ring[tail];
                v3 dir = port_ball_sample(&rng, 1.f);
                if (facing != (0 <= dot(dir, n))) dir = V(-dir.x, -dir.y, -dir.z);
                q->depth = 0; q->skip = tri; q->o = at; q->d = dir; q->weight = w; q->from_camera = 0; q->lo = 0.f; q->hi = INFINITY;
                tail = (tail + 1) % RING;
                if ((tail + 1) % RING == head) continue;
            }
            w = V(s.weight.x * albedo.x * mirror.x, s.weight.y * albedo.y * mirror.y, s.weight.z * albedo.z * mirror.z);
            if (3.f / 256.f <= w.x + w.y + w.z) { /* mirror */
                segment* q = &ring[tail];
                const float twice = -2.f * dot(n, s.d);
                q->depth = s.depth - 1; q->skip = tri; q->o = at; q->d = V(s.d.x + twice * n.x, s.d.y + twice * n.y, s.d.z + twice * n.z);
                q->weight = w; q->from_camera = 0; q->lo = 0.f; q->hi = INFINITY;
                tail = (tail + 1) % RING;
                if ((tail + 1) % RING == head) continue;
            }
            w = V(s.weight.x * albedo.x * glass.x, s.weight.y * albedo.y * glass.y, s.weight.z * albedo.z * glass.z);
            if (3.f / 256.f <= w.x + w.y + w.z) { /* glass: same ray continued behind the surface */
                segment* q = &ring[tail];
                q->depth = s.depth - 1; q->skip = tri; q->o = s.o; q->d = s.d; q->weight = w; q->from_camera = s.from_camera; q->lo = t; q->hi = INFINITY;
                tail = (tail + 1) % RING;
                if ((tail + 1) % RING == head) continue;
            }
        }
    }
    {
        const float scale = (float)0xFFFF / (float)j->samples;
        j->out_r[pixel] = add16(j->out_r[pixel], colour.x, scale);
        j->out_g[pixel] = add16(j->out_g[pixel], colour.y, scale);
        j->out_b[pixel] = add16(j->out_b[pixel], colour.z, scale);
    }
}

/* ---- flat entry points for ctypes --------------------------------------------------------------------------------- */
typedef struct {
    const port_job* job;
    uint32_t row_begin, row_end, row_step;
    volatile uint32_t* next_row;
} port_worker;

static void* port_worker_main(void* arg) {
    port_worker* w = (port_worker*)arg;
    for (;;) {
        const uint32_t row = __sync_fetch_and_add(w->next_row, w->row_step);
        uint32_t x, s;
        if (row >= w->row_end) break;
        for (x = 0; x < w->job->width; ++x)
            for (s = 1; s <= w->job->samples; ++s) port_pixel_sample(w->job, row * w->job->width + x, s);
    }
    return 0;
}

/* Renders rows row_begin, row_begin + row_step, ... below row_end into planes that the caller zeroed; `threads` host threads. */
void port_render_rows_step(const port_job* job, uint32_t row_begin, uint32_t row_end, uint32_t row_step, int threads) {
    volatile uint32_t next = row_begin;
    port_worker w;
    pthread_t tid[256];
    int i;
    if (row_end > job->height) row_end = job->height;
    w.job = job; w.row_begin = row_begin; w.row_end = row_end; w.row_step = row_step ? row_step : 1u; w.next_row = &next;
    if (threads > 256) threads = 256;
    if (threads <= 1) { port_worker_main(&w); return; }
    for (i = 0; i < threads; ++i) pthread_create(&tid[i], 0, port_worker_main, &w);
    for (i = 0; i < threads; ++i) pthread_join(tid[i], 0);
}

void port_render_rows(const port_job* job, uint32_t row_begin, uint32_t row_end, int threads) {
    port_render_rows_step(job, row_begin, row_end, 1u, threads);
}

size_t port_job_size(void) { return sizeof(port_job); }

/* KAT doors */
float port_kat_rand(uint64_t* state, float lo, float hi) { return port_rand(state, lo, hi); }
int port_kat_hit(const float* o, const float* d, float lo, float hi, const float* a, const float* b, const float* c, float* out3) {
    return port_hit_triangle(load3(o), load3(d), lo, hi, load3(a), load3(b), load3(c), out3, out3 + 1, out3 + 2);
}
void port_kat_ball(uint64_t* state, float radius, float* out3) {
    const v3 p = port_ball_sample(state, radius);
    out3[0] = p.x; out3[1] = p.y; out3[2] = p.z;
}
