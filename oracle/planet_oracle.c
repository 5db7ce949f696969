/* TEST INFRASTRUCTURE ONLY -- see planet_oracle.h for scope and parity status.
 *
 * CPU restatement of the pgcomp/planet terrain hot path in plain C.  Build with
 * `-O2 -ffp-contract=off` (oracle/Makefile): the reference's own results are
 * defined by unfused IEEE arithmetic (its build.bat uses plain g++; SURVEY.md
 * section 8c), and FMA contraction changes low bits of fade and lerp.
 */
#include "planet_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* perlin.h:10-36 -- the two constant tables (data, regenerated from the      */
/* compiled reference by oracle/gen_golden.py and checked in the tests)       */
/* ------------------------------------------------------------------------- */
static const unsigned char orc_table[256] = {
    0xd3, 0xde, 0x5a, 0x2a, 0x88, 0x25, 0xcc, 0x7e, 0x16, 0x65, 0xd5, 0x89, 0xfb, 0x1c, 0xf7, 0xcd,
    0xb9, 0xb0, 0xc8, 0xce, 0xf3, 0x82, 0xfc, 0xbc, 0x13, 0xeb, 0xe7, 0x01, 0xaa, 0x6d, 0x0b, 0x1f,
    0x3a, 0x86, 0xe6, 0x94, 0x41, 0xb8, 0xfa, 0xe2, 0x81, 0xc5, 0x87, 0x63, 0xc9, 0x05, 0x28, 0xdc,
    0x84, 0xda, 0x0f, 0x6e, 0x78, 0xef, 0x97, 0x23, 0x8d, 0x46, 0xd9, 0x07, 0x6b, 0x96, 0xb2, 0xa2,
    0xa0, 0x5d, 0xa4, 0x76, 0xae, 0x1d, 0x2d, 0x54, 0xcf, 0x51, 0x08, 0x40, 0x2b, 0xf4, 0xcb, 0x43,
    0x5f, 0x19, 0x45, 0x03, 0xb7, 0xf2, 0x5e, 0xac, 0x79, 0x90, 0x7a, 0xf9, 0x3d, 0x9f, 0xf0, 0x3b,
    0xc1, 0x9d, 0xe0, 0x34, 0x47, 0x70, 0x20, 0xa7, 0x9b, 0xa5, 0xb1, 0xff, 0x4e, 0x0a, 0x1a, 0x95,
    0x7c, 0x85, 0x8c, 0xbd, 0xe9, 0x3c, 0x60, 0xfe, 0x32, 0xec, 0x83, 0xd7, 0x31, 0x4f, 0x36, 0xd6,
    0xc4, 0x68, 0xea, 0x12, 0xb5, 0x35, 0x98, 0x74, 0x7f, 0x1e, 0xb6, 0x06, 0x62, 0x92, 0xd0, 0x66,
    0xdd, 0xf1, 0x30, 0xe4, 0x49, 0x52, 0xf5, 0x8e, 0x69, 0x50, 0x22, 0xf6, 0x17, 0x8b, 0xee, 0x61,
    0x33, 0xbe, 0xba, 0xe8, 0x2c, 0x5b, 0x57, 0xad, 0x10, 0xa8, 0x2e, 0x4b, 0xc7, 0x8a, 0xc6, 0x21,
    0x18, 0x42, 0xe1, 0xc3, 0xa9, 0x64, 0x58, 0xed, 0x26, 0x39, 0x00, 0x04, 0x56, 0x0e, 0xfd, 0x73,
    0x2f, 0xd4, 0xb4, 0xab, 0xa3, 0x3f, 0xc2, 0xe3, 0xd2, 0x3e, 0x0c, 0x59, 0xa1, 0xc0, 0x27, 0xa6,
    0x80, 0x7b, 0x11, 0xdf, 0x6a, 0x75, 0xe5, 0x6c, 0x4c, 0x91, 0x7d, 0xdb, 0xaf, 0x24, 0xca, 0x72,
    0x99, 0x48, 0xd1, 0x1b, 0x53, 0x55, 0x0d, 0x44, 0x93, 0x9e, 0xbb, 0xb3, 0x9c, 0x9a, 0x38, 0x4d,
    0x14, 0x8f, 0x77, 0x67, 0x71, 0xbf, 0x09, 0x29, 0x4a, 0xd8, 0x02, 0x6f, 0x15, 0x5c, 0xf8, 0x37
};

/* perlin.h:30-36: twelve cube-edge directions, rows 12..15 repeat rows 0, 1, 9, 11 */
static const float orc_vectors[16][3] = {
    { 1,  1,  0}, {-1,  1,  0}, { 1, -1,  0}, {-1, -1,  0},
    { 1,  0,  1}, {-1,  0,  1}, { 1,  0, -1}, {-1,  0, -1},
    { 0,  1,  1}, { 0, -1,  1}, { 0,  1, -1}, { 0, -1, -1},
    { 1,  1,  0}, {-1,  1,  0}, { 0, -1,  1}, { 0, -1, -1}
};

void orc_perlin_tables(unsigned char *table256, float *vectors48)
{
    memcpy(table256, orc_table, 256);
    memcpy(vectors48, orc_vectors, sizeof orc_vectors);
}

/* perlin.h:38-41 -- two's-complement mask, so negative seeds are valid */
int orc_perlin_random(int seed)
{
    return orc_table[seed & 255];
}

/* perlin.h:43-48 -- three chained table lookups, then a dot with a {0,+-1} vector.
 * The sum is evaluated left to right in float: (x*v0 + y*v1) + z*v2. */
float orc_perlin_gradient(float x, float y, float z, int ix, int iy, int iz)
{
    int h = orc_perlin_random(orc_perlin_random(orc_perlin_random(ix) + iy) + iz);
    const float *g = orc_vectors[h & 15];
    float s = x * g[0];
    s = s + y * g[1];
    s = s + z * g[2];
    return s;
}

/* perlin.h:52-55 -- NOT floor(): truncation of (x-1) for negative x, so a negative
 * integer coordinate lands one cell lower with fraction exactly 1.0 */
static int orc_cell(double x)
{
    return (int)((x < 0.0) ? (x - 1.0) : x);
}

/* perlin.h:62 -- quintic fade with float literals promoted to double; the whole
 * polynomial is double, left-associated, rounded to float once on assignment */
static float orc_fade(double t)
{
    double c = (t * 6.0f - 15.0f) * t + 10.0f;
    c = c * t;
    c = c * t;
    c = c * t;
    return (float)c;
}

static float orc_lerp(float a, float b, float t)   /* perlin.h:77 */
{
    float d = b - a;
    float m = d * t;
    return a + m;
}

/* perlin.h:50-88 */
float orc_perlin_noise3(double x, double y, double z)
{
    int ix = orc_cell(x), iy = orc_cell(y), iz = orc_cell(z);

    x -= ix;            /* perlin.h:58-60: fraction stays double */
    y -= iy;
    z -= iz;

    float u = orc_fade(x), v = orc_fade(y), w = orc_fade(z);

    /* perlin.h:68-75: `x - 1` is formed in double and only then narrowed to the
     * float parameter -- float(x_d - 1.0), not float(x_d) - 1.0f */
    float x0 = (float)x, x1 = (float)(x - 1);
    float y0 = (float)y, y1 = (float)(y - 1);
    float z0 = (float)z, z1 = (float)(z - 1);

    float g0 = orc_perlin_gradient(x0, y0, z0, ix,     iy,     iz);
    float g1 = orc_perlin_gradient(x1, y0, z0, ix + 1, iy,     iz);
    float g2 = orc_perlin_gradient(x0, y1, z0, ix,     iy + 1, iz);
    float g3 = orc_perlin_gradient(x1, y1, z0, ix + 1, iy + 1, iz);
    float g4 = orc_perlin_gradient(x0, y0, z1, ix,     iy,     iz + 1);
    float g5 = orc_perlin_gradient(x1, y0, z1, ix + 1, iy,     iz + 1);
    float g6 = orc_perlin_gradient(x0, y1, z1, ix,     iy + 1, iz + 1);
    float g7 = orc_perlin_gradient(x1, y1, z1, ix + 1, iy + 1, iz + 1);

    float a0 = orc_lerp(g0, g1, u), a1 = orc_lerp(g2, g3, u);   /* perlin.h:78-81 */
    float a2 = orc_lerp(g4, g5, u), a3 = orc_lerp(g6, g7, u);
    float b0 = orc_lerp(a0, a1, v), b1 = orc_lerp(a2, a3, v);   /* perlin.h:83-84 */
    return orc_lerp(b0, b1, w);                                 /* perlin.h:86 */
}

void orc_noise3_batch(const double *xyz, long n, float *out)
{
    for (long i = 0; i < n; i++)
        out[i] = orc_perlin_noise3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

/* main.cpp:689-707 -- frequency double, amplitude/value float, octave 0 first */
float orc_perlin_fbm(double x, double y, double z, double lacunarity, float gain, int octaves)
{
    double frequency = 1.0;
    float amplitude = 1.0f, value = 0.0f;
    for (int i = 0; i < octaves; ++i) {
        float n = orc_perlin_noise3(x * frequency, y * frequency, z * frequency);
        float t = n * amplitude;
        value = value + t;
        frequency *= lacunarity;
        amplitude = amplitude * gain;
    }
    return value;
}

/* main.cpp:709-734 -- (1-|n|)^2, weighted by the previous octave's value, no clamp */
float orc_perlin_ridged(double x, double y, double z, double lacunarity, float gain, int octaves)
{
    const float offset = 1.0f;
    double frequency = 1.0;
    float amplitude = 1.0f, weight = 1.0f, value = 0.0f;
    for (int i = 0; i < octaves; ++i) {
        float v = orc_perlin_noise3(x * frequency, y * frequency, z * frequency);
        v = (v < 0.0f) ? -v : v;
        v = offset - v;
        v = v * v;
        float t = v * amplitude;     /* main.cpp:727: (v*amplitude)*weight */
        t = t * weight;
        value = value + t;
        weight = v;
        frequency *= lacunarity;
        amplitude = amplitude * gain;
    }
    return value;
}

void orc_fractal_batch(const double *xyz, long n, int kind, double lacunarity, float gain,
                       int octaves, float *out)
{
    for (long i = 0; i < n; i++) {
        const double *p = xyz + 3 * i;
        out[i] = (kind == ORC_RIDGED) ? orc_perlin_ridged(p[0], p[1], p[2], lacunarity, gain, octaves)
                                      : orc_perlin_fbm(p[0], p[1], p[2], lacunarity, gain, octaves);
    }
}

/* ------------------------------------------------------------------------- */
/* height functor + generator seam                                            */
/* ------------------------------------------------------------------------- */
void orc_default_height_params(orc_height_params *p)
{
    p->kind = ORC_RIDGED;        /* main.cpp:829 */
    p->lacunarity = 2.0;
    p->gain = 0.55f;
    p->fixed_octaves = 0;        /* main.cpp:827 */
    p->coord_scale = 0.00001;    /* main.cpp:828 */
    p->height_scale = 8848.0f;   /* main.cpp:831 */
}

/* main.cpp:825-832 (Perlin::operator()) and :837-840 (ConstantZero) */
static float orc_height(const orc_height_params *hp, orc_vec3d p, int depth, int max_depth)
{
    if (hp->kind == ORC_ZERO) return 0.0f;
    int octaves = (hp->fixed_octaves > 0) ? hp->fixed_octaves : 6 + 12 * depth / max_depth;
    p.x *= hp->coord_scale; p.y *= hp->coord_scale; p.z *= hp->coord_scale;
    float h = (hp->kind == ORC_RIDGED)
        ? orc_perlin_ridged(p.x, p.y, p.z, hp->lacunarity, hp->gain, octaves)
        : orc_perlin_fbm(p.x, p.y, p.z, hp->lacunarity, hp->gain, octaves);
    return h * hp->height_scale;
}

float orc_get_height_at(const orc_height_params *hp, const double *p, int depth, int max_depth)
{
    orc_vec3d v = { p[0], p[1], p[2] };          /* main.cpp:118-121 */
    return orc_height(hp, v, depth, max_depth);
}

static orc_vec3d v_add(orc_vec3d a, orc_vec3d b) { orc_vec3d r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static orc_vec3d v_sub(orc_vec3d a, orc_vec3d b) { orc_vec3d r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static orc_vec3d v_mul(orc_vec3d a, double s)    { orc_vec3d r = { a.x * s, a.y * s, a.z * s }; return r; }
static orc_vec3d v_div(orc_vec3d a, double s)    { orc_vec3d r = { a.x / s, a.y / s, a.z / s }; return r; }
/* vec3.h:46-49: Dot is (x*x + y*y) + z*z; Normalize divides each component by the length */
static orc_vec3d v_normalize(orc_vec3d v)
{
    double d = v.x * v.x + v.y * v.y;
    d = d + v.z * v.z;
    return v_div(v, sqrt(d));
}

/* main.cpp:123-151 -- bilinear sample positions on the flat quad, one texel of
 * border all round: u = (x-1)/(dim-3) */
void orc_generate_height_map(const orc_height_params *hp, float *data, int dim,
                             const orc_quad *q, int max_depth)
{
    int depth = (int)orc_get_depth(q->id);
    orc_vec3d v0 = v_sub(q->p[1], q->p[0]);
    orc_vec3d v1 = v_sub(q->p[3], q->p[2]);
    double div = 1.0 / (dim - 3);
    for (int y = 0; y < dim; y++) {
        for (int x = 0; x < dim; x++) {
            double u = (x - 1) * div;
            double v = (y - 1) * div;
            orc_vec3d p0 = v_add(q->p[0], v_mul(v0, u));
            orc_vec3d p1 = v_add(q->p[2], v_mul(v1, u));
            orc_vec3d v2 = v_sub(p1, p0);
            orc_vec3d p = v_add(p0, v_mul(v2, v));
            data[y * dim + x] = orc_height(hp, p, depth, max_depth);
        }
    }
}

typedef struct {
    const orc_height_params *hp; const orc_quad *quads; long nquads;
    int dim, max_depth, tid, nthreads; float *out;
} orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    size_t per = (size_t)j->dim * j->dim;
    for (long i = j->tid; i < j->nquads; i += j->nthreads)
        orc_generate_height_map(j->hp, j->out + i * per, j->dim, j->quads + i, j->max_depth);
    return 0;
}

/* the reference calls GenerateHeightMap once per new leaf quad (main.cpp:244);
 * patches are independent, so nthreads > 1 stripes them over host threads */
void orc_generate_height_maps(const orc_height_params *hp, const orc_quad *quads, long nquads,
                              int dim, int max_depth, float *out, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; orc_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        orc_job j = { hp, quads, nquads, dim, max_depth, t, nthreads, out };
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], 0, orc_worker, &jobs[t]);
    }
    orc_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], 0);
}

/* ------------------------------------------------------------------------- */
/* QuadID (main.cpp:19-65): bit 63 valid, 60-62 root face, 55-59 depth,       */
/* bits 0-54 path with 2 bits per level, level-1 child in the lowest bits     */
/* ------------------------------------------------------------------------- */
#define ORC_BITS(v, first, count) (((v) >> (uint64_t)(first)) & ((((uint64_t)1) << (uint64_t)(count)) - 1))

uint64_t orc_get_root(uint64_t id)  { return ORC_BITS(id, 60, 3); }
uint64_t orc_get_depth(uint64_t id) { return ORC_BITS(id, 55, 5); }
uint64_t orc_get_index(uint64_t id) { return ORC_BITS(id, 0, 55); }

uint64_t orc_make_root_id(uint64_t root)
{
    return ((uint64_t)1 << 63) | (root << 60);
}

uint64_t orc_make_child_id(uint64_t id, uint64_t child)
{
    uint64_t depth = orc_get_depth(id);
    return (id + ((uint64_t)1 << 55)) | (child << (2 * depth));
}

uint64_t orc_get_child_index(uint64_t id)
{
    uint64_t depth = orc_get_depth(id);
    return (id >> (2 * (depth - 1))) & 3;
}

uint64_t orc_get_parent_id(uint64_t id)
{
    uint64_t depth = orc_get_depth(id);
    uint64_t mask = ~((uint64_t)3 << (2 * (depth - 1)));
    return (id - ((uint64_t)1 << 55)) & mask;
}

/* ------------------------------------------------------------------------- */
/* cube -> sphere subdivision geometry                                        */
/* ------------------------------------------------------------------------- */
/* main.cpp:604-624: eight normalised cube corners times radius; QUAD(a,b,c,d)
 * stores {v[a], v[b], v[d], v[c]} -- note the c/d swap at main.cpp:605 */
void orc_root_quads(double radius, orc_quad *out6)
{
    static const double c[8][3] = {
        {-1, -1, -1}, { 1, -1, -1}, { 1,  1, -1}, {-1,  1, -1},
        {-1, -1,  1}, { 1, -1,  1}, { 1,  1,  1}, {-1,  1,  1}
    };
    static const int f[6][4] = {
        {0, 1, 2, 3}, {1, 5, 6, 2}, {5, 4, 7, 6}, {4, 0, 3, 7}, {3, 2, 6, 7}, {4, 5, 1, 0}
    };
    orc_vec3d v[8];
    for (int i = 0; i < 8; i++) {
        orc_vec3d t = { c[i][0], c[i][1], c[i][2] };
        v[i] = v_mul(v_normalize(t), radius);
    }
    for (int r = 0; r < 6; r++) {
        out6[r].p[0] = v[f[r][0]];
        out6[r].p[1] = v[f[r][1]];
        out6[r].p[2] = v[f[r][3]];
        out6[r].p[3] = v[f[r][2]];
        out6[r].id = orc_make_root_id((uint64_t)r);
    }
}

/* main.cpp:546-547 (centre) and :581-594 (edge midpoints, child order) */
void orc_split_quad(double radius, const orc_quad *q, orc_quad *out4)
{
    orc_vec3d s = v_add(v_add(v_add(q->p[0], q->p[1]), q->p[2]), q->p[3]);
    orc_vec3d mid = v_mul(v_normalize(s), radius);
    orc_vec3d g[9];
    g[0] = q->p[0];
    g[1] = v_mul(v_normalize(v_add(q->p[0], q->p[1])), radius);
    g[2] = q->p[1];
    g[3] = v_mul(v_normalize(v_add(q->p[0], q->p[2])), radius);
    g[4] = mid;
    g[5] = v_mul(v_normalize(v_add(q->p[1], q->p[3])), radius);
    g[6] = q->p[2];
    g[7] = v_mul(v_normalize(v_add(q->p[2], q->p[3])), radius);
    g[8] = q->p[3];
    static const int k[4][4] = { {0, 1, 3, 4}, {1, 2, 4, 5}, {3, 4, 6, 7}, {4, 5, 7, 8} };
    for (int c = 0; c < 4; c++) {
        for (int j = 0; j < 4; j++) out4[c].p[j] = g[k[c][j]];
        out4[c].id = orc_make_child_id(q->id, (uint64_t)c);
    }
}

static void orc_uniform_rec(double radius, const orc_quad *q, int levels, orc_quad *out, long *n)
{
    if (levels == 0) { out[(*n)++] = *q; return; }
    orc_quad kids[4];
    orc_split_quad(radius, q, kids);
    for (int c = 0; c < 4; c++) orc_uniform_rec(radius, &kids[c], levels - 1, out, n);
}

/* every quad of one root face at uniform depth, in the depth-first order the
 * reference's recursion (main.cpp:589-592) appends leaves */
long orc_uniform_quads(double radius, int face, int depth, orc_quad *out)
{
    orc_quad roots[6];
    long n = 0;
    orc_root_quads(radius, roots);
    orc_uniform_rec(radius, &roots[face], depth, out, &n);
    return n;
}

/* walk the id's path from its root: level l's child index sits at bits 2(l-1) */
int orc_quad_from_id(double radius, uint64_t id, orc_quad *out)
{
    if (!(id >> 63)) return 0;
    uint64_t root = orc_get_root(id), depth = orc_get_depth(id);
    if (root >= 6) return 0;
    orc_quad roots[6], q, kids[4];
    orc_root_quads(radius, roots);
    q = roots[root];
    for (uint64_t l = 0; l < depth; l++) {
        orc_split_quad(radius, &q, kids);
        q = kids[(id >> (2 * l)) & 3];
    }
    *out = q;
    return q.id == id;
}

/* ------------------------------------------------------------------------- */
/* patch mesh (main.cpp:391-474)                                              */
/* ------------------------------------------------------------------------- */
int orc_patch_vertex_count(int n) { return n * n + 4 * n; }           /* main.cpp:393-394 */
int orc_patch_index_count(int n)                                       /* main.cpp:395-400 */
{
    int quads = n - 1, per_reset = 2;
    int per_strip = 2 + quads * 2 + per_reset;
    int skirt = quads * 4 + 2 * per_strip;
    return quads * per_strip - per_reset + skirt;
}

/* main.cpp:402-425: skirt row, then per row skirt + n verts + skirt, then skirt row.
 * div is double; each product is narrowed to float by V3(float,float,float). */
void orc_patch_vertices(int n, float *o)
{
    double div = 1.0 / (n - 1);
    int k = 0;
    for (int x = 0; x < n; ++x) { o[k++] = (float)(x * div); o[k++] = 0.0f; o[k++] = 1.0f; }
    for (int y = 0; y < n; ++y) {
        o[k++] = 0.0f; o[k++] = (float)(y * div); o[k++] = 1.0f;
        for (int x = 0; x < n; ++x) { o[k++] = (float)(x * div); o[k++] = (float)(y * div); o[k++] = 0.0f; }
        o[k++] = 1.0f; o[k++] = (float)(y * div); o[k++] = 1.0f;
    }
    for (int x = 0; x < n; ++x) { o[k++] = (float)(x * div); o[k++] = 1.0f; o[k++] = 1.0f; }
}

/* main.cpp:427-474: one triangle strip, degenerate pairs between rows; the two
 * running cursors advance asymmetrically around the skirt rows */
void orc_patch_indices(int n, uint32_t *o)
{
    int k = 0, quads = n - 1;
    uint32_t v0 = 0, v1 = (uint32_t)n + 1;
    for (int x = 0; x < n; ++x) { o[k++] = v0++; o[k++] = v1++; }
    o[k++] = v1 - 1; o[k++] = v0; v1++;
    for (int y = 0; y < quads; ++y) {
        for (int x = 0; x < n + 2; ++x) { o[k++] = v0++; o[k++] = v1++; }
        if (y + 1 < quads) { o[k++] = v1 - 1; o[k++] = v0; }
    }
    v0++;
    o[k++] = v1 - 1; o[k++] = v0;
    for (int x = 0; x < n; ++x) { o[k++] = v0++; o[k++] = v1++; }
}

#define ORC_PI 3.1415926535897932384626433832795      /* math.h:7 */

int orc_max_lod(double radius, int n)                  /* main.cpp:497 */
{
    return (int)(log2(2.0 * ORC_PI * radius / (n - 1)) - 2);
}

float orc_max_skirt_size(double radius, int n)         /* main.cpp:500 */
{
    return (float)((2 * ORC_PI * radius) / (4 * (n - 1)) * 0.00001 * 8 * 8848.0);
}

float orc_skirt_size_for_quad(float max_skirt, uint64_t id)   /* main.cpp:674-677 */
{
    float s = max_skirt;
    int depth = (int)orc_get_depth(id) - 1;
    if (depth > 0) s /= (float)(2 << depth);
    return s;
}

/* ------------------------------------------------------------------------- */
/* GLSL stage restated in fp32 (main.cpp:286-380) -- parity unpinned          */
/* ------------------------------------------------------------------------- */
typedef struct { float x, y, z; } f3;
typedef struct { f3 p, n; } glsl_V;                    /* main.cpp:298 */

static f3 f3_add(f3 a, f3 b) { f3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static f3 f3_sub(f3 a, f3 b) { f3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static f3 f3_scale(f3 a, float s) { f3 r = { a.x * s, a.y * s, a.z * s }; return r; }
static float f3_dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static float f3_length(f3 a) { return sqrtf(f3_dot(a, a)); }
static f3 f3_normalize(f3 a) { float l = f3_length(a); f3 r = { a.x / l, a.y / l, a.z / l }; return r; }
static f3 f3_cross(f3 a, f3 b)
{
    f3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
/* GLSL mix(x, y, a) = x*(1-a) + y*a */
static f3 f3_mix(f3 a, f3 b, float t) { return f3_add(f3_scale(a, 1.0f - t), f3_scale(b, t)); }

static glsl_V glsl_interpolate_linear(glsl_V v0, glsl_V v1, float t)   /* main.cpp:300-308 */
{
    glsl_V r;
    r.n = f3_normalize(f3_mix(v0.n, v1.n, t));
    r.p = f3_mix(v0.p, v1.p, t);
    return r;
}

static glsl_V glsl_interpolate(glsl_V v0, glsl_V v1, float t)          /* main.cpp:310-332 */
{
    float d = f3_dot(v0.n, v1.n);
    if (1.0f - d < 0.001f) return glsl_interpolate_linear(v0, v1, t);

    float theta2 = acosf(d);
    float k = 1.0f - t;
    f3 n = f3_normalize(f3_add(f3_scale(v0.n, sinf(k * theta2)), f3_scale(v1.n, sinf(t * theta2))));

    float theta = theta2 * 0.5f;
    float gamma = theta - theta2 * t;
    float tan_theta = tanf(theta);
    float x = 1.0f - tanf(gamma) / tan_theta;
    float y = 1.0f / sinf(theta) - 1.0f / (cosf(gamma) * tan_theta);
    f3 v = f3_scale(f3_sub(v1.p, v0.p), 0.5f);
    glsl_V r;
    r.p = f3_add(f3_add(v0.p, f3_scale(v, x)), f3_scale(n, y * f3_length(v)));
    r.n = n;
    return r;
}

/* One shader invocation.  UV = (ux, uy, skirt) is the vertex attribute exactly as
 * the patch vertex buffer holds it; (tx, ty) is the texel the sampler coordinate
 * lands on.  With the quad's own height map, mix(corners0, corners1, UV.xy) =
 * ((1.5+vx)/dim, (1.5+vy)/dim) is the centre of texel (vx+1, vy+1)
 * (main.cpp:196-199, 358), so GL_LINEAR filtering reduces to a direct read and
 * the four normal taps are the 4-neighbours (main.cpp:338-346). */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* texture(HeightMap, uv).r with GL_LINEAR / GL_CLAMP_TO_EDGE (render.cpp:429-433) in exact
 * fp32 (real GPUs quantise the weights to 8 bits; unpinned like the rest of the shader) */
static float glsl_texture(const float *H, int dim, float u, float v)
{
    float x = u * (float)dim - 0.5f, y = v * (float)dim - 0.5f;
    float x0 = floorf(x), y0 = floorf(y);
    float fx = x - x0, fy = y - y0;
    int ix0 = clampi((int)x0, 0, dim - 1), ix1 = clampi((int)x0 + 1, 0, dim - 1);
    int iy0 = clampi((int)y0, 0, dim - 1), iy1 = clampi((int)y0 + 1, 0, dim - 1);
    float a = H[iy0 * dim + ix0], b = H[iy0 * dim + ix1], c = H[iy1 * dim + ix0], d = H[iy1 * dim + ix1];
    float top = a + (b - a) * fx, bot = c + (d - c) * fx;
    return top + (bot - top) * fy;
}

/* rect == NULL: the quad's own map, direct texel reads.  Otherwise rect = {corners0.xy,
 * corners1.xy, pixel_size.xy} as GetHeightMapForQuad returns them (main.cpp:196-199, 214-235). */
static void glsl_vertex(const glsl_V c[4], const float *H, int dim, int n,
                        float ux, float uy, float skirt, int tx, int ty,
                        float skirt_size, const float *rect, float *pos4, float *nrm4)
{
    glsl_V p = glsl_interpolate(c[0], c[1], ux);              /* main.cpp:354 */
    glsl_V q = glsl_interpolate(c[2], c[3], ux);              /* main.cpp:355 */
    glsl_V v = glsl_interpolate(p, q, uy);                    /* main.cpp:356 */

    float hc, x0, x1, y0, y1;
    if (rect) {                                                /* main.cpp:339-344, 358 */
        float tu = rect[0] * (1.0f - ux) + rect[2] * ux, tv = rect[1] * (1.0f - uy) + rect[3] * uy;
        hc = glsl_texture(H, dim, tu, tv);
        x0 = glsl_texture(H, dim, tu - rect[4], tv); x1 = glsl_texture(H, dim, tu + rect[4], tv);
        y0 = glsl_texture(H, dim, tu, tv - rect[5]); y1 = glsl_texture(H, dim, tu, tv + rect[5]);
    } else {
        hc = H[ty * dim + tx];
        x0 = H[ty * dim + tx - 1]; x1 = H[ty * dim + tx + 1];
        y0 = H[(ty - 1) * dim + tx]; y1 = H[(ty + 1) * dim + tx];
    }
    float height = hc - skirt_size * skirt;                    /* main.cpp:360 */
    f3 pq = f3_sub(q.p, p.p);
    float xyscale = f3_length(pq) / (float)(n - 1);            /* main.cpp:361 (29.0 = n-1) */
    f3 nt = { x0 - x1, 2.0f * xyscale, y0 - y1 };
    nt = f3_normalize(nt);                                     /* main.cpp:345 */

    f3 nn = v.n;
    f3 t = f3_normalize(f3_cross(nn, pq));                     /* main.cpp:363 */
    f3 bi = f3_normalize(f3_cross(t, nn));                     /* main.cpp:364 */
    /* mat3(t, n, bi) * normal: columns t, n, bi (main.cpp:365) */
    f3 N = f3_add(f3_add(f3_scale(t, nt.x), f3_scale(nn, nt.y)), f3_scale(bi, nt.z));
    N = f3_normalize(N);

    f3 pos = f3_add(v.p, f3_scale(v.n, height));               /* main.cpp:366 */
    pos4[0] = pos.x; pos4[1] = pos.y; pos4[2] = pos.z; pos4[3] = height;

    /* fragment stage evaluated at the vertex (main.cpp:373-380) */
    f3 l = { 0.0f, 1.0f, -1.0f };
    l = f3_normalize(l);
    float lambert = f3_dot(N, l);
    float light = 0.001f + (lambert > 0.0f ? lambert : 0.0f);
    nrm4[0] = N.x; nrm4[1] = N.y; nrm4[2] = N.z; nrm4[3] = sqrtf(light);
}

/* the per-quad uniforms of a draw, main.cpp:666-672: P[j] = float(q.p[j] - cam), N[j] =
 * float(Normalize(q.p[j])).  PINNED: tests compare them bit for bit with the values the reference
 * passes to glUniform (captured by oracle/ref_oracle.cpp's recording GL). */
void orc_quad_uniforms(const orc_quad *q, const double *cam_pos, float *PN24)
{
    orc_vec3d cam = { cam_pos[0], cam_pos[1], cam_pos[2] };
    for (int j = 0; j < 4; j++) {
        orc_vec3d rel = v_sub(q->p[j], cam);
        orc_vec3d nd = v_normalize(q->p[j]);
        PN24[3 * j] = (float)rel.x; PN24[3 * j + 1] = (float)rel.y; PN24[3 * j + 2] = (float)rel.z;
        PN24[12 + 3 * j] = (float)nd.x; PN24[12 + 3 * j + 1] = (float)nd.y; PN24[12 + 3 * j + 2] = (float)nd.z;
    }
}

static void shade_patch(const orc_quad *q, const double *cam_pos, const float *heights,
                        int n, float skirt_size, const float *rect, float *pos4, float *nrm4)
{
    int dim = n + 2, nv = orc_patch_vertex_count(n);
    glsl_V c[4];
    float PN[24];
    orc_quad_uniforms(q, cam_pos, PN);
    for (int j = 0; j < 4; j++) {
        c[j].p.x = PN[3 * j]; c[j].p.y = PN[3 * j + 1]; c[j].p.z = PN[3 * j + 2];
        c[j].n.x = PN[12 + 3 * j]; c[j].n.y = PN[12 + 3 * j + 1]; c[j].n.z = PN[12 + 3 * j + 2];
    }
    float *uv = (float *)malloc(sizeof(float) * 3 * nv);
    int *tex = (int *)malloc(sizeof(int) * 2 * nv);
    orc_patch_vertices(n, uv);
    int k = 0;
    /* texel under each vertex, in the vertex order of main.cpp:406-422 */
    for (int x = 0; x < n; ++x, ++k) { tex[2 * k] = x + 1; tex[2 * k + 1] = 1; }
    for (int y = 0; y < n; ++y) {
        tex[2 * k] = 1; tex[2 * k + 1] = y + 1; ++k;
        for (int x = 0; x < n; ++x, ++k) { tex[2 * k] = x + 1; tex[2 * k + 1] = y + 1; }
        tex[2 * k] = n; tex[2 * k + 1] = y + 1; ++k;
    }
    for (int x = 0; x < n; ++x, ++k) { tex[2 * k] = x + 1; tex[2 * k + 1] = n; }
    for (k = 0; k < nv; k++)
        glsl_vertex(c, heights, dim, n, uv[3 * k], uv[3 * k + 1], uv[3 * k + 2],
                    tex[2 * k], tex[2 * k + 1], skirt_size, rect, pos4 + 4 * k, nrm4 + 4 * k);
    free(uv); free(tex);
}

void orc_shade_patch(const orc_quad *q, const double *cam_pos, const float *heights,
                     int n, float skirt_size, float *pos4, float *nrm4)
{
    shade_patch(q, cam_pos, heights, n, skirt_size, 0, pos4, nrm4);
}

/* the cache path: quad i samples pool slot slots[i] through rects[6*i..] (main.cpp:191-237) */
void orc_shade_patches_rect(const orc_quad *quads, long nquads, const double *cam_pos,
                            const float *pool, const int *slots, const float *rects, int n,
                            float max_skirt, float *pos4, float *nrm4)
{
    int dim = n + 2, nv = orc_patch_vertex_count(n);
    for (long i = 0; i < nquads; i++)
        shade_patch(quads + i, cam_pos, pool + (size_t)slots[i] * dim * dim, n,
                    orc_skirt_size_for_quad(max_skirt, quads[i].id), rects + 6 * i,
                    pos4 + (size_t)i * nv * 4, nrm4 + (size_t)i * nv * 4);
}

void orc_shade_patches(const orc_quad *quads, long nquads, const double *cam_pos,
                       const float *heights, int n, float max_skirt, float *pos4, float *nrm4)
{
    int dim = n + 2, nv = orc_patch_vertex_count(n);
    for (long i = 0; i < nquads; i++)
        orc_shade_patch(quads + i, cam_pos, heights + (size_t)i * dim * dim, n,
                        orc_skirt_size_for_quad(max_skirt, quads[i].id),
                        pos4 + (size_t)i * nv * 4, nrm4 + (size_t)i * nv * 4);
}
