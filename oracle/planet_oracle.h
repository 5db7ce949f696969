/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the pgcomp/planet terrain hot path.
 *
 * Plain C restatement (gcc, -O2 -ffp-contract=off) of the reference algorithm,
 * function by function, each citing the reference file:line it follows.  It is
 * the checker the CUDA path is compared against in tests/, in
 * __graft_entry__.smoke() and (timed, as `cpu_baseline.kind == "port"`) in
 * bench.py when oracle/_ref is not available.  Nothing in planet_b200/ or
 * include/ may include, link or load it; the product has no CPU path.
 *
 * Parity status:
 *   PINNED  (noise, fBm/ridged, height functor, GenerateHeightMap, QuadID,
 *            root faces, child split, patch vertex grid, strip index buffer):
 *            checked bit-for-bit against the reference's own source compiled
 *            here (oracle/_ref, see oracle/ref_oracle.cpp) and against the
 *            golden vectors that library produced (tests/golden/, written by
 *            oracle/gen_golden.py).  The reference ships no tests of its own.
 *   UNPINNED (orc_shade_*: displacement + normals + Lambert term): the
 *            reference does this in a GLSL 1.40 vertex/fragment shader
 *            (main.cpp:282-382); there is no CPU implementation to execute and
 *            no GL driver here.  The restatement follows the shader text line
 *            by line in fp32 -- "parity unpinned" for that part.  Everything
 *            the shader is FED is pinned: the vertex attributes (patch vertex
 *            buffer), the per-quad uniforms P/N/SkirtSize (orc_quad_uniforms,
 *            orc_skirt_size_for_quad) and the texture rects, each compared bit
 *            for bit with what the reference hands to GL.
 */
#ifndef PLANET_ORACLE_H
#define PLANET_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double x, y, z; } orc_vec3d;

/* main.cpp:68-72: struct Quad { Vec3d p[4]; QuadID id; } -- 104 bytes */
typedef struct { orc_vec3d p[4]; uint64_t id; } orc_quad;

enum { ORC_RIDGED = 0, ORC_FBM = 1, ORC_ZERO = 2 };

/* constants of Perlin::operator() (main.cpp:823-833) lifted into parameters */
typedef struct {
    int kind;             /* ORC_RIDGED (main.cpp:829, default) / ORC_FBM (main.cpp:830) */
    double lacunarity;    /* 2.0 */
    float gain;           /* 0.55f */
    int fixed_octaves;    /* <=0: 6 + 12*depth/max_depth (main.cpp:827) */
    double coord_scale;   /* 0.00001 (main.cpp:828) */
    float height_scale;   /* 8848.0f (main.cpp:831) */
} orc_height_params;

void orc_default_height_params(orc_height_params *p);

/* perlin.h */
void  orc_perlin_tables(unsigned char *table256, float *vectors48);
int   orc_perlin_random(int seed);
float orc_perlin_gradient(float x, float y, float z, int ix, int iy, int iz);
float orc_perlin_noise3(double x, double y, double z);
void  orc_noise3_batch(const double *xyz, long n, float *out);

/* main.cpp:689-734 */
float orc_perlin_fbm(double x, double y, double z, double lacunarity, float gain, int octaves);
float orc_perlin_ridged(double x, double y, double z, double lacunarity, float gain, int octaves);
void  orc_fractal_batch(const double *xyz, long n, int kind, double lacunarity, float gain,
                        int octaves, float *out);

/* main.cpp:107-158, 823-833 */
float orc_get_height_at(const orc_height_params *hp, const double *p, int depth, int max_depth);
void  orc_generate_height_map(const orc_height_params *hp, float *data, int dim,
                              const orc_quad *q, int max_depth);
void  orc_generate_height_maps(const orc_height_params *hp, const orc_quad *quads, long nquads,
                               int dim, int max_depth, float *out, int nthreads);

/* main.cpp:19-65 */
uint64_t orc_make_root_id(uint64_t root);
uint64_t orc_make_child_id(uint64_t id, uint64_t child);
uint64_t orc_get_parent_id(uint64_t id);
uint64_t orc_get_root(uint64_t id);
uint64_t orc_get_depth(uint64_t id);
uint64_t orc_get_index(uint64_t id);
uint64_t orc_get_child_index(uint64_t id);

/* main.cpp:546-547, 581-594, 604-624 */
void orc_root_quads(double radius, orc_quad *out6);
void orc_split_quad(double radius, const orc_quad *q, orc_quad *out4);
long orc_uniform_quads(double radius, int face, int depth, orc_quad *out);
int  orc_quad_from_id(double radius, uint64_t id, orc_quad *out);

/* main.cpp:391-474, 497, 500 */
int  orc_patch_vertex_count(int n);              /* n*n + 4n */
int  orc_patch_index_count(int n);               /* 2n^2 + 8n - 4 */
void orc_patch_vertices(int n, float *out_xyz);  /* (u, v, skirt) triples */
void orc_patch_indices(int n, uint32_t *out);
int   orc_max_lod(double radius, int n);
float orc_max_skirt_size(double radius, int n);
float orc_skirt_size_for_quad(float max_skirt, uint64_t id);   /* main.cpp:674-677 */

/* GLSL vertex + fragment stage (main.cpp:286-380), restated in fp32.  For every
 * vertex of the reference patch (order of orc_patch_vertices) of one quad:
 *   pos4  = (v.p + v.n*height, height)   camera-relative position, w = height
 *   nrm4  = (Normal, sqrt(0.001 + max(0, dot(Normal, l))))  w = Lambert colour
 * `heights` is the quad's own dim x dim map, dim == n + 2. */
/* per-quad draw uniforms P[4], N[4] (main.cpp:666-672) as 24 floats -- pinned against the
 * reference's captured glUniform values */
void orc_quad_uniforms(const orc_quad *q, const double *cam_pos, float *PN24);
void orc_shade_patch(const orc_quad *q, const double *cam_pos, const float *heights,
                     int n, float skirt_size, float *pos4, float *nrm4);
void orc_shade_patches(const orc_quad *quads, long nquads, const double *cam_pos,
                       const float *heights, int n, float max_skirt, float *pos4, float *nrm4);
/* same through the cache's texture rects: quad i samples pool slot slots[i] with GL_LINEAR /
 * CLAMP_TO_EDGE at mix(corners0, corners1, UV.xy); rects = 6 floats per quad
 * (corners0.xy, corners1.xy, pixel_size.xy) as GetHeightMapForQuad returns (main.cpp:191-237) */
void orc_shade_patches_rect(const orc_quad *quads, long nquads, const double *cam_pos,
                            const float *pool, const int *slots, const float *rects, int n,
                            float max_skirt, float *pos4, float *nrm4);

#ifdef __cplusplus
}
#endif
#endif
