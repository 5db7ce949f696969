// planet_common.cuh -- shared definitions of the sm_100a terrain kernels.
//
// Layout-compatible views of the reference's data (Quad 104 B, QuadID bit fields),
// the two constant tables of perlin.h, and the *exact* device restatement of
// PerlinNoise3 / PerlinfBm / PerlinRidged in which every IEEE rounding of the
// reference (perlin.h:50-88, main.cpp:689-734) is reproduced with _rn intrinsics,
// so nvcc's default FMA contraction cannot change a bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/planet_gpu.h"

namespace planet {

// ---- error plumbing (defined in planet_api.cu) ---------------------------------------
int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
void count_launch(int n = 1);
bool ensure_init();

#define PLANET_CUDA(call) do { int _rc = ::planet::check_cuda((call), #call); if (_rc) return _rc; } while (0)

// ---- perlin.h:10-36 -- constant tables -----------------------------------------------
// (values checked against the reference-built golden fixture in tests/)
#define PLANET_PERM_TABLE \
    0xd3, 0xde, 0x5a, 0x2a, 0x88, 0x25, 0xcc, 0x7e, 0x16, 0x65, 0xd5, 0x89, 0xfb, 0x1c, 0xf7, 0xcd, \
    0xb9, 0xb0, 0xc8, 0xce, 0xf3, 0x82, 0xfc, 0xbc, 0x13, 0xeb, 0xe7, 0x01, 0xaa, 0x6d, 0x0b, 0x1f, \
    0x3a, 0x86, 0xe6, 0x94, 0x41, 0xb8, 0xfa, 0xe2, 0x81, 0xc5, 0x87, 0x63, 0xc9, 0x05, 0x28, 0xdc, \
    0x84, 0xda, 0x0f, 0x6e, 0x78, 0xef, 0x97, 0x23, 0x8d, 0x46, 0xd9, 0x07, 0x6b, 0x96, 0xb2, 0xa2, \
    0xa0, 0x5d, 0xa4, 0x76, 0xae, 0x1d, 0x2d, 0x54, 0xcf, 0x51, 0x08, 0x40, 0x2b, 0xf4, 0xcb, 0x43, \
    0x5f, 0x19, 0x45, 0x03, 0xb7, 0xf2, 0x5e, 0xac, 0x79, 0x90, 0x7a, 0xf9, 0x3d, 0x9f, 0xf0, 0x3b, \
    0xc1, 0x9d, 0xe0, 0x34, 0x47, 0x70, 0x20, 0xa7, 0x9b, 0xa5, 0xb1, 0xff, 0x4e, 0x0a, 0x1a, 0x95, \
    0x7c, 0x85, 0x8c, 0xbd, 0xe9, 0x3c, 0x60, 0xfe, 0x32, 0xec, 0x83, 0xd7, 0x31, 0x4f, 0x36, 0xd6, \
    0xc4, 0x68, 0xea, 0x12, 0xb5, 0x35, 0x98, 0x74, 0x7f, 0x1e, 0xb6, 0x06, 0x62, 0x92, 0xd0, 0x66, \
    0xdd, 0xf1, 0x30, 0xe4, 0x49, 0x52, 0xf5, 0x8e, 0x69, 0x50, 0x22, 0xf6, 0x17, 0x8b, 0xee, 0x61, \
    0x33, 0xbe, 0xba, 0xe8, 0x2c, 0x5b, 0x57, 0xad, 0x10, 0xa8, 0x2e, 0x4b, 0xc7, 0x8a, 0xc6, 0x21, \
    0x18, 0x42, 0xe1, 0xc3, 0xa9, 0x64, 0x58, 0xed, 0x26, 0x39, 0x00, 0x04, 0x56, 0x0e, 0xfd, 0x73, \
    0x2f, 0xd4, 0xb4, 0xab, 0xa3, 0x3f, 0xc2, 0xe3, 0xd2, 0x3e, 0x0c, 0x59, 0xa1, 0xc0, 0x27, 0xa6, \
    0x80, 0x7b, 0x11, 0xdf, 0x6a, 0x75, 0xe5, 0x6c, 0x4c, 0x91, 0x7d, 0xdb, 0xaf, 0x24, 0xca, 0x72, \
    0x99, 0x48, 0xd1, 0x1b, 0x53, 0x55, 0x0d, 0x44, 0x93, 0x9e, 0xbb, 0xb3, 0x9c, 0x9a, 0x38, 0x4d, \
    0x14, 0x8f, 0x77, 0x67, 0x71, 0xbf, 0x09, 0x29, 0x4a, 0xd8, 0x02, 0x6f, 0x15, 0x5c, 0xf8, 0x37

// perlin.h:30-36, one row per (hash & 15): twelve cube-edge directions + 4 repeats
#define PLANET_GRAD_VECTORS \
    { 1,  1,  0}, {-1,  1,  0}, { 1, -1,  0}, {-1, -1,  0}, \
    { 1,  0,  1}, {-1,  0,  1}, { 1,  0, -1}, {-1,  0, -1}, \
    { 0,  1,  1}, { 0, -1,  1}, { 0,  1, -1}, { 0, -1, -1}, \
    { 1,  1,  0}, {-1,  1,  0}, { 0, -1,  1}, { 0, -1, -1}

// device copies in global memory (one per translation unit, no -rdc needed);
// kernels stage them into shared memory
#ifdef __CUDACC__
static __device__ const unsigned char g_perm[256] = { PLANET_PERM_TABLE };
static __device__ const float g_grad[16][3] = { PLANET_GRAD_VECTORS };
#endif

// ---- reference-layout views ---------------------------------------------------------
struct d3 { double x, y, z; };
struct Quad { d3 p[4]; uint64_t id; };            // main.cpp:68-72
static_assert(sizeof(Quad) == 104, "Quad must match the reference layout");
static_assert(sizeof(planet_gpu_quad) == 104, "C-ABI quad must match the reference layout");
static_assert(sizeof(planet_gpu_params) == 72, "planet_gpu_params layout is part of the ABI");

// main.cpp:24-28
__host__ __device__ inline uint64_t quad_root(uint64_t id)  { return (id >> 60) & 7; }
__host__ __device__ inline uint64_t quad_depth(uint64_t id) { return (id >> 55) & 31; }
__host__ __device__ inline uint64_t quad_index(uint64_t id) { return id & ((1ull << 55) - 1); }
// main.cpp:32-39, 41-49
__host__ __device__ inline uint64_t make_root_id(uint64_t root) { return (1ull << 63) | (root << 60); }
__host__ __device__ inline uint64_t make_child_id(uint64_t id, uint64_t child)
{
    return (id + (1ull << 55)) | (child << (2 * quad_depth(id)));
}

// main.cpp:827 -- integer arithmetic, (12*depth)/max_depth truncating
__host__ __device__ inline int octaves_for(int fixed_octaves, int depth, int max_depth)
{
    return fixed_octaves > 0 ? fixed_octaves : 6 + 12 * depth / max_depth;
}

// The functor constants, as passed by value to kernels.
struct HeightCfg {
    int kind;            // PLANET_NOISE_*
    int fixed_octaves;
    int max_depth;
    float gain;
    float height_scale;
    double lacunarity;
    double coord_scale;
    double seed[3];
    int has_seed;        // 0: seed is {0,0,0} and is not added (keeps -0.0 coordinates intact)
};

HeightCfg make_cfg(const planet_gpu_params *p, int max_depth);

// Fused gather: is quad q (index in the gathered buffer) one of the 1-in-`every` whose map the shade
// kernel pushes instead of the height-map kernel?  The q >> 2 term rotates the residue so that the
// pushing does not always fall on the same warp of the shade kernel's 4-warp CTAs.
// every >= 2: one map in `every`; every in [-7, -1]: -every maps in 8 (finer shares above one half)
__host__ __device__ inline bool shade_pushes_quad(int64_t q, int every)
{
    if (every < 0) return (int)(((q >> 3) + q) & 7) < -every;
    return every > 0 && ((q >> 2) + q) % every == every - 1;
}

// Fused gather (multi-GPU, K4): besides its own buffer a height-map kernel can store every finished
// tile straight into the same position of up to 7 peers' buffers (CUDA-IPC mapped, the stores
// travel over NVLink), so the all-gather of finished patches rides under the arithmetic instead of
// being a collective after the kernel.  `release` (optional): flags in LOCAL memory, one per rank,
// that the peers bump when they are done READING the buffer this launch is about to overwrite
// (k4_gather.cu); the kernel waits for release[r] >= release_min for every peer r before it
// computes.  `error`: set to 1 if that wait times out (a peer died); the kernel then runs on.
struct PeerOut {
    float *ptr[7];
    int n;
    const uint32_t *release;
    uint32_t release_min;
    int rank, world;
    uint32_t *error;
    // split of the pushing between K2 and K3: with k3_every > 0, quads whose index in the gathered
    // buffer is congruent to k3_every - 1 modulo k3_every are NOT pushed by the height-map kernel;
    // the shade kernel, which stages every map in shared memory anyway, pushes them as one 4 KB
    // bulk copy per peer (the NVLink transfer is then spread over both kernels)
    int64_t quad0;      // index in the gathered buffer of the launch's first quad
    int k3_every;
    // Progress mode (k4_gather.cu, PLANET_GATHER_PUSH_CONCURRENT): the kernel pushes nothing itself; every
    // warp publishes, after each finished 128-sample tile, (tag << 32 | tiles done) in progress[its index
    // in the grid] with release semantics, and a pusher kernel running BESIDE it on the same SMs sends the
    // finished tiles on.  n == 0 then; the release wait above still guards the peers' buffers.
    unsigned long long *progress;
    uint32_t progress_tag;
    uint32_t progress_mask;     // publish when (tiles done & mask) == 0 and at the warp's last tile (2^k - 1)
};
int validate_params(const planet_gpu_params *p);

#ifdef __CUDACC__
// Spin until a flag another GPU writes into this GPU's memory reaches `want` (wrap-safe compare).
// Bounded: after ~2 s without progress the peer is taken for dead, *error is set and the caller
// runs on (a hung kernel would take the whole box with it).
__device__ __forceinline__ void gather_wait_flag(const uint32_t *flag, uint32_t want, uint32_t *error)
{
    uint64_t t0 = 0;
    for (uint32_t spins = 0;; spins++) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - want) >= 0) return;
        if ((spins & 1023u) == 1023u) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            else if (now - t0 > 2000000000ull) { if (error) atomicExch(error, 1u); return; }
        }
    }
}

// =====================================================================================
// exact arithmetic: unfused IEEE ops, in the reference's evaluation order
// =====================================================================================
namespace exact {

__device__ __forceinline__ d3 add(d3 a, d3 b) { return { __dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z) }; }
__device__ __forceinline__ d3 sub(d3 a, d3 b) { return { __dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z) }; }
__device__ __forceinline__ d3 mul(d3 a, double s) { return { __dmul_rn(a.x, s), __dmul_rn(a.y, s), __dmul_rn(a.z, s) }; }
__device__ __forceinline__ d3 div(d3 a, double s) { return { __ddiv_rn(a.x, s), __ddiv_rn(a.y, s), __ddiv_rn(a.z, s) }; }
// vec3.h:46-49: Dot = (x*x + y*y) + z*z, Normalize = v / sqrt(Dot)
__device__ __forceinline__ d3 normalize(d3 v)
{
    double d = __dadd_rn(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)), __dmul_rn(v.z, v.z));
    return div(v, __dsqrt_rn(d));
}

// perlin.h:52-55 -- truncation of (x-1) for negative x; NOT floor()
__device__ __forceinline__ int cell(double x)
{
    return __double2int_rz((x < 0.0) ? __dadd_rn(x, -1.0) : x);
}

// perlin.h:62 -- quintic in double, left-associated, one rounding to float
__device__ __forceinline__ float fade(double t)
{
    double c = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(t, 6.0), -15.0), t), 10.0);
    c = __dmul_rn(c, t);
    c = __dmul_rn(c, t);
    c = __dmul_rn(c, t);
    return __double2float_rn(c);
}

__device__ __forceinline__ float lerp(float a, float b, float t)     // perlin.h:77
{
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

// perlin.h:43-48; `perm` and `grad` point at the shared-memory copies of the tables
__device__ __forceinline__ float gradient(const unsigned char *perm, const float *grad,
                                          float x, float y, float z, int ix, int iy, int iz)
{
    int h = perm[(perm[(perm[ix & 255] + iy) & 255] + iz) & 255];
    const float *g = grad + 3 * (h & 15);
    float s = __fmul_rn(x, g[0]);
    s = __fadd_rn(s, __fmul_rn(y, g[1]));
    s = __fadd_rn(s, __fmul_rn(z, g[2]));
    return s;
}

// perlin.h:50-88
__device__ __forceinline__ float noise3(const unsigned char *perm, const float *grad,
                                        double x, double y, double z)
{
    int ix = cell(x), iy = cell(y), iz = cell(z);
    x = __dsub_rn(x, (double)ix);
    y = __dsub_rn(y, (double)iy);
    z = __dsub_rn(z, (double)iz);
    float u = fade(x), v = fade(y), w = fade(z);
    // perlin.h:68-75: x-1 formed in double, then narrowed
    float x0 = __double2float_rn(x), x1 = __double2float_rn(__dadd_rn(x, -1.0));
    float y0 = __double2float_rn(y), y1 = __double2float_rn(__dadd_rn(y, -1.0));
    float z0 = __double2float_rn(z), z1 = __double2float_rn(__dadd_rn(z, -1.0));
    float g0 = gradient(perm, grad, x0, y0, z0, ix,     iy,     iz);
    float g1 = gradient(perm, grad, x1, y0, z0, ix + 1, iy,     iz);
    float g2 = gradient(perm, grad, x0, y1, z0, ix,     iy + 1, iz);
    float g3 = gradient(perm, grad, x1, y1, z0, ix + 1, iy + 1, iz);
    float g4 = gradient(perm, grad, x0, y0, z1, ix,     iy,     iz + 1);
    float g5 = gradient(perm, grad, x1, y0, z1, ix + 1, iy,     iz + 1);
    float g6 = gradient(perm, grad, x0, y1, z1, ix,     iy + 1, iz + 1);
    float g7 = gradient(perm, grad, x1, y1, z1, ix + 1, iy + 1, iz + 1);
    float a0 = lerp(g0, g1, u), a1 = lerp(g2, g3, u), a2 = lerp(g4, g5, u), a3 = lerp(g6, g7, u);
    float b0 = lerp(a0, a1, v), b1 = lerp(a2, a3, v);
    return lerp(b0, b1, w);
}

// main.cpp:689-707
__device__ __forceinline__ float fbm(const unsigned char *perm, const float *grad,
                                     double x, double y, double z, double lacunarity, float gain, int octaves)
{
    double frequency = 1.0;
    float amplitude = 1.0f, value = 0.0f;
    for (int i = 0; i < octaves; ++i) {
        float n = noise3(perm, grad, __dmul_rn(x, frequency), __dmul_rn(y, frequency), __dmul_rn(z, frequency));
        value = __fadd_rn(value, __fmul_rn(n, amplitude));
        frequency = __dmul_rn(frequency, lacunarity);
        amplitude = __fmul_rn(amplitude, gain);
    }
    return value;
}

// main.cpp:709-734
__device__ __forceinline__ float ridged(const unsigned char *perm, const float *grad,
                                        double x, double y, double z, double lacunarity, float gain, int octaves)
{
    double frequency = 1.0;
    float amplitude = 1.0f, weight = 1.0f, value = 0.0f;
    for (int i = 0; i < octaves; ++i) {
        float v = noise3(perm, grad, __dmul_rn(x, frequency), __dmul_rn(y, frequency), __dmul_rn(z, frequency));
        v = (v < 0.0f) ? -v : v;
        v = __fsub_rn(1.0f, v);
        v = __fmul_rn(v, v);
        value = __fadd_rn(value, __fmul_rn(__fmul_rn(v, amplitude), weight));
        weight = v;
        frequency = __dmul_rn(frequency, lacunarity);
        amplitude = __fmul_rn(amplitude, gain);
    }
    return value;
}

// Perlin::operator() (main.cpp:825-832) / ConstantZero (main.cpp:837-840)
__device__ __forceinline__ float height(const unsigned char *perm, const float *grad,
                                        const HeightCfg &c, d3 p, int depth)
{
    if (c.kind == PLANET_NOISE_ZERO) return 0.0f;
    int octaves = octaves_for(c.fixed_octaves, depth, c.max_depth);
    p = mul(p, c.coord_scale);
    if (c.has_seed) {
        p.x = __dadd_rn(p.x, c.seed[0]); p.y = __dadd_rn(p.y, c.seed[1]); p.z = __dadd_rn(p.z, c.seed[2]);
    }
    float h = (c.kind == PLANET_NOISE_RIDGED) ? ridged(perm, grad, p.x, p.y, p.z, c.lacunarity, c.gain, octaves)
                                              : fbm(perm, grad, p.x, p.y, p.z, c.lacunarity, c.gain, octaves);
    return __fmul_rn(h, c.height_scale);
}

// sample position of texel (x, y), main.cpp:132-146
__device__ __forceinline__ d3 sample_point(const Quad &q, int x, int y, double div)
{
    d3 v0 = sub(q.p[1], q.p[0]);
    d3 v1 = sub(q.p[3], q.p[2]);
    double u = __dmul_rn((double)(x - 1), div);
    double v = __dmul_rn((double)(y - 1), div);
    d3 p0 = add(q.p[0], mul(v0, u));
    d3 p1 = add(q.p[2], mul(v1, u));
    d3 v2 = sub(p1, p0);
    return add(p0, mul(v2, v));
}

} // namespace exact
#endif // __CUDACC__

} // namespace planet
