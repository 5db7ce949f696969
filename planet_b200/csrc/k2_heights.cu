// k2_heights.cu -- K2: fused sample-position + multi-octave noise + height kernels.
//
// Replaces, for whole batches of quads in one launch, the reference's hot loop
//   Gen::GenerateHeightMap (main.cpp:123-151) -> Perlin::operator() (main.cpp:825-832)
//   -> PerlinRidged / PerlinfBm (main.cpp:689-734) -> PerlinNoise3 (perlin.h:50-88),
// plus batched Gen::GetHeightAt (main.cpp:118-121) and raw noise evaluation.
//
// Two arithmetic modes (include/planet_gpu.h):
//   EXACT  every rounding of the reference reproduced -> bit-identical heights.
//   FAST   the throughput kernel, described below.
//
// FAST kernel design (B200, sm_100a; no tensor cores -- this is hashing + FP32 FMA):
//  * Lattice split in 64-bit fixed point.  A coordinate c (double) is reduced mod 256
//    (only cell & 255 ever reaches the permutation table, perlin.h:40) and held as
//    X = floor(c * 2^55) in two 32-bit registers.  With lacunarity 2.0 octave k scales
//    by 2^k exactly, so its cell byte and the top 23 fraction bits are one 32-bit
//    window of X: W = funnelshift_l(lo, hi, k) -> cell = W[30:23], fraction = W[22:0].
//    One SHF per axis per octave replaces the reference's fp64 multiply, floor and
//    subtract; the fraction becomes a float with one LOP3 (mantissa splice) and one FADD.
//  * The permutation table is staged in shared memory in a lane-replicated layout:
//    row r of a table holds the same entry 32 times, one copy per bank, so the 14
//    dependent lookups per octave-sample never bank-conflict no matter how random the
//    cells are.  Entries are stored pre-scaled as byte offsets of the next row, so a
//    chained lookup is one IADD + one LDS.  T12 (levels 1,2: 512 rows x 128 B) packs
//    both scalings as two u16; T3 (level 3: 512 rows x 256 B) holds the decoded
//    gradient of perlin.h:30-36 as {half2(gx, gy), float gz}, read with one LDS.64 and
//    unpacked on the FMA pipe (HADD2.F32), keeping the half-rate ALU pipe for hashing.
//  * 512-thread CTAs, one per SM (192 KB of tables), persistent over tiles of 1024
//    consecutive samples; each thread owns 2 consecutive texels (float2 store) for ILP.
//  * Per-tile prologue: three threads per touched quad turn the 104-byte Quad into the
//    bilinear form P = A + B x + y (C + D x) per axis in doubles pre-scaled by 2^55
//    (same sample points as main.cpp:132-146 up to 1 ulp of double), so a sample's
//    position costs 3 DFMA + one F2I per axis.
#include "planet_common.cuh"

#include <algorithm>

#include <cuda_fp16.h>

namespace planet {

// =====================================================================================
// EXACT kernels
// =====================================================================================
__device__ __forceinline__ void stage_small_tables(unsigned char *s_perm, float *s_grad)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_perm[i] = g_perm[i];
    for (int i = threadIdx.x; i < 48; i += blockDim.x) s_grad[i] = (&g_grad[0][0])[i];
    __syncthreads();
}

__global__ void __launch_bounds__(256)
k_height_maps_exact(const Quad *__restrict__ quads, int64_t total, int dim, HeightCfg cfg,
                    float *__restrict__ out)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    stage_small_tables(s_perm, s_grad);
    const int dim2 = dim * dim;
    const double div = __ddiv_rn(1.0, (double)(dim - 3));            // main.cpp:134
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = i / dim2;
        int r = (int)(i - q * dim2);
        int y = r / dim, x = r - y * dim;
        Quad quad = quads[q];
        d3 p = exact::sample_point(quad, x, y, div);
        out[i] = exact::height(s_perm, s_grad, cfg, p, (int)quad_depth(quad.id));
    }
}

__global__ void __launch_bounds__(256)
k_heights_at_exact(const double *__restrict__ xyz, int64_t n, int depth, HeightCfg cfg,
                   float *__restrict__ out)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    stage_small_tables(s_perm, s_grad);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        d3 p = { xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] };
        out[i] = exact::height(s_perm, s_grad, cfg, p, depth);
    }
}

// raw PerlinNoise3 (octaves == 0) / PerlinfBm / PerlinRidged on points
__global__ void __launch_bounds__(256)
k_noise_exact(const double *__restrict__ xyz, int64_t n, int kind, double lacunarity, float gain,
              int octaves, float *__restrict__ out)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    stage_small_tables(s_perm, s_grad);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        float v;
        if (octaves == 0) v = exact::noise3(s_perm, s_grad, x, y, z);
        else if (kind == PLANET_NOISE_RIDGED) v = exact::ridged(s_perm, s_grad, x, y, z, lacunarity, gain, octaves);
        else v = exact::fbm(s_perm, s_grad, x, y, z, lacunarity, gain, octaves);
        out[i] = v;
    }
}

// =====================================================================================
// FAST kernel
// =====================================================================================
namespace fast {

constexpr int THREADS = 512;
constexpr int S = 2;                          // consecutive samples per thread
constexpr int TILE = THREADS * S;             // samples per CTA iteration
constexpr int ROWS = 512;                     // table index range: perm (<=255) + cell (<=255) + 1
constexpr int T12_ROW = 128;                  // 32 lanes x u32
constexpr int T3_ROW = 256;                   // 32 lanes x {u32, f32}
constexpr int T12_BYTES = ROWS * T12_ROW;     // 64 KB
constexpr int T3_BYTES = ROWS * T3_ROW;       // 128 KB
constexpr int MAX_TILE_QUADS = TILE / 16 + 2; // dim >= 4 -> at most this many quads per tile
constexpr double FIX_ONE = 36028797018963968.0;           // 2^55
constexpr double FIX_WRAP = 256.0 * 36028797018963968.0;  // 2^63: one period of cell & 255

// per-quad, per-axis bilinear form of the sample position, pre-scaled by 2^55
struct AxisCoef { double a, b, c, d; };
struct TileQuad { AxisCoef ax[3]; int octaves; int wide; };

constexpr int SMEM_BYTES = T12_BYTES + T3_BYTES + MAX_TILE_QUADS * (int)sizeof(TileQuad);

// shared-memory table reads; `base` is a pointer into the dynamic shared array, so the
// compiler emits LDS.U16 / LDS.64 with 32-bit addressing and immediate offsets
__device__ __forceinline__ uint32_t lds_u16(const unsigned char *base, uint32_t off)
{
    return *reinterpret_cast<const unsigned short *>(base + off);
}
__device__ __forceinline__ uint2 lds_v2(const unsigned char *base, uint32_t off)
{
    return *reinterpret_cast<const uint2 *>(base + off);
}

// Build the lane-replicated tables.  T12[i][lane] = (perm*128) | (perm*256) << 16,
// T3[i][lane] = { half2(gx, gy), gz } of perlin_vectors[perm & 15]; perm = table[i & 255].
__device__ void build_tables(unsigned char *smem)
{
    uint32_t *t12 = reinterpret_cast<uint32_t *>(smem);
    uint2 *t3 = reinterpret_cast<uint2 *>(smem + T12_BYTES);
    for (int w = threadIdx.x; w < ROWS * 32; w += blockDim.x) {
        int i = w >> 5;
        uint32_t p = g_perm[i & 255];
        t12[w] = (p << 7) | (p << 24);
        const float *g = g_grad[p & 15];
        __half2 h = __floats2half2_rn(g[0], g[1]);
        uint2 e;
        e.x = *reinterpret_cast<uint32_t *>(&h);
        e.y = __float_as_uint(g[2]);
        t3[w] = e;
    }
}

// one gradient corner: perlin.h:43-48 with the vector pre-decoded in T3.  Exactly one
// of (gx, gy, gz) is zero and the others are +-1, so the sum has a single rounding and
// equals the reference's (x*v0 + y*v1) + z*v2 bit for bit.
__device__ __forceinline__ float corner(const unsigned char *t3, uint32_t off, float X, float Y, float Z)
{
    uint2 e = lds_v2(t3, off);
    __half2 h = *reinterpret_cast<__half2 *>(&e.x);
    float gx = __low2float(h), gy = __high2float(h), gz = __uint_as_float(e.y);
    return fmaf(gz, Z, fmaf(gy, Y, gx * X));
}

__device__ __forceinline__ float fade(float t)        // perlin.h:62 in fp32 + FMA
{
    return t * t * t * fmaf(t, fmaf(t, 6.0f, -15.0f), 10.0f);
}
__device__ __forceinline__ float lerp(float a, float b, float t) { return fmaf(b - a, t, a); }

// PerlinNoise3 for octave k of a fixed-point coordinate.  t12/t3 point at this lane's
// slot of row 0 of each table (lane*4 / lane*8 bytes in).
__device__ __forceinline__ float noise_octave(const unsigned char *t12_lane, const unsigned char *t3_lane,
                                              uint32_t xlo, uint32_t xhi, uint32_t ylo, uint32_t yhi,
                                              uint32_t zlo, uint32_t zhi, int k)
{
    // 32-bit windows: bits 30..23 = cell & 255, bits 22..0 = top of the fraction
    uint32_t wx = __funnelshift_l(xlo, xhi, k);
    uint32_t wy = __funnelshift_l(ylo, yhi, k);
    uint32_t wz = __funnelshift_l(zlo, zhi, k);
    float fx = __uint_as_float((wx & 0x007FFFFFu) | 0x3F800000u) - 1.0f;   // [0,1), 23 bits
    float fy = __uint_as_float((wy & 0x007FFFFFu) | 0x3F800000u) - 1.0f;
    float fz = __uint_as_float((wz & 0x007FFFFFu) | 0x3F800000u) - 1.0f;
    float gx1 = fx - 1.0f, gy1 = fy - 1.0f, gz1 = fz - 1.0f;

    // hash chain R(R(R(ix)+iy)+iz) for the 8 corners (perlin.h:45), 2 + 4 + 8 lookups
    uint32_t cx = (wx >> 16) & 0x7F80u;                 // (cell & 255) * 128
    uint32_t cy = (wy >> 16) & 0x7F80u;
    uint32_t cz = (wz >> 15) & 0xFF00u;                 // (cell & 255) * 256
    uint32_t a0 = lds_u16(t12_lane, cx), a1 = lds_u16(t12_lane, cx + T12_ROW);   // R(ix), R(ix+1) (*128)
    uint32_t b00 = lds_u16(t12_lane, a0 + cy + 2), b01 = lds_u16(t12_lane, a0 + cy + 2 + T12_ROW);
    uint32_t b10 = lds_u16(t12_lane, a1 + cy + 2), b11 = lds_u16(t12_lane, a1 + cy + 2 + T12_ROW);
    float g0 = corner(t3_lane, b00 + cz,          fx,  fy,  fz);     // R(R(R(ix)+iy)+iz) -> vector
    float g1 = corner(t3_lane, b10 + cz,          gx1, fy,  fz);
    float g2 = corner(t3_lane, b01 + cz,          fx,  gy1, fz);
    float g3 = corner(t3_lane, b11 + cz,          gx1, gy1, fz);
    float g4 = corner(t3_lane, b00 + cz + T3_ROW, fx,  fy,  gz1);
    float g5 = corner(t3_lane, b10 + cz + T3_ROW, gx1, fy,  gz1);
    float g6 = corner(t3_lane, b01 + cz + T3_ROW, fx,  gy1, gz1);
    float g7 = corner(t3_lane, b11 + cz + T3_ROW, gx1, gy1, gz1);

    float u = fade(fx), v = fade(fy), w = fade(fz);
    float l0 = lerp(g0, g1, u), l1 = lerp(g2, g3, u), l2 = lerp(g4, g5, u), l3 = lerp(g6, g7, u);
    return lerp(lerp(l0, l1, v), lerp(l2, l3, v), w);
}

// fractal sum over octaves for S samples at once (main.cpp:689-734 with FMA)
struct Fixed3 { uint32_t xlo, xhi, ylo, yhi, zlo, zhi; };

template <bool GUARD>
__device__ __forceinline__ void fractal_loop(const unsigned char *t12_lane, const unsigned char *t3_lane,
                                             const Fixed3 (&p)[S], const int (&octaves)[S], int omax,
                                             int kind, float gain, float (&value)[S])
{
    float amp = 1.0f;
    float weight[S];
#pragma unroll
    for (int s = 0; s < S; s++) { value[s] = 0.0f; weight[s] = 1.0f; }
    for (int k = 0; k < omax; k++) {
#pragma unroll
        for (int s = 0; s < S; s++) {
            float n = noise_octave(t12_lane, t3_lane, p[s].xlo, p[s].xhi, p[s].ylo, p[s].yhi,
                                   p[s].zlo, p[s].zhi, k);
            bool live = !GUARD || k < octaves[s];
            if (kind == PLANET_NOISE_RIDGED) {        // main.cpp:722-728
                float v = 1.0f - fabsf(n);
                v = v * v;
                float nv = fmaf(v * amp, weight[s], value[s]);
                if (live) { value[s] = nv; weight[s] = v; }
            } else {                                  // main.cpp:701
                float nv = fmaf(n, amp, value[s]);
                if (live) value[s] = nv;
            }
        }
        amp *= gain;                                  // main.cpp:703 / 730
    }
}

__device__ __forceinline__ void fractal(const unsigned char *t12_lane, const unsigned char *t3_lane,
                                        const Fixed3 (&p)[S], const int (&octaves)[S], int kind,
                                        float gain, float (&value)[S])
{
    int omax = octaves[0];
    bool same = true;
#pragma unroll
    for (int s = 1; s < S; s++) { omax = max(omax, octaves[s]); same = same && octaves[s] == octaves[0]; }
    if (same) fractal_loop<false>(t12_lane, t3_lane, p, octaves, omax, kind, gain, value);
    else      fractal_loop<true>(t12_lane, t3_lane, p, octaves, omax, kind, gain, value);
}

// reduce a scaled coordinate (units of 2^-55) to one period and convert to fixed point
__device__ __forceinline__ void to_fixed(double v, uint32_t &lo, uint32_t &hi)
{
    long long f = __double2ll_rd(v);
    lo = (uint32_t)(unsigned long long)f;
    hi = (uint32_t)((unsigned long long)f >> 32);
}
__device__ __forceinline__ double wrap_period(double v)     // v - 2^63 * rint(v / 2^63): exact
{
    return fma(-FIX_WRAP, rint(v * (1.0 / FIX_WRAP)), v);
}

// ---- height maps --------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
k_height_maps_fast(const Quad *__restrict__ quads, int64_t total, int dim, HeightCfg cfg,
                   float *__restrict__ out, int64_t ntiles, int out_aligned8)
{
    extern __shared__ __align__(16) unsigned char smem[];
    TileQuad *tq = reinterpret_cast<TileQuad *>(smem + T12_BYTES + T3_BYTES);
    build_tables(smem);

    const int lane = threadIdx.x & 31;
    const unsigned char *t12_lane = smem + lane * 4;
    const unsigned char *t3_lane = smem + T12_BYTES + lane * 8;
    const int dim2 = dim * dim;
    const double div = 1.0 / (double)(dim - 3);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const int64_t last = min(base + TILE, total) - 1;
        const int64_t q_first = base / dim2;
        const int nq = (int)(last / dim2 - q_first) + 1;
        __syncthreads();                  // tables built / previous tile done with tq[]

        // prologue: quad -> per-axis bilinear coefficients (main.cpp:130-146 regrouped)
        for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {
            const double *qp = reinterpret_cast<const double *>(quads + q_first + qi);
            const double s = cfg.coord_scale * FIX_ONE;
            double span = 0.0;
#pragma unroll
            for (int axis = 0; axis < 3; axis++) {
                double p0 = qp[axis], p1 = qp[3 + axis], p2 = qp[6 + axis], p3 = qp[9 + axis];
                double v0 = p1 - p0, v1 = p3 - p2;
                AxisCoef c;
                c.a = wrap_period((p0 * cfg.coord_scale + (cfg.has_seed ? cfg.seed[axis] : 0.0)) * FIX_ONE);
                c.b = v0 * (s * div);
                c.c = (p2 - p0) * (s * div);
                c.d = (v1 - v0) * (s * div * div);
                tq[qi].ax[axis] = c;
                span = fmax(span, (fabs(c.b) + fabs(c.c)) * (double)dim + fabs(c.d) * (double)dim * (double)dim);
            }
            uint64_t id = quads[q_first + qi].id;
            tq[qi].octaves = octaves_for(cfg.fixed_octaves, (int)quad_depth(id), cfg.max_depth);
            // a quad whose scaled extent nears half a period cannot keep |P| < 2^63 from a
            // reduced corner alone; such quads take a per-sample reduction instead
            tq[qi].wide = span > 0.4 * FIX_WRAP;
        }
        __syncthreads();

        const int64_t lin0 = base + (int64_t)threadIdx.x * S;
        Fixed3 p[S];
        int oct[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            int64_t lin = min(lin0 + s, last);
            int64_t q = lin / dim2;
            int r = (int)(lin - q * dim2);
            int y = r / dim, x = r - y * dim;
            const TileQuad &c = tq[(int)(q - q_first)];
            double xd = (double)(x - 1), yd = (double)(y - 1);
            double P[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                P[a] = fma(yd, fma(c.ax[a].d, xd, c.ax[a].c), fma(c.ax[a].b, xd, c.ax[a].a));
                if (c.wide) P[a] = wrap_period(P[a]);
            }
            to_fixed(P[0], p[s].xlo, p[s].xhi);
            to_fixed(P[1], p[s].ylo, p[s].yhi);
            to_fixed(P[2], p[s].zlo, p[s].zhi);
            oct[s] = c.octaves;
        }

        float value[S];
        if (cfg.kind == PLANET_NOISE_ZERO) {
#pragma unroll
            for (int s = 0; s < S; s++) value[s] = 0.0f;
        } else {
            fractal(t12_lane, t3_lane, p, oct, cfg.kind, cfg.gain, value);
        }

        if (out_aligned8 && lin0 + S - 1 <= last) {          // lin0 is even by construction
            float2 h = make_float2(value[0] * cfg.height_scale, value[1] * cfg.height_scale);
            *reinterpret_cast<float2 *>(out + lin0) = h;
        } else {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (lin0 + s <= last) out[lin0 + s] = value[s] * cfg.height_scale;
        }
    }
}

// ---- points: batched GetHeightAt and raw noise ----------------------------------------
// scale / seed / height_scale == 1 / 0 / 1 and octaves0 == 1 give PerlinNoise3 itself.
__global__ void __launch_bounds__(THREADS, 1)
k_points_fast(const double *__restrict__ xyz, int64_t n, int kind, float gain, int octaves,
              double coord_scale, double sx, double sy, double sz, float height_scale,
              float *__restrict__ out, int64_t ntiles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    build_tables(smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned char *t12_lane = smem + lane * 4;
    const unsigned char *t3_lane = smem + T12_BYTES + lane * 8;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // samples of one thread are strided by THREADS so the 24-byte point loads and the
        // 4-byte stores of a warp stay contiguous
        Fixed3 p[S];
        int oct[S];
        int64_t idx[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            idx[s] = tile * TILE + s * THREADS + threadIdx.x;
            int64_t i = min(idx[s], n - 1);
            double c[3] = { xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] };
            c[0] = fma(c[0], coord_scale, sx); c[1] = fma(c[1], coord_scale, sy); c[2] = fma(c[2], coord_scale, sz);
            // reduce mod 256 in double (exact: power-of-two period), then scale to 2^55
            double r0 = fma(-256.0, rint(c[0] * (1.0 / 256.0)), c[0]);
            double r1 = fma(-256.0, rint(c[1] * (1.0 / 256.0)), c[1]);
            double r2 = fma(-256.0, rint(c[2] * (1.0 / 256.0)), c[2]);
            to_fixed(r0 * FIX_ONE, p[s].xlo, p[s].xhi);
            to_fixed(r1 * FIX_ONE, p[s].ylo, p[s].yhi);
            to_fixed(r2 * FIX_ONE, p[s].zlo, p[s].zhi);
            oct[s] = octaves;
        }
        float value[S];
        fractal(t12_lane, t3_lane, p, oct, kind, gain, value);
#pragma unroll
        for (int s = 0; s < S; s++)
            if (idx[s] < n) out[idx[s]] = value[s] * height_scale;
    }
}

} // namespace fast

// =====================================================================================
// host-side launchers
// =====================================================================================
static int g_sm_count = 0;
static bool g_fast_attr_set = false;

static int sm_count()
{
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

static int prepare_fast()
{
    if (!g_fast_attr_set) {
        PLANET_CUDA(cudaFuncSetAttribute(fast::k_height_maps_fast,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
        PLANET_CUDA(cudaFuncSetAttribute(fast::k_points_fast,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
        g_fast_attr_set = true;
    }
    return 0;
}

// FAST handles the lattice split as shifts, which needs lacunarity == 2 and every octave's
// window inside the 64-bit word; anything else runs the EXACT kernel on the GPU.
static bool fast_applicable(double lacunarity, int max_octaves)
{
    return lacunarity == 2.0 && max_octaves <= 32;
}

int launch_height_maps(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, int dim,
                       int max_depth, float *d_out, cudaStream_t stream)
{
    HeightCfg cfg = make_cfg(p, max_depth);
    int64_t total = nquads * (int64_t)dim * dim;
    if (total == 0) return 0;
    int max_oct = octaves_for(cfg.fixed_octaves, 31, max_depth > 0 ? max_depth : 1);
    if (cfg.fixed_octaves <= 0) max_oct = 6 + 12 * 31 / (max_depth > 0 ? max_depth : 1);
    bool use_fast = p->precision == PLANET_PRECISION_FAST &&
                    fast_applicable(cfg.lacunarity, cfg.fixed_octaves > 0 ? cfg.fixed_octaves : 32) &&
                    (cfg.fixed_octaves > 0 || max_depth >= 12);   // 6 + 12*31/max_depth <= 32
    (void)max_oct;
    if (use_fast) {
        int rc = prepare_fast();
        if (rc) return rc;
        int64_t ntiles = (total + fast::TILE - 1) / fast::TILE;
        int grid = (int)std::min<int64_t>(ntiles, sm_count());
        fast::k_height_maps_fast<<<grid, fast::THREADS, fast::SMEM_BYTES, stream>>>(
            d_quads, total, dim, cfg, d_out, ntiles, (reinterpret_cast<uintptr_t>(d_out) & 7) == 0);
    } else {
        int64_t blocks = (total + 255) / 256;
        int grid = (int)std::min<int64_t>(blocks, (int64_t)sm_count() * 8);
        k_height_maps_exact<<<grid, 256, 0, stream>>>(d_quads, total, dim, cfg, d_out);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "height map kernel launch");
}

int launch_heights_at(const planet_gpu_params *p, const double *d_xyz, int64_t n, int depth,
                      int max_depth, float *d_out, cudaStream_t stream)
{
    if (n == 0) return 0;
    HeightCfg cfg = make_cfg(p, max_depth);
    int octaves = octaves_for(cfg.fixed_octaves, depth, max_depth);
    bool use_fast = p->precision == PLANET_PRECISION_FAST && cfg.kind != PLANET_NOISE_ZERO &&
                    fast_applicable(cfg.lacunarity, octaves);
    if (use_fast) {
        int rc = prepare_fast();
        if (rc) return rc;
        int64_t ntiles = (n + fast::TILE - 1) / fast::TILE;
        int grid = (int)std::min<int64_t>(ntiles, sm_count());
        fast::k_points_fast<<<grid, fast::THREADS, fast::SMEM_BYTES, stream>>>(
            d_xyz, n, cfg.kind, cfg.gain, octaves, cfg.coord_scale,
            cfg.has_seed ? cfg.seed[0] : 0.0, cfg.has_seed ? cfg.seed[1] : 0.0,
            cfg.has_seed ? cfg.seed[2] : 0.0, cfg.height_scale, d_out, ntiles);
    } else {
        int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
        k_heights_at_exact<<<grid, 256, 0, stream>>>(d_xyz, n, depth, cfg, d_out);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "heights_at kernel launch");
}

int launch_noise(const double *d_xyz, int64_t n, int kind, double lacunarity, float gain, int octaves,
                 int precision, float *d_out, cudaStream_t stream)
{
    if (n == 0) return 0;
    if (precision == PLANET_PRECISION_FAST && fast_applicable(octaves == 0 ? 2.0 : lacunarity, octaves)) {
        int rc = prepare_fast();
        if (rc) return rc;
        int64_t ntiles = (n + fast::TILE - 1) / fast::TILE;
        int grid = (int)std::min<int64_t>(ntiles, sm_count());
        // a single PerlinNoise3 is a 1-octave fBm with amplitude 1
        fast::k_points_fast<<<grid, fast::THREADS, fast::SMEM_BYTES, stream>>>(
            d_xyz, n, octaves == 0 ? PLANET_NOISE_FBM : kind, gain, octaves == 0 ? 1 : octaves,
            1.0, 0.0, 0.0, 0.0, 1.0f, d_out, ntiles);
    } else {
        int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
        k_noise_exact<<<grid, 256, 0, stream>>>(d_xyz, n, kind, lacunarity, gain, octaves, d_out);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "noise kernel launch");
}

} // namespace planet
