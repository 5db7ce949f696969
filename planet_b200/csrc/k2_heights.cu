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
//  * The tables are staged in shared memory in a lane-replicated layout, so the 7 dependent
//    lookups per octave-sample never bank-conflict no matter how random the cells are.
//    Entries are stored pre-scaled as byte offsets of the next row and the lane's copy offset is
//    OR-ed into the cell offset once per axis, so a chained lookup is one IADD3 + one
//    LDS [R + UR + imm].  Every entry carries its +1 neighbour: P1 and P2 (levels 1, 2) hold
//    {R(i), R(i+1)} as two pre-scaled u32 (LDS.64), T3 (level 3) holds the gradients of BOTH
//    z-neighbours as FINISHED floats {gx, gy, gx', gy'} = 2*v in {0, +-2} (LDS.128), so the 8
//    corners of a cell cost 1 + 2 + 4 loads and gx, gy go straight into the FMAs.  gz rides in
//    the two lowest mantissa bits of gx and costs one shift.  An LDS.64 / LDS.128 is served a
//    half / quarter warp at a time, hence 16 / 8 copies per row.  22 shared-memory wavefronts per
//    octave-sample buy 16 fewer instructions than byte codes decoded by AND/PRMT/shift
//    (tools/microbench3.cu).  Measured cost model of this kernel: an SM sub-partition spends one
//    cycle per 32-bit register a warp instruction writes -- one per ordinary instruction, one per
//    word a load returns (tools/microbench5.cu): a load costs what the arithmetic that could
//    replace it would cost, and only fewer results make the loop faster.
//  * 768-thread CTAs, one per SM (160 KB of tables), persistent; every warp owns a contiguous
//    run of 128-sample tiles; each thread owns 2 consecutive texels whose two independent dependency chains
//    interleave in the instruction stream.  Plain FP32 FMA: the packed f32x2 forms
//    (FFMA2/FMUL2/FADD2) hold the issue port for two cycles (tools/microbench2.cu), so they save
//    no issue time and would need register-pair moves here.  91 instructions per octave-sample.
//  * Tiles that lie inside one quad (every tile of a 32 x 32 map) are walked as runs with no
//    per-tile setup; the noise kind is a template parameter.
//  * Per-tile prologue: one thread per touched quad turns the 104-byte Quad into the
//    bilinear form P = A + B x + y (C + D x) per axis in doubles pre-scaled by 2^55
//    (same sample points as main.cpp:132-146 up to 1 ulp of double), so a sample's
//    position costs 3 DFMA + one F2I per axis.
#include "planet_common.cuh"
#include "planet_call_surface.cuh"
#include "planet_tma.cuh"

#include <algorithm>

#include <cstdlib>

namespace planet {

// (PeerOut, the fused multi-GPU gather, is declared in planet_common.cuh)

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
                    float *__restrict__ out, PeerOut peers)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    stage_small_tables(s_perm, s_grad);
    const int64_t dim2 = (int64_t)dim * dim;                          // dim <= 32768 (check_height_args)
    const double div = __ddiv_rn(1.0, (double)(dim - 3));            // main.cpp:134
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = i / dim2;
        int r = (int)(i - q * dim2);
        int y = r / dim, x = r - y * dim;
        Quad quad = quads[q];
        d3 p = exact::sample_point(quad, x, y, div);
        float h = exact::height(s_perm, s_grad, cfg, p, (int)quad_depth(quad.id));
        out[i] = h;
#pragma unroll
        for (int r = 0; r < 7; r++)
            if (r < peers.n) peers.ptr[r][i] = h;
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

// ---- the reference's seam, one call at a time (main.cpp:107-111) -----------------------------
// A synchronous GenerateHeightMap / GetHeightAt through the two function pointers is pure
// latency: the quad / point travels as a kernel ARGUMENT (no H2D copy) and the kernel stores its
// result straight into pinned host memory (no D2H copy), so a call is one launch + one sync.
// A single map is 1 024 samples x 6..18 octaves: too little work to fill the chip one thread per
// sample, and a serial octave loop is the whole latency.  So 8 lanes share a sample: lane `sub`
// evaluates octaves sub, sub + 8, ... (the noise of an octave does not depend on the others) and
// the group's lanes then accumulate them in the reference's order (main.cpp:699-704 / 716-731),
// which keeps the float sum bit-identical to the sequential loop.
__global__ void __launch_bounds__(256)
k_height_map_seam_exact(Quad quad, int dim, HeightCfg cfg, float *__restrict__ host_out)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    stage_small_tables(s_perm, s_grad);
    const int dim2 = dim * dim;
    const double div = __ddiv_rn(1.0, (double)(dim - 3));            // main.cpp:134
    const int octaves = cfg.kind == PLANET_NOISE_ZERO ? 0 : octaves_for(cfg.fixed_octaves, (int)quad_depth(quad.id), cfg.max_depth);
    const int lane = threadIdx.x & 31, sub = lane & 7, first = lane & ~7;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, ngroups = (gridDim.x * blockDim.x) >> 3;
    for (int i0 = group - (lane >> 3); i0 < dim2; i0 += ngroups) {   // i0: the warp's first sample (warp-uniform trip count)
        const int i = min(i0 + (lane >> 3), dim2 - 1);
        const int y = i / dim, x = i - y * dim;
        d3 p = exact::mul(exact::sample_point(quad, x, y, div), cfg.coord_scale);   // main.cpp:132-146, 828
        if (cfg.has_seed) { p.x = __dadd_rn(p.x, cfg.seed[0]); p.y = __dadd_rn(p.y, cfg.seed[1]); p.z = __dadd_rn(p.z, cfg.seed[2]); }
        float amplitude = 1.0f, weight = 1.0f, value = 0.0f;
        double frequency = 1.0;
        for (int k0 = 0; k0 < octaves; k0 += 8) {
            double f = frequency;                                    // lacunarity^k by the reference's repeated product
            for (int j = 0; j < sub; j++) f = __dmul_rn(f, cfg.lacunarity);
            float n = 0.0f;
            if (k0 + sub < octaves)
                n = exact::noise3(s_perm, s_grad, __dmul_rn(p.x, f), __dmul_rn(p.y, f), __dmul_rn(p.z, f));
            for (int j = 0; j < 8 && k0 + j < octaves; j++) {
                const float nj = __shfl_sync(0xffffffffu, n, first + j);
                if (cfg.kind == PLANET_NOISE_RIDGED) {
                    float v = (nj < 0.0f) ? -nj : nj;
                    v = __fsub_rn(1.0f, v);
                    v = __fmul_rn(v, v);
                    value = __fadd_rn(value, __fmul_rn(__fmul_rn(v, amplitude), weight));
                    weight = v;
                } else {
                    value = __fadd_rn(value, __fmul_rn(nj, amplitude));
                }
                amplitude = __fmul_rn(amplitude, cfg.gain);
                frequency = __dmul_rn(frequency, cfg.lacunarity);
            }
        }
        if (sub == 0 && i0 + (lane >> 3) < dim2) host_out[i] = __fmul_rn(value, cfg.height_scale);   // main.cpp:831
    }
}

__global__ void __launch_bounds__(32)
k_height_at_seam_exact(double px, double py, double pz, int depth, HeightCfg cfg, float *__restrict__ host_out)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    stage_small_tables(s_perm, s_grad);
    if (threadIdx.x == 0) host_out[0] = exact::height(s_perm, s_grad, cfg, d3{ px, py, pz }, depth);
}

// raw PerlinNoise3 (octaves == 0) / PerlinfBm / PerlinRidged on points, through the
// reference-named device functions of planet_call_surface.cuh
__global__ void __launch_bounds__(256)
k_noise_exact(const double *__restrict__ xyz, int64_t n, int kind, double lacunarity, float gain,
              int octaves, float *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        float v;
        if (octaves == 0) v = PerlinNoise3(x, y, z);
        else if (kind == PLANET_NOISE_RIDGED) v = PerlinRidged(x, y, z, lacunarity, gain, octaves);
        else v = PerlinfBm(x, y, z, lacunarity, gain, octaves);
        out[i] = v;
    }
}

// flags[r] >= want for every rank r != rank (one thread per rank)
__global__ void k_gather_wait(const uint32_t *flags, uint32_t want, int rank, int world, uint32_t *error)
{
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) gather_wait_flag(flags + threadIdx.x, want, error);
}

int launch_gather_wait(const uint32_t *flags, uint32_t want, int rank, int world, uint32_t *error, cudaStream_t stream)
{
    k_gather_wait<<<1, 32, 0, stream>>>(flags, want, rank, world, error);
    count_launch();
    return check_cuda(cudaGetLastError(), "gather wait launch");
}

// =====================================================================================
// FAST kernel
// =====================================================================================
namespace fast {

#ifndef PLANET_K2_ALIGN
#define PLANET_K2_ALIGN 1
#endif
#ifndef PLANET_K2_UNROLL
#define PLANET_K2_UNROLL 1
#endif
constexpr int K2_UNROLL = PLANET_K2_UNROLL;   // octave-loop unroll factor (build-time tuning knob)
constexpr int THREADS = 512;
constexpr int S = 2;                          // consecutive samples per thread (one 8-byte store)
constexpr int TILE = THREADS * S;             // samples per CTA iteration
constexpr int ROWS = 512;                     // table index range: perm (<=255) + cell (<=255) + 1
// Table layout for a replication factor REPL.  REPL = 32 is the throughput layout (160 KB): an
// LDS.64 is served a half-warp at a time and an LDS.128 a quarter-warp at a time, so 16 copies of
// an 8-byte entry and 8 copies of a 16-byte entry put the lanes of every phase in distinct bank
// groups whatever rows they ask for (tools/microbench3.cu).  REPL = 1 is the compact 14 KB
// layout whose build costs nothing, for small batches where latency matters.
//   P1 (level 1, 256 rows): { R(i), R(i+1) } * P_ROW      -- byte offsets of the next P2 row
//   P2 (level 2, 512 rows): { R(i), R(i+1) } * T3_ROW     -- byte offsets of the next T3 row
//   T3 (level 3, 512 rows): { gx(i)|zc, gy(i), gx(i+1)|zc, gy(i+1) }
// where R = perlin_random_table[i & 255].  Every entry carries its +1 neighbour, so the 8
// corners of a cell cost 1 + 2 + 4 loads.
template <int REPL> struct Layout {
    static_assert(REPL == 32 || REPL == 1, "layouts in use");
    static constexpr int P_COPIES = REPL == 32 ? 16 : 1;
    static constexpr int T3_COPIES = REPL == 32 ? 8 : 1;
    static constexpr int P_ROW = P_COPIES * 8, T3_ROW = T3_COPIES * 16;
    static constexpr int LOGP = REPL == 32 ? 7 : 3, LOG3 = REPL == 32 ? 7 : 4;
    static constexpr int P1_BYTES = 256 * P_ROW;          // 32 KB at REPL 32
    static constexpr int P2_BYTES = ROWS * P_ROW;         // 64 KB
    static constexpr int T3_BYTES = ROWS * T3_ROW;        // 64 KB
    static constexpr int P2_AT = P1_BYTES, T3_AT = P1_BYTES + P2_BYTES;
    static constexpr int TABLES = P1_BYTES + P2_BYTES + T3_BYTES;
};
constexpr uint32_t ONE_BITS = 0x3F800000u;                 // float 1.0, see splice_mantissa
constexpr double FIX_ONE = 36028797018963968.0;           // 2^55
constexpr double FIX_WRAP = 256.0 * 36028797018963968.0;  // 2^63: one period of cell & 255

// per-quad, per-axis bilinear form of the sample position, pre-scaled by 2^55
struct AxisCoef { double a, b, c, d; };
struct TileQuad { AxisCoef ax[3]; int octaves; int wide; };

template <int REPL> constexpr int smem_bytes(int nthreads)
{
    // tables + per-warp quad scratch (never less than the 2 KB the table build stages through)
    // + up to one P1 of slack in front: the tables start on a multiple of P1's size (tables_at)
    return (PLANET_K2_ALIGN ? Layout<REPL>::P1_BYTES : 0) + Layout<REPL>::TABLES +
           ((nthreads / 32) * (128 / 16 + 2) * (int)sizeof(TileQuad) > 2048
                ? (nthreads / 32) * (128 / 16 + 2) * (int)sizeof(TileQuad) : 2048);
}
// Start of the tables inside the dynamic shared array: the first address whose shared-window
// address is a multiple of P1's size.  A level-1 address is then (absolute P1 address of the
// lane's copy) | (cell offset) -- the same LOP3 that merges the lane offset -- and the load
// needs no base register added (one instruction per octave-sample in an issue-bound loop).
template <int REPL> __device__ __forceinline__ unsigned char *tables_at(unsigned char *smem)
{
    if (!PLANET_K2_ALIGN) return smem;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    return smem + ((0u - base) & (uint32_t)(Layout<REPL>::P1_BYTES - 1));
}
constexpr int SMEM_BYTES = smem_bytes<32>(768);

// shared-memory table reads; `base` is a pointer into the dynamic shared array, so the
// compiler emits LDS.64 / LDS.128 with 32-bit addressing and immediate offsets
__device__ __forceinline__ uint2 lds_v2(const unsigned char *base, uint32_t off)
{
    return *reinterpret_cast<const uint2 *>(base + off);
}
__device__ __forceinline__ uint4 lds_v4(const unsigned char *base, uint32_t off)
{
    return *reinterpret_cast<const uint4 *>(base + off);
}
// A lane's view of the tables: the (warp-uniform) start of the dynamic shared array plus the
// byte offsets of the lane's own copies inside a P1/P2 row and a T3 row.  Keeping the base uniform and the
// lane part a 32-bit offset lets every lookup be LDS [R + UR + imm]: the lane offset is OR-ed
// into the cell offset once per axis instead of being added to every address.
struct LaneTab { const unsigned char *base; uint32_t lp, l3, l1abs; };
template <int REPL> __device__ __forceinline__ LaneTab lane_tab(const unsigned char *tabs, int lane, uint32_t zero = 0)
{
    using L = Layout<REPL>;
    const uint32_t lp = (uint32_t)(lane % L::P_COPIES) * 8u;
    // `zero` is a run-time 0 the compiler cannot see through (one_bits ^ ONE_BITS): without it ptxas
    // does not keep base | lp in a register but re-derives it with a second LOP3 at every use
    return { tabs, lp, (uint32_t)(lane % L::T3_COPIES) * 16u, ((uint32_t)__cvta_generic_to_shared(tabs) | lp) ^ zero };
}
// level-1 read from an absolute shared-window address (tables_at)
__device__ __forceinline__ uint2 lds_abs_v2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// A gradient vector of perlin.h:30-36 as three byte codes, one per component, each the TOP
// byte of the float 2*v: 0x00 -> 0.0, 0x40 -> +2.0, 0xC0 -> -2.0 (x in byte 3, y in byte 1, z in
// byte 0).  The factor 2 is exact and is taken back out of the octave amplitude.
__device__ __forceinline__ uint32_t grad_code(int h)
{
    const float *g = g_grad[h & 15];
    auto byte = [](float v) -> uint32_t { return v > 0.f ? 0x40u : (v < 0.f ? 0xC0u : 0x00u); };
    return (byte(g[0]) << 24) | (byte(g[1]) << 8) | byte(g[2]);
}
// The table form of a code: gx and gy as finished floats (no decode instruction at the point of
// use), the z code in the two lowest mantissa bits of gx, so gz = word << 30 is one shift.
// FAST uses gx as it is (the stowaway bits move it by <= 3 * 2^-23 relative); EXACT masks them.
__device__ __forceinline__ uint32_t gx_word(uint32_t code) { return (code & 0xFF000000u) | ((code & 0xC0u) >> 6); }
__device__ __forceinline__ uint32_t gy_word(uint32_t code) { return (code & 0x0000FF00u) << 16; }

// Build the (lane-replicated) tables described at Layout
template <int REPL>
__device__ void build_tables(unsigned char *smem)
{
    using L = Layout<REPL>;
    uint2 *p1 = reinterpret_cast<uint2 *>(smem);
    uint2 *p2 = reinterpret_cast<uint2 *>(smem + L::P2_AT);
    uint4 *t3 = reinterpret_cast<uint4 *>(smem + L::T3_AT);
    // stage {perm, code(perm)} for the 256 table entries in the (not yet used) scratch area
    uint2 *stage = reinterpret_cast<uint2 *>(smem + L::TABLES);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t p = g_perm[i];
        stage[i] = make_uint2(p, grad_code(p));
    }
    __syncthreads();
    for (int w = threadIdx.x; w < ROWS * L::P_COPIES; w += blockDim.x) {
        const int i = w / L::P_COPIES;
        const uint32_t r0 = stage[i & 255].x, r1 = stage[(i + 1) & 255].x;
        if (i < 256) p1[w] = make_uint2(r0 << L::LOGP, r1 << L::LOGP);
        p2[w] = make_uint2(r0 << L::LOG3, r1 << L::LOG3);
    }
    for (int w = threadIdx.x; w < ROWS * L::T3_COPIES; w += blockDim.x) {
        const int i = w / L::T3_COPIES;
        const uint32_t c0 = stage[i & 255].y, c1 = stage[(i + 1) & 255].y;
        t3[w] = make_uint4(gx_word(c0), gy_word(c0), gx_word(c1), gy_word(c1));
    }
}

// dot of one corner's gradient with the corner-relative position (perlin.h:43-48).  Exactly one
// component is zero and the others are +-2, so the sum has a single rounding: 2 * the
// reference's (x*v0 + y*v1) + z*v2.
__device__ __forceinline__ float corner(uint32_t gxw, uint32_t gyw, float X, float Y, float Z)
{
    return fmaf(__uint_as_float(gxw << 30), Z, fmaf(__uint_as_float(gyw), Y, __uint_as_float(gxw) * X));
}

// the same dot minus `a`, as one chain: (gx*X - a) + gy*Y + gz*Z
__device__ __forceinline__ float corner_minus(uint32_t gxw, uint32_t gyw, float X, float Y, float Z, float a)
{
    return fmaf(__uint_as_float(gxw << 30), Z, fmaf(__uint_as_float(gyw), Y, fmaf(__uint_as_float(gxw), X, -a)));
}

__device__ __forceinline__ float fade1(float t)       // perlin.h:62 in fp32 + FMA
{
    return (t * t * t) * fmaf(fmaf(t, 6.0f, -15.0f), t, 10.0f);
}
__device__ __forceinline__ float lerp1(float a, float b, float t) { return fmaf(b - a, t, a); }

struct Fixed3 { uint32_t xlo, xhi, ylo, yhi, zlo, zhi; };

// (w & 0x007FFFFF) | one_bits as ONE LOP3.  one_bits is 0x3F800000 (float 1.0) handed down from a
// kernel argument: written as two immediates the compiler needs two LOP3 (one immediate slot per
// instruction), and this runs three times per octave-sample in an issue-bound loop.
__device__ __forceinline__ float splice_mantissa(uint32_t w, uint32_t one_bits)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(d) : "r"(w), "r"(one_bits));
    return __uint_as_float(d);
}

// hash chain R(R(R(ix)+iy)+iz) of one sample for octave k (perlin.h:45): returns the four
// {z, z+1} gradient pairs of the (x, y) columns 00, 10, 01, 11 and the three fractions
struct Hashed { uint4 e00, e10, e01, e11; float mx, my, mz; };

template <int REPL>
__device__ __forceinline__ Hashed hash_octave(const LaneTab &tab, const Fixed3 &p, int k, uint32_t one_bits)
{
    using L = Layout<REPL>;
    // 32-bit windows of the fixed-point coordinate: bits 30..23 = cell & 255, 22..0 = fraction
    uint32_t wx = __funnelshift_l(p.xlo, p.xhi, k);
    uint32_t wy = __funnelshift_l(p.ylo, p.yhi, k);
    uint32_t wz = __funnelshift_l(p.zlo, p.zhi, k);
    Hashed h;
    h.mx = splice_mantissa(wx, one_bits);                           // 1 + fraction, 23 bits
    h.my = splice_mantissa(wy, one_bits);
    h.mz = splice_mantissa(wz, one_bits);
    // (cell & 255) * row size, with the lane's copy offset OR-ed into the (zero) low bits
    const uint32_t cx = ((wx >> (23 - L::LOGP)) & (255u << L::LOGP)) | (PLANET_K2_ALIGN ? tab.l1abs : tab.lp);   // absolute P1 address
    const uint32_t cy = ((wy >> (23 - L::LOGP)) & (255u << L::LOGP)) | tab.lp;
    const uint32_t cz = ((wz >> (23 - L::LOG3)) & (255u << L::LOG3)) | tab.l3;
    const unsigned char *t = tab.base;
    const uint2 a = PLANET_K2_ALIGN ? lds_abs_v2(cx) : lds_v2(t, cx);   // R(ix), R(ix+1)
    const uint2 b0 = lds_v2(t, a.x + cy + L::P2_AT);                // R(R(ix)+iy), R(R(ix)+iy+1)
    const uint2 b1 = lds_v2(t, a.y + cy + L::P2_AT);                // the same for ix+1
    h.e00 = lds_v4(t, b0.x + cz + L::T3_AT);                        // R(R(R(ix)+iy)+iz), ..+iz+1
    h.e10 = lds_v4(t, b1.x + cz + L::T3_AT);
    h.e01 = lds_v4(t, b0.y + cz + L::T3_AT);
    h.e11 = lds_v4(t, b1.y + cz + L::T3_AT);
    return h;
}

// 2 * PerlinNoise3 from the hashed gradients and fractions of one sample (perlin.h:50-88).
// The x+1 corner of each pair is formed as (its dot) - (the x corner's dot) in one FMA chain
// seeded with the negated first dot, so the lerp along x is a single FMA: 7 instructions per
// corner pair instead of 8.  (Also folding x - 1 into the chain start, gx*x - gx, and lerping
// with 1 - u removes two more instructions per iteration and no time: measured, not kept.)
__device__ __forceinline__ float noise_from(const Hashed &h)
{
    const float x0 = h.mx - 1.0f, x1 = h.mx - 2.0f;                 // fraction, fraction - 1 (exact)
    const float y0 = h.my - 1.0f, y1 = h.my - 2.0f;
    const float z0 = h.mz - 1.0f, z1 = h.mz - 2.0f;
    const float u = fade1(x0), v = fade1(y0), w = fade1(z0);
    const float g0 = corner(h.e00.x, h.e00.y, x0, y0, z0);          // perlin.h:68-75
    const float l0 = fmaf(corner_minus(h.e10.x, h.e10.y, x1, y0, z0, g0), u, g0);
    const float g2 = corner(h.e01.x, h.e01.y, x0, y1, z0);
    const float l1 = fmaf(corner_minus(h.e11.x, h.e11.y, x1, y1, z0, g2), u, g2);
    const float g4 = corner(h.e00.z, h.e00.w, x0, y0, z1);
    const float l2 = fmaf(corner_minus(h.e10.z, h.e10.w, x1, y0, z1, g4), u, g4);
    const float g6 = corner(h.e01.z, h.e01.w, x0, y1, z1);
    const float l3 = fmaf(corner_minus(h.e11.z, h.e11.w, x1, y1, z1, g6), u, g6);
    return lerp1(lerp1(l0, l1, v), lerp1(l2, l3, v), w);            // perlin.h:77-86
}
template <int REPL>
__device__ __forceinline__ float noise_octave(const LaneTab &tab, const Fixed3 &p, int k, uint32_t one_bits)
{
    return noise_from(hash_octave<REPL>(tab, p, k, one_bits));
}

// fractal sum over octaves for the thread's two samples (main.cpp:689-734 with FMA); the two
// independent chains interleave in the instruction stream.  `half_amp` carries amplitude/2
// because noise_octave returns 2*noise.
template <int REPL, int N, bool GUARD>
__device__ __forceinline__ void fractal_loop(const LaneTab &tab, const Fixed3 (&p)[N], const int (&octaves)[N], int omax,
                                             int kind, float gain, uint32_t one_bits, float (&value)[N])
{
    float half_amp = 0.5f;
#pragma unroll
    for (int s = 0; s < N; s++) value[s] = 0.0f;
    if (kind == PLANET_NOISE_RIDGED) {                               // main.cpp:716-731
        float weight[N];
#pragma unroll
        for (int s = 0; s < N; s++) weight[s] = 1.0f;
#pragma unroll K2_UNROLL
        for (int k = 0; k < omax; k++) {
#pragma unroll
            for (int s = 0; s < N; s++) {
                const float n = noise_octave<REPL>(tab, p[s], k, one_bits);
                float v = fmaf(-0.5f, fabsf(n), 1.0f);              // offset - |noise|
                v = v * v;
                const float nv = fmaf(v * (2.0f * half_amp), weight[s], value[s]);
                if (!GUARD || k < octaves[s]) { value[s] = nv; weight[s] = v; }
            }
            half_amp *= gain;
        }
    } else {                                                         // main.cpp:699-704
#pragma unroll K2_UNROLL
        for (int k = 0; k < omax; k++) {
#pragma unroll
            for (int s = 0; s < N; s++) {
                const float n = noise_octave<REPL>(tab, p[s], k, one_bits);
                if (!GUARD || k < octaves[s]) value[s] = fmaf(n, half_amp, value[s]);
            }
            half_amp *= gain;                                        // main.cpp:703
        }
    }
}

template <int REPL, int N>
__device__ __forceinline__ void fractal(const LaneTab &tab, const Fixed3 (&p)[N], const int (&octaves)[N], int kind,
                                        float gain, uint32_t one_bits, float (&value)[N])
{
    int omax = octaves[0];
    bool same = true;
#pragma unroll
    for (int s = 1; s < N; s++) { omax = max(omax, octaves[s]); same = same && octaves[s] == octaves[0]; }
    if (same) fractal_loop<REPL, N, false>(tab, p, octaves, omax, kind, gain, one_bits, value);
    else      fractal_loop<REPL, N, true>(tab, p, octaves, omax, kind, gain, one_bits, value);
}

// ---- EXACT arithmetic on the replicated tables ------------------------------------------
// Every rounding of perlin.h:50-88 as in exact::noise3 (planet_common.cuh) -- custom floor,
// double fraction, quintic in double, x-1 formed in double, unfused float lerps -- but the hash
// chain and the gradient fetch come from the lane-replicated tables above: 10 conflict-free
// LDS instead of 14 byte loads + 24 float loads with ~3-way conflicts, no `& 255` / `* 3` index
// arithmetic.  Two identities keep the bits: (1) gradient components are 0 or +-1, so every
// product of the dot is exact and x*g0 + y*g1 + z*g2 with separate roundings equals the same sum
// written as two FMAs; (2) the table holds 2*v, so the value is 2*noise, and scaling
// by two commutes with every rounding on the way (no value here is near the denormal range).
// Returns 2 * PerlinNoise3(x, y, z).
__device__ __forceinline__ float corner_exact2(uint32_t gxw, uint32_t gyw, float x, float y, float z)
{
    const float gx = __uint_as_float(gxw & 0xFF000000u);             // drop the stowaway z code
    const float gy = __uint_as_float(gyw);
    const float gz = __uint_as_float(gxw << 30);
    return __fmaf_rn(z, gz, __fmaf_rn(y, gy, __fmul_rn(x, gx)));     // perlin.h:47
}

template <int REPL>
__device__ __forceinline__ float noise3_exact2(const LaneTab &tab, double x, double y, double z)
{
    using L = Layout<REPL>;
    const int ix = exact::cell(x), iy = exact::cell(y), iz = exact::cell(z);     // perlin.h:52-55
    x = __dsub_rn(x, (double)ix);
    y = __dsub_rn(y, (double)iy);
    z = __dsub_rn(z, (double)iz);
    const float u = exact::fade(x), v = exact::fade(y), w = exact::fade(z);
    const float x0 = __double2float_rn(x), x1 = __double2float_rn(__dadd_rn(x, -1.0));   // perlin.h:68-75
    const float y0 = __double2float_rn(y), y1 = __double2float_rn(__dadd_rn(y, -1.0));
    const float z0 = __double2float_rn(z), z1 = __double2float_rn(__dadd_rn(z, -1.0));
    const uint32_t cx = (((uint32_t)ix & 255u) << L::LOGP) | tab.lp;    // PerlinRandom's `& 255`, perlin.h:40
    const uint32_t cy = (((uint32_t)iy & 255u) << L::LOGP) | tab.lp;
    const uint32_t cz = (((uint32_t)iz & 255u) << L::LOG3) | tab.l3;
    const unsigned char *t = tab.base;
    const uint2 a = lds_v2(t, cx);
    const uint2 b0 = lds_v2(t, a.x + cy + L::P2_AT), b1 = lds_v2(t, a.y + cy + L::P2_AT);
    const uint4 e00 = lds_v4(t, b0.x + cz + L::T3_AT), e10 = lds_v4(t, b1.x + cz + L::T3_AT);
    const uint4 e01 = lds_v4(t, b0.y + cz + L::T3_AT), e11 = lds_v4(t, b1.y + cz + L::T3_AT);
    const float g0 = corner_exact2(e00.x, e00.y, x0, y0, z0), g1 = corner_exact2(e10.x, e10.y, x1, y0, z0);
    const float g2 = corner_exact2(e01.x, e01.y, x0, y1, z0), g3 = corner_exact2(e11.x, e11.y, x1, y1, z0);
    const float g4 = corner_exact2(e00.z, e00.w, x0, y0, z1), g5 = corner_exact2(e10.z, e10.w, x1, y0, z1);
    const float g6 = corner_exact2(e01.z, e01.w, x0, y1, z1), g7 = corner_exact2(e11.z, e11.w, x1, y1, z1);
    const float l0 = exact::lerp(g0, g1, u), l1 = exact::lerp(g2, g3, u);
    const float l2 = exact::lerp(g4, g5, u), l3 = exact::lerp(g6, g7, u);
    return exact::lerp(exact::lerp(l0, l1, v), exact::lerp(l2, l3, v), w);   // perlin.h:77-86
}

// Perlin::operator() (main.cpp:825-832) over noise3_exact2; same operation order as exact::height
template <int REPL>
__device__ __forceinline__ float height_exact_tab(const LaneTab &tab, const HeightCfg &c, d3 p, int depth)
{
    if (c.kind == PLANET_NOISE_ZERO) return 0.0f;
    const int octaves = octaves_for(c.fixed_octaves, depth, c.max_depth);
    p = exact::mul(p, c.coord_scale);
    if (c.has_seed) {
        p.x = __dadd_rn(p.x, c.seed[0]); p.y = __dadd_rn(p.y, c.seed[1]); p.z = __dadd_rn(p.z, c.seed[2]);
    }
    double frequency = 1.0;
    float amplitude = 1.0f, weight = 1.0f, value = 0.0f;
    const bool ridged = c.kind == PLANET_NOISE_RIDGED;
    for (int i = 0; i < octaves; ++i) {
        float n = __fmul_rn(0.5f, noise3_exact2<REPL>(tab, __dmul_rn(p.x, frequency),
                                                      __dmul_rn(p.y, frequency), __dmul_rn(p.z, frequency)));
        if (ridged) {                                                // main.cpp:716-731
            n = fabsf(n);
            n = __fsub_rn(1.0f, n);
            n = __fmul_rn(n, n);
            value = __fadd_rn(value, __fmul_rn(__fmul_rn(n, amplitude), weight));
            weight = n;
        } else {                                                     // main.cpp:699-704
            value = __fadd_rn(value, __fmul_rn(n, amplitude));
        }
        frequency = __dmul_rn(frequency, c.lacunarity);
        amplitude = __fmul_rn(amplitude, c.gain);
    }
    return __fmul_rn(value, c.height_scale);
}

// GenerateHeightMap (main.cpp:123-151) in EXACT arithmetic for large batches: persistent CTAs,
// one per SM, one sample per thread per pass, consecutive threads on consecutive texels.
template <int NTHREADS, bool GATHER>
__global__ void __launch_bounds__(NTHREADS, 1)
k_height_maps_exact_tab(const Quad *__restrict__ quads, int64_t total, int dim, HeightCfg cfg,
                        float *__restrict__ out, uint64_t magic_dim, PeerOut peers)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using L = Layout<32>;
    unsigned char *smem = tables_at<32>(smem_raw);
    build_tables<32>(smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const LaneTab tab = lane_tab<32>(smem, lane);
    const uint32_t dim2 = (uint32_t)dim * (uint32_t)dim;
    const double div = __ddiv_rn(1.0, (double)(dim - 3));            // main.cpp:134
    // i = q * dim2 + r is advanced incrementally: one 64-bit division per thread, none per sample
    const int64_t i0 = blockIdx.x * (int64_t)NTHREADS + threadIdx.x;
    const uint32_t step = gridDim.x * NTHREADS;
    const uint32_t step_q = step / dim2, step_r = step - step_q * dim2;
    int64_t q = i0 / dim2;
    uint32_t r = (uint32_t)(i0 - q * dim2);
    for (int64_t i = i0; i < total; i += step) {
        const uint32_t y = (uint32_t)(((uint64_t)r * magic_dim) >> 40), x = r - y * (uint32_t)dim;   // exact: r < dim^2, dim <= 8192
        const Quad *quad = quads + q;
        Quad qd;
        qd.p[0] = quad->p[0]; qd.p[1] = quad->p[1]; qd.p[2] = quad->p[2]; qd.p[3] = quad->p[3];
        const d3 p = exact::sample_point(qd, (int)x, (int)y, div);
        const float h = height_exact_tab<32>(tab, cfg, p, (int)quad_depth(quad->id));
        out[i] = h;
        if constexpr (GATHER) {
#pragma unroll
            for (int k = 0; k < 7; k++)
                if (k < peers.n) peers.ptr[k][i] = h;
        }
        q += step_q; r += step_r;
        if (r >= dim2) { r -= dim2; q++; }
    }
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

// exact r / d for r*d < 2^40 with m = ceil(2^40 / d) (host-computed)
__device__ __forceinline__ uint32_t div_magic(uint32_t r, uint64_t m) { return (uint32_t)(((uint64_t)r * m) >> 40); }

// ---- height maps --------------------------------------------------------------------
// Warps are autonomous: after the one-time table build there is no block-wide barrier.
// A warp owns "warp tiles" of WTILE consecutive samples; for each it turns the quads the
// tile touches into bilinear coefficients in its private shared-memory scratch (3 lanes per
// quad, one per axis), then walks the tile in SUB passes of 64 samples, 2 per lane.
constexpr int SUB = 2;
constexpr int WTILE = 32 * S * SUB;                      // 128 samples per warp tile
constexpr int MAX_WTILE_QUADS = WTILE / 16 + 2;          // dim >= 4

// GATHER is a template flag, not a run-time test of peers.n: with the peer pointers live in the
// plain kernel the octave loop grew by 14 address adds (register pressure), 6 % of the run time.
// KIND is the noise kind as a template parameter (fBm or ridged; ZERO runs as fBm with no octaves):
// the per-sub-tile dispatch on it was 3 % of the kernel's time in branches and reconvergence.
template <int NTHREADS, int REPL, bool GATHER, int KIND>
__global__ void __launch_bounds__(NTHREADS, REPL == 32 ? 1 : 2)
k_height_maps_fast(const Quad *__restrict__ quads, int64_t total, int dim, HeightCfg cfg,
                   float *__restrict__ out, int64_t nwtiles, int out_aligned8,
                   uint64_t magic_dim, uint64_t magic_dim2, uint32_t one_bits, PeerOut peers, int bulk_stores)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using L = Layout<REPL>;
    unsigned char *smem = tables_at<REPL>(smem_raw);
    // the peer pointers wait in shared memory until the stores: held in registers across the
    // octave loop they cost it 14 rematerialised address adds.  s_peer[0] is the kernel's own
    // buffer: destination d of a finished tile is s_peer[d], d = 0 .. peers.n
    __shared__ float *s_peer[8];
    if (GATHER && threadIdx.x < 8) s_peer[threadIdx.x] = threadIdx.x == 0 ? out : peers.ptr[threadIdx.x - 1];
    if constexpr (GATHER) {
        // the buffer this launch overwrites on the peers must have been released by them
        // (k4_gather.cu: double-buffered gathered buffers, no host barrier inside a step)
        if (peers.release && (int)threadIdx.x < peers.world && (int)threadIdx.x != peers.rank)
            gather_wait_flag(peers.release + threadIdx.x, peers.release_min, peers.error);
    }
    build_tables<REPL>(smem);
    __syncthreads();
    // Staging for the fused gather: a finished 128-sample tile (512 B) is assembled in shared
    // memory and leaves the SM as ONE bulk copy per destination (planet_tma.cuh) instead of a
    // float2 store per lane per destination.  The buffers live in the slack the 32 KB alignment
    // of the tables leaves in front of (or behind) them: two per warp when it is large enough
    // (always, in practice: the dynamic array starts ~1 KB into the shared window), else one.
    float *stage = nullptr;
    int stage_bufs = 0;
    if constexpr (GATHER) {
        const uint32_t front = (uint32_t)(smem - smem_raw);
        const uint32_t back = (uint32_t)(PLANET_K2_ALIGN ? L::P1_BYTES : 0) - front;
        const uint32_t per_warp = WTILE * sizeof(float);
        unsigned char *region = front >= back ? smem_raw : smem + L::TABLES + (NTHREADS / 32) * MAX_WTILE_QUADS * sizeof(TileQuad);
        const uint32_t room = front >= back ? front : back;
        stage_bufs = bulk_stores ? (int)min(2u, room / ((NTHREADS / 32) * per_warp)) : 0;
        region += (16u - ((uint32_t)__cvta_generic_to_shared(region) & 15u)) & 15u;
        stage = reinterpret_cast<float *>(region) + (threadIdx.x >> 5) * stage_bufs * WTILE;
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WARPS = NTHREADS / 32;
    // progress mode of the gather: lane 0 publishes the warp's finished-tile count (release: the tile's
    // stores, ordered before it by the __syncwarp, are visible to whoever acquires the counter)
    auto publish = [&](uint32_t tiles_done, bool last) {
        if constexpr (GATHER) {
            if (peers.progress && (last || (tiles_done & peers.progress_mask) == 0)) {
                __syncwarp();
                if (lane == 0) {
                    const unsigned long long v = ((unsigned long long)peers.progress_tag << 32) | tiles_done;
                    unsigned long long *at = peers.progress + (blockIdx.x * WARPS + warp);
                    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(at), "l"(v) : "memory");
                }
            }
        }
    };
    TileQuad *tq = reinterpret_cast<TileQuad *>(smem + L::TABLES) + warp * MAX_WTILE_QUADS;
    const LaneTab tab = lane_tab<REPL>(smem, lane, one_bits ^ ONE_BITS);
    const uint32_t dim2 = (uint32_t)dim * (uint32_t)dim;
    const bool small_maps = dim2 < (uint32_t)WTILE;          // a warp tile may then span > 2 quads
    const double div = 1.0 / (double)(dim - 3);
    const double s = cfg.coord_scale * FIX_ONE;

    // Every warp of the grid owns one contiguous run of warp tiles, so consecutive tiles mostly
    // stay inside one quad (8 tiles per 32 x 32 map) and its coefficients are built once;
    // (q_first, r_base) = divmod(tile start, dim2) is advanced incrementally, so the only 64-bit
    // divisions happen once, here
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    const int64_t per_warp = (nwtiles + nwarps - 1) / nwarps;
    int64_t wt = ((int64_t)blockIdx.x * WARPS + warp) * per_warp;
    const int64_t wt_end = min(nwtiles, wt + per_warp);
    const uint32_t wt_first = (uint32_t)wt;                          // (progress mode: tiles done = wt - wt_first)
    int64_t q_first = (wt * WTILE) / dim2;
    uint32_t r_base = (uint32_t)(wt * WTILE - q_first * dim2);
    int64_t q_built = -1;                                            // quad whose coefficients sit in tq[0]

    // the lane's two texels of a sub-tile of a REGULAR tile (below): one quad, same row
    auto positions_regular = [&](uint32_t r, Fixed3 *p) {
        const uint32_t yy = div_magic(r, magic_dim), xx = r - yy * (uint32_t)dim;   // xx even: xx + 1 is on the same row
        const TileQuad &c = tq[0];
        const double xd = (double)((int)xx - 1), yd = (double)((int)yy - 1);
        double P0[3], P1[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double step = fma(c.ax[a].d, yd, c.ax[a].b);
            P0[a] = fma(yd, c.ax[a].c, fma(step, xd, c.ax[a].a));
            P1[a] = P0[a] + step;
        }
        to_fixed(P0[0], p[0].xlo, p[0].xhi); to_fixed(P0[1], p[0].ylo, p[0].yhi); to_fixed(P0[2], p[0].zlo, p[0].zhi);
        to_fixed(P1[0], p[1].xlo, p[1].xhi); to_fixed(P1[1], p[1].ylo, p[1].yhi); to_fixed(P1[2], p[1].zlo, p[1].zhi);
    };
    // plain (write-back, evict-normal) store: the maps are K3's input and a C2 batch (67 MB)
    // fits the 126 MB L2; a streaming store made K3 re-read them from HBM (K3 0.117 -> 0.100 ms)
    auto store_pair = [&](int64_t o, const float *value) {
        const float2 h2 = make_float2(value[0] * cfg.height_scale, value[1] * cfg.height_scale);
        *reinterpret_cast<float2 *>(out + o) = h2;
        if constexpr (GATHER) {
#pragma unroll
            for (int r = 0; r < 7; r++)                              // NVLink peer stores (tiles the bulk path cannot take)
                if (r < peers.n) __stcs(reinterpret_cast<float2 *>(((float *volatile *)s_peer)[r + 1] + o), h2);
        }
    };

    int buf = 0;                                                     // staging buffer of the fused gather in use
    while (wt < wt_end) {
        const int64_t base = wt * WTILE;
        const int n_here = (int)min((int64_t)WTILE, total - base);           // samples in this warp tile
        const uint32_t r_end = r_base + (uint32_t)n_here - 1;
        const int nq = (int)(small_maps ? div_magic(r_end, magic_dim2) : (uint32_t)(r_end >= dim2)) + 1;

        // prologue: quad -> per-axis bilinear coefficients (main.cpp:130-146 regrouped); skipped when
        // the tile lies in the one quad the previous tile left in tq[0]
        const bool have = nq == 1 && q_first == q_built;                     // warp-uniform
        q_built = nq == 1 ? q_first : -1;
        __syncwarp();
        if (!have && lane < nq * 3) {
            const int qi = lane / 3, axis = lane - qi * 3;
            const double *qp = reinterpret_cast<const double *>(quads + q_first + qi);
            double p0 = qp[axis], p1 = qp[3 + axis], p2 = qp[6 + axis], p3 = qp[9 + axis];
            double v0 = p1 - p0, v1 = p3 - p2;
            AxisCoef c;
            const double seed = !cfg.has_seed ? 0.0 : axis == 0 ? cfg.seed[0] : axis == 1 ? cfg.seed[1] : cfg.seed[2];
            c.a = wrap_period((p0 * cfg.coord_scale + seed) * FIX_ONE);
            c.b = v0 * (s * div);
            c.c = (p2 - p0) * (s * div);
            c.d = (v1 - v0) * (s * div * div);
            tq[qi].ax[axis] = c;
            // a quad whose scaled extent nears half a period cannot keep |P| < 2^63 from a
            // reduced corner alone; such quads take a per-sample reduction instead
            double span = (fabs(c.b) + fabs(c.c)) * (double)dim + fabs(c.d) * (double)dim * (double)dim;
            unsigned wide = __ballot_sync(__activemask(), span > 0.4 * FIX_WRAP);
            if (axis == 0) {
                uint64_t id = reinterpret_cast<const uint64_t *>(qp)[12];
                tq[qi].octaves = cfg.kind == PLANET_NOISE_ZERO ? 0 : octaves_for(cfg.fixed_octaves, (int)quad_depth(id), cfg.max_depth);
                tq[qi].wide = (wide >> lane) & 7u;                           // any of this quad's 3 axes
            }
        }
        __syncwarp();

        // Regular tiles -- a full tile inside one quad of even dim, not wide, 8-byte aligned output:
        // every tile of the 32 x 32 maps of the reference -- skip the ragged-edge bookkeeping
        // (second texel on another row / quad, per-sample range reduction, partial tiles), and so
        // do the full tiles that follow in the same quad: the whole run is walked here with no
        // per-tile setup at all (15 % of the kernel's time was outside the octave loops).
        const bool regular = out_aligned8 && n_here == WTILE && nq == 1 && !(dim & 1) && tq[0].wide == 0;
        if (regular) {
            const int run = (int)min(wt_end - wt, (int64_t)((dim2 - r_base) / (uint32_t)WTILE));   // >= 1
            const int octaves = tq[0].octaves;
            uint32_t r = r_base + (uint32_t)(lane * S);
            int64_t o = base + lane * S;                                     // even
            // (fused gather) maps left to the shade kernel's push are only stored locally
            bool push_this_quad = true;
            if constexpr (GATHER)
                if (stage_bufs && peers.k3_every != 0)
                    push_this_quad = !shade_pushes_quad(peers.quad0 + q_first, peers.k3_every);
#pragma unroll 1
            for (int t = 0; t < run * SUB; t++, r += 32 * S, o += 32 * S) {
                Fixed3 p[S];
                const int oct[S] = { octaves, octaves };
                float value[S];
                positions_regular(r, p);
                fractal_loop<REPL, S, false>(tab, p, oct, octaves, KIND, cfg.gain, one_bits, value);
                if constexpr (GATHER) {
                    if (stage_bufs && push_this_quad) {
                        // fused gather: the tile is assembled in the warp's staging buffer and pushed to
                        // every destination (this GPU + the peers over NVLink) by lanes 0 .. n, one
                        // 512-byte bulk copy each
                        float *sb = stage + buf * WTILE;
                        *reinterpret_cast<float2 *>(sb + (t & (SUB - 1)) * (32 * S) + lane * S) =
                            make_float2(value[0] * cfg.height_scale, value[1] * cfg.height_scale);
                        if ((t & (SUB - 1)) == SUB - 1) {
                            tma::fence_smem_writes();
                            __syncwarp();
                            if (lane <= peers.n) {
                                float *dst = ((float *volatile *)s_peer)[lane] + (o - lane * S - (SUB - 1) * 32 * S);
                                tma::store_bulk(dst, sb, WTILE * sizeof(float));
                                tma::commit();
                                // the buffer written next was handed over one tile ago (two buffers) / just now (one)
                                if (stage_bufs == 2) tma::wait_read<1>(); else tma::wait_read<0>();
                            }
                            buf ^= stage_bufs - 1;
                            __syncwarp();
                        }
                        continue;
                    }
                    if (stage_bufs) {                                        // this quad's map travels from the shade kernel
                        *reinterpret_cast<float2 *>(out + o) = make_float2(value[0] * cfg.height_scale, value[1] * cfg.height_scale);
                        continue;
                    }
                }
                store_pair(o, value);
                if constexpr (GATHER)
                    if ((t & (SUB - 1)) == SUB - 1) publish((uint32_t)wt - wt_first + (uint32_t)(t / SUB) + 1u, wt + t / SUB + 1 == wt_end);
            }
            wt += run;
            r_base += (uint32_t)run * WTILE;
            if (r_base >= dim2) { r_base -= dim2; q_first++; }               // the run ended with the quad
            continue;
        }

        // the lane's two texels of sub-tile `sub`: fixed-point positions and octave counts
        auto positions = [&](int sub, Fixed3 *p, int *oct) {
            const int i0 = sub * (32 * S) + lane * S;                        // first sample of this lane
            // texel (q, y, x) of the lane's first sample; the second one follows by increment
            uint32_t r = r_base + (uint32_t)min(i0, n_here - 1);
            uint32_t q[S], y[S], x[S];
            q[0] = small_maps ? div_magic(r, magic_dim2) : (uint32_t)(r >= dim2);
            r -= q[0] * dim2;
            y[0] = div_magic(r, magic_dim);
            x[0] = r - y[0] * (uint32_t)dim;
            q[1] = q[0]; y[1] = y[0]; x[1] = x[0];
            if (i0 + 1 < n_here) {
                if (++x[1] == (uint32_t)dim) { x[1] = 0; if (++y[1] == (uint32_t)dim) { y[1] = 0; ++q[1]; } }
            }
            // sample 0: P = A + B x + y (C + D x) per axis (coefficients pre-scaled by 2^55)
            const TileQuad &c = tq[q[0]];
            const double xd = (double)((int)x[0] - 1), yd = (double)((int)y[0] - 1);
            double P0[3], P1[3], step[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                step[a] = fma(c.ax[a].d, yd, c.ax[a].b);                     // dP/dx along this row
                P0[a] = fma(yd, c.ax[a].c, fma(step[a], xd, c.ax[a].a));
                P1[a] = P0[a] + step[a];                                     // the neighbour texel, same row
            }
            int wide = c.wide;
            oct[0] = oct[1] = c.octaves;
            // warp-uniform slow paths: the second texel starts a new row / quad (odd dim), or a
            // quad too wide for the per-quad reduction (then every sample is reduced on its own)
            if (__any_sync(0xffffffffu, (q[1] != q[0]) | (y[1] != y[0]))) {
                if (q[1] != q[0] || y[1] != y[0]) {
                    const TileQuad &c1 = tq[q[1]];
                    const double xd1 = (double)((int)x[1] - 1), yd1 = (double)((int)y[1] - 1);
#pragma unroll
                    for (int a = 0; a < 3; a++)
                        P1[a] = fma(yd1, fma(c1.ax[a].d, xd1, c1.ax[a].c), fma(c1.ax[a].b, xd1, c1.ax[a].a));
                    wide |= c1.wide;
                    oct[1] = c1.octaves;
                }
            }
            if (__any_sync(0xffffffffu, wide)) {
                if (wide) {
#pragma unroll
                    for (int a = 0; a < 3; a++) { P0[a] = wrap_period(P0[a]); P1[a] = wrap_period(P1[a]); }
                }
            }
            to_fixed(P0[0], p[0].xlo, p[0].xhi); to_fixed(P0[1], p[0].ylo, p[0].yhi); to_fixed(P0[2], p[0].zlo, p[0].zhi);
            to_fixed(P1[0], p[1].xlo, p[1].xhi); to_fixed(P1[1], p[1].ylo, p[1].yhi); to_fixed(P1[2], p[1].zlo, p[1].zhi);
        };
        auto store = [&](int sub, const float *value) {
            const int i0 = sub * (32 * S) + lane * S;
            const int64_t o = base + i0;                                     // even
            if (out_aligned8 && i0 + 1 < n_here) {
                store_pair(o, value);
            } else {
#pragma unroll
                for (int sidx = 0; sidx < S; sidx++)
                    if (i0 + sidx < n_here) {
                        const float h = value[sidx] * cfg.height_scale;
                        out[o + sidx] = h;
                        if constexpr (GATHER) {
#pragma unroll
                            for (int r = 0; r < 7; r++)
                                if (r < peers.n) ((float *volatile *)s_peer)[r + 1][o + sidx] = h;
                        }
                    }
            }
        };
        // (four chains per thread -- both sub-tiles in one octave loop, 127 registers at 512 threads --
        // were measured too: 0.476 ms against 0.463 ms for this form at 768 threads)
#pragma unroll 1
        for (int sub = 0; sub < SUB; sub++) {
            if (sub * (32 * S) >= n_here) break;
            Fixed3 p[S];
            int oct[S];
            float value[S] = { 0.0f, 0.0f };
            positions(sub, p, oct);
            fractal<REPL, S>(tab, p, oct, KIND, cfg.gain, one_bits, value);
            store(sub, value);
        }

        wt++;
        publish((uint32_t)wt - wt_first, wt == wt_end);
        r_base += WTILE;                                                     // the next tile of this warp
        if (small_maps) { const uint32_t dq = div_magic(r_base, magic_dim2); q_first += dq; r_base -= dq * dim2; }
        else if (r_base >= dim2) { r_base -= dim2; q_first++; }
    }
    if constexpr (GATHER) {
        if (stage_bufs && lane <= peers.n) tma::wait_all<0>();               // every pushed tile has been written
    }
}

// ---- points: batched GetHeightAt and raw noise ----------------------------------------
// scale / seed / height_scale == 1 / 0 / 1 and octaves0 == 1 give PerlinNoise3 itself.
template <int REPL>
__global__ void __launch_bounds__(THREADS, REPL == 32 ? 1 : 2)
k_points_fast(const double *__restrict__ xyz, int64_t n, int kind, float gain, int octaves,
              double coord_scale, double sx, double sy, double sz, float height_scale,
              float *__restrict__ out, int64_t ntiles, uint32_t one_bits)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = tables_at<REPL>(smem_raw);
    build_tables<REPL>(smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const LaneTab tab = lane_tab<REPL>(smem, lane, one_bits ^ ONE_BITS);

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
        fractal<REPL, S>(tab, p, oct, kind, gain, one_bits, value);
#pragma unroll
        for (int s = 0; s < S; s++)
            if (idx[s] < n) out[idx[s]] = value[s] * height_scale;
    }
}

} // namespace fast

// =====================================================================================
// host-side launchers
// =====================================================================================
// per-device caches (a process may drive several GPUs)
static int g_sm_count[64] = {};
static bool g_fast_attr_set[64] = {};

static int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & 63;
}

static int sm_count()
{
    const int dev = current_device();
    if (!g_sm_count[dev]) {
        cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count[dev] <= 0) g_sm_count[dev] = 148;
    }
    return g_sm_count[dev];
}

static int k2_threads()
{
    static int t = 0;
    if (!t) {
        const char *e = getenv("PLANET_K2_THREADS");            // tuning knob: 512 or 768
        t = e ? atoi(e) : 768;                                  // 768 measured 2 % faster than 512 on C2
        if (t != 512 && t != 768) t = 768;
    }
    return t;
}

// batches of at most this many samples use the compact-table kernels (no 192 KB table build)
static int64_t k2_small_max()
{
    static int64_t v = -1;
    if (v < 0) {
        const char *e = getenv("PLANET_K2_SMALL_MAX");
        v = e ? atoll(e) : (1 << 20);
    }
    return v;
}

// the twelve instantiations of the height-map kernel: (threads, table layout) x gather x kind
template <int NT, int REPL> static const void *fast_kernel(bool gather, bool ridged)
{
    constexpr int F = PLANET_NOISE_FBM, R = PLANET_NOISE_RIDGED;
    if (gather) return ridged ? (const void *)fast::k_height_maps_fast<NT, REPL, true, R> : (const void *)fast::k_height_maps_fast<NT, REPL, true, F>;
    return ridged ? (const void *)fast::k_height_maps_fast<NT, REPL, false, R> : (const void *)fast::k_height_maps_fast<NT, REPL, false, F>;
}

static int prepare_fast()
{
    const int dev = current_device();
    if (!g_fast_attr_set[dev]) {
        for (int g = 0; g < 2; g++)
            for (int r = 0; r < 2; r++) {
                PLANET_CUDA(cudaFuncSetAttribute(fast_kernel<512, 32>(g, r), cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
                PLANET_CUDA(cudaFuncSetAttribute(fast_kernel<768, 32>(g, r), cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
            }
        PLANET_CUDA(cudaFuncSetAttribute(fast::k_points_fast<32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
        PLANET_CUDA(cudaFuncSetAttribute(fast::k_height_maps_exact_tab<1024, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
        PLANET_CUDA(cudaFuncSetAttribute(fast::k_height_maps_exact_tab<1024, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, fast::SMEM_BYTES));
        g_fast_attr_set[dev] = true;
    }
    return 0;
}

template <int REPL>
static void launch_points(int grid, cudaStream_t stream, const double *d_xyz, int64_t n, int kind, float gain,
                          int octaves, double scale, double sx, double sy, double sz, float hs, float *d_out,
                          int64_t ntiles)
{
    fast::k_points_fast<REPL><<<grid, fast::THREADS, fast::smem_bytes<REPL>(fast::THREADS), stream>>>(
        d_xyz, n, kind, gain, octaves, scale, sx, sy, sz, hs, d_out, ntiles, fast::ONE_BITS);
}

// FAST handles the lattice split as shifts, which needs lacunarity == 2 and every octave's
// window inside the 64-bit word; anything else runs the EXACT kernel on the GPU.
static bool fast_applicable(double lacunarity, int max_octaves)
{
    return lacunarity == 2.0 && max_octaves <= 32;
}

int launch_height_maps_gathered(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, int dim,
                                int max_depth, float *d_out, const PeerOut &peers, cudaStream_t stream);

int launch_height_maps(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, int dim,
                       int max_depth, float *d_out, cudaStream_t stream)
{
    PeerOut none = {};
    return launch_height_maps_gathered(p, d_quads, nquads, dim, max_depth, d_out, none, stream);
}

// does a gathered launch with these arguments push its tiles as bulk copies (the FAST kernel on the
// replicated tables, every destination 16-byte aligned)?  Only then can part of the pushing be
// left to the shade kernel (PeerOut::k3_every).
bool height_maps_push_in_bulk(const planet_gpu_params *p, int64_t nquads, int dim, int max_depth, const float *d_out,
                              const PeerOut &peers)
{
    HeightCfg cfg = make_cfg(p, max_depth);
    const int64_t total = nquads * (int64_t)dim * dim;
    const int max_oct = octaves_for(cfg.fixed_octaves, 31, max_depth != 0 ? max_depth : 1);
    bool ok = p->precision == PLANET_PRECISION_FAST && dim <= 8192 && fast_applicable(cfg.lacunarity, max_oct) &&
              total > k2_small_max() && !(dim & 1) && (dim * dim) % fast::WTILE == 0 &&
              (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && !getenv("PLANET_K2_NO_BULK");
    for (int r = 0; r < peers.n; r++) ok = ok && (reinterpret_cast<uintptr_t>(peers.ptr[r]) & 15) == 0;
    return ok;
}

// The tile ownership of the big FAST kernel for a batch of `total` samples -- warps in the grid, tiles
// per warp, tiles in all -- for the pusher kernel that follows the kernel's progress counters
// (k4_gather.cu).  False when this batch would not run on that kernel (EXACT arithmetic, the
// small-batch kernel, unaligned output): progress mode is then not available.
bool height_maps_progress_layout(const planet_gpu_params *p, int64_t nquads, int dim, int max_depth, const float *d_out,
                                 int *nwarps, int64_t *per_warp, int64_t *nwtiles)
{
    HeightCfg cfg = make_cfg(p, max_depth);
    const int64_t total = nquads * (int64_t)dim * dim;
    const int max_oct = octaves_for(cfg.fixed_octaves, 31, max_depth != 0 ? max_depth : 1);
    if (!(p->precision == PLANET_PRECISION_FAST && dim <= 8192 && fast_applicable(cfg.lacunarity, max_oct) &&
          total > k2_small_max() && (total & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0))
        return false;
    const int nt = k2_threads();
    *nwtiles = (total + fast::WTILE - 1) / fast::WTILE;
    const int grid = (int)std::min<int64_t>((*nwtiles + nt / 32 - 1) / (nt / 32), sm_count());
    *nwarps = grid * (nt / 32);
    *per_warp = (*nwtiles + *nwarps - 1) / *nwarps;
    return true;
}

int launch_height_maps_gathered(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, int dim,
                                int max_depth, float *d_out, const PeerOut &peers, cudaStream_t stream)
{
    HeightCfg cfg = make_cfg(p, max_depth);
    int64_t total = nquads * (int64_t)dim * dim;
    if (total == 0) return 0;
    // FAST needs every octave's window inside the 64-bit fixed-point word (<= 32 octaves) and the
    // magic-number row division exact (dim^3 < 2^40).  The quads live on the device, so the octave
    // bound is taken at the deepest id the 5-bit depth field can hold (main.cpp:27): with the
    // reference's max_depth = 18 that is 6 + 12*31/18 = 26.  Anything else runs EXACT on the GPU.
    const int max_oct = octaves_for(cfg.fixed_octaves, 31, max_depth != 0 ? max_depth : 1);
    bool use_fast = p->precision == PLANET_PRECISION_FAST && dim <= 8192 &&
                    fast_applicable(cfg.lacunarity, max_oct);
    if (use_fast) {
        int rc = prepare_fast();
        if (rc) return rc;
        int64_t nwtiles = (total + fast::WTILE - 1) / fast::WTILE;
        const uint64_t one40 = 1ull << 40;
        const uint64_t m1 = (one40 + dim - 1) / dim, m2 = (one40 + (uint64_t)dim * dim - 1) / ((uint64_t)dim * dim);
        int al = (reinterpret_cast<uintptr_t>(d_out) & 7) == 0;
        for (int r = 0; r < peers.n; r++) al = al && (reinterpret_cast<uintptr_t>(peers.ptr[r]) & 7) == 0;
        const bool gather = peers.n > 0 || peers.release != nullptr || peers.progress != nullptr, ridged = cfg.kind == PLANET_NOISE_RIDGED;
        uint32_t one_bits = fast::ONE_BITS;
        PeerOut peers_arg = peers;
        // tiles leave the SM as 512-byte bulk copies when every destination is 16-byte aligned
        int bulk = gather && !peers.progress && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && !getenv("PLANET_K2_NO_BULK");
        for (int r = 0; r < peers.n; r++) bulk = bulk && (reinterpret_cast<uintptr_t>(peers.ptr[r]) & 15) == 0;
        void *args[] = { (void *)&d_quads, &total, &dim, &cfg, (void *)&d_out, &nwtiles, &al, (void *)&m1, (void *)&m2, &one_bits, &peers_arg, &bulk };
        auto launch_fast = [&](const void *kern, int grid, int threads, size_t smem, cudaStream_t st) -> int {
            return check_cuda(cudaLaunchKernel(kern, dim3(grid), dim3(threads), args, smem, st), "height map kernel launch");
        };
        if (total <= k2_small_max()) {
            // latency path: compact tables (6 KB), 256-thread CTAs spread over the whole chip
            int grid = (int)std::min<int64_t>((nwtiles + 7) / 8, (int64_t)sm_count() * 8);
            rc = launch_fast(fast_kernel<256, 1>(gather, ridged), grid, 256, fast::smem_bytes<1>(256), stream);
        } else {
            const int nt = k2_threads();
            int grid = (int)std::min<int64_t>((nwtiles + nt / 32 - 1) / (nt / 32), sm_count());
            const size_t sm = fast::smem_bytes<32>(nt);
            rc = launch_fast(nt == 512 ? fast_kernel<512, 32>(gather, ridged) : fast_kernel<768, 32>(gather, ridged), grid, nt, sm, stream);
        }
        if (rc) return rc;
    } else if (total > k2_small_max() && dim <= 8192) {
        // EXACT arithmetic, large batch: same roundings on the replicated tables (1 CTA per SM)
        int rc = prepare_fast();
        if (rc) return rc;
        if (peers.release) { rc = launch_gather_wait(peers.release, peers.release_min, peers.rank, peers.world, peers.error, stream); if (rc) return rc; }
        constexpr int NT = 1024;
        const uint64_t m1 = ((1ull << 40) + dim - 1) / dim;
        int grid = (int)std::min<int64_t>((total + NT - 1) / NT, sm_count());
        auto kern = peers.n > 0 ? fast::k_height_maps_exact_tab<NT, true> : fast::k_height_maps_exact_tab<NT, false>;
        kern<<<grid, NT, fast::smem_bytes<32>(0), stream>>>(d_quads, total, dim, cfg, d_out, m1, peers);
    } else {
        if (peers.release) { int rc = launch_gather_wait(peers.release, peers.release_min, peers.rank, peers.world, peers.error, stream); if (rc) return rc; }
        int64_t blocks = (total + 255) / 256;
        int grid = (int)std::min<int64_t>(blocks, (int64_t)sm_count() * 8);
        k_height_maps_exact<<<grid, 256, 0, stream>>>(d_quads, total, dim, cfg, d_out, peers);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "height map kernel launch");
}

// The two seam calls in EXACT arithmetic (what the reference's defaults select); `host_out` is
// pinned host memory.  Returns PLANET_E_UNSUPPORTED when the params ask for FAST: the caller then
// takes the batched path with a batch of one.
int launch_height_map_seam(const planet_gpu_params *p, const Quad *h_quad, int dim, int max_depth,
                           float *host_out, cudaStream_t stream)
{
    if (p->precision != PLANET_PRECISION_EXACT || dim > 2048) return PLANET_E_UNSUPPORTED;
    HeightCfg cfg = make_cfg(p, max_depth);
    const int grid = std::min((dim * dim * 8 + 255) / 256, sm_count() * 8);   // 8 lanes per sample
    k_height_map_seam_exact<<<grid, 256, 0, stream>>>(*h_quad, dim, cfg, host_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "height map (seam) launch");
}

int launch_height_at_seam(const planet_gpu_params *p, const double *h_xyz, int depth, int max_depth,
                          float *host_out, cudaStream_t stream)
{
    if (p->precision != PLANET_PRECISION_EXACT) return PLANET_E_UNSUPPORTED;
    HeightCfg cfg = make_cfg(p, max_depth);
    k_height_at_seam_exact<<<1, 32, 0, stream>>>(h_xyz[0], h_xyz[1], h_xyz[2], depth, cfg, host_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "height at (seam) launch");
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
        const double s0 = cfg.has_seed ? cfg.seed[0] : 0.0, s1 = cfg.has_seed ? cfg.seed[1] : 0.0,
                     s2 = cfg.has_seed ? cfg.seed[2] : 0.0;
        if (n <= k2_small_max())
            launch_points<1>((int)std::min<int64_t>(ntiles, (int64_t)sm_count() * 2), stream, d_xyz, n, cfg.kind, cfg.gain,
                             octaves, cfg.coord_scale, s0, s1, s2, cfg.height_scale, d_out, ntiles);
        else
            launch_points<32>((int)std::min<int64_t>(ntiles, sm_count()), stream, d_xyz, n, cfg.kind, cfg.gain,
                              octaves, cfg.coord_scale, s0, s1, s2, cfg.height_scale, d_out, ntiles);
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
        // a single PerlinNoise3 is a 1-octave fBm with amplitude 1
        const int k = octaves == 0 ? PLANET_NOISE_FBM : kind, o = octaves == 0 ? 1 : octaves;
        if (n <= k2_small_max())
            launch_points<1>((int)std::min<int64_t>(ntiles, (int64_t)sm_count() * 2), stream, d_xyz, n, k, gain, o,
                             1.0, 0.0, 0.0, 0.0, 1.0f, d_out, ntiles);
        else
            launch_points<32>((int)std::min<int64_t>(ntiles, sm_count()), stream, d_xyz, n, k, gain, o,
                              1.0, 0.0, 0.0, 0.0, 1.0f, d_out, ntiles);
    } else {
        int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
        k_noise_exact<<<grid, 256, 0, stream>>>(d_xyz, n, kind, lacunarity, gain, octaves, d_out);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "noise kernel launch");
}

} // namespace planet
