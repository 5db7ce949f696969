// tools/microbench5.cu -- marginal cost of each part of the K2 octave loop.  The stripped loop of
// microbench3.cu (float level-3 table, one sample per thread, 768 threads per SM) with one part
// removed at a time; the drop in SMSP cycles per warp-octave is what that part really costs,
// whatever a count of issue slots says.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench5 tools/microbench5.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int ROWS = 512, ITERS = 2048;
enum { BASE = 0, NO_FADE = 1, NO_LERP = 2, NO_LDS3 = 4, NO_LDS12 = 8, NO_GZ_TERM = 16, NO_LCG = 32, NO_X1 = 64, NO_DOTS = 128 };

__device__ __forceinline__ uint32_t perm(int i) { return (uint32_t)((i * 167 + 13) & 255); }
__device__ __forceinline__ uint32_t code_of(uint32_t h)
{
    const int g = h & 15;
    const int zero = g % 3;
    const uint32_t s0 = (g & 4) ? 0xC0u : 0x40u, s1 = (g & 8) ? 0xC0u : 0x40u;
    uint32_t c[3]; int k = 0;
    for (int a = 0; a < 3; a++) c[a] = (a == zero) ? 0u : (k++ ? s1 : s0);
    return (c[0] << 24) | (c[1] << 8) | c[2];
}

template <int V, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_octaves(long long *cyc, float *sink, uint32_t seed)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *t12 = reinterpret_cast<uint32_t *>(smem);
    unsigned char *t3 = smem + ROWS * 128;
    for (int w = threadIdx.x; w < ROWS * 32; w += THREADS) {
        const int i = w >> 5;
        const uint32_t p = perm(i);
        t12[w] = (p << 7) | (p << (16 + 7));
        if ((w & 31) < 8) {
            const uint32_t c0 = code_of(perm(i)), c1 = code_of(perm(i + 1));
            uint4 e;
            e.x = (c0 & 0xFF000000u) | ((c0 & 0xC0u) >> 6);
            e.y = (c0 & 0x0000FF00u) << 16;
            e.z = (c1 & 0xFF000000u) | ((c1 & 0xC0u) >> 6);
            e.w = (c1 & 0x0000FF00u) << 16;
            reinterpret_cast<uint4 *>(t3)[i * 8 + (w & 31)] = e;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned char *t12_lane = smem + lane * 4;
    const unsigned char *t3_lane = t3 + (lane & 7) * 16;
    uint32_t sx = seed + threadIdx.x * 2654435761u, sy = sx * 747796405u + 1u, sz = sy * 2891336453u + 7u;
    float acc = 0.f, amp = 0.5f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (V & NO_LCG) { sx += 0x01234567u; sy += 0x00ABCDEFu; sz += 0x00654321u; }
        else { sx = sx * 1664525u + 1013904223u; sy = sy * 22695477u + 1u; sz = sz * 1103515245u + 12345u; }
        const float mx = __uint_as_float((sx & 0x007FFFFFu) | 0x3F800000u);
        const float my = __uint_as_float((sy & 0x007FFFFFu) | 0x3F800000u);
        const float mz = __uint_as_float((sz & 0x007FFFFFu) | 0x3F800000u);
        const float x0 = mx - 1.f, y0 = my - 1.f, z0 = mz - 1.f;
        const float x1 = (V & NO_X1) ? mx : mx - 2.f, y1 = (V & NO_X1) ? my : my - 2.f, z1 = (V & NO_X1) ? mz : mz - 2.f;
        const uint32_t cx = (sx >> 16) & 0x7F80u, cy = (sy >> 16) & 0x7F80u, cz = (sz >> 16) & 0x7F80u;
        auto u16 = [&](uint32_t off) -> uint32_t {
            if (V & NO_LDS12) return (off * 0x9E37u) & 0x7F80u;
            return *reinterpret_cast<const unsigned short *>(t12_lane + off);
        };
        const uint32_t a0 = u16(cx), a1 = u16(cx + 128);
        const uint32_t b00 = u16(a0 + cy + 2), b01 = u16(a0 + cy + 130), b10 = u16(a1 + cy + 2), b11 = u16(a1 + cy + 130);
        float g[8];
        auto corner2 = [&](uint32_t off, float X, float Y, float &d0, float &d1) {
            uint4 e;
            if (V & NO_LDS3) { e.x = off << 17; e.y = off << 19; e.z = off << 18; e.w = off << 16; }
            else e = *reinterpret_cast<const uint4 *>(t3_lane + off);
            if (V & NO_DOTS) {
                d0 = __uint_as_float(e.x ^ e.y) + X; d1 = __uint_as_float(e.z ^ e.w) + Y;
            } else if (V & NO_GZ_TERM) {
                d0 = fmaf(__uint_as_float(e.y), Y, __uint_as_float(e.x) * X);
                d1 = fmaf(__uint_as_float(e.w), Y, __uint_as_float(e.z) * X);
            } else {
                d0 = fmaf(__uint_as_float(e.x << 30), z0, fmaf(__uint_as_float(e.y), Y, __uint_as_float(e.x) * X));
                d1 = fmaf(__uint_as_float(e.z << 30), z1, fmaf(__uint_as_float(e.w), Y, __uint_as_float(e.z) * X));
            }
        };
        corner2(b00 + cz, x0, y0, g[0], g[4]); corner2(b10 + cz, x1, y0, g[1], g[5]);
        corner2(b01 + cz, x0, y1, g[2], g[6]); corner2(b11 + cz, x1, y1, g[3], g[7]);
        auto fade = [](float t) { return (V & NO_FADE) ? t : t * t * t * fmaf(fmaf(t, 6.f, -15.f), t, 10.f); };
        auto lerp = [](float a, float b, float t) { return (V & NO_LERP) ? a + b : fmaf(b - a, t, a); };
        const float u = fade(x0), v = fade(y0), w = fade(z0);
        float n = lerp(lerp(lerp(g[0], g[1], u), lerp(g[2], g[3], u), v),
                       lerp(lerp(g[4], g[5], u), lerp(g[6], g[7], u), v), w);
        if (V & NO_LERP) n = fmaf(n, u, v * w);
        acc = fmaf(n, amp, acc);
        amp = amp * 0.999f;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 1234.5f) sink[0] = acc;
}

template <int V, int THREADS> static void run(const char *name, int sms, long long *d_cyc, float *d_sink)
{
    const size_t smem = (size_t)ROWS * 128 + ROWS * 128;
    cudaFuncSetAttribute(k_octaves<V, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_octaves<V, THREADS><<<sms, THREADS, smem>>>(d_cyc, d_sink, 1);
    k_octaves<V, THREADS><<<sms, THREADS, smem>>>(d_cyc, d_sink, 2);
    cudaError_t err = cudaDeviceSynchronize();
    std::vector<long long> cyc(sms);
    cudaMemcpy(cyc.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[sms / 2];
    printf("{\"variant\": \"%s\", \"threads\": %d, \"smsp_cycles_per_warp_octave\": %.1f, \"err\": \"%s\"}\n", name, THREADS,
           med / ITERS / (THREADS / 128.0), cudaGetErrorString(err));
}

#define RUN(V) run<V, 768>(#V, sms, d_cyc, d_sink)
int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *d_cyc; float *d_sink;
    cudaMalloc(&d_cyc, sms * sizeof(long long)); cudaMalloc(&d_sink, 4);
    RUN(BASE); RUN(NO_LCG); RUN(NO_LCG | NO_FADE); RUN(NO_LCG | NO_LERP); RUN(NO_LCG | NO_LDS3); RUN(NO_LCG | NO_LDS12);
    RUN(NO_LCG | NO_LDS3 | NO_LDS12); RUN(NO_LCG | NO_GZ_TERM); RUN(NO_LCG | NO_X1); RUN(NO_LCG | NO_DOTS);
    RUN(NO_LCG | NO_DOTS | NO_FADE | NO_LERP); RUN(NO_LCG | NO_FADE | NO_LERP);
    run<NO_LCG, 256>("NO_LCG", sms, d_cyc, d_sink);
    run<NO_LCG, 512>("NO_LCG", sms, d_cyc, d_sink);
    run<NO_LCG, 1024>("NO_LCG", sms, d_cyc, d_sink);
    return 0;
}
