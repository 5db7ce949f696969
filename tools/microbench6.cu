// tools/microbench6.cu -- K2's real octave loop (k2_heights.cu's fractal<>) timed in isolation:
// N samples per thread, T threads per SM, random fixed-point positions, 8 fBm octaves, no
// position prologue and no stores.  Answers: what does the loop alone cost per warp-octave, and
// how much of K2's time is outside it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I include -o tools/microbench6 tools/microbench6.cu
#include "../planet_b200/csrc/k2_heights.cu"
#include <cstdio>
#include <vector>
#include <algorithm>

namespace planet { int set_error(int c, const char *, ...) { return c; } int check_cuda(cudaError_t e, const char *) { return e != cudaSuccess; }
void count_launch(int) {}
HeightCfg make_cfg(const planet_gpu_params *, int) { return HeightCfg(); } }

using namespace planet::fast;
constexpr int ITERS = 256, OCT = 8;

template <int N, int T>
__global__ void __launch_bounds__(T, 1) k_loop(long long *cyc, float *sink, uint32_t seed, int octaves, float gain, uint32_t one_bits, int kind)
{
    extern __shared__ __align__(16) unsigned char smem[];
    build_tables<32>(smem);
    __syncthreads();
    const LaneTab tab = lane_tab<32>(smem, threadIdx.x & 31);
    uint32_t s = seed + threadIdx.x * 2654435761u;
    float acc = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        Fixed3 p[N]; int oct[N]; float value[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            s = s * 1664525u + 1013904223u; p[i].xhi = s >> 1; p[i].xlo = s * 747796405u;
            s = s * 1664525u + 1013904223u; p[i].yhi = s >> 1; p[i].ylo = s * 747796405u;
            s = s * 1664525u + 1013904223u; p[i].zhi = s >> 1; p[i].zlo = s * 747796405u;
            oct[i] = octaves;
        }
        fractal<32, N>(tab, p, oct, kind, gain, one_bits, value);
#pragma unroll
        for (int i = 0; i < N; i++) acc += value[i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 1234.5f) sink[0] = acc;
}

template <int N, int T> static void run(int sms, long long *d_cyc, float *d_sink)
{
    const size_t smem = Layout<32>::TABLES + 4096;
    cudaFuncSetAttribute(k_loop<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_loop<N, T>);
    k_loop<N, T><<<sms, T, smem>>>(d_cyc, d_sink, 1, OCT, 0.5f, 0x3F800000u, PLANET_NOISE_FBM);
    k_loop<N, T><<<sms, T, smem>>>(d_cyc, d_sink, 2, OCT, 0.5f, 0x3F800000u, PLANET_NOISE_FBM);
    cudaError_t err = cudaDeviceSynchronize();
    std::vector<long long> cyc(sms);
    cudaMemcpy(cyc.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[sms / 2];
    printf("{\"samples_per_thread\": %d, \"threads\": %d, \"regs\": %d, \"smsp_cycles_per_warp_octave_sample\": %.1f, \"err\": \"%s\"}\n",
           N, T, fa.numRegs, med / ITERS / OCT / N / (T / 128.0), cudaGetErrorString(err));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *d_cyc; float *d_sink;
    cudaMalloc(&d_cyc, sms * sizeof(long long)); cudaMalloc(&d_sink, 4);
    run<1, 512>(sms, d_cyc, d_sink); run<1, 768>(sms, d_cyc, d_sink); run<1, 1024>(sms, d_cyc, d_sink);
    run<2, 512>(sms, d_cyc, d_sink); run<2, 768>(sms, d_cyc, d_sink);
    run<3, 512>(sms, d_cyc, d_sink);
    return 0;
}
