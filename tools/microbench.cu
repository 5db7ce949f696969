// tools/microbench.cu -- per-instruction throughput on one B200 (thread-ops / clk / SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// One CTA of 1024 threads per SM; every thread runs ITERS x 16 independent ops of one kind;
// cycles from clock64() inside the kernel (SM clock domain, independent of DVFS).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 2048

#define BENCH_KERNEL(NAME, DECL, BODY, SINK)                                              \
__global__ void __launch_bounds__(1024, 1) k_##NAME(long long *cyc, unsigned *sink, unsigned seed) { \
    DECL;                                                                                  \
    __syncthreads();                                                                       \
    long long t0 = clock64();                                                              \
    _Pragma("unroll 8") for (int it = 0; it < ITERS; it++) { BODY; }                                           \
    long long t1 = clock64();                                                              \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                       \
    SINK;                                                                                  \
}

// 16 independent registers r0..r15
#define R16(T, init) T r0=init+0, r1=init+1, r2=init+2, r3=init+3, r4=init+4, r5=init+5, r6=init+6, r7=init+7, \
                       r8=init+8, r9=init+9, r10=init+10, r11=init+11, r12=init+12, r13=init+13, r14=init+14, r15=init+15
#define FOR16(OP) OP(r0) OP(r1) OP(r2) OP(r3) OP(r4) OP(r5) OP(r6) OP(r7) OP(r8) OP(r9) OP(r10) OP(r11) OP(r12) OP(r13) OP(r14) OP(r15)
#define SUMF (r0+r1+r2+r3+r4+r5+r6+r7+r8+r9+r10+r11+r12+r13+r14+r15)
#define SINKF if (SUMF == (decltype(r0))123456) sink[0] = 1
#define SINKU if ((r0^r1^r2^r3^r4^r5^r6^r7^r8^r9^r10^r11^r12^r13^r14^r15) == 0x12345u) sink[0] = 1

#define OP_FFMA(r) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r) : "f"(a), "f"(b));
BENCH_KERNEL(ffma, float a = 1.0001f + seed; float b = 0.5f; R16(float, (float)threadIdx.x), FOR16(OP_FFMA), SINKF)
#define OP_FADD(r) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r) : "f"(a));
BENCH_KERNEL(fadd, float a = 1.0001f + seed; R16(float, (float)threadIdx.x), FOR16(OP_FADD), SINKF)
#define OP_FMUL(r) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r) : "f"(a));
BENCH_KERNEL(fmul, float a = 1.0001f + seed; R16(float, (float)threadIdx.x), FOR16(OP_FMUL), SINKF)

#define OP_FFMA2(r) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r) : "l"(a), "l"(b));
BENCH_KERNEL(ffma2, unsigned long long a = 0x3f8003473f800347ull + seed; unsigned long long b = 0x3f0000003f000000ull; R16(unsigned long long, 0x3f8000003f800000ull + threadIdx.x), FOR16(OP_FFMA2), SINKU)
#define OP_FADD2(r) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r) : "l"(a));
BENCH_KERNEL(fadd2, unsigned long long a = 0x3f8003473f800347ull + seed; R16(unsigned long long, 0x3f8000003f800000ull + threadIdx.x), FOR16(OP_FADD2), SINKU)

#define OP_LOP3(r) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(a), "r"(b));
BENCH_KERNEL(lop3, unsigned a = 0x9e3779b9u + seed; unsigned b = threadIdx.x * 77u; R16(unsigned, threadIdx.x * 3u), FOR16(OP_LOP3), SINKU)
#define OP_IADD(r) asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(a));
BENCH_KERNEL(iadd, unsigned a = 0x9e3779b9u + seed; R16(unsigned, threadIdx.x * 3u), FOR16(OP_IADD), SINKU)
#define OP_SHF(r) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
BENCH_KERNEL(shf, unsigned a = 0x9e3779b9u + seed; unsigned b = 5 + (seed & 3); R16(unsigned, threadIdx.x * 3u), FOR16(OP_SHF), SINKU)
#define OP_PRMT(r) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
BENCH_KERNEL(prmt, unsigned a = 0x9e3779b9u + seed; unsigned b = 0x1230 + (seed & 3); R16(unsigned, threadIdx.x * 3u), FOR16(OP_PRMT), SINKU)
#define OP_IMAD(r) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
BENCH_KERNEL(imad, unsigned a = 0x9e3779b9u + seed; unsigned b = 12345u; R16(unsigned, threadIdx.x * 3u), FOR16(OP_IMAD), SINKU)
#define OP_SHL(r) asm volatile("shl.b32 %0, %0, 3;" : "+r"(r));
BENCH_KERNEL(shl_imm, R16(unsigned, threadIdx.x * 3u + seed), FOR16(OP_SHL), SINKU)
#define OP_FSEL(r) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p; }" : "+f"(r) : "f"(a), "f"(b));
BENCH_KERNEL(setp_selp, float a = 3.0f + seed; float b = 0.5f; R16(float, (float)threadIdx.x), FOR16(OP_FSEL), SINKF)
#define OP_H2F(r) asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, hi; }" : "+r"(r));
BENCH_KERNEL(hadd2_f32, R16(unsigned, 0x3c003c00u + threadIdx.x + seed), FOR16(OP_H2F), SINKU)
#define OP_I2F(r) asm volatile("{ .reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t; }" : "+r"(r));
BENCH_KERNEL(i2f, R16(unsigned, threadIdx.x * 3u + seed), FOR16(OP_I2F), SINKU)
#define OP_DFMA(r) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r) : "d"(a), "d"(b));
BENCH_KERNEL(dfma, double a = 1.0001 + seed; double b = 0.5; R16(double, (double)threadIdx.x), FOR16(OP_DFMA), SINKF)
#define OP_DADD(r) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(r) : "d"(a));
BENCH_KERNEL(dadd, double a = 1.0001 + seed; R16(double, (double)threadIdx.x), FOR16(OP_DADD), SINKF)
#define OP_D2LL(r) asm volatile("{ .reg .s64 t; cvt.rmi.s64.f64 t, %0; cvt.rn.f64.s64 %0, t; }" : "+d"(r));
BENCH_KERNEL(d2ll_ll2d, R16(double, (double)threadIdx.x * 1.37 + seed), FOR16(OP_D2LL), SINKF)

// shared-memory loads: address pattern `mode`: 0 = conflict-free lane slot, random row (like K2's tables);
// the loaded value feeds the next address so loads stay dependent per chain but 8 chains are independent.
template <int BYTES, int CONFLICT>
__global__ void __launch_bounds__(1024, 1) k_lds(long long *cyc, unsigned *sink, unsigned seed)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int ROWS = 256, ROWB = 32 * BYTES;
    for (int i = threadIdx.x; i < ROWS * 32 * BYTES / 4; i += blockDim.x) {
        unsigned row = (i * 4 / ROWB);
        reinterpret_cast<unsigned *>(smem)[i] = ((row * 167u + 13u) & 255u) * ROWB;   // next row offset
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned lane_off = CONFLICT ? (lane & ~(CONFLICT - 1)) * BYTES : lane * BYTES;  // CONFLICT lanes share a bank set
    unsigned o[8];
    for (int j = 0; j < 8; j++) o[j] = ((threadIdx.x * 31u + j * 57u + seed) & 255u) * ROWB;
    if (CONFLICT) for (int j = 0; j < 8; j++) o[j] = (((threadIdx.x + j) * 131u + seed) & 255u) * ROWB;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned char *p = smem + o[j] + lane_off;
            if (BYTES == 2) o[j] = (*reinterpret_cast<const unsigned short *>(p)) & 0xFFFFu;
            else if (BYTES == 4) o[j] = *reinterpret_cast<const unsigned *>(p);
            else if (BYTES == 8) { uint2 v = *reinterpret_cast<const uint2 *>(p); o[j] = v.x ^ (v.y & 0); }
            else { uint4 v = *reinterpret_cast<const uint4 *>(p); o[j] = v.x ^ ((v.y | v.z | v.w) & 0); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    unsigned s = 0;
    for (int j = 0; j < 8; j++) s ^= o[j];
    if (s == 0x12345u) sink[0] = 1;
}

template <class K>
static void run(const char *name, K kern, int ops_per_iter, size_t smem, int sms, long long *d_cyc, unsigned *d_sink)
{
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms, 1024, smem>>>(d_cyc, d_sink, 0);                   // warm-up
    cudaEventRecord(e0);
    kern<<<sms, 1024, smem>>>(d_cyc, d_sink, 1);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(sms);
    cudaMemcpy(cyc.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(cyc.begin(), cyc.end());
    double med = (double)cyc[sms / 2];
    double ops = 1024.0 * ITERS * ops_per_iter;
    printf("{\"op\": \"%s\", \"thread_ops_per_clk_per_sm\": %.2f, \"warp_instr_per_clk_per_sm\": %.3f, "
           "\"median_cycles\": %.0f, \"ms\": %.3f, \"implied_mhz\": %.0f, \"err\": \"%s\"}\n",
           name, ops / med, ops / med / 32.0, med, ms, med / (ms * 1e3), cudaGetErrorString(err));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *d_cyc; unsigned *d_sink;
    cudaMalloc(&d_cyc, 1024 * sizeof(long long)); cudaMalloc(&d_sink, 4);
    printf("{\"sms\": %d}\n", sms);
#define RUN(NAME) run(#NAME, k_##NAME, 16, 0, sms, d_cyc, d_sink)
    RUN(ffma); RUN(fadd); RUN(fmul); RUN(ffma2); RUN(fadd2); RUN(lop3); RUN(iadd); RUN(shf); RUN(prmt);
    RUN(imad); RUN(shl_imm); RUN(setp_selp); RUN(hadd2_f32); RUN(i2f); RUN(dfma); RUN(dadd); RUN(d2ll_ll2d);
    run("lds_u16_conflict_free", k_lds<2, 0>, 8, 256 * 32 * 2, sms, d_cyc, d_sink);
    run("lds_32_conflict_free", k_lds<4, 0>, 8, 256 * 32 * 4, sms, d_cyc, d_sink);
    run("lds_64_conflict_free", k_lds<8, 0>, 8, 256 * 32 * 8, sms, d_cyc, d_sink);
    run("lds_128_conflict_free", k_lds<16, 0>, 8, 256 * 32 * 16, sms, d_cyc, d_sink);
    run("lds_32_4way_conflict", k_lds<4, 4>, 8, 256 * 32 * 4, sms, d_cyc, d_sink);
    return 0;
}
