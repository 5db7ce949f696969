// nvlink_push.cu -- what egress rate can one B200 sustain into its peers, by store method?
//
// One process, device 0 sends, devices 1..N-1 receive (peer access enabled, no IPC needed).  Each
// variant moves TOTAL bytes, destinations interleaved chunk by chunk over the peers, and is timed
// with CUDA events on device 0:
//   ce        cudaMemcpyPeerAsync, one stream per peer
//   st16      one 16-byte st.global per thread per iteration (streaming), grid = SMs x ctas_per_sm
//   bulk<B>   B-byte cp.async.bulk shared -> global, NB in flight per CTA (the K2 gather uses B = 512)
// Output: one JSON line per variant.  Build: nvcc -arch=sm_100a -O3 -o nvlink_push nvlink_push.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../planet_b200/csrc/planet_tma.cuh"

#define OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Peers { char *ptr[7]; int n; };

__global__ void k_st16(Peers peers, size_t bytes_per_peer)
{
    const size_t nvec = bytes_per_peer / 16;
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec * peers.n; i += (size_t)gridDim.x * blockDim.x) {
        // consecutive 512-byte groups rotate over the peers (what the fused gather does per tile)
        const size_t group = i / 32, lane = i % 32;
        const int p = (int)(group % peers.n);
        const size_t g = group / peers.n;
        __stcs(reinterpret_cast<uint4 *>(peers.ptr[p]) + g * 32 + lane, v);
    }
}

template <int NB>
__global__ void k_bulk(Peers peers, size_t bytes_per_peer, int chunk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < NB * chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i;
    planet::tma::fence_smem_writes();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = bytes_per_peer / chunk * peers.n;
    int it = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, it++) {
        const int p = (int)(c % peers.n);
        const size_t g = c / peers.n;
        planet::tma::store_bulk(peers.ptr[p] + g * chunk, smem + (it % NB) * chunk, (uint32_t)chunk);
        planet::tma::commit();
        planet::tma::wait_read<NB - 1>();
    }
    planet::tma::wait_all<0>();
}

// the fused gather's pattern without the arithmetic: every warp of a 768-thread CTA owns a run of
// `chunk`-byte tiles; per tile lanes 0..n push it to the n peers (lane 0: to a local buffer, as K2
// does for its own copy), two staging buffers per warp, wait_read<1> before a buffer is reused
template <bool ONE_LANE, bool WAIT_LATE>
__global__ void __launch_bounds__(768, 1) k_warp8(Peers peers, char *local, size_t bytes_per_peer, int chunk, int spin)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    unsigned char *stage = smem + (size_t)warp * 2 * chunk;
    const size_t ntiles = bytes_per_peer / chunk;
    const size_t per_warp = (ntiles + (size_t)gridDim.x * warps - 1) / ((size_t)gridDim.x * warps);
    size_t t = ((size_t)blockIdx.x * warps + warp) * per_warp;
    const size_t t_end = t + per_warp < ntiles ? t + per_warp : ntiles;
    int buf = 0;
    float acc = lane;
    const bool issuer = ONE_LANE ? lane == 0 : lane <= peers.n;
    for (; t < t_end; t++) {
        for (int i = 0; i < spin; i++) acc = fmaf(acc, 1.0001f, 0.5f);                 // stand-in for the octave loops
        if (WAIT_LATE) {                                                               // the buffer about to be written was pushed two tiles ago
            if (issuer) planet::tma::wait_read<1>();
            __syncwarp();
        }
        for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<uint4 *>(stage + buf * chunk + i) = make_uint4(__float_as_uint(acc), 1u, 2u, 3u);
        planet::tma::fence_smem_writes();
        __syncwarp();
        if (issuer) {
            if (ONE_LANE) {
                planet::tma::store_bulk(local + t * chunk, stage + buf * chunk, (uint32_t)chunk);
                for (int d = 0; d < peers.n; d++) planet::tma::store_bulk(peers.ptr[d] + t * chunk, stage + buf * chunk, (uint32_t)chunk);
            } else {
                char *dst = (lane == 0 ? local : peers.ptr[lane - 1]) + t * chunk;
                planet::tma::store_bulk(dst, stage + buf * chunk, (uint32_t)chunk);
            }
            planet::tma::commit();
            if (!WAIT_LATE) planet::tma::wait_read<1>();
        }
        buf ^= 1;
        if (!WAIT_LATE) __syncwarp();
    }
    if (issuer) planet::tma::wait_all<0>();
}

// E2: the same pattern with plain 16-byte global stores from registers (LSU path, no staging, no TMA)
__global__ void __launch_bounds__(768, 1) k_warp8_lsu(Peers peers, char *local, size_t bytes_per_peer, int spin)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const size_t ntiles = bytes_per_peer / 512;
    const size_t per_warp = (ntiles + (size_t)gridDim.x * warps - 1) / ((size_t)gridDim.x * warps);
    size_t t = ((size_t)blockIdx.x * warps + warp) * per_warp;
    const size_t t_end = t + per_warp < ntiles ? t + per_warp : ntiles;
    float acc = lane;
    for (; t < t_end; t++) {
        for (int i = 0; i < spin; i++) acc = fmaf(acc, 1.0001f, 0.5f);
        const uint4 v = make_uint4(__float_as_uint(acc), 1u, 2u, 3u);
        __stcs(reinterpret_cast<uint4 *>(local + t * 512) + lane, v);
#pragma unroll
        for (int d = 0; d < 7; d++)
            if (d < peers.n) __stcs(reinterpret_cast<uint4 *>(peers.ptr[d] + t * 512) + lane, v);
    }
}

// E4: NBUF staging buffers per warp (tiles in flight per warp), lanes 0..n issue, wait before the buffer is rewritten
template <int NBUF>
__global__ void __launch_bounds__(768, 1) k_warp8_deep(Peers peers, char *local, size_t bytes_per_peer, int chunk, int spin)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    unsigned char *stage = smem + (size_t)warp * NBUF * chunk;
    const size_t ntiles = bytes_per_peer / chunk;
    const size_t per_warp = (ntiles + (size_t)gridDim.x * warps - 1) / ((size_t)gridDim.x * warps);
    size_t t = ((size_t)blockIdx.x * warps + warp) * per_warp;
    const size_t t_end = t + per_warp < ntiles ? t + per_warp : ntiles;
    int buf = 0;
    float acc = lane;
    for (; t < t_end; t++) {
        for (int i = 0; i < spin; i++) acc = fmaf(acc, 1.0001f, 0.5f);
        if (lane <= peers.n) planet::tma::wait_read<NBUF - 1>();
        __syncwarp();
        for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<uint4 *>(stage + buf * chunk + i) = make_uint4(__float_as_uint(acc), 1u, 2u, 3u);
        planet::tma::fence_smem_writes();
        __syncwarp();
        if (lane <= peers.n) {
            char *dst = (lane == 0 ? local : peers.ptr[lane - 1]) + t * chunk;
            planet::tma::store_bulk(dst, stage + buf * chunk, (uint32_t)chunk);
            planet::tma::commit();
        }
        buf = buf + 1 == NBUF ? 0 : buf + 1;
    }
    if (lane <= peers.n) planet::tma::wait_all<0>();
}

// E1: warp specialisation.  24 compute warps stage their tiles (two buffers each) and raise a flag; a 25th warp
// does nothing but hand staged tiles to the TMA unit, round-robin over the compute warps, and frees the buffers
// one round later.  If a blocked bulk-copy issue only stalls the issuing warp, compute and push now overlap.
__global__ void __launch_bounds__(800, 1) k_warp8_spec(Peers peers, char *local, size_t bytes_per_peer, int chunk, int spin)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ volatile uint32_t ready[24], freed[24];                // tiles staged / tiles whose buffer is free again, per compute warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int CW = 24;
    if (threadIdx.x < CW) { ready[threadIdx.x] = 0; freed[threadIdx.x] = 0; }
    __syncthreads();
    const size_t ntiles = bytes_per_peer / chunk;
    const size_t per_warp = (ntiles + (size_t)gridDim.x * CW - 1) / ((size_t)gridDim.x * CW);
    if (warp < CW) {
        unsigned char *stage = smem + (size_t)warp * 2 * chunk;
        size_t t = ((size_t)blockIdx.x * CW + warp) * per_warp;
        const size_t t_end = t + per_warp < ntiles ? t + per_warp : ntiles;
        float acc = lane;
        uint32_t k = 0;
        for (; t < t_end; t++, k++) {
            for (int i = 0; i < spin; i++) acc = fmaf(acc, 1.0001f, 0.5f);
            while (k >= 2 && freed[warp] < k - 1) { }                 // buffer k & 1 was used by tile k - 2
            for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<uint4 *>(stage + (k & 1) * chunk + i) = make_uint4(__float_as_uint(acc), 1u, 2u, 3u);
            planet::tma::fence_smem_writes();
            __syncwarp();
            if (lane == 0) { __threadfence_block(); ready[warp] = k + 1; }
        }
        return;
    }
    // the pusher warp: rounds over the compute warps
    for (uint32_t k = 0; k < per_warp; k++) {
        for (int w = 0; w < CW; w++) {
            const size_t t = ((size_t)blockIdx.x * CW + w) * per_warp + k;
            if (t >= ntiles || t >= ((size_t)blockIdx.x * CW + w + 1) * per_warp) continue;
            while (ready[w] < k + 1) { }
            __threadfence_block();
            if (lane <= peers.n) {
                char *dst = (lane == 0 ? local : peers.ptr[lane - 1]) + t * chunk;
                planet::tma::store_bulk(dst, smem + (size_t)w * 2 * chunk + (k & 1) * chunk, (uint32_t)chunk);
            }
        }
        if (lane <= peers.n) { planet::tma::commit(); planet::tma::wait_read<1>(); }    // round k - 1 has been read out
        __syncwarp();
        if (k >= 1 && lane < CW) freed[lane] = k;                                       // tiles 0 .. k-1 of every warp are free
    }
    if (lane <= peers.n) planet::tma::wait_all<0>();
}

int main(int argc, char **argv)
{
    int ndev = 0;
    OK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { fprintf(stderr, "needs >= 2 GPUs\n"); return 2; }
    const int npeers = argc > 1 ? atoi(argv[1]) : ndev - 1;
    const size_t total = (size_t)(argc > 2 ? atof(argv[2]) : 448.0) * (1 << 20);
    const size_t per_peer = total / npeers / 32768 * 32768;
    Peers peers = {};
    peers.n = npeers;
    const int nrecv = ndev - 1 < npeers ? ndev - 1 : npeers;             // peers beyond the device count share receivers
    for (int p = 0; p < npeers; p++) {
        OK(cudaSetDevice(1 + p % nrecv));
        OK(cudaMalloc((void **)&peers.ptr[p], per_peer));
        OK(cudaMemset(peers.ptr[p], 0, per_peer));
        OK(cudaDeviceSynchronize());
    }
    OK(cudaSetDevice(0));
    for (int p = 0; p < nrecv; p++) OK(cudaDeviceEnablePeerAccess(p + 1, 0));
    char *src = nullptr;
    OK(cudaMalloc((void **)&src, per_peer));
    OK(cudaMemset(src, 1, per_peer));
    int sms = 148;
    OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    OK(cudaEventCreate(&e0)); OK(cudaEventCreate(&e1));
    auto report = [&](const char *name, int a, int b, float ms) {
        printf("{\"variant\": \"%s\", \"param\": %d, \"ctas\": %d, \"peers\": %d, \"MB\": %.1f, \"ms\": %.4f, \"egress_GBs\": %.1f}\n",
               name, a, b, npeers, per_peer * npeers / 1e6, ms, per_peer * npeers / 1e6 / ms);
        fflush(stdout);
    };
    // ---- copy engines ----
    {
        std::vector<cudaStream_t> st(npeers);
        for (auto &s : st) OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        for (int rep = 0; rep < 3; rep++) {
            OK(cudaDeviceSynchronize());
            OK(cudaEventRecord(e0, st[0]));
            std::vector<cudaEvent_t> done(npeers);
            for (int p = 0; p < npeers; p++) {
                OK(cudaMemcpyPeerAsync(peers.ptr[p], 1 + p % nrecv, src, 0, per_peer, st[p]));
                OK(cudaEventCreateWithFlags(&done[p], cudaEventDisableTiming));
                OK(cudaEventRecord(done[p], st[p]));
                OK(cudaStreamWaitEvent(st[0], done[p], 0));
            }
            OK(cudaEventRecord(e1, st[0]));
            OK(cudaDeviceSynchronize());
            float ms = 0; OK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 2) report("ce", 0, npeers, ms);
            for (auto &d : done) cudaEventDestroy(d);
        }
    }
    auto timeit = [&](auto launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            OK(cudaDeviceSynchronize());
            OK(cudaEventRecord(e0));
            launch();
            OK(cudaEventRecord(e1));
            OK(cudaDeviceSynchronize());
            OK(cudaGetLastError());
            float ms = 0; OK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        return best;
    };
    const bool quick = argc > 3;
    for (int per_sm : { 1, 2, 4, 8 })
        if (!quick) report("st16", 256, sms * per_sm, timeit([&] { k_st16<<<sms * per_sm, 256>>>(peers, per_peer); }));
    OK(cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OK(cudaFuncSetAttribute(k_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    for (int chunk : { 512, 2048, 8192, 32768 })
        for (int per_sm : { 1, 4 }) {
            if (quick || chunk * 4 * per_sm > 200 * 1024) continue;
            report("bulk_nb2", chunk, sms * per_sm, timeit([&] { k_bulk<2><<<sms * per_sm, 32, 2 * chunk>>>(peers, per_peer, chunk); }));
            report("bulk_nb4", chunk, sms * per_sm, timeit([&] { k_bulk<4><<<sms * per_sm, 32, 4 * chunk>>>(peers, per_peer, chunk); }));
            if (chunk * 8 * per_sm <= 200 * 1024)
                report("bulk_nb8", chunk, sms * per_sm, timeit([&] { k_bulk<8><<<sms * per_sm, 32, 8 * chunk>>>(peers, per_peer, chunk); }));
        }
    auto warp8 = [&](auto kern, const char *name) {
        OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int chunk : { 512, 2048 })
            for (int spin : { 0, 2000, 4000, 8000 })
                report(name, chunk, spin, timeit([&] { kern<<<sms, 768, 24 * 2 * chunk>>>(peers, src, per_peer, chunk, spin); }));
    };
    warp8(k_warp8<false, false>, "warp8_8lanes_wait_after_commit");
    for (int spin : { 0, 2000, 4000, 8000 })
        report("warp8_compute_only_no_push", 512, spin, timeit([&] { Peers none = {}; none.n = -1; k_warp8_deep<2><<<sms, 768, 24 * 2 * 512>>>(none, src, per_peer, 512, spin); }));
    for (int spin : { 0, 2000, 4000, 8000 })
        report("warp8_lsu_st16", 512, spin, timeit([&] { k_warp8_lsu<<<sms, 768>>>(peers, src, per_peer, spin); }));
    auto deep = [&](auto kern, const char *name, int nbuf) {
        OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int spin : { 0, 2000, 4000, 8000 })
            report(name, 512, spin, timeit([&] { kern<<<sms, 768, 24 * nbuf * 512>>>(peers, src, per_peer, 512, spin); }));
    };
    deep(k_warp8_deep<2>, "warp8_deep_nbuf2", 2);
    deep(k_warp8_deep<4>, "warp8_deep_nbuf4", 4);
    deep(k_warp8_deep<8>, "warp8_deep_nbuf8", 8);
    OK(cudaFuncSetAttribute(k_warp8_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int chunk : { 512, 2048 })
        for (int spin : { 0, 2000, 4000, 8000 })
            report("warp8_pusher_warp", chunk, spin, timeit([&] { k_warp8_spec<<<sms, 800, 24 * 2 * chunk>>>(peers, src, per_peer, chunk, spin); }));
    for (int ctas : { 8, 16, 32, 64 })
        report("bulk_nb4_fewctas", 8192, ctas, timeit([&] { k_bulk<4><<<ctas, 32, 4 * 8192>>>(peers, per_peer, 8192); }));
    return 0;
}
