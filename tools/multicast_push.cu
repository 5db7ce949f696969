// multicast_push.cu -- does an NVSwitch multicast object raise the all-gather rate over unicast peer stores?
//
// One process drives all N GPUs of the box.  Every GPU owns a shard of SHARD bytes and must end up
// with all N shards in its own buffer (the K4 exchange: every rank's height maps on every rank).
// All GPUs push at the same time (the kernels spin on a host flag and start together); per GPU the
// time is taken on the device with %globaltimer from the common start to the GPU's last store, and
// the figure reported is bytes RECEIVED per GPU / the slowest GPU's time.  Variants:
//   uni_st16    every 16-byte vector stored to the 7 peers' buffers (st.global.v4) + the local one
//   uni_bulk    512-byte cp.async.bulk shared->global per destination (what the fused K2 gather does)
//   mc_st16     ONE multimem.st.global.v4.f32 per vector to the multicast address: the switch
//               replicates it into all N buffers (egress = 1 shard instead of 7)
//   mc_bulk     512-byte cp.async.bulk to the multicast address (the PTX ISA only promises multimem.*
//               on such addresses; tried because it is what the kernel would want -- checked, not trusted)
// After every variant each GPU's buffer is checked against the expected pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o multicast_push multicast_push.cu   (no -lcuda: the
// driver entry points are fetched through cudaGetDriverEntryPoint)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../planet_b200/csrc/planet_tma.cuh"

#define OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(1); } } while (0)
#define DRV(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { fprintf(stderr, "%s: CUresult %d\n", #call, (int)r_); exit(1); } } while (0)

template <class F> static F drv(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { fprintf(stderr, "no driver entry point %s\n", name); exit(1); }
    return (F)fn;
}

struct Dest { char *ptr[8]; int n; };       // destinations of one GPU's shard (already offset to the shard)

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void start_together(volatile int *go)
{
    if (threadIdx.x == 0) while (*go == 0) { }
    __syncthreads();
}
__device__ __forceinline__ void finish(unsigned long long *times, unsigned int *done)
{
    __threadfence_system();                                               // this thread's stores are performed everywhere
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done, 1u) == gridDim.x - 1) times[1] = gtime();
    }
}
__device__ __forceinline__ uint4 pattern(int rank, size_t vec) { return make_uint4((unsigned)rank, (unsigned)vec, (unsigned)(vec >> 32) ^ 0x5a5au, 0x12345678u); }

__global__ void __launch_bounds__(512) k_uni_st16(Dest d, size_t shard, int rank, volatile int *go, unsigned long long *times, unsigned int *done)
{
    start_together(go);
    if (blockIdx.x == 0 && threadIdx.x == 0) times[0] = gtime();
    const size_t nvec = shard / 16;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = pattern(rank, v);
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < d.n) __stcs(reinterpret_cast<uint4 *>(d.ptr[k]) + v, x);
    }
    finish(times, done);
}

__global__ void __launch_bounds__(512) k_mc_st16(char *mc, size_t shard, int rank, volatile int *go, unsigned long long *times, unsigned int *done)
{
    start_together(go);
    if (blockIdx.x == 0 && threadIdx.x == 0) times[0] = gtime();
    const size_t nvec = shard / 16;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = pattern(rank, v);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     :: "l"(mc + v * 16), "f"(__uint_as_float(x.x)), "f"(__uint_as_float(x.y)), "f"(__uint_as_float(x.z)), "f"(__uint_as_float(x.w)) : "memory");
    }
    finish(times, done);
}

// every warp stages 512-byte tiles in shared memory (two buffers) and lanes 0..n-1 push them, one bulk copy per destination
__global__ void __launch_bounds__(768, 1) k_bulk(Dest d, size_t shard, int rank, volatile int *go, unsigned long long *times, unsigned int *done)
{
    extern __shared__ __align__(128) unsigned char smem[];
    start_together(go);
    if (blockIdx.x == 0 && threadIdx.x == 0) times[0] = gtime();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    unsigned char *stage = smem + (size_t)warp * 1024;
    const size_t ntiles = shard / 512;
    const size_t per_warp = (ntiles + (size_t)gridDim.x * warps - 1) / ((size_t)gridDim.x * warps);
    size_t t = ((size_t)blockIdx.x * warps + warp) * per_warp;
    const size_t t_end = t + per_warp < ntiles ? t + per_warp : ntiles;
    int buf = 0;
    for (; t < t_end; t++) {
        if (lane < d.n) planet::tma::wait_read<1>();
        __syncwarp();
        *reinterpret_cast<uint4 *>(stage + buf * 512 + lane * 16) = pattern(rank, t * 32 + lane);
        planet::tma::fence_smem_writes();
        __syncwarp();
        if (lane < d.n) {
            planet::tma::store_bulk(d.ptr[lane] + t * 512, stage + buf * 512, 512);
            planet::tma::commit();
        }
        buf ^= 1;
    }
    if (lane < d.n) planet::tma::wait_all<0>();
    finish(times, done);
}

__global__ void k_check(const char *buf, size_t shard, int world, unsigned long long *bad)
{
    const size_t nvec = shard / 16;
    unsigned long long mine = 0;
    for (int r = 0; r < world; r++)
        for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
            const uint4 got = *(reinterpret_cast<const uint4 *>(buf + (size_t)r * shard) + v), want = pattern(r, v);
            mine += got.x != want.x || got.y != want.y || got.z != want.z || got.w != want.w;
        }
    if (mine) atomicAdd(bad, mine);
}

int main(int argc, char **argv)
{
    int ndev = 0;
    OK(cudaGetDeviceCount(&ndev));
    const int world = argc > 1 ? atoi(argv[1]) : ndev;
    if (world < 2 || world > ndev || world > 8) { fprintf(stderr, "world %d of %d devices\n", world, ndev); return 2; }
    const double shard_mb = argc > 2 ? atof(argv[2]) : 48.0;              // C3 on 8 GPUs: 12 288 maps x 4 KB = 48 MiB per rank

    auto cuMemCreate_ = drv<CUresult (*)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long)>("cuMemCreate");
    auto cuMemAddressReserve_ = drv<CUresult (*)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long)>("cuMemAddressReserve");
    auto cuMemMap_ = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>("cuMemMap");
    auto cuMemSetAccess_ = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t)>("cuMemSetAccess");
    auto cuMemGetAllocationGranularity_ = drv<CUresult (*)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags)>("cuMemGetAllocationGranularity");
    auto cuMulticastCreate_ = drv<CUresult (*)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *)>("cuMulticastCreate");
    auto cuMulticastAddDevice_ = drv<CUresult (*)(CUmemGenericAllocationHandle, CUdevice)>("cuMulticastAddDevice");
    auto cuMulticastBindMem_ = drv<CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long)>("cuMulticastBindMem");
    auto cuMulticastGetGranularity_ = drv<CUresult (*)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags)>("cuMulticastGetGranularity");
    auto cuDeviceGetAttribute_ = drv<CUresult (*)(int *, CUdevice_attribute, CUdevice)>("cuDeviceGetAttribute");

    for (int d = 0; d < world; d++) { OK(cudaSetDevice(d)); OK(cudaFree(0)); }
    int mc_ok = 0;
    DRV(cuDeviceGetAttribute_(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, 0));
    printf("{\"multicast_supported\": %d, \"world\": %d}\n", mc_ok, world);
    if (!mc_ok) return 3;

    CUmulticastObjectProp mprop = {};
    mprop.numDevices = world; mprop.handleTypes = 0; mprop.flags = 0;
    size_t gran = 0;
    mprop.size = 2 << 20;
    DRV(cuMulticastGetGranularity_(&gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    size_t shard = (size_t)(shard_mb * (1 << 20)) / 32768 * 32768;
    size_t total = ((shard * world) + gran - 1) / gran * gran;
    mprop.size = total;
    CUmemGenericAllocationHandle mc;
    DRV(cuMulticastCreate_(&mc, &mprop));
    for (int d = 0; d < world; d++) DRV(cuMulticastAddDevice_(mc, d));

    std::vector<CUmemGenericAllocationHandle> mem(world);
    std::vector<CUdeviceptr> uni(world);
    std::vector<CUmemAccessDesc> access(world);
    for (int d = 0; d < world; d++) { access[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE; access[d].location.id = d; access[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; }
    for (int d = 0; d < world; d++) {
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = d;
        size_t g2 = 0;
        DRV(cuMemGetAllocationGranularity_(&g2, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        if (total % g2) { fprintf(stderr, "granularity %zu vs %zu\n", g2, total); return 1; }
        OK(cudaSetDevice(d));
        DRV(cuMemCreate_(&mem[d], total, &prop, 0));
        DRV(cuMulticastBindMem_(mc, 0, mem[d], 0, total, 0));
        DRV(cuMemAddressReserve_(&uni[d], total, gran, 0, 0));
        DRV(cuMemMap_(uni[d], total, 0, mem[d], 0));
        DRV(cuMemSetAccess_(uni[d], total, access.data(), world));        // every GPU may store into every buffer (NVLink peer access)
    }
    CUdeviceptr mc_va;
    DRV(cuMemAddressReserve_(&mc_va, total, gran, 0, 0));
    DRV(cuMemMap_(mc_va, total, 0, mc, 0));
    DRV(cuMemSetAccess_(mc_va, total, access.data(), world));

    int *go_h = nullptr;
    OK(cudaHostAlloc((void **)&go_h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    std::vector<unsigned long long *> times(world), bad(world);
    std::vector<unsigned int *> done(world);
    std::vector<cudaStream_t> st(world);
    int sms = 148;
    for (int d = 0; d < world; d++) {
        OK(cudaSetDevice(d));
        OK(cudaMalloc((void **)&times[d], 16)); OK(cudaMalloc((void **)&done[d], 4)); OK(cudaMalloc((void **)&bad[d], 8));
        OK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
        OK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024));
    }
    OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    auto run = [&](const char *name, int variant, int ctas_per_sm) {
        double best_ms = 1e30, best_first = 0;
        unsigned long long errors = 0;
        for (int rep = 0; rep < 4; rep++) {
            for (int d = 0; d < world; d++) {
                OK(cudaSetDevice(d));
                OK(cudaMemsetAsync((void *)uni[d], 0xff, shard * world, st[d]));
                OK(cudaMemsetAsync(done[d], 0, 4, st[d])); OK(cudaMemsetAsync(bad[d], 0, 8, st[d]));
                OK(cudaStreamSynchronize(st[d]));
            }
            *go_h = 0;
            for (int d = 0; d < world; d++) {
                OK(cudaSetDevice(d));
                Dest dest = {};
                if (variant == 0 || variant == 1) {                       // unicast: own buffer + the peers', rotated so GPUs do not gang up on one target
                    dest.n = world;
                    for (int k = 0; k < world; k++) dest.ptr[k] = (char *)uni[(d + k) % world] + (size_t)d * shard;
                } else {                                                  // one destination: the multicast address
                    dest.n = 1;
                    dest.ptr[0] = (char *)mc_va + (size_t)d * shard;
                }
                const int grid = sms * ctas_per_sm;
                if (variant == 0) k_uni_st16<<<grid, 512, 0, st[d]>>>(dest, shard, d, go_h, times[d], done[d]);
                else if (variant == 2) k_mc_st16<<<grid, 512, 0, st[d]>>>(dest.ptr[0], shard, d, go_h, times[d], done[d]);
                else k_bulk<<<sms, 768, 24 * 1024, st[d]>>>(dest, shard, d, go_h, times[d], done[d]);
                OK(cudaGetLastError());
            }
            *(volatile int *)go_h = 1;
            for (int d = 0; d < world; d++) { OK(cudaSetDevice(d)); OK(cudaStreamSynchronize(st[d])); }
            // %globaltimer is per GPU (the epochs differ): a GPU's time is its own end - its own start
            double slow = 0, fast = 1e30;
            for (int d = 0; d < world; d++) {
                unsigned long long t[2];
                OK(cudaSetDevice(d));
                OK(cudaMemcpy(t, times[d], 16, cudaMemcpyDeviceToHost));
                const double ms_d = (t[1] - t[0]) * 1e-6;
                if (ms_d > slow) slow = ms_d;
                if (ms_d < fast) fast = ms_d;
            }
            // multicast stores are posted: give the fabric a moment, then check every buffer
            for (int d = 0; d < world; d++) {
                OK(cudaSetDevice(d));
                k_check<<<sms * 2, 256, 0, st[d]>>>((const char *)uni[d], shard, world, bad[d]);
                unsigned long long b = 0;
                OK(cudaMemcpyAsync(&b, bad[d], 8, cudaMemcpyDeviceToHost, st[d]));
                OK(cudaStreamSynchronize(st[d]));
                errors += b;
            }
            if (slow < best_ms) { best_ms = slow; best_first = fast; }
        }
        const double recv_mb = shard * (double)(world - 1) / 1e6;
        printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"world\": %d, \"shard_MB\": %.1f, \"received_per_gpu_MB\": %.1f, \"ms_slowest_gpu\": %.4f, "
               "\"ms_fastest_gpu\": %.4f, \"ingress_GBs_per_gpu\": %.1f, \"wrong_vectors\": %llu}\n",
               name, ctas_per_sm, world, shard / 1e6, recv_mb, best_ms, best_first, recv_mb / best_ms, errors);
        fflush(stdout);
    };
    run("uni_st16", 0, 2);
    run("uni_st16", 0, 4);
    run("uni_bulk", 1, 1);
    run("mc_st16", 2, 1);
    run("mc_st16", 2, 2);
    run("mc_st16", 2, 4);
    run("mc_bulk", 3, 1);
    return 0;
}
