// multicast_ipc_probe.cu -- can an NVSwitch multicast object be shared between PROCESSES on this box, and how?
//
// The product runs one process per GPU (torchrun), so a multicast object created by rank 0 has to
// reach the other ranks as a shareable handle.  Two ways exist: a fabric handle (64 plain bytes, needs
// the IMEX service) and a POSIX file descriptor (has to travel over a Unix socket with SCM_RIGHTS).
// This probe forks N processes (before any CUDA call), rank r on GPU r, tries both, and then checks
// the thing end to end: every rank binds its own memory, rank r writes pattern(r) into shard r through
// the multicast address with multimem.st AND with a 512-byte cp.async.bulk, and every rank checks all
// shards in its LOCAL memory.  Prints one JSON line per rank.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o multicast_ipc_probe multicast_ipc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <sys/socket.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../planet_b200/csrc/planet_tma.cuh"

#define OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "[%d] %s: %s\n", g_rank, #call, cudaGetErrorString(e_)); _exit(1); } } while (0)
#define DRV(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { fprintf(stderr, "[%d] %s: CUresult %d\n", g_rank, #call, (int)r_); _exit(1); } } while (0)
static int g_rank = -1;

template <class F> static F drv(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { fprintf(stderr, "no driver entry point %s\n", name); _exit(1); }
    return (F)fn;
}

// ---- a file descriptor over a Unix socket (SCM_RIGHTS) ----
static int send_fd(int sock, int fd)
{
    char byte = 'F', ctl[CMSG_SPACE(sizeof(int))] = {};
    iovec iov = { &byte, 1 };
    msghdr msg = {};
    msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}
static int recv_fd(int sock)
{
    char byte = 0, ctl[CMSG_SPACE(sizeof(int))] = {};
    iovec iov = { &byte, 1 };
    msghdr msg = {};
    msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
    if (recvmsg(sock, &msg, 0) != 1) return -1;
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_type != SCM_RIGHTS) return -1;
    int fd = -1;
    memcpy(&fd, CMSG_DATA(c), sizeof(int));
    return fd;
}

__device__ __forceinline__ uint4 pattern(int rank, unsigned v, unsigned salt) { return make_uint4((unsigned)rank, v, salt, 0x12345678u); }

__global__ void k_write_mc(char *mc_shard, size_t shard, int rank)          // first half multimem.st, second half bulk copies
{
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t nvec = shard / 16, half = nvec / 2;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < half; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = pattern(rank, (unsigned)v, 1u);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     :: "l"(mc_shard + v * 16), "f"(__uint_as_float(x.x)), "f"(__uint_as_float(x.y)), "f"(__uint_as_float(x.z)), "f"(__uint_as_float(x.w)) : "memory");
    }
    // bulk half: each CTA stages 512-byte tiles
    const size_t ntiles = (nvec - half) / 32;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (threadIdx.x < 32) reinterpret_cast<uint4 *>(smem)[threadIdx.x] = pattern(rank, (unsigned)(half + t * 32 + threadIdx.x), 2u);
        planet::tma::fence_smem_writes();
        __syncthreads();
        if (threadIdx.x == 0) {
            planet::tma::store_bulk(mc_shard + (half + t * 32) * 16, smem, 512);
            planet::tma::commit();
            planet::tma::wait_all<0>();
        }
        __syncthreads();
    }
    __threadfence_system();
}

__global__ void k_check(const char *buf, size_t shard, int world, unsigned long long *bad)
{
    const size_t nvec = shard / 16, half = nvec / 2, bulk_end = half + (nvec - half) / 32 * 32;
    unsigned long long mine = 0;
    for (int r = 0; r < world; r++)
        for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < bulk_end; v += (size_t)gridDim.x * blockDim.x) {
            const uint4 got = *(reinterpret_cast<const uint4 *>(buf + (size_t)r * shard) + v), want = pattern(r, (unsigned)v, v < half ? 1u : 2u);
            mine += got.x != want.x || got.y != want.y || got.z != want.z || got.w != want.w;
        }
    if (mine) atomicAdd(bad, mine);
}

int main(int argc, char **argv)
{
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    const bool try_fabric = argc > 2 && !strcmp(argv[2], "fabric");
    // star of socket pairs around rank 0 (the product would use an abstract-namespace listening socket instead)
    std::vector<int> to_child(world, -1);
    int to_parent = -1;
    int rank = 0;
    for (int r = 1; r < world; r++) {
        int sv[2];
        if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv)) { perror("socketpair"); return 1; }
        pid_t pid = fork();
        if (pid == 0) { rank = r; to_parent = sv[1]; close(sv[0]); break; }
        to_child[r] = sv[0]; close(sv[1]);
    }
    g_rank = rank;
    auto barrier = [&]() {                                                    // through rank 0
        char b = 'b';
        if (rank == 0) { for (int r = 1; r < world; r++) if (read(to_child[r], &b, 1) != 1) _exit(1); for (int r = 1; r < world; r++) if (write(to_child[r], &b, 1) != 1) _exit(1); }
        else { if (write(to_parent, &b, 1) != 1) _exit(1); if (read(to_parent, &b, 1) != 1) _exit(1); }
    };

    OK(cudaSetDevice(rank));
    OK(cudaFree(0));
    auto cuMemCreate_ = drv<CUresult (*)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long)>("cuMemCreate");
    auto cuMemAddressReserve_ = drv<CUresult (*)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long)>("cuMemAddressReserve");
    auto cuMemMap_ = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>("cuMemMap");
    auto cuMemSetAccess_ = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t)>("cuMemSetAccess");
    auto cuMemExport_ = drv<CUresult (*)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long)>("cuMemExportToShareableHandle");
    auto cuMemImport_ = drv<CUresult (*)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType)>("cuMemImportFromShareableHandle");
    auto cuMulticastCreate_ = drv<CUresult (*)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *)>("cuMulticastCreate");
    auto cuMulticastAddDevice_ = drv<CUresult (*)(CUmemGenericAllocationHandle, CUdevice)>("cuMulticastAddDevice");
    auto cuMulticastBindMem_ = drv<CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long)>("cuMulticastBindMem");
    auto cuMulticastGetGranularity_ = drv<CUresult (*)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags)>("cuMulticastGetGranularity");
    auto cuDeviceGet_ = drv<CUresult (*)(CUdevice *, int)>("cuDeviceGet");

    const CUmemAllocationHandleType htype = try_fabric ? CU_MEM_HANDLE_TYPE_FABRIC : CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmulticastObjectProp mprop = {};
    mprop.numDevices = world; mprop.handleTypes = htype; mprop.flags = 0; mprop.size = 2 << 20;
    size_t gran = 0;
    DRV(cuMulticastGetGranularity_(&gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    const size_t shard = 4 << 20;
    const size_t total = (shard * world + gran - 1) / gran * gran;
    mprop.size = total;

    CUmemGenericAllocationHandle mc = 0;
    if (rank == 0) {
        DRV(cuMulticastCreate_(&mc, &mprop));
        if (try_fabric) {
            CUmemFabricHandle fh;
            CUresult r = cuMemExport_(&fh, mc, htype, 0);
            if (r != CUDA_SUCCESS) { printf("{\"rank\": 0, \"handle\": \"fabric\", \"export\": \"CUresult %d\"}\n", (int)r); fflush(stdout); for (int k = 1; k < world; k++) { char z[64] = {}; if (write(to_child[k], z, 64) != 64) _exit(1); } _exit(3); }
            for (int k = 1; k < world; k++) if (write(to_child[k], &fh, sizeof fh) != (ssize_t)sizeof fh) _exit(1);
        } else {
            int fd = -1;
            DRV(cuMemExport_(&fd, mc, htype, 0));
            for (int k = 1; k < world; k++) if (send_fd(to_child[k], fd)) { perror("send_fd"); _exit(1); }
        }
    } else {
        if (try_fabric) {
            CUmemFabricHandle fh;
            if (read(to_parent, &fh, sizeof fh) != (ssize_t)sizeof fh) _exit(1);
            char z[64] = {};
            if (!memcmp(&fh, z, 64)) _exit(3);
            DRV(cuMemImport_(&mc, &fh, htype));
        } else {
            int fd = recv_fd(to_parent);
            if (fd < 0) { fprintf(stderr, "[%d] recv_fd failed\n", rank); _exit(1); }
            DRV(cuMemImport_(&mc, (void *)(uintptr_t)fd, htype));
            close(fd);
        }
    }
    CUdevice dev;
    DRV(cuDeviceGet_(&dev, rank));
    DRV(cuMulticastAddDevice_(mc, dev));
    barrier();                                                                // every device added before anyone binds

    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = rank;
    prop.requestedHandleTypes = htype;
    CUmemGenericAllocationHandle mem;
    DRV(cuMemCreate_(&mem, total, &prop, 0));
    DRV(cuMulticastBindMem_(mc, 0, mem, 0, total, 0));
    CUmemAccessDesc access = {};
    access.location.type = CU_MEM_LOCATION_TYPE_DEVICE; access.location.id = rank; access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUdeviceptr uni, mc_va;
    DRV(cuMemAddressReserve_(&uni, total, gran, 0, 0));
    DRV(cuMemMap_(uni, total, 0, mem, 0));
    DRV(cuMemSetAccess_(uni, total, &access, 1));
    DRV(cuMemAddressReserve_(&mc_va, total, gran, 0, 0));
    DRV(cuMemMap_(mc_va, total, 0, mc, 0));
    DRV(cuMemSetAccess_(mc_va, total, &access, 1));
    OK(cudaMemset((void *)uni, 0xff, total));
    OK(cudaDeviceSynchronize());
    barrier();                                                                // everyone bound and cleared

    k_write_mc<<<64, 256, 512>>>((char *)mc_va + (size_t)rank * shard, shard, rank);
    OK(cudaGetLastError());
    OK(cudaDeviceSynchronize());
    barrier();                                                                // everyone has written
    usleep(2000);
    unsigned long long *bad, h_bad = 0;
    OK(cudaMalloc((void **)&bad, 8)); OK(cudaMemset(bad, 0, 8));
    k_check<<<128, 256>>>((const char *)uni, shard, world, bad);
    OK(cudaMemcpy(&h_bad, bad, 8, cudaMemcpyDeviceToHost));
    printf("{\"rank\": %d, \"world\": %d, \"handle\": \"%s\", \"granularity\": %zu, \"wrong_vectors\": %llu}\n", rank, world,
           try_fabric ? "fabric" : "posix_fd over SCM_RIGHTS", gran, h_bad);
    fflush(stdout);
    barrier();
    if (rank == 0) for (int r = 1; r < world; r++) { int st; wait(&st); }
    return h_bad ? 4 : 0;
}
