// k4_gather.cu -- K4: gathering the finished patches of all GPUs into one buffer on every GPU.
//
// The reference is one process on one CPU; the hot path here is sharded by patch range over the
// GPUs of a box, one process per GPU, and the only exchange on the path is this one:
// every rank's finished height maps (the compact form of a patch, 4 B per vertex) end up in ONE
// buffer, in the reference's emission order (main.cpp:589-592, 604-624), on every rank.
//
// Plumbing, all of it here in C++ behind the C-ABI (include/planet_gpu.h, planet_gpu_gather_*):
//   * NCCL (dlopen'ed: the library loads without it on a single-GPU box) provides the bootstrap
//     -- ncclCommInitRank from a 128-byte unique id the caller distributes by any means -- the
//     exchange of the CUDA IPC handles, the barriers of create/destroy, and the plain collective
//     (planet_gpu_gather_nccl: ncclAllGather, or grouped ncclBroadcast for ragged shards);
//   * every rank owns one allocation [flags | gathered buffer 0 | gathered buffer 1] and maps the
//     peers' allocations with cudaIpcOpenMemHandle (NVLink peer access through NVSwitch);
//   * planet_gpu_gather_height_maps is K2 with the gather fused in: each finished 128-sample tile
//     leaves the SM as one 512-byte bulk copy per destination (planet_tma.cuh) -- this GPU's
//     buffer and the same offset of every peer's -- so the all-gather rides under the arithmetic;
//   * completion is signalled GPU to GPU: after its kernel a rank bumps arrive[rank] in every
//     peer's flag block (st.release.sys), planet_gpu_gather_wait spins on the local flags
//     (ld.acquire.sys, bounded: a dead peer sets an error instead of hanging the box) and bumps
//     release[rank] on the peers once the caller is done reading; a launch that overwrites a
//     buffer first waits for the peers' releases of the step that filled it.  With two buffers
//     nothing in a step waits on the host: no host barrier, no stream synchronisation.
#include "planet_common.cuh"
#include "planet_tma.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace planet {

int launch_height_maps_gathered(const planet_gpu_params *, const Quad *, int64_t, int, int, float *, const PeerOut &, cudaStream_t);
int launch_gather_wait(const uint32_t *, uint32_t, int, int, uint32_t *, cudaStream_t);
bool height_maps_push_in_bulk(const planet_gpu_params *, int64_t, int, int, const float *, const PeerOut &);
bool shade_can_push(const planet_gpu_params *, const float *);
bool height_maps_progress_layout(const planet_gpu_params *, int64_t, int, int, const float *, int *, int64_t *, int64_t *);
int launch_shade_push(const planet_gpu_params *, const Quad *, int64_t, const double *, const float *,
                      const planet_gpu_texrect *, float, float *, float *, const PeerOut *, cudaStream_t);

namespace {

// ---- NCCL, resolved at run time ----------------------------------------------------------------
struct NcclApi {
    void *so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a process that already carries NCCL (torch bundles one) gets that copy back by soname
        for (const char *name : { "libnccl.so.2", "libnccl.so" }) {
            api.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.so) break;
        }
        if (!api.so) return;
        auto sym = [&](const char *n) { return dlsym(api.so, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast &&
                 api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

int check_nccl(ncclResult_t r, const char *what)
{
    if (r == ncclSuccess) return 0;
    return set_error(PLANET_E_CUDA, "%s: NCCL: %s", what, nccl_api().GetErrorString(r));
}
#define PLANET_NCCL(call) do { int _rc = check_nccl((call), #call); if (_rc) return _rc; } while (0)

// ---- the flag block at the head of every rank's allocation (uint32 words) ----------------------
constexpr int FLAG_ARRIVE = 0;      // arrive[r]:  rank r's shard of step `value` is complete in THIS rank's buffer
constexpr int FLAG_RELEASE = 16;    // release[r]: rank r is done reading ITS buffer of step `value`
constexpr int FLAG_ERROR = 32;      // set by a wait that timed out
constexpr size_t FLAG_BYTES = 4096;

struct PeerFlags { uint32_t *ptr[7]; int n; };

// after this rank's K2 + pushes: tell every peer that our shard of step `value` has landed
__global__ void k_gather_signal(PeerFlags pf, int slot, uint32_t value)
{
    if ((int)threadIdx.x < pf.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pf.ptr[threadIdx.x] + slot), "r"(value) : "memory");
    }
}

// arrive[r] >= step for every peer, then (optionally) release[rank] = step on every peer, in one launch
__global__ void k_gather_wait_release(const uint32_t *flags, uint32_t step, int rank, int world, uint32_t *error,
                                      PeerFlags pf, int release)
{
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank)
        gather_wait_flag(flags + FLAG_ARRIVE + threadIdx.x, step, error);
    __syncwarp();
    if (release && (int)threadIdx.x < pf.n)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pf.ptr[threadIdx.x] + FLAG_RELEASE + rank), "r"(step) : "memory");
}

// The pushing done by a kernel of its own, beside the compute kernels: 16-byte vectors of a finished
// range of the LOCAL buffer are loaded (L2 hits: K2 has just written them) and stored to the same
// offset of every peer's buffer.  128 threads and 32 registers per CTA, no shared memory: one such
// CTA fits on an SM next to a resident 768-thread K2 CTA (61 440 of the 65 536 registers, 217 of the
// 228 KB of shared memory), so the transfer runs on the SMs that are computing the next chunk and
// costs them ~8 issue slots per 112 bytes sent -- while no compute warp ever waits for the link.
struct PushDst { float4 *ptr[7]; int n; };

__global__ void __launch_bounds__(128, 16)
k_push_range(const float4 *__restrict__ src, PushDst dst, size_t nvec)
{
    const size_t stride = (size_t)gridDim.x * 128;
    size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    for (; i + stride < nvec; i += 2 * stride) {                         // two loads in flight per thread
        const float4 a = __ldcg(src + i), b = __ldcg(src + i + stride);
#pragma unroll
        for (int k = 0; k < 7; k++)
            if (k < dst.n) { __stcs(dst.ptr[k] + i, a); __stcs(dst.ptr[k] + i + stride, b); }
    }
    if (i < nvec) {
        const float4 a = __ldcg(src + i);
#pragma unroll
        for (int k = 0; k < 7; k++)
            if (k < dst.n) __stcs(dst.ptr[k] + i, a);
    }
}

// The pusher that runs BESIDE one un-chunked K2 launch (PLANET_GATHER_PUSH_CONCURRENT).  K2's warps own
// contiguous runs of 512-byte tiles and publish how many they have finished (PeerOut::progress); every
// pusher warp looks after a fixed set of K2 warps, round robin: it acquires the counter, and sends the
// tiles finished since its last visit -- one 16-byte vector per lane loaded from the local buffer (L2),
// stored to the same offset of every peer's buffer -- until all its warps are done.  A counter carries
// the step in its upper half, so one left over from an earlier step reads as "nothing yet" and one from
// a later step (this pusher lagging behind the next K2) as "all done".
constexpr int PUSH_MAX_OWNED = 32;     // K2 warps per pusher warp

__global__ void __launch_bounds__(128, 16)
k_push_progress(const float4 *__restrict__ src, PushDst dst, const unsigned long long *progress, uint32_t tag,
                int nwarps, long long per_warp, long long nwtiles, long long total_vec, uint32_t *error)
{
    __shared__ uint32_t s_sent[4][PUSH_MAX_OWNED];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pw = blockIdx.x * 4 + w, npw = gridDim.x * 4;
    const int owned = pw < nwarps ? (nwarps - pw + npw - 1) / npw : 0;       // K2 warps pw, pw + npw, ...
    for (int j = lane; j < PUSH_MAX_OWNED; j += 32) s_sent[w][j] = 0;
    __syncwarp();
    int open = owned;
    unsigned long long idle_since = 0;
    while (open > 0) {
        open = 0;
        bool moved = false;
        for (int j = 0; j < owned; j++) {
            const int gw = pw + j * npw;
            const long long first = (long long)gw * per_warp;
            const long long mine = min(per_warp, max(0ll, nwtiles - first));  // tiles this K2 warp owns
            const uint32_t sent = s_sent[w][j];
            if (sent >= (uint32_t)mine) continue;
            // polled with a RELAXED load: an acquire load invalidates the SM's L1 every time, and L1 is the same
            // array K2's table lookups live in (K2 ran at half speed beside a pusher that polled with acquire);
            // the acquire fence is paid once per observed advance instead
            unsigned long long v;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(progress + gw) : "memory");
            const uint32_t vt = (uint32_t)(v >> 32);
            uint32_t done = vt == tag ? (uint32_t)v : ((int32_t)(vt - tag) > 0 ? (uint32_t)mine : 0u);
            done = min(done, (uint32_t)mine);
            if (done > sent) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (uint32_t t = sent; t < done; t++) {
                const long long vec = (first + t) * 32 + lane;               // 32 vectors of 16 bytes per tile
                if (vec < total_vec) {
                    const float4 x = __ldcg(src + vec);
#pragma unroll
                    for (int k = 0; k < 7; k++)
                        if (k < dst.n) __stcs(dst.ptr[k] + vec, x);
                }
            }
            if (done > sent) { moved = true; __syncwarp(); if (lane == 0) s_sent[w][j] = done; __syncwarp(); }
            if (done < (uint32_t)mine) open++;
        }
        if (open && !moved) {
            // bounded like every wait in this file: 2 s without any progress of K2 and the pusher gives up
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!idle_since) idle_since = now;
            else if (now - idle_since > 2000000000ull) { if (lane == 0 && error) atomicExch(error, 1u); return; }
            __nanosleep(1000);
        } else {
            idle_since = 0;
        }
    }
}

// The same pusher moving the bytes with the TMA unit instead of the load/store unit: K2's octave loop
// keeps the SM's shared-memory pipe ~75 % busy, and 16-byte stores that wait for NVLink in that pipe
// hold K2's table lookups up behind them (K2 ran at half speed beside k_push_progress).  Here one lane
// per warp only issues bulk copies: finished tiles come from the local buffer into a 1 KB staging slot
// (global -> shared, completion on an mbarrier) and leave for every peer (shared -> global), two slots
// per warp in flight; the data never passes through a register.
constexpr int PUSH_SLOT = 1024;        // bytes per staging slot: 2 consecutive tiles of one K2 warp (8 KB per CTA: what K2 leaves of the SM)

__global__ void __launch_bounds__(128, 16)
k_push_progress_tma(const char *__restrict__ src, PushDst dst, const unsigned long long *progress, uint32_t tag,
                    int nwarps, long long per_warp, long long nwtiles, long long total_bytes, uint32_t *error)
{
    __shared__ __align__(128) unsigned char s_stage[4][2][PUSH_SLOT];
    __shared__ __align__(8) uint64_t s_bar[4][2];
    __shared__ uint32_t s_sent[4][PUSH_MAX_OWNED];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane != 0) return;                                                   // one issuing lane per warp; nothing here is warp-collective
    const int pw = blockIdx.x * 4 + w, npw = gridDim.x * 4;
    const int owned = pw < nwarps ? (nwarps - pw + npw - 1) / npw : 0;
    for (int j = 0; j < owned; j++) s_sent[w][j] = 0;
    tma::mbar_init(&s_bar[w][0], 1); tma::mbar_init(&s_bar[w][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t phase[2] = { 0, 0 };
    int slot = 0;
    int open = owned;
    unsigned long long idle_since = 0;
    while (open > 0) {
        open = 0;
        bool moved = false;
        for (int j = 0; j < owned; j++) {
            const int gw = pw + j * npw;
            const long long first = (long long)gw * per_warp;
            const long long mine = min(per_warp, max(0ll, nwtiles - first));
            uint32_t sent = s_sent[w][j];
            if (sent >= (uint32_t)mine) continue;
            unsigned long long v;                                            // relaxed poll, see k_push_progress
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(progress + gw) : "memory");
            const uint32_t vt = (uint32_t)(v >> 32);
            uint32_t done = vt == tag ? (uint32_t)v : ((int32_t)(vt - tag) > 0 ? (uint32_t)mine : 0u);
            done = min(done, (uint32_t)mine);
            if (done > sent) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); tma::fence_before_async_reads(); }
            while (sent < done) {
                const uint32_t ntile = min(done - sent, (uint32_t)(PUSH_SLOT / 512));
                const long long at = (first + sent) * 512;
                const uint32_t bytes = (uint32_t)min((long long)ntile * 512, total_bytes - at);
                tma::wait_read<1>();                                         // the slot's previous contents have left it
                tma::mbar_expect_tx(&s_bar[w][slot], bytes);
                tma::load_bulk(s_stage[w][slot], src + at, bytes, &s_bar[w][slot]);
                while (!tma::mbar_try_wait(&s_bar[w][slot], phase[slot])) { }
                phase[slot] ^= 1;
#pragma unroll
                for (int k = 0; k < 7; k++)
                    if (k < dst.n) tma::store_bulk(reinterpret_cast<char *>(dst.ptr[k]) + at, s_stage[w][slot], bytes);
                tma::commit();
                slot ^= 1;
                sent += ntile;
                moved = true;
            }
            s_sent[w][j] = sent;
            if (sent < (uint32_t)mine) open++;
        }
        if (open && !moved) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!idle_since) idle_since = now;
            else if (now - idle_since > 2000000000ull) { if (error) atomicExch(error, 1u); break; }
            __nanosleep(1000);
        } else {
            idle_since = 0;
        }
    }
    tma::wait_all<0>();
}

struct Gather {
    int rank = 0, world = 1, nbuf = 1, device = 0;
    size_t bytes = 0, stride = 0;
    ncclComm_t comm = nullptr;
    char *base = nullptr;               // [flags | buffer 0 | buffer 1]
    char *peer_base[8] = {};            // the peers' allocations mapped here; peer_base[rank] == base
    uint32_t step = 0;                  // gathers issued so far (fused or copy-engine)
    int last_buffer = 0;
    // copy-engine path: one stream per peer and a ring of events that order the copies behind the
    // kernels of the caller's stream
    int push_mode = PLANET_GATHER_PUSH_COPY_ENGINES;   // who moves the bytes of planet_gpu_gather_push
    int sm_count = 148;
    int k3_every = 0;                   // > 0: the shade kernel pushes every k3_every-th quad (planet_gpu_gather_set_shade_share)
    unsigned long long *progress = nullptr;   // PLANET_GATHER_PUSH_CONCURRENT: one counter per K2 warp
    int progress_cap = 0;
    bool carveout_set = false;
    PeerOut pending = {};               // the peers of the step whose shade-kernel share is still to be pushed
    bool shade_pending = false;
    cudaStream_t push_stream[7] = {};
    cudaEvent_t ring[32] = {};
    int ring_at = 0;

    float *buffer(int r, int b) const { return reinterpret_cast<float *>(peer_base[r] + FLAG_BYTES + (size_t)b * stride); }
    uint32_t *flags(int r) const { return reinterpret_cast<uint32_t *>(peer_base[r]); }
    PeerFlags peer_flags() const
    {
        PeerFlags pf = {};
        pf.n = world - 1;
        for (int k = 0; k < world - 1; k++) pf.ptr[k] = flags((rank + 1 + k) % world);
        return pf;
    }
};

int nccl_barrier(Gather *g, cudaStream_t stream)
{
    if (g->world == 1) return check_cuda(cudaStreamSynchronize(stream), "sync");
    int *word = reinterpret_cast<int *>(g->base + FLAG_BYTES - 64);      // scratch inside the flag block
    PLANET_NCCL(nccl_api().AllReduce(word, word, 1, ncclInt, ncclSum, g->comm, stream));
    return check_cuda(cudaStreamSynchronize(stream), "barrier sync");
}

} // namespace
} // namespace planet

using namespace planet;

extern "C" {

int planet_gpu_gather_unique_id(void *id)
{
    if (!id) return set_error(PLANET_E_INVALID, "id is NULL");
    NcclApi &api = nccl_api();
    if (!api.ok) return set_error(PLANET_E_UNSUPPORTED, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    static_assert(sizeof(ncclUniqueId) == PLANET_GATHER_ID_BYTES, "unique id size is part of the ABI");
    ncclUniqueId uid;
    PLANET_NCCL(api.GetUniqueId(&uid));
    memcpy(id, &uid, sizeof uid);
    return 0;
}

void *planet_gpu_gather_create(const void *id, int rank, int world, int64_t bytes, int n_buffers)
{
    if (!ensure_init()) return nullptr;
    if (world < 1 || world > 8 || rank < 0 || rank >= world || bytes <= 0 || n_buffers < 1 || n_buffers > 2 ||
        (world > 1 && !id)) {
        set_error(PLANET_E_INVALID, "gather_create(rank=%d, world=%d, bytes=%lld, n_buffers=%d): world in [1, 8], "
                  "rank in [0, world), n_buffers 1 or 2, id required when world > 1", rank, world, (long long)bytes, n_buffers);
        return nullptr;
    }
    Gather *g = new Gather();
    g->rank = rank; g->world = world; g->nbuf = n_buffers; g->bytes = (size_t)bytes;
    g->stride = ((size_t)bytes + 4095) & ~(size_t)4095;
    cudaGetDevice(&g->device);
    auto fail = [&](const char *what) -> void * {
        char msg[400];
        snprintf(msg, sizeof msg, "%s", planet_gpu_last_error());
        for (int r = 0; r < 8; r++) if (r != g->rank && g->peer_base[r]) cudaIpcCloseMemHandle(g->peer_base[r]);
        for (auto &st : g->push_stream) if (st) cudaStreamDestroy(st);
        for (auto &e : g->ring) if (e) cudaEventDestroy(e);
        if (g->base) cudaFree(g->base);
        if (g->comm) nccl_api().CommDestroy(g->comm);
        delete g;
        set_error(PLANET_E_CUDA, "gather_create: %s: %s", what, msg);
        return nullptr;
    };
    if (world > 1) {
        NcclApi &api = nccl_api();
        if (!api.ok) { set_error(PLANET_E_UNSUPPORTED, "libnccl.so.2 could not be loaded"); return fail("NCCL"); }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        if (check_nccl(api.CommInitRank(&g->comm, world, uid, rank), "ncclCommInitRank")) return fail("communicator");
    }
    if (check_cuda(cudaMalloc(&g->base, FLAG_BYTES + (size_t)n_buffers * g->stride), "cudaMalloc(gathered buffers)")) return fail("allocation");
    if (check_cuda(cudaMemset(g->base, 0, FLAG_BYTES), "cudaMemset(flags)")) return fail("flags");
    g->peer_base[rank] = g->base;
    if (world > 1) {
        // exchange the IPC handles of the allocations with an all-gather on the new communicator
        cudaIpcMemHandle_t mine, all[8];
        char *d_x = nullptr;
        if (check_cuda(cudaIpcGetMemHandle(&mine, g->base), "cudaIpcGetMemHandle")) return fail("IPC handle");
        if (check_cuda(cudaMalloc(&d_x, sizeof mine * 9), "cudaMalloc(handles)")) return fail("IPC handle exchange");
        bool ok = !check_cuda(cudaMemcpy(d_x + 8 * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice), "H2D handle") &&
                  !check_nccl(nccl_api().AllGather(d_x + 8 * sizeof mine, d_x, sizeof mine, ncclChar, g->comm, nullptr), "ncclAllGather(handles)") &&
                  !check_cuda(cudaStreamSynchronize(nullptr), "handle exchange") &&
                  !check_cuda(cudaMemcpy(all, d_x, sizeof mine * world, cudaMemcpyDeviceToHost), "D2H handles");
        cudaFree(d_x);
        if (!ok) return fail("IPC handle exchange");
        for (int r = 0; r < world; r++) {
            if (r == rank) continue;
            void *p = nullptr;
            if (check_cuda(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return fail("peer mapping");
            g->peer_base[r] = (char *)p;
        }
        // The side streams keep the DEFAULT priority.  Measured (profiles/r02y_*): with the pusher on a
        // high-priority stream the K2 kernel it shares the SMs with runs at half speed for as long as the
        // pusher is resident, whatever the pusher does (even only sleeping between polls)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, g->device);
        for (int k = 0; k < world - 1; k++)
            if (check_cuda(cudaStreamCreateWithPriority(&g->push_stream[k], cudaStreamNonBlocking, getenv("PLANET_PUSH_HIGHPRIO") ? prio_hi : prio_lo), "push stream")) return fail("streams");
        for (auto &e : g->ring)
            if (check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")) return fail("events");
        if (nccl_barrier(g, nullptr)) return fail("barrier");            // every rank has mapped every buffer
    }
    return g;
}

void planet_gpu_gather_destroy(void *gather)
{
    Gather *g = (Gather *)gather;
    if (!g) return;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    // Order matters: nobody may still be writing into a buffer when its mappings go, and the
    // owner frees only after every peer has closed its mapping.
    if (g->world > 1) nccl_barrier(g, nullptr);
    for (int r = 0; r < g->world; r++)
        if (r != g->rank && g->peer_base[r]) cudaIpcCloseMemHandle(g->peer_base[r]);
    if (g->world > 1) nccl_barrier(g, nullptr);
    for (auto &st : g->push_stream) if (st) cudaStreamDestroy(st);
    for (auto &e : g->ring) if (e) cudaEventDestroy(e);
    if (g->progress) cudaFree(g->progress);
    if (g->base) cudaFree(g->base);
    if (g->comm) nccl_api().CommDestroy(g->comm);
    delete g;
}

float *planet_gpu_gather_buffer(void *gather, int which)
{
    Gather *g = (Gather *)gather;
    if (!g || which < 0 || which >= g->nbuf) { set_error(PLANET_E_INVALID, "gather_buffer(%d)", which); return nullptr; }
    return g->buffer(g->rank, which);
}

int planet_gpu_gather_last_buffer(const void *gather) { return gather ? ((const Gather *)gather)->last_buffer : 0; }

int planet_gpu_gather_height_maps(void *gather, const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                  int64_t nquads, int64_t first_quad, int dim, int max_depth, void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (dim <= 3 || dim > 32768) return set_error(PLANET_E_INVALID, "dim %d outside [4, 32768]", dim);
    if (p->fixed_octaves <= 0 && max_depth == 0) return set_error(PLANET_E_INVALID, "max_depth == 0 divides by zero at main.cpp:827");
    const size_t texels = (size_t)dim * dim;
    if (nquads < 0 || first_quad < 0 || (nquads > 0 && !d_quads) || (size_t)(first_quad + nquads) * texels * sizeof(float) > g->bytes)
        return set_error(PLANET_E_INVALID, "quads [%lld, %lld) x %d^2 floats do not fit the %zu-byte gathered buffer",
                         (long long)first_quad, (long long)(first_quad + nquads), dim, g->bytes);
    cudaStream_t stream = (cudaStream_t)stream_;
    const uint32_t step = ++g->step;
    const int b = (int)((step - 1) % (uint32_t)g->nbuf);
    g->last_buffer = b;
    const size_t off = (size_t)first_quad * texels;
    PeerOut peers = {};
    peers.n = g->world - 1;
    for (int k = 0; k < peers.n; k++) peers.ptr[k] = g->buffer((g->rank + 1 + k) % g->world, b) + off;
    if (g->world > 1) {
        // step s overwrites what step s - nbuf put into the peers' buffers: wait for their releases
        peers.release = g->flags(g->rank) + FLAG_RELEASE;
        peers.release_min = step > (uint32_t)g->nbuf ? step - (uint32_t)g->nbuf : 0u;
        peers.rank = g->rank; peers.world = g->world;
        peers.error = g->flags(g->rank) + FLAG_ERROR;
    }
    g->shade_pending = false;
    peers.quad0 = first_quad;
    // One un-chunked K2 launch that pushes nothing itself, and the pusher kernel beside it on the side stream
    int k2_warps = 0;
    int64_t per_warp = 0, nwtiles = 0;
    if (g->world > 1 && g->push_mode == PLANET_GATHER_PUSH_CONCURRENT && nquads > 0 &&
        height_maps_progress_layout(p, nquads, dim, max_depth, g->buffer(g->rank, b) + off, &k2_warps, &per_warp, &nwtiles) &&
        (off * sizeof(float)) % 16 == 0) {
        if (k2_warps > g->progress_cap) {
            if (g->progress) { PLANET_CUDA(cudaStreamSynchronize(g->push_stream[0])); cudaFree(g->progress); g->progress = nullptr; g->progress_cap = 0; }
            PLANET_CUDA(cudaMalloc(&g->progress, (size_t)k2_warps * sizeof(unsigned long long)));
            PLANET_CUDA(cudaMemset(g->progress, 0, (size_t)k2_warps * sizeof(unsigned long long)));
            g->progress_cap = k2_warps;
        }
        PeerOut local = peers;
        local.n = 0;                                                      // K2 stores to this rank's buffer only ...
        local.progress = g->progress;                                     // ... and says how far it is
        local.progress_tag = step;
        {   // how often a K2 warp publishes (every 2^k tiles; tuning knob PLANET_PUSH_EVERY, used by tools/gather_modes.py)
            const char *e = getenv("PLANET_PUSH_EVERY");
            int every = e ? atoi(e) : 1;
            while (every & (every - 1)) every &= every - 1;
            local.progress_mask = (uint32_t)(every > 0 ? every - 1 : 0);
        }
        PushDst dst = {};
        dst.n = g->world - 1;
        for (int k = 0; k < dst.n; k++) dst.ptr[k] = reinterpret_cast<float4 *>(peers.ptr[k]);
        // the pusher may start as soon as the stream reaches this point (it only ever follows K2's counters,
        // and K2 itself waits for the peers' release of the buffer before its first tile)
        cudaEvent_t ev = g->ring[g->ring_at++ & 31];
        PLANET_CUDA(cudaEventRecord(ev, stream));
        PLANET_CUDA(cudaStreamWaitEvent(g->push_stream[0], ev, 0));
        rc = launch_height_maps_gathered(p, (const Quad *)d_quads, nquads, dim, max_depth, g->buffer(g->rank, b) + off, local, stream);
        if (rc) return rc;
        // the pusher's CTAs must be able to share an SM with K2's, whichever arrives first: an SM keeps one
        // shared-memory carve-out while anything is resident on it, and K2 needs the largest
        if (!g->carveout_set) {
            PLANET_CUDA(cudaFuncSetAttribute(k_push_progress, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            PLANET_CUDA(cudaFuncSetAttribute(k_push_progress_tma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            g->carveout_set = true;
        }
        int pushers = std::min(g->sm_count, (k2_warps + 3) / 4);
        if (getenv("PLANET_PUSH_CTAS")) pushers = std::max(1, std::min(pushers, atoi(getenv("PLANET_PUSH_CTAS"))));
        const int owned = (k2_warps + pushers * 4 - 1) / (pushers * 4);
        if (owned > PUSH_MAX_OWNED) return set_error(PLANET_E_UNSUPPORTED, "gather: %d K2 warps per pusher warp", owned);
        const long long total_vec = (long long)nquads * (long long)texels / 4;
        if (getenv("PLANET_PUSH_LSU"))
            k_push_progress<<<pushers, 128, 0, g->push_stream[0]>>>(reinterpret_cast<const float4 *>(g->buffer(g->rank, b) + off), dst,
                                                                    g->progress, step, k2_warps, per_warp, nwtiles, total_vec, g->flags(g->rank) + FLAG_ERROR);
        else
            k_push_progress_tma<<<pushers, 128, 0, g->push_stream[0]>>>(reinterpret_cast<const char *>(g->buffer(g->rank, b) + off), dst,
                                                                        g->progress, step, k2_warps, per_warp, nwtiles, total_vec * 16, g->flags(g->rank) + FLAG_ERROR);
        count_launch();
        PLANET_CUDA(cudaGetLastError());
        k_gather_signal<<<1, 32, 0, g->push_stream[0]>>>(g->peer_flags(), FLAG_ARRIVE + g->rank, step);
        count_launch();
        return check_cuda(cudaGetLastError(), "gather signal launch");
    }
    // part of the pushing may be left to the shade kernel (planet_gpu_gather_shade must then follow)
    if (g->world > 1 && (g->k3_every > 1 || g->k3_every < 0) && nquads > 0 && dim == p->patch_verts + 2 &&
        height_maps_push_in_bulk(p, nquads, dim, max_depth, g->buffer(g->rank, b) + off, peers) &&
        shade_can_push(p, g->buffer(g->rank, b) + off)) {
        peers.k3_every = g->k3_every;
        g->pending = peers;
        g->shade_pending = true;
    }
    if (nquads > 0) {
        rc = launch_height_maps_gathered(p, (const Quad *)d_quads, nquads, dim, max_depth, g->buffer(g->rank, b) + off, peers, stream);
        if (rc) return rc;
    } else if (g->world > 1) {
        rc = launch_gather_wait(peers.release, peers.release_min, g->rank, g->world, peers.error, stream);
        if (rc) return rc;
    }
    if (g->world > 1 && !g->shade_pending) {
        k_gather_signal<<<1, 32, 0, stream>>>(g->peer_flags(), FLAG_ARRIVE + g->rank, step);
        count_launch();
        PLANET_CUDA(cudaGetLastError());
    }
    return 0;
}

int planet_gpu_gather_set_shade_share(void *gather, int every)
{
    Gather *g = (Gather *)gather;
    if (!g || every < -7 || every == 1) return set_error(PLANET_E_INVALID, "gather_set_shade_share(%d): 0 (off), >= 2 (one map in `every`) or -1 .. -7 (that many maps in 8)", every);
    g->k3_every = every;
    return 0;
}

// K3 for the quads of the last planet_gpu_gather_height_maps, reading their maps from the gathered
// buffer; pushes the share of the maps that call left to it, then signals the peers
int planet_gpu_gather_shade(void *gather, const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t nquads,
                            int64_t first_quad, const double *cam_pos, float max_skirt, float *d_pos4, float *d_nrm4,
                            void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (!cam_pos || nquads < 0 || (nquads > 0 && !d_quads)) return set_error(PLANET_E_INVALID, "NULL argument");
    const int dim = p->patch_verts + 2;
    const size_t texels = (size_t)dim * dim;
    if (first_quad < 0 || (size_t)(first_quad + nquads) * texels * sizeof(float) > g->bytes)
        return set_error(PLANET_E_INVALID, "quads [%lld, %lld) outside the gathered buffer", (long long)first_quad, (long long)(first_quad + nquads));
    if (g->shade_pending && g->pending.quad0 != first_quad)
        return set_error(PLANET_E_INVALID, "gather_shade must cover the quads of the preceding gather_height_maps");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (max_skirt < 0.0f) max_skirt = planet_gpu_max_skirt_size(p->radius, p->patch_verts);
    const float *maps = g->buffer(g->rank, g->last_buffer) + (size_t)first_quad * texels;
    rc = launch_shade_push(p, (const Quad *)d_quads, nquads, cam_pos, maps, nullptr, max_skirt, d_pos4, d_nrm4,
                           g->shade_pending ? &g->pending : nullptr, stream);
    if (rc) return rc;
    if (g->shade_pending) {
        g->shade_pending = false;
        k_gather_signal<<<1, 32, 0, stream>>>(g->peer_flags(), FLAG_ARRIVE + g->rank, g->step);
        count_launch();
        PLANET_CUDA(cudaGetLastError());
    }
    return 0;
}

// ---- the copy-engine path: begin / push / publish -------------------------------------------------
// The fused kernel above couples the NVLink transfer to the arithmetic: a warp that has to wait for
// the link cannot compute, and the transfer ends with the kernel, so nothing after K2 overlaps it.
// Here the height-map kernels run undisturbed in chunks and every finished chunk is handed to the
// copy engines (one cudaMemcpyAsync per peer on the gather's own streams, ordered behind the chunk
// by an event), so the transfer also runs under the kernels that FOLLOW K2 (K3, the index stream).
int planet_gpu_gather_begin(void *gather, void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    const uint32_t step = ++g->step;
    g->last_buffer = (int)((step - 1) % (uint32_t)g->nbuf);
    if (g->world > 1 && step > (uint32_t)g->nbuf)                      // the peers have released the buffer this step overwrites
        return launch_gather_wait(g->flags(g->rank) + FLAG_RELEASE, step - (uint32_t)g->nbuf, g->rank, g->world,
                                  g->flags(g->rank) + FLAG_ERROR, (cudaStream_t)stream_);
    return 0;
}

int planet_gpu_gather_push(void *gather, int64_t offset_bytes, int64_t size_bytes, void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g || offset_bytes < 0 || size_bytes < 0 || (size_t)(offset_bytes + size_bytes) > g->bytes)
        return set_error(PLANET_E_INVALID, "gather_push: range outside the buffer");
    if (g->world == 1 || size_bytes == 0) return 0;
    cudaEvent_t ev = g->ring[g->ring_at++ & 31];
    PLANET_CUDA(cudaEventRecord(ev, (cudaStream_t)stream_));
    const char *src = reinterpret_cast<const char *>(g->buffer(g->rank, g->last_buffer)) + offset_bytes;
    if (g->push_mode == PLANET_GATHER_PUSH_SM_KERNEL) {
        if ((offset_bytes | size_bytes) & 15) return set_error(PLANET_E_INVALID, "gather_push: the SM pusher moves 16-byte vectors (offset %lld, size %lld)", (long long)offset_bytes, (long long)size_bytes);
        PushDst dst = {};
        dst.n = g->world - 1;
        for (int k = 0; k < dst.n; k++)
            dst.ptr[k] = reinterpret_cast<float4 *>(reinterpret_cast<char *>(g->buffer((g->rank + 1 + k) % g->world, g->last_buffer)) + offset_bytes);
        PLANET_CUDA(cudaStreamWaitEvent(g->push_stream[0], ev, 0));
        const size_t nvec = (size_t)size_bytes / 16;
        const int grid = (int)std::min<size_t>((size_t)g->sm_count, (nvec + 127) / 128);
        k_push_range<<<grid, 128, 0, g->push_stream[0]>>>(reinterpret_cast<const float4 *>(src), dst, nvec);
        count_launch();
        return check_cuda(cudaGetLastError(), "gather push launch");
    }
    for (int k = 0; k < g->world - 1; k++) {
        char *dst = reinterpret_cast<char *>(g->buffer((g->rank + 1 + k) % g->world, g->last_buffer)) + offset_bytes;
        PLANET_CUDA(cudaStreamWaitEvent(g->push_stream[k], ev, 0));
        PLANET_CUDA(cudaMemcpyAsync(dst, src, (size_t)size_bytes, cudaMemcpyDeviceToDevice, g->push_stream[k]));
    }
    return 0;
}

// every push so far has been queued: each peer is told (arrive flag) as soon as ITS copies are done
int planet_gpu_gather_publish(void *gather)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    if (g->world > 1 && g->push_mode == PLANET_GATHER_PUSH_SM_KERNEL) {      // every push sits on one stream: one signal to all peers
        k_gather_signal<<<1, 32, 0, g->push_stream[0]>>>(g->peer_flags(), FLAG_ARRIVE + g->rank, g->step);
        count_launch();
        return check_cuda(cudaGetLastError(), "gather publish launch");
    }
    for (int k = 0; k < g->world - 1; k++) {
        PeerFlags one = {};
        one.n = 1;
        one.ptr[0] = g->flags((g->rank + 1 + k) % g->world);
        k_gather_signal<<<1, 32, 0, g->push_stream[k]>>>(one, FLAG_ARRIVE + g->rank, g->step);
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "gather publish launch");
}

int planet_gpu_gather_set_push_mode(void *gather, int mode)
{
    Gather *g = (Gather *)gather;
    if (!g || mode < PLANET_GATHER_PUSH_COPY_ENGINES || mode > PLANET_GATHER_PUSH_CONCURRENT)
        return set_error(PLANET_E_INVALID, "gather_set_push_mode(%d)", mode);
    g->push_mode = mode;
    return 0;
}

int planet_gpu_gather_wait(void *gather, int release, void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    if (g->world == 1 || g->step == 0) return 0;
    k_gather_wait_release<<<1, 32, 0, (cudaStream_t)stream_>>>(g->flags(g->rank), g->step, g->rank, g->world,
                                                                g->flags(g->rank) + FLAG_ERROR, g->peer_flags(), release);
    count_launch();
    return check_cuda(cudaGetLastError(), "gather wait launch");
}

int planet_gpu_gather_nccl(void *gather, int which, const int64_t *offset_bytes, const int64_t *size_bytes, void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g || which < 0 || which >= g->nbuf || !offset_bytes || !size_bytes) return set_error(PLANET_E_INVALID, "gather_nccl: bad argument");
    if (g->world == 1) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    char *buf = reinterpret_cast<char *>(g->buffer(g->rank, which));
    bool equal = true;
    for (int r = 0; r < g->world; r++) {
        if (offset_bytes[r] < 0 || size_bytes[r] < 0 || (size_t)(offset_bytes[r] + size_bytes[r]) > g->bytes)
            return set_error(PLANET_E_INVALID, "gather_nccl: shard %d outside the buffer", r);
        equal = equal && size_bytes[r] == size_bytes[0] && offset_bytes[r] == offset_bytes[0] + r * size_bytes[0];
    }
    NcclApi &api = nccl_api();
    if (equal) {                                                          // equal shards in rank order: one in-place all-gather
        PLANET_NCCL(api.AllGather(buf + offset_bytes[g->rank], buf + offset_bytes[0], (size_t)size_bytes[0], ncclChar, g->comm, stream));
    } else {                                                              // ragged shards: every rank broadcasts its own range
        PLANET_NCCL(api.GroupStart());
        for (int r = 0; r < g->world; r++)
            if (size_bytes[r] > 0) {
                ncclResult_t res = api.Broadcast(buf + offset_bytes[r], buf + offset_bytes[r], (size_t)size_bytes[r], ncclChar, r, g->comm, stream);
                if (res != ncclSuccess) { api.GroupEnd(); return check_nccl(res, "ncclBroadcast"); }
            }
        PLANET_NCCL(api.GroupEnd());
    }
    return 0;
}

int planet_gpu_gather_barrier(void *gather, void *stream_)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    return nccl_barrier(g, (cudaStream_t)stream_);
}

int planet_gpu_gather_error(void *gather)
{
    Gather *g = (Gather *)gather;
    if (!g) return set_error(PLANET_E_INVALID, "gather is NULL");
    uint32_t e = 0;
    PLANET_CUDA(cudaMemcpy(&e, g->flags(g->rank) + FLAG_ERROR, sizeof e, cudaMemcpyDeviceToHost));
    if (e) return set_error(PLANET_E_CUDA, "a peer did not signal within 2 s (gather wait timed out)");
    return 0;
}

} // extern "C"
