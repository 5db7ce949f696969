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

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

namespace planet {

int launch_height_maps_gathered(const planet_gpu_params *, const Quad *, int64_t, int, int, float *, const PeerOut &, cudaStream_t);
int launch_gather_wait(const uint32_t *, uint32_t, int, int, uint32_t *, cudaStream_t);
bool height_maps_push_in_bulk(const planet_gpu_params *, int64_t, int, int, const float *, const PeerOut &);
bool shade_can_push(const planet_gpu_params *, const float *);
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
    int k3_every = 0;                   // > 0: the shade kernel pushes every k3_every-th quad (planet_gpu_gather_set_shade_share)
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
        for (int k = 0; k < world - 1; k++)
            if (check_cuda(cudaStreamCreateWithFlags(&g->push_stream[k], cudaStreamNonBlocking), "push stream")) return fail("streams");
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
    // part of the pushing may be left to the shade kernel (planet_gpu_gather_shade must then follow)
    g->shade_pending = false;
    peers.quad0 = first_quad;
    if (g->world > 1 && g->k3_every > 1 && nquads > 0 && dim == p->patch_verts + 2 &&
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
    if (!g || every < 0 || every == 1) return set_error(PLANET_E_INVALID, "gather_set_shade_share(%d): 0 (off) or >= 2", every);
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
    for (int k = 0; k < g->world - 1; k++) {
        PeerFlags one = {};
        one.n = 1;
        one.ptr[0] = g->flags((g->rank + 1 + k) % g->world);
        k_gather_signal<<<1, 32, 0, g->push_stream[k]>>>(one, FLAG_ARRIVE + g->rank, g->step);
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "gather publish launch");
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
