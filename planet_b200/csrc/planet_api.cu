// planet_api.cu -- the C-ABI of include/planet_gpu.h: lifecycle, validation, error
// reporting, the two reference-shaped entry points (main.cpp:107-111) with their pinned
// staging, the host-buffer batch path, and the FP32 peak probe.  No CPU compute path
// exists in this library: every entry point either launches a kernel or fails.
#include "planet_common.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace planet {

// launchers (k1_tessellate.cu, k2_heights.cu, k3_shade.cu)
int launch_height_maps(const planet_gpu_params *, const Quad *, int64_t, int, int, float *, cudaStream_t);
int launch_height_maps_gathered(const planet_gpu_params *, const Quad *, int64_t, int, int, float *, const PeerOut &, cudaStream_t);
int launch_heights_at(const planet_gpu_params *, const double *, int64_t, int, int, float *, cudaStream_t);
int launch_height_map_seam(const planet_gpu_params *, const Quad *, int, int, float *, cudaStream_t);
int launch_height_at_seam(const planet_gpu_params *, const double *, int, int, float *, cudaStream_t);
int launch_noise(const double *, int64_t, int, double, float, int, int, float *, cudaStream_t);
int launch_tessellate_uniform(const planet_gpu_params *, int, int64_t, int64_t, Quad *, uint32_t *, cudaStream_t);
int launch_quads_from_ids(const planet_gpu_params *, const uint64_t *, int64_t, Quad *, cudaStream_t);
int launch_index_stream_beside(const planet_gpu_params *, int64_t, uint32_t *, cudaStream_t);
int launch_patch_mesh(int, float *, uint32_t *, cudaStream_t);
int launch_shade(const planet_gpu_params *, const Quad *, int64_t, const double *, const float *,
                 const planet_gpu_texrect *, float, float *, float *, cudaStream_t);
int launch_select_lod(const planet_gpu_params *, const double *, int, Quad *, int64_t, int64_t *, cudaStream_t);
uint32_t host_strip_index(int, int);
uint64_t host_uniform_leaf_id(int64_t, int);
void release_lod_scratch();
void release_strip_cache();

// ---- state ------------------------------------------------------------------------------
static thread_local char t_error[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::mutex g_mutex;
static bool g_ready = false;
static int g_device = -1;
static planet_gpu_params g_params;          // used by the two legacy-shaped entry points
static bool g_params_set = false;

// staging for the synchronous host-pointer paths (grown on demand, freed at shutdown)
static struct Staging {
    void *h_pinned = nullptr; size_t h_cap = 0;
    void *d_in = nullptr;     size_t d_in_cap = 0;
    void *d_out = nullptr;    size_t d_out_cap = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;           // D2H of finished chunks overlaps the next chunk's kernel
    cudaStream_t up_stream = nullptr;             // H2D of the quads behind the first chunk's
    cudaEvent_t chunk_done[16] = {};
    cudaEvent_t quads_up = nullptr;
} g_stage;

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof t_error, fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    return set_error(PLANET_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int do_init(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(PLANET_E_NO_DEVICE, "no usable CUDA device (%s); this library has no CPU path",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count)
        return set_error(PLANET_E_INVALID, "device %d out of range [0, %d)", device, count);
    PLANET_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PLANET_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_error(PLANET_E_NO_DEVICE, "device %d (%s, sm_%d%d) is not Blackwell; kernels are built for sm_100a only",
                         device, prop.name, prop.major, prop.minor);
    g_device = device;
    g_ready = true;
    return 0;
}

bool ensure_init()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_ready) return true;
    int cur = 0;                              // lazily adopt the caller's current device
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = 0; }
    return do_init(cur) == 0;
}

int validate_params(const planet_gpu_params *p)
{
    if (!p) return set_error(PLANET_E_INVALID, "params is NULL");
    if (p->noise_kind < PLANET_NOISE_RIDGED || p->noise_kind > PLANET_NOISE_ZERO)
        return set_error(PLANET_E_INVALID, "noise_kind %d unknown", p->noise_kind);
    if (p->precision != PLANET_PRECISION_EXACT && p->precision != PLANET_PRECISION_FAST)
        return set_error(PLANET_E_INVALID, "precision %d unknown", p->precision);
    if (p->patch_verts < 2) return set_error(PLANET_E_INVALID, "patch_verts %d < 2", p->patch_verts);
    if (p->fixed_octaves > 64) return set_error(PLANET_E_INVALID, "fixed_octaves %d > 64", p->fixed_octaves);
    if (!(p->radius > 0.0)) return set_error(PLANET_E_INVALID, "radius must be positive");
    return 0;
}

HeightCfg make_cfg(const planet_gpu_params *p, int max_depth)
{
    HeightCfg c;
    c.kind = p->noise_kind;
    c.fixed_octaves = p->fixed_octaves;
    c.max_depth = max_depth;
    c.gain = p->gain;
    c.height_scale = p->height_scale;
    c.lacunarity = p->lacunarity;
    c.coord_scale = p->coord_scale;
    c.seed[0] = p->seed_offset[0]; c.seed[1] = p->seed_offset[1]; c.seed[2] = p->seed_offset[2];
    c.has_seed = (c.seed[0] != 0.0 || c.seed[1] != 0.0 || c.seed[2] != 0.0);
    return c;
}

static int check_height_args(const planet_gpu_params *p, int dim, int max_depth)
{
    int rc = validate_params(p);
    if (rc) return rc;
    if (dim <= 3) return set_error(PLANET_E_INVALID, "dim %d <= 3 (main.cpp:128 asserts dim > 3)", dim);
    if (dim > 32768) return set_error(PLANET_E_UNSUPPORTED, "dim %d > 32768 (a single map would exceed 4 GiB)", dim);
    if (p->fixed_octaves <= 0 && max_depth == 0)
        return set_error(PLANET_E_INVALID, "max_depth == 0 divides by zero at main.cpp:827");
    return 0;
}

static int grow(void **ptr, size_t *cap, size_t need, bool pinned)
{
    if (need <= *cap) return 0;
    if (*ptr) { if (pinned) cudaFreeHost(*ptr); else cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
    size_t want = need + need / 4 + 4096;
    PLANET_CUDA(pinned ? cudaMallocHost(ptr, want) : cudaMalloc(ptr, want));
    *cap = want;
    return 0;
}

static int stage_stream()
{
    if (!g_stage.stream) PLANET_CUDA(cudaStreamCreateWithFlags(&g_stage.stream, cudaStreamNonBlocking));
    if (!g_stage.copy_stream) {
        PLANET_CUDA(cudaStreamCreateWithFlags(&g_stage.copy_stream, cudaStreamNonBlocking));
        PLANET_CUDA(cudaStreamCreateWithFlags(&g_stage.up_stream, cudaStreamNonBlocking));
        for (auto &e : g_stage.chunk_done) PLANET_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PLANET_CUDA(cudaEventCreateWithFlags(&g_stage.quads_up, cudaEventDisableTiming));
    }
    return 0;
}

// ---- FP32 peak probe ----------------------------------------------------------------------
// 8 independent FFMA chains per thread, 1024 threads per SM, no memory traffic.
__global__ void __launch_bounds__(256) k_ffma_probe(float *sink, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) sink[0] = s;
}

} // namespace planet

using namespace planet;

extern "C" {

int planet_gpu_abi_version(void) { return PLANET_GPU_ABI_VERSION; }

void planet_gpu_default_params(planet_gpu_params *out)
{
    if (!out) return;
    memset(out, 0, sizeof *out);
    out->radius = 6371000.0;                 // main.cpp:821
    out->patch_verts = 30;                   // main.cpp:391
    out->noise_kind = PLANET_NOISE_RIDGED;   // main.cpp:829
    out->lacunarity = 2.0;
    out->gain = 0.55f;
    out->fixed_octaves = 0;                  // main.cpp:827
    out->coord_scale = 0.00001;              // main.cpp:828
    out->height_scale = 8848.0f;             // main.cpp:831
    out->precision = PLANET_PRECISION_EXACT;
}

int planet_gpu_init(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    return do_init(device);
}

void planet_gpu_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_stage.h_pinned) cudaFreeHost(g_stage.h_pinned);
    if (g_stage.d_in) cudaFree(g_stage.d_in);
    if (g_stage.d_out) cudaFree(g_stage.d_out);
    if (g_stage.stream) cudaStreamDestroy(g_stage.stream);
    if (g_stage.copy_stream) {
        cudaStreamDestroy(g_stage.copy_stream);
        if (g_stage.up_stream) cudaStreamDestroy(g_stage.up_stream);
        for (auto &e : g_stage.chunk_done) if (e) cudaEventDestroy(e);
        if (g_stage.quads_up) { cudaEventDestroy(g_stage.quads_up); g_stage.quads_up = nullptr; }
    }
    g_stage = Staging();
    release_lod_scratch();
    release_strip_cache();
    g_ready = false;
}

const char *planet_gpu_last_error(void) { return t_error; }

int planet_gpu_device_info(char *name, int name_cap, int *sm_count, int *clock_khz, int *fp32_lanes_per_sm)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    cudaDeviceProp prop;
    PLANET_CUDA(cudaGetDeviceProperties(&prop, g_device));
    if (name && name_cap > 0) { strncpy(name, prop.name, name_cap - 1); name[name_cap - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device);
    if (clock_khz) *clock_khz = khz;
    if (fp32_lanes_per_sm) *fp32_lanes_per_sm = 128;     // 4 SMSPs x 32 FP32 lanes on sm_100
    return 0;
}

int planet_gpu_set_params(const planet_gpu_params *p)
{
    int rc = validate_params(p);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    g_params = *p;
    g_params_set = true;
    return 0;
}

static const planet_gpu_params *legacy_params()
{
    if (!g_params_set) { planet_gpu_default_params(&g_params); g_params_set = true; }
    return &g_params;
}

float planet_gpu_get_height_at(const double *p, int depth, int max_depth)
{
    float result = NAN;
    int rc = PLANET_E_NO_DEVICE;
    if (ensure_init()) {
        std::lock_guard<std::mutex> lock(g_mutex);
        const planet_gpu_params *prm = legacy_params();
        rc = (prm->fixed_octaves <= 0 && max_depth == 0)
                 ? set_error(PLANET_E_INVALID, "max_depth == 0 divides by zero at main.cpp:827") : 0;
        if (!rc) rc = stage_stream();
        if (!rc) rc = grow(&g_stage.h_pinned, &g_stage.h_cap, 64, true);
        if (!rc) {
            // fast lane: the point is a kernel argument, the height lands in pinned host memory
            rc = launch_height_at_seam(prm, p, depth, max_depth, (float *)g_stage.h_pinned, g_stage.stream);
            if (rc == PLANET_E_UNSUPPORTED) {                            // FAST params: batch of one
                rc = grow(&g_stage.d_in, &g_stage.d_in_cap, 64, false);
                if (!rc) rc = grow(&g_stage.d_out, &g_stage.d_out_cap, 64, false);
                if (!rc) {
                    memcpy((char *)g_stage.h_pinned + 32, p, 24);
                    rc = check_cuda(cudaMemcpyAsync(g_stage.d_in, (char *)g_stage.h_pinned + 32, 24, cudaMemcpyHostToDevice, g_stage.stream), "H2D point");
                }
                if (!rc) rc = launch_heights_at(prm, (const double *)g_stage.d_in, 1, depth, max_depth,
                                                (float *)g_stage.d_out, g_stage.stream);
                if (!rc) rc = check_cuda(cudaMemcpyAsync(g_stage.h_pinned, g_stage.d_out, 4, cudaMemcpyDeviceToHost, g_stage.stream), "D2H height");
            }
        }
        if (!rc) rc = check_cuda(cudaStreamSynchronize(g_stage.stream), "sync");
        if (!rc) memcpy(&result, g_stage.h_pinned, 4);
    }
    if (rc) fprintf(stderr, "[ERROR] planet_gpu_get_height_at: %s\n", t_error);
    return result;
}

// shared body of the two host-buffer calls: H2D quads, K2 in chunks with the D2H of finished
// chunks overlapped on the copy stream, optionally K3 on the resident maps while the last
// chunks still drain
struct ShadeArgs { const double *cam_pos; float max_skirt; float *d_pos4, *d_nrm4; };

static int host_pipeline(const planet_gpu_params *p, const planet_gpu_quad *h_quads, int64_t nquads, int dim,
                         int max_depth, float *h_out, float *d_mirror, const ShadeArgs *shade)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = check_height_args(p, dim, max_depth);
    if (rc) return rc;
    if (nquads < 0 || (nquads > 0 && (!h_quads || !h_out))) return set_error(PLANET_E_INVALID, "NULL buffer");
    if (nquads == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    const size_t in_bytes = (size_t)nquads * sizeof(Quad);
    const size_t out_bytes = (size_t)nquads * dim * dim * sizeof(float);
    rc = stage_stream();
    if (!rc) rc = grow(&g_stage.d_in, &g_stage.d_in_cap, in_bytes, false);
    if (!rc && !d_mirror) rc = grow(&g_stage.d_out, &g_stage.d_out_cap, out_bytes, false);
    if (rc) return rc;
    float *d_out = d_mirror ? d_mirror : (float *)g_stage.d_out;
    // Caller memory may be pageable; cudaMemcpyAsync then stages through the driver's own
    // pinned buffers.  Callers that want full PCIe rate pass cudaHostRegister'ed memory.
    // The output (dim*dim*4 bytes per quad) dominates the PCIe traffic, so large batches run as
    // a pipeline: kernel on chunk c while chunk c-1 drains to the host on the copy stream.
    // Chunk boundaries.  What the caller waits for is the PCIe drain (67 MB at ~55 GB/s for a C2 batch against
    // 0.43 ms of K2), so the first bytes should start crossing as early as possible and the link must
    // never run dry afterwards: the first chunk is ONE wave of the height-map kernel (a 128-sample tile for
    // each of its resident warps, ~20 us), then 2, 4, 8, 8, ... waves -- every chunk's copy is at least as
    // long as the next chunk's kernel.  Whole waves keep every warp of a chunk equally loaded.  The quads
    // are uploaded in two copies, the first chunk's ahead of the rest.
    const size_t per_quad = (size_t)dim * dim;
    int device = 0, sms = 148;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t wave_quads = std::max<int64_t>(1, ((int64_t)sms * 24 * 128 + (int64_t)per_quad - 1) / (int64_t)per_quad);
    int64_t bounds[17] = { 0 };
    int chunks = 0;
    if (nquads < 256) {
        bounds[++chunks] = nquads;
    } else {
        int64_t waves = 1;
        while (bounds[chunks] < nquads && chunks < 16) {
            int64_t hi = std::min(nquads, bounds[chunks] + waves * wave_quads);
            if (chunks == 15 || nquads - hi < wave_quads / 2) hi = nquads;        // last slot, or only a crumb left
            bounds[++chunks] = hi;
            waves = std::min<int64_t>(waves * 2, 8);
        }
    }
    auto enqueue = [&]() -> int {
        const size_t first_in = (size_t)bounds[1] * sizeof(Quad);
        PLANET_CUDA(cudaMemcpyAsync(g_stage.d_in, h_quads, first_in, cudaMemcpyHostToDevice, g_stage.stream));
        if (in_bytes > first_in) {                                    // the other quads go up beside chunk 0, on their own stream
            PLANET_CUDA(cudaMemcpyAsync((char *)g_stage.d_in + first_in, (const char *)h_quads + first_in, in_bytes - first_in,
                                        cudaMemcpyHostToDevice, g_stage.up_stream));
            PLANET_CUDA(cudaEventRecord(g_stage.quads_up, g_stage.up_stream));
        }
        for (int c = 0; c < chunks; c++) {
            const int64_t lo = bounds[c], hi = bounds[c + 1];
            if (hi == lo) continue;
            if (c == 1) PLANET_CUDA(cudaStreamWaitEvent(g_stage.stream, g_stage.quads_up, 0));   // chunk 1's kernel is the first to need them
            int r = launch_height_maps(p, (const Quad *)g_stage.d_in + lo, hi - lo, dim, max_depth,
                                       d_out + lo * per_quad, g_stage.stream);
            if (r) return r;
            cudaStream_t cs = chunks > 1 ? g_stage.copy_stream : g_stage.stream;
            if (chunks > 1) {
                PLANET_CUDA(cudaEventRecord(g_stage.chunk_done[c], g_stage.stream));
                PLANET_CUDA(cudaStreamWaitEvent(cs, g_stage.chunk_done[c], 0));
            }
            PLANET_CUDA(cudaMemcpyAsync(h_out + lo * per_quad, d_out + lo * per_quad,
                                        (size_t)(hi - lo) * per_quad * sizeof(float), cudaMemcpyDeviceToHost, cs));
        }
        if (shade)                                                    // K3 reads the resident maps; the PCIe drain goes on beside it
            return launch_shade(p, (const Quad *)g_stage.d_in, nquads, shade->cam_pos, d_out, nullptr,
                                shade->max_skirt, shade->d_pos4, shade->d_nrm4, g_stage.stream);
        return 0;
    };
    rc = enqueue();
    // Drain both streams whatever happened: copies already queued write into the caller's
    // buffer, which must not be touched after this call returns.
    cudaError_t e1 = cudaStreamSynchronize(g_stage.stream);
    cudaError_t e2 = chunks > 1 ? cudaStreamSynchronize(g_stage.copy_stream) : cudaSuccess;
    if (g_stage.up_stream) cudaStreamSynchronize(g_stage.up_stream);
    if (rc) return rc;
    if (e1 != cudaSuccess) return check_cuda(e1, "height maps (host path)");
    return check_cuda(e2, "height maps D2H");
}

int planet_gpu_generate_height_maps_host(const planet_gpu_params *p, const planet_gpu_quad *h_quads,
                                         int64_t nquads, int dim, int max_depth, float *h_out,
                                         float *d_mirror)
{
    return host_pipeline(p, h_quads, nquads, dim, max_depth, h_out, d_mirror, nullptr);
}

int planet_gpu_terrain_host(const planet_gpu_params *p, const planet_gpu_quad *h_quads, int64_t nquads,
                            int max_depth, const double *cam_pos, float max_skirt, float *h_heights,
                            float *d_heights, float *d_pos4, float *d_nrm4)
{
    int rc = validate_params(p);
    if (rc) return rc;
    if (!cam_pos || !d_heights) return set_error(PLANET_E_INVALID, "NULL argument");
    if (max_skirt < 0.0f) max_skirt = planet_gpu_max_skirt_size(p->radius, p->patch_verts);
    const ShadeArgs shade = { cam_pos, max_skirt, d_pos4, d_nrm4 };
    return host_pipeline(p, h_quads, nquads, p->patch_verts + 2, max_depth, h_heights, d_heights, &shade);
}

void planet_gpu_generate_height_map(float *data, int dim, const void *quad, int max_depth)
{
    int rc = PLANET_E_NO_DEVICE;
    if (ensure_init()) {
        planet_gpu_params prm;
        { std::lock_guard<std::mutex> lock(g_mutex); prm = *legacy_params(); }
        rc = check_height_args(&prm, dim, max_depth);
        if (!rc && (!data || !quad)) rc = set_error(PLANET_E_INVALID, "NULL buffer");
        if (!rc) {
            // fast lane: the quad is a kernel argument, the map lands in pinned host memory
            std::lock_guard<std::mutex> lock(g_mutex);
            const size_t bytes = (size_t)dim * dim * sizeof(float);
            rc = stage_stream();
            if (!rc) rc = grow(&g_stage.h_pinned, &g_stage.h_cap, bytes, true);
            if (!rc) rc = launch_height_map_seam(&prm, (const Quad *)quad, dim, max_depth, (float *)g_stage.h_pinned, g_stage.stream);
            if (!rc) rc = check_cuda(cudaStreamSynchronize(g_stage.stream), "sync");
            if (!rc) memcpy(data, g_stage.h_pinned, bytes);
        }
        if (rc == PLANET_E_UNSUPPORTED)                                  // FAST params or a huge map: batch of one
            rc = planet_gpu_generate_height_maps_host(&prm, (const planet_gpu_quad *)quad, 1, dim, max_depth, data, nullptr);
    }
    if (rc) {
        fprintf(stderr, "[ERROR] planet_gpu_generate_height_map: %s\n", t_error);
        if (data && dim > 0) for (int i = 0; i < dim * dim; i++) data[i] = NAN;
    }
}

int planet_gpu_generate_height_maps(const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                    int64_t nquads, int dim, int max_depth, float *d_out, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = check_height_args(p, dim, max_depth);
    if (rc) return rc;
    if (nquads < 0 || (nquads > 0 && (!d_quads || !d_out))) return set_error(PLANET_E_INVALID, "NULL buffer");
    return launch_height_maps(p, (const Quad *)d_quads, nquads, dim, max_depth, d_out, (cudaStream_t)stream);
}

int planet_gpu_generate_height_maps_gathered(const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                             int64_t nquads, int dim, int max_depth, float *d_out,
                                             float *const *peer_out, int n_peers, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = check_height_args(p, dim, max_depth);
    if (rc) return rc;
    if (nquads < 0 || (nquads > 0 && (!d_quads || !d_out))) return set_error(PLANET_E_INVALID, "NULL buffer");
    if (n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_out))
        return set_error(PLANET_E_INVALID, "n_peers %d outside [0, 7]", n_peers);
    PeerOut peers = {};
    peers.n = n_peers;
    for (int r = 0; r < n_peers; r++) {
        if (!peer_out[r]) return set_error(PLANET_E_INVALID, "peer_out[%d] is NULL", r);
        peers.ptr[r] = peer_out[r];
    }
    return launch_height_maps_gathered(p, (const Quad *)d_quads, nquads, dim, max_depth, d_out, peers, (cudaStream_t)stream);
}

int planet_gpu_heights_at(const planet_gpu_params *p, const double *d_xyz, int64_t n, int depth,
                          int max_depth, float *d_out, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (p->fixed_octaves <= 0 && max_depth == 0)
        return set_error(PLANET_E_INVALID, "max_depth == 0 divides by zero at main.cpp:827");
    if (n < 0 || (n > 0 && (!d_xyz || !d_out))) return set_error(PLANET_E_INVALID, "NULL buffer");
    return launch_heights_at(p, d_xyz, n, depth, max_depth, d_out, (cudaStream_t)stream);
}

int planet_gpu_noise(const double *d_xyz, int64_t n, int kind, double lacunarity, float gain, int octaves,
                     int precision, float *d_out, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    if (octaves < 0 || octaves > 64) return set_error(PLANET_E_INVALID, "octaves %d outside [0, 64]", octaves);
    if (kind != PLANET_NOISE_FBM && kind != PLANET_NOISE_RIDGED && octaves != 0)
        return set_error(PLANET_E_INVALID, "kind %d is not fBm or ridged", kind);
    if (n < 0 || (n > 0 && (!d_xyz || !d_out))) return set_error(PLANET_E_INVALID, "NULL buffer");
    return launch_noise(d_xyz, n, kind, lacunarity, gain, octaves, precision, d_out, (cudaStream_t)stream);
}

int planet_gpu_tessellate_uniform(const planet_gpu_params *p, int depth, int64_t first, int64_t nquads,
                                  planet_gpu_quad *d_quads, uint32_t *d_indices, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (depth < 0 || depth > 27) return set_error(PLANET_E_INVALID, "depth %d outside [0, 27]", depth);
    const int64_t leaves = (int64_t)6 << (2 * depth);
    if (first < 0 || nquads < 0 || first + nquads > leaves)
        return set_error(PLANET_E_INVALID, "leaf range [%lld, %lld) outside [0, %lld)", (long long)first,
                         (long long)(first + nquads), (long long)leaves);
    if (nquads > 0 && !d_quads && !d_indices) return set_error(PLANET_E_INVALID, "both d_quads and d_indices are NULL");
    return launch_tessellate_uniform(p, depth, first, nquads, (Quad *)d_quads, d_indices, (cudaStream_t)stream);
}

int planet_gpu_merged_indices_beside(const planet_gpu_params *p, int64_t nquads, uint32_t *d_indices, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nquads < 0 || (nquads > 0 && !d_indices)) return set_error(PLANET_E_INVALID, "NULL buffer");
    if ((reinterpret_cast<uintptr_t>(d_indices) & 7) != 0) return set_error(PLANET_E_INVALID, "index buffer must be 8-byte aligned");
    return launch_index_stream_beside(p, nquads, d_indices, (cudaStream_t)stream);
}

int planet_gpu_quads_from_ids(const planet_gpu_params *p, const uint64_t *d_ids, int64_t n,
                              planet_gpu_quad *d_quads, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_ids || !d_quads))) return set_error(PLANET_E_INVALID, "NULL buffer");
    return launch_quads_from_ids(p, d_ids, n, (Quad *)d_quads, (cudaStream_t)stream);
}

uint32_t planet_gpu_strip_index(int k, int n) { return planet::host_strip_index(k, n); }
uint64_t planet_gpu_uniform_leaf_id(int64_t leaf, int depth) { return planet::host_uniform_leaf_id(leaf, depth); }
int planet_gpu_patch_vertex_count(int n) { return n * n + 4 * n; }
int planet_gpu_patch_index_count(int n) { return 2 * n * n + 8 * n - 4; }

int planet_gpu_patch_mesh(int patch_verts, float *d_vertices, uint32_t *d_indices, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    if (patch_verts < 2) return set_error(PLANET_E_INVALID, "patch_verts %d < 2", patch_verts);
    return launch_patch_mesh(patch_verts, d_vertices, d_indices, (cudaStream_t)stream);
}

// main.cpp:497 / :500 -- host scalars of InitPlanet (fp64 <cmath>, exactly as the reference)
int planet_gpu_max_lod(double radius, int n)
{
    const double pi = 3.1415926535897932384626433832795;          // math.h:7
    return (int)(std::log2(2.0 * pi * radius / (n - 1)) - 2);
}
float planet_gpu_max_skirt_size(double radius, int n)
{
    const double pi = 3.1415926535897932384626433832795;
    return (float)((2 * pi * radius) / (4 * (n - 1)) * 0.00001 * 8 * 8848.0);
}

int planet_gpu_shade(const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t nquads,
                     const double *cam_pos, const float *d_heights, float max_skirt, float *d_pos4,
                     float *d_nrm4, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nquads < 0 || (nquads > 0 && (!d_quads || !d_heights || !cam_pos)))
        return set_error(PLANET_E_INVALID, "NULL buffer");
    if (max_skirt < 0.0f) max_skirt = planet_gpu_max_skirt_size(p->radius, p->patch_verts);
    return launch_shade(p, (const Quad *)d_quads, nquads, cam_pos, d_heights, nullptr, max_skirt, d_pos4, d_nrm4,
                        (cudaStream_t)stream);
}

int planet_gpu_shade_cached(const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t nquads,
                            const double *cam_pos, const float *d_pool, const planet_gpu_texrect *d_rects,
                            float max_skirt, float *d_pos4, float *d_nrm4, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nquads < 0 || (nquads > 0 && (!d_quads || !d_pool || !d_rects || !cam_pos)))
        return set_error(PLANET_E_INVALID, "NULL buffer");
    if (max_skirt < 0.0f) max_skirt = planet_gpu_max_skirt_size(p->radius, p->patch_verts);
    return launch_shade(p, (const Quad *)d_quads, nquads, cam_pos, d_pool, d_rects, max_skirt, d_pos4, d_nrm4,
                        (cudaStream_t)stream);
}

int planet_gpu_select_lod(const planet_gpu_params *p, const double *cam_pos, int max_lod,
                          planet_gpu_quad *d_quads, int64_t capacity, int64_t *count, void *stream)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (!cam_pos || !d_quads || !count) return set_error(PLANET_E_INVALID, "NULL argument");
    if (max_lod < 0 || max_lod > 27) return set_error(PLANET_E_INVALID, "max_lod %d outside [0, 27]", max_lod);
    return launch_select_lod(p, cam_pos, max_lod, (Quad *)d_quads, capacity, count, (cudaStream_t)stream);
}

int planet_gpu_measure_fp32_peak(double ms, double *tflops, double *elapsed_ms)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device);
    float *sink = nullptr;
    PLANET_CUDA(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    PLANET_CUDA(cudaEventCreate(&e0));
    PLANET_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8, block = 256;            // 2048 threads per SM
    int iters = 2000;
    float t = 0.f;
    // calibrate, then run for about `ms`
    for (int pass = 0; pass < 2; pass++) {
        k_ffma_probe<<<grid, block>>>(sink, 200, 1.0001f, 0.0001f);      // warm-up
        cudaEventRecord(e0);
        k_ffma_probe<<<grid, block>>>(sink, iters, 1.0001f, 0.0001f);
        cudaEventRecord(e1);
        count_launch(2);
        PLANET_CUDA(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&t, e0, e1);
        if (pass == 0 && t > 0.f) iters = (int)fmin(2.0e6, fmax(200.0, iters * ms / t));
    }
    double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * block;
    if (tflops) *tflops = flops / (t * 1e-3) / 1e12;
    if (elapsed_ms) *elapsed_ms = t;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 0;
}

int64_t planet_gpu_launch_count(void) { return g_launches.load(); }

} // extern "C"
