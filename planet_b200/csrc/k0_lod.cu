// k0_lod.cu -- camera-driven LOD selection: the caller of the hot path.
//
// Replaces the reference's serial recursion
//   RenderPlanet: 6 root quads -> ProcessQuad(root, cam, max_lod)          main.cpp:604-624
//   ProcessQuad:  5 displaced sample points, split test, 4 children        main.cpp:537-598
// with a level-synchronous frontier expansion inside ONE cooperative launch (a grid barrier
// between quadtree levels): a warp owns one frontier quad, its lanes evaluate GetHeightAt(p, 0, 1)
// for the four corners and the centre with the fractal's octaves spread across lanes (EXACT
// arithmetic, accumulated in the reference's order, so every split decision is the
// reference's), the warp votes, and lane 0 appends either the leaf or the four children
// (midpoint rule in the reference's fp64 operation order -> bit-identical corners).  Leaves are
// finally sorted by their depth-first key, which reproduces the order in which the recursion
// appends them to planet.quads.
#include "planet_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include <cooperative_groups.h>
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace planet {

namespace lod {

// depth-first position of a quad: face, then the child index of every level, most significant
// first.  Leaves are prefix-free, so zero padding keeps the recursion's order.
__host__ __device__ inline uint64_t dfs_key(uint64_t id)
{
    const int depth = (int)quad_depth(id);
    uint64_t key = quad_root(id) << 54;
    for (int l = 1; l <= depth; l++) key |= ((id >> (2 * (l - 1))) & 3) << (54 - 2 * l);
    return key;
}

__device__ __forceinline__ double length_sq(d3 v)                     // vec3.h:46-47
{
    return __dadd_rn(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)), __dmul_rn(v.z, v.z));
}
__device__ __forceinline__ double shfl_d(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ d3 shfl_d3(d3 v, int src) { return { shfl_d(v.x, src), shfl_d(v.y, src), shfl_d(v.z, src) }; }

// GetHeightAt(p, 0, 1) for the quad's 5 sample points with the fractal's octaves spread over
// the lanes of one warp: lane = 6*point + sub evaluates octaves sub, sub+6, ...; the point's
// leader lane (sub == 0) then accumulates them in the reference's order (main.cpp:699-704 /
// 716-731), so the float sum is bit-identical to the sequential loop.
__device__ __forceinline__ float height_split_octaves(const unsigned char *s_perm, const float *s_grad,
                                                      const HeightCfg &cfg, d3 p, int lane)
{
    if (cfg.kind == PLANET_NOISE_ZERO) return 0.0f;
    const int octaves = octaves_for(cfg.fixed_octaves, 0, cfg.max_depth);
    const int pt = min(lane / 6, 4), sub = lane - pt * 6;            // lanes 30, 31 shadow point 4
    p = exact::mul(p, cfg.coord_scale);                               // main.cpp:828
    if (cfg.has_seed) { p.x = __dadd_rn(p.x, cfg.seed[0]); p.y = __dadd_rn(p.y, cfg.seed[1]); p.z = __dadd_rn(p.z, cfg.seed[2]); }
    float amplitude = 1.0f, weight = 1.0f, value = 0.0f;              // leader-lane state
    double frequency = 1.0;
    for (int k0 = 0; k0 < octaves; k0 += 6) {
        // this lane's octave of the round: frequency = lacunarity^k by the reference's repeated product
        const int k = k0 + sub;
        double f = frequency;
        for (int j = 0; j < sub; j++) f = __dmul_rn(f, cfg.lacunarity);
        float n = 0.0f;
        if (k < octaves && sub < 6)
            n = exact::noise3(s_perm, s_grad, __dmul_rn(p.x, f), __dmul_rn(p.y, f), __dmul_rn(p.z, f));
        for (int j = 0; j < 6; j++) {                                 // leader accumulates octaves k0 .. k0+5 in order
            float nj = __shfl_sync(0xffffffffu, n, pt * 6 + j);
            if (k0 + j < octaves) {
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
    }
    return __fmul_rn(value, cfg.height_scale);                        // main.cpp:831 (valid on every lane of the point)
}

// One frontier quad per WARP.  Leaves go to leaves[*n_leaves], children to next[*n_next].
__device__ __forceinline__ void lod_group(const unsigned char *s_perm, const float *s_grad, const Quad *frontier,
                                          int qi, int n, int lod, int max_lod, double radius, d3 cam,
                                          const HeightCfg &cfg, Quad *leaves, uint64_t *keys, Quad *next,
                                          int capacity, int *n_leaves, int *n_next)
{
    const int lane = threadIdx.x & 31;
    const int pt = min(lane / 6, 4);                                  // sample point of this lane: corners 0..3, centre
    const bool live = qi < n;
    Quad q;
    if (live) q = frontier[qi];
    else { q.id = 0; for (int j = 0; j < 4; j++) q.p[j] = { 1.0, 1.0, 1.0 }; }

    bool split = false;
    d3 mid = { 0, 0, 0 };
    if (lod != 0) {                                                   // main.cpp:539 (uniform per level)
        // main.cpp:546-547
        d3 sum = exact::add(exact::add(exact::add(q.p[0], q.p[1]), q.p[2]), q.p[3]);
        d3 mid_n = exact::normalize(sum);
        mid = exact::mul(mid_n, radius);
        // main.cpp:549-556: four displaced corners and the displaced centre
        d3 base = pt == 0 ? q.p[0] : pt == 1 ? q.p[1] : pt == 2 ? q.p[2] : pt == 3 ? q.p[3] : mid;
        d3 dir = pt < 4 ? exact::normalize(base) : mid_n;
        float h = height_split_octaves(s_perm, s_grad, cfg, base, lane);   // GetHeightAt(p, 0, 1)
        d3 p = exact::add(base, exact::mul(dir, (double)h));
        // main.cpp:560-562: d = (|p3-p0|^2 + |p2-p1|^2) / (1 + 2.5*lod/max_lod); point i lives on lane 6i
        d3 p0 = shfl_d3(p, 0), p1 = shfl_d3(p, 6), p2 = shfl_d3(p, 12), p3 = shfl_d3(p, 18);
        double denom = __dadd_rn(1.0, __ddiv_rn(__dmul_rn(2.5, (double)lod), (double)max_lod));
        double d = __ddiv_rn(__dadd_rn(length_sq(exact::sub(p3, p0)), length_sq(exact::sub(p2, p1))), denom);
        // main.cpp:564-571
        bool near = __dmul_rn(length_sq(exact::sub(p, cam)), 2.0) < d;
        split = __any_sync(0xffffffffu, near);
    }
    if (!live) return;                                                // (warp-uniform: one frontier quad per warp)

    if (!split) {                                                     // main.cpp:541-543, 573-577
        if (lane == 0) {
            int idx = atomicAdd(n_leaves, 1);
            if (idx < capacity) { leaves[idx] = q; keys[idx] = dfs_key(q.id); }
        }
        return;
    }
    // main.cpp:581-592: 3x3 grid p0, V(0,1), p1, V(0,2), mid, V(1,3), p2, V(2,3), p3.  The four edge
    // midpoints are four independent fp64 normalisations (a division chain each): lanes 0..3 take one
    // each instead of lane 0 running all four, and lane c then writes child c.
    int idx = 0;
    if (lane == 0) idx = atomicAdd(n_next, 4);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx + 3 >= capacity) return;
    const int e = lane & 3;                                           // edge: V(0,1) V(0,2) V(1,3) V(2,3)
    const d3 ea = e < 2 ? q.p[0] : e == 2 ? q.p[1] : q.p[2];
    const d3 eb = e == 0 ? q.p[1] : e == 1 ? q.p[2] : q.p[3];
    const d3 ev = exact::mul(exact::normalize(exact::add(ea, eb)), radius);
    const d3 v01 = shfl_d3(ev, 0), v02 = shfl_d3(ev, 1), v13 = shfl_d3(ev, 2), v23 = shfl_d3(ev, 3);
    if (lane < 4) {
        Quad c;
        c.id = make_child_id(q.id, (uint64_t)lane);
        if (lane == 0)      { c.p[0] = q.p[0]; c.p[1] = v01; c.p[2] = v02; c.p[3] = mid; }
        else if (lane == 1) { c.p[0] = v01; c.p[1] = q.p[1]; c.p[2] = mid; c.p[3] = v13; }
        else if (lane == 2) { c.p[0] = v02; c.p[1] = mid; c.p[2] = q.p[2]; c.p[3] = v23; }
        else                { c.p[0] = mid; c.p[1] = v13; c.p[2] = v23; c.p[3] = q.p[3]; }
        next[idx + lane] = c;
    }
}

// All levels in ONE cooperative launch: a grid-wide barrier separates the levels.
// counters[0] = leaves; counters[1 + level] = size of the frontier entering that level.
__global__ void __launch_bounds__(256)
k_lod_all_levels(Quad *fa, Quad *fb, int max_lod, double radius, double cam_x, double cam_y, double cam_z,
                 HeightCfg cfg, Quad *leaves, uint64_t *keys, int capacity, int *counters)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_perm[i] = g_perm[i];
    for (int i = threadIdx.x; i < 48; i += blockDim.x) s_grad[i] = (&g_grad[0][0])[i];
    __syncthreads();
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const d3 cam = { cam_x, cam_y, cam_z };
    const int groups = (gridDim.x * blockDim.x) >> 5;                 // one warp per frontier quad
    int level = 0;
    for (int lod = max_lod; lod >= 0; lod--, level++) {
        const int n = min(((volatile int *)counters)[1 + level], capacity);   // written before the last barrier
        if (n == 0) break;                                            // uniform across the grid
        for (int base = 0; base < n; base += groups) {                // whole warps iterate together
            int qi = base + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
            lod_group(s_perm, s_grad, fa, qi, n, lod, max_lod, radius, cam, cfg, leaves, keys, fb, capacity,
                      &counters[0], &counters[2 + level]);
        }
        grid.sync();
        Quad *t = fa; fa = fb; fb = t;
    }
}

__global__ void __launch_bounds__(256)
k_lod_level(const Quad *__restrict__ frontier, int n, int lod, int max_lod, double radius,
            double cam_x, double cam_y, double cam_z, HeightCfg cfg, Quad *__restrict__ leaves,
            uint64_t *__restrict__ keys, Quad *__restrict__ next, int capacity, int *counters)
{
    __shared__ unsigned char s_perm[256];
    __shared__ float s_grad[48];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_perm[i] = g_perm[i];
    for (int i = threadIdx.x; i < 48; i += blockDim.x) s_grad[i] = (&g_grad[0][0])[i];
    __syncthreads();
    const d3 cam = { cam_x, cam_y, cam_z };
    lod_group(s_perm, s_grad, frontier, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n, lod, max_lod, radius, cam, cfg,
              leaves, keys, next, capacity, &counters[0], &counters[1]);
}

// Up to 4 096 leaves (the usual frame has a few hundred): sort by depth-first key and gather in
// ONE CTA, with the count read on the device.  counters[31] = 1 tells the host to fall back to
// the device-wide sort.
constexpr int BS_THREADS = 1024, BS_ITEMS = 4;
__global__ void __launch_bounds__(BS_THREADS)
k_block_sort_gather(const Quad *__restrict__ leaves, const uint64_t *__restrict__ keys, int *counters,
                    int capacity, int max_items, Quad *__restrict__ out)
{
    typedef cub::BlockRadixSort<uint64_t, BS_THREADS, BS_ITEMS, int> Sort;
    __shared__ typename Sort::TempStorage temp;
    const int n = counters[0];
    if (n > max_items || n > capacity) { if (threadIdx.x == 0) counters[31] = 1; return; }
    if (n <= BS_THREADS && max_items >= BS_THREADS) {
        // The usual frame (a few hundred leaves): every thread ranks its own key against all others
        // straight out of shared memory -- n broadcast reads and compares, ~1 us -- instead of the 15
        // four-bit passes of a 57-bit radix sort over 4 096 slots (46 us, a third of the frame's LOD
        // selection).  Keys are unique (no leaf is another leaf's ancestor), so ranks are a permutation.
        uint64_t *sk = reinterpret_cast<uint64_t *>(&temp);
        const uint64_t mine = (int)threadIdx.x < n ? keys[threadIdx.x] : ~0ull;
        sk[threadIdx.x] = mine;
        __syncthreads();
        if ((int)threadIdx.x < n) {
            int rank = 0;
            for (int j = 0; j < n; j++) rank += sk[j] < mine;
            out[rank] = leaves[threadIdx.x];
        }
        return;
    }
    uint64_t k[BS_ITEMS];
    int v[BS_ITEMS];
#pragma unroll
    for (int j = 0; j < BS_ITEMS; j++) {
        const int i = threadIdx.x * BS_ITEMS + j;
        k[j] = i < n ? keys[i] : ~0ull;
        v[j] = i;
    }
    Sort(temp).Sort(k, v, 0, 57);
#pragma unroll
    for (int j = 0; j < BS_ITEMS; j++) {
        const int i = threadIdx.x * BS_ITEMS + j;                     // blocked arrangement: rank of item j
        if (i < n) out[i] = leaves[v[j]];
    }
}

__global__ void k_pad_keys(uint64_t *keys, int *order, const int *counters, int capacity)
{
    // entries past the leaf count sort to the end
    const int n = min(counters[0], capacity);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += gridDim.x * blockDim.x) {
        if (i >= n) keys[i] = ~0ull;
        order[i] = i;
    }
}

__global__ void k_gather_sorted(const Quad *__restrict__ leaves, const int *__restrict__ order, int n,
                                const int *counters, Quad *__restrict__ out)
{
    if (counters) n = min(n, counters[0]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = leaves[order[i]];
}

__global__ void k_iota(int *v, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

} // namespace lod

int launch_tessellate_uniform(const planet_gpu_params *, int, int64_t, int64_t, Quad *, uint32_t *, cudaStream_t);

// scratch owned by the library (grown on demand, reused across calls; calls are serialised)
static std::mutex g_lod_mutex;
static struct LodScratch {
    void *buf = nullptr; size_t cap = 0;
} g_lod;

void release_lod_scratch()                                            // planet_gpu_shutdown
{
    std::lock_guard<std::mutex> lock(g_lod_mutex);
    if (g_lod.buf) cudaFree(g_lod.buf);
    g_lod.buf = nullptr; g_lod.cap = 0;
}

int launch_select_lod(const planet_gpu_params *p, const double *cam, int max_lod, Quad *d_out,
                      int64_t capacity, int64_t *count, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_lod_mutex);
    const int cap = (int)std::min<int64_t>(capacity, 1 << 24);
    if (cap < 6) return set_error(PLANET_E_INVALID, "capacity %lld < 6 root quads", (long long)capacity);
    // layout: frontier A | frontier B | leaves | keys | keys_sorted | order | order_sorted | counters | cub temp
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (int *)nullptr, (int *)nullptr, cap, 0, 64, stream);
    const size_t quads_bytes = (size_t)cap * sizeof(Quad);
    const size_t need = 3 * quads_bytes + 2 * (size_t)cap * 8 + 2 * (size_t)cap * 4 + 256 + cub_bytes + 1024;
    if (need > g_lod.cap) {
        if (g_lod.buf) cudaFree(g_lod.buf);
        g_lod.buf = nullptr; g_lod.cap = 0;
        PLANET_CUDA(cudaMalloc(&g_lod.buf, need));
        g_lod.cap = need;
    }
    char *base = (char *)g_lod.buf;
    Quad *fa = (Quad *)base, *fb = (Quad *)(base + quads_bytes), *leaves = (Quad *)(base + 2 * quads_bytes);
    uint64_t *keys = (uint64_t *)(base + 3 * quads_bytes), *keys_sorted = keys + cap;
    int *order = (int *)(keys_sorted + cap), *order_sorted = order + cap;
    int *counters = (int *)(((uintptr_t)(order_sorted + cap) + 127) & ~(uintptr_t)127);
    void *cub_temp = (void *)(counters + 64);

    HeightCfg cfg = make_cfg(p, 1);                                   // GetHeightAt(p, 0, 1): max_depth = 1
    PLANET_CUDA(cudaMemsetAsync(counters, 0, 32 * sizeof(int), stream));
    // main.cpp:604-624: the six root quads (depth-0 leaves of the uniform tree)
    int rc = launch_tessellate_uniform(p, 0, 0, 6, fa, nullptr, stream);
    if (rc) return rc;

    // Fast path: every level in one cooperative launch (a grid barrier between levels) and a
    // fixed-size sort, so the host synchronises once, at the end, to read the leaf count.
    int dev = 0, coop = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (coop && cap <= (1 << 16) && max_lod + 3 <= 31 && !getenv("PLANET_K0_LEVEL_LAUNCHES")) {
        const int six = 6;
        PLANET_CUDA(cudaMemcpyAsync(counters + 1, &six, sizeof(int), cudaMemcpyHostToDevice, stream));
        int grid = sms;                                               // 148 CTAs x 8 warps = 1 184 quads per pass
        double cx = cam[0], cy = cam[1], cz = cam[2], radius = p->radius;
        int ml = max_lod, capi = cap;
        void *args[] = { &fa, &fb, &ml, &radius, &cx, &cy, &cz, &cfg, &leaves, &keys, &capi, &counters };
        // (One CTA walking the levels with block barriers instead of grid barriers was measured too: 93-113 us
        //  per frame's selection at 512 / 768 / 1 024 threads against 91 us for this kernel -- a level costs the
        //  latency of its fp64 chains and 6-octave samples, not its barrier, and one CTA needs more rounds.)
        PLANET_CUDA(cudaLaunchCooperativeKernel((void *)lod::k_lod_all_levels, dim3(grid), dim3(256), args, 0, stream));
        int block_sort_max = lod::BS_THREADS * lod::BS_ITEMS;
        if (const char *e = getenv("PLANET_K0_BLOCK_SORT_MAX")) block_sort_max = std::min(block_sort_max, atoi(e));   // test knob
        lod::k_block_sort_gather<<<1, lod::BS_THREADS, 0, stream>>>(leaves, keys, counters, cap, block_sort_max, d_out);
        count_launch(2);
        PLANET_CUDA(cudaGetLastError());
        int h[32];
        PLANET_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, stream));
        PLANET_CUDA(cudaStreamSynchronize(stream));
        int64_t worst = h[0];
        for (int l = 1; l < 31; l++) worst = std::max<int64_t>(worst, h[l]);
        if (count) *count = h[0];
        if (worst > cap) {
            if (count) *count = worst;
            return set_error(PLANET_E_INVALID, "LOD selection needs more than the %d quads of capacity given", cap);
        }
        if (h[31]) {                                                  // more leaves than one CTA sorts: device-wide sort
            lod::k_pad_keys<<<64, 256, 0, stream>>>(keys, order, counters, cap);
            PLANET_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, keys, keys_sorted, order, order_sorted,
                                                        cap, 0, 64, stream));
            lod::k_gather_sorted<<<64, 256, 0, stream>>>(leaves, order_sorted, cap, counters, d_out);
            count_launch(3);
            PLANET_CUDA(cudaGetLastError());
        }
        return 0;
    }

    // General path: one launch and one host synchronisation per level.
    int n = 6, total_leaves = 0;
    for (int lod = max_lod; lod >= 0 && n > 0; lod--) {
        int groups_per_block = 256 / 32;
        int grid = (n + groups_per_block - 1) / groups_per_block;
        lod::k_lod_level<<<grid, 256, 0, stream>>>(fa, n, lod, max_lod, p->radius, cam[0], cam[1], cam[2], cfg,
                                                   leaves, keys, fb, cap, counters);
        count_launch();
        PLANET_CUDA(cudaGetLastError());
        int h[2];
        PLANET_CUDA(cudaMemcpyAsync(h, counters, 8, cudaMemcpyDeviceToHost, stream));
        PLANET_CUDA(cudaStreamSynchronize(stream));
        total_leaves = h[0];
        n = h[1];
        if (total_leaves > cap || n > cap) {
            if (count) *count = (int64_t)total_leaves + n;
            return set_error(PLANET_E_INVALID, "LOD selection needs more than the %d quads of capacity given", cap);
        }
        PLANET_CUDA(cudaMemsetAsync(counters + 1, 0, 4, stream));
        std::swap(fa, fb);
    }
    if (count) *count = total_leaves;
    if (total_leaves == 0) return 0;
    lod::k_iota<<<(total_leaves + 255) / 256, 256, 0, stream>>>(order, total_leaves);
    PLANET_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, keys, keys_sorted, order, order_sorted,
                                                total_leaves, 0, 57, stream));
    lod::k_gather_sorted<<<(total_leaves + 255) / 256, 256, 0, stream>>>(leaves, order_sorted, total_leaves, nullptr, d_out);
    count_launch(3);
    return check_cuda(cudaGetLastError(), "LOD gather launch");
}

} // namespace planet
