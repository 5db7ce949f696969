// k0_lod.cu -- camera-driven LOD selection: the caller of the hot path.
//
// Replaces the reference's serial recursion
//   RenderPlanet: 6 root quads -> ProcessQuad(root, cam, max_lod)          main.cpp:604-624
//   ProcessQuad:  5 displaced sample points, split test, 4 children        main.cpp:537-598
// with a level-synchronous frontier expansion: one launch per quadtree level; an 8-lane group
// owns one frontier quad, lanes 0..4 evaluate GetHeightAt(p, 0, 1) for its four corners and its
// centre in parallel (the functor in EXACT arithmetic, so every split decision is the
// reference's), the group votes, and lane 0 appends either the leaf or the four children
// (midpoint rule in the reference's fp64 operation order -> bit-identical corners).  Leaves are
// finally sorted by their depth-first key, which reproduces the order in which the recursion
// appends them to planet.quads.
#include "planet_common.cuh"

#include <algorithm>
#include <mutex>

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
    return __shfl_sync(0xffffffffu, v, src, 8);
}
__device__ __forceinline__ d3 shfl_d3(d3 v, int src) { return { shfl_d(v.x, src), shfl_d(v.y, src), shfl_d(v.z, src) }; }

// counters[0] = leaves so far, counters[1] = size of the next frontier
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

    const int g = threadIdx.x & 7;                                    // lane inside the quad's group
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = qi < n;
    Quad q;
    if (live) q = frontier[qi];
    else { q.id = 0; for (int j = 0; j < 4; j++) q.p[j] = { 1.0, 1.0, 1.0 }; }

    bool split = false;
    d3 mid = { 0, 0, 0 };
    if (lod != 0) {                                                   // main.cpp:539 (uniform per launch)
        // main.cpp:546-547
        d3 sum = exact::add(exact::add(exact::add(q.p[0], q.p[1]), q.p[2]), q.p[3]);
        d3 mid_n = exact::normalize(sum);
        mid = exact::mul(mid_n, radius);
        // main.cpp:549-556: lanes 0..3 displace a corner, lane 4 the centre
        d3 base = g < 4 ? q.p[g & 3] : mid;
        d3 dir = g < 4 ? exact::normalize(base) : mid_n;
        float h = exact::height(s_perm, s_grad, cfg, base, 0);        // GetHeightAt(p, 0, 1)
        d3 p = exact::add(base, exact::mul(dir, (double)h));
        // main.cpp:560-562: d = (|p3-p0|^2 + |p2-p1|^2) / (1 + 2.5*lod/max_lod)
        d3 p3 = shfl_d3(p, 3), p2 = shfl_d3(p, 2);
        double part = g == 0 ? length_sq(exact::sub(p3, p)) : length_sq(exact::sub(p2, p));   // lane 0 / lane 1
        double b = shfl_d(part, 1);
        double denom = __dadd_rn(1.0, __ddiv_rn(__dmul_rn(2.5, (double)lod), (double)max_lod));
        double d = shfl_d(__ddiv_rn(__dadd_rn(part, b), denom), 0);
        // main.cpp:564-571
        d3 cam = { cam_x, cam_y, cam_z };
        bool near = g < 5 && __dmul_rn(length_sq(exact::sub(p, cam)), 2.0) < d;
        unsigned vote = __ballot_sync(0xffffffffu, near);
        split = ((vote >> (threadIdx.x & 24)) & 0xffu) != 0;         // any lane of this 8-lane group
    }
    if (!live || g != 0) return;

    if (!split) {                                                     // main.cpp:541-543, 573-577
        int idx = atomicAdd(&counters[0], 1);
        if (idx < capacity) { leaves[idx] = q; keys[idx] = dfs_key(q.id); }
        return;
    }
    // main.cpp:581-592: 3x3 grid p0, V(0,1), p1, V(0,2), mid, V(1,3), p2, V(2,3), p3
    int idx = atomicAdd(&counters[1], 4);
    if (idx + 3 >= capacity) return;
    d3 v01 = exact::mul(exact::normalize(exact::add(q.p[0], q.p[1])), radius);
    d3 v02 = exact::mul(exact::normalize(exact::add(q.p[0], q.p[2])), radius);
    d3 v13 = exact::mul(exact::normalize(exact::add(q.p[1], q.p[3])), radius);
    d3 v23 = exact::mul(exact::normalize(exact::add(q.p[2], q.p[3])), radius);
    Quad c;
    c.p[0] = q.p[0]; c.p[1] = v01; c.p[2] = v02; c.p[3] = mid; c.id = make_child_id(q.id, 0); next[idx + 0] = c;
    c.p[0] = v01; c.p[1] = q.p[1]; c.p[2] = mid; c.p[3] = v13; c.id = make_child_id(q.id, 1); next[idx + 1] = c;
    c.p[0] = v02; c.p[1] = mid; c.p[2] = q.p[2]; c.p[3] = v23; c.id = make_child_id(q.id, 2); next[idx + 2] = c;
    c.p[0] = mid; c.p[1] = v13; c.p[2] = v23; c.p[3] = q.p[3]; c.id = make_child_id(q.id, 3); next[idx + 3] = c;
}

__global__ void k_gather_sorted(const Quad *__restrict__ leaves, const int *__restrict__ order, int n,
                                Quad *__restrict__ out)
{
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

int launch_select_lod(const planet_gpu_params *p, const double *cam, int max_lod, Quad *d_out,
                      int64_t capacity, int64_t *count, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_lod_mutex);
    const int cap = (int)std::min<int64_t>(capacity, 1 << 24);
    if (cap < 6) return set_error(PLANET_E_INVALID, "capacity %lld < 6 root quads", (long long)capacity);
    // layout: frontier A | frontier B | leaves | keys | keys_sorted | order | order_sorted | counters | cub temp
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (int *)nullptr, (int *)nullptr, cap, 0, 57, stream);
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
    void *cub_temp = (void *)(counters + 32);

    HeightCfg cfg = make_cfg(p, 1);                                   // GetHeightAt(p, 0, 1): max_depth = 1
    PLANET_CUDA(cudaMemsetAsync(counters, 0, 8, stream));
    // main.cpp:604-624: the six root quads (depth-0 leaves of the uniform tree)
    int rc = launch_tessellate_uniform(p, 0, 0, 6, fa, nullptr, stream);
    if (rc) return rc;
    int n = 6, total_leaves = 0;
    for (int lod = max_lod; lod >= 0 && n > 0; lod--) {
        int groups_per_block = 256 / 8;
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
    lod::k_gather_sorted<<<(total_leaves + 255) / 256, 256, 0, stream>>>(leaves, order_sorted, total_leaves, d_out);
    count_launch(3);
    return check_cuda(cudaGetLastError(), "LOD gather launch");
}

} // namespace planet
