// k1_tessellate.cu -- K1: subdivision geometry and index buffers, written straight to HBM.
//
// Replaces the serial CPU code of the reference:
//   root faces            RenderPlanet,  main.cpp:604-624
//   child split           ProcessQuad,   main.cpp:546-547, 581-594
//   QuadID arithmetic                    main.cpp:19-65
//   patch vertex grid     InitPlanet,    main.cpp:402-425
//   strip index buffer    InitPlanet,    main.cpp:427-474
//
// The reference reaches a leaf by recursion from its root face; every leaf's corners are
// a pure function of its QuadID, so here one thread walks one id from the root (depth
// <= 27 levels of 5 normalisations) with the reference's exact fp64 operation order
// (IEEE add/mul/div/sqrt, no FMA) -- corners come out bit-identical, and shared edges of
// neighbouring quads stay bit-identical too because both sides compute Normalize(a+b).
// The strip index buffer has a closed form per output slot (the reference's loop has
// asymmetric cursor resets); the strip of one patch is kept in device memory,
// held in registers by the index-stream CTAs and written with 16-byte stores, rebased per
// quad so that all patches index one merged vertex buffer.
#include "planet_common.cuh"
#include "planet_tma.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace planet {

// main.cpp:604-624: corner signs of the cube, and the six faces.  QUAD(a,b,c,d) stores
// {v[a], v[b], v[d], v[c]} (main.cpp:605), applied here to the face table directly.
__constant__ signed char c_cube[8][3] = {
    {-1, -1, -1}, { 1, -1, -1}, { 1,  1, -1}, {-1,  1, -1},
    {-1, -1,  1}, { 1, -1,  1}, { 1,  1,  1}, {-1,  1,  1}
};
__constant__ unsigned char c_face[6][4] = {      // stored order p[0], p[1], p[2], p[3]
    {0, 1, 3, 2}, {1, 5, 2, 6}, {5, 4, 6, 7}, {4, 0, 7, 3}, {3, 2, 7, 6}, {4, 5, 0, 1}
};

__device__ __forceinline__ d3 cube_corner(int i, double radius)
{
    d3 v = { (double)c_cube[i][0], (double)c_cube[i][1], (double)c_cube[i][2] };
    return exact::mul(exact::normalize(v), radius);                   // main.cpp:607
}

// descend from the root face along the id's path; child order main.cpp:589-592
__device__ Quad quad_from_id(uint64_t id, double radius)
{
    Quad q;
    const int root = (int)quad_root(id);
    const int depth = (int)quad_depth(id);
#pragma unroll
    for (int j = 0; j < 4; j++) q.p[j] = cube_corner(c_face[root][j], radius);
    for (int l = 0; l < depth; l++) {
        const int child = (int)((id >> (2 * l)) & 3);
        // main.cpp:546-547: centre; :581 VERT(i,j) = Normalize(p[i]+p[j]) * radius
        d3 s = exact::add(exact::add(exact::add(q.p[0], q.p[1]), q.p[2]), q.p[3]);
        d3 mid = exact::mul(exact::normalize(s), radius);
        // 3x3 grid g0..g8 = p0, V(0,1), p1, V(0,2), mid, V(1,3), p2, V(2,3), p3; child c takes
        // (0,1,3,4) (1,2,4,5) (3,4,6,7) (4,5,7,8): only the two midpoints it touches are needed.
        // Which two is a SELECT on the child index, not a branch: the lanes of a warp are consecutive
        // leaves, i.e. all four children at the deepest levels, and four divergent paths of two
        // normalisations each made those levels cost 9 normalisations instead of 3.
        //   first midpoint  m1 = V(0,1) V(0,1) V(0,2) V(1,3)   for child 0 1 2 3
        //   second midpoint m2 = V(0,2) V(1,3) V(2,3) V(2,3)
        auto pick = [](bool c, const d3 &a, const d3 &b) { return d3{ c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z }; };
        const bool hi = (child & 2) != 0, odd = (child & 1) != 0;
        const d3 a1 = pick(hi && odd, q.p[1], q.p[0]);                    // 0 0 0 1
        const d3 b1 = pick(hi, pick(odd, q.p[3], q.p[2]), q.p[1]);        // 1 1 2 3
        const d3 a2 = pick(hi, q.p[2], pick(odd, q.p[1], q.p[0]));        // 0 1 2 2
        const d3 b2 = pick(hi || odd, q.p[3], q.p[2]);                    // 2 3 3 3
        const d3 m1 = exact::mul(exact::normalize(exact::add(a1, b1)), radius);
        const d3 m2 = exact::mul(exact::normalize(exact::add(a2, b2)), radius);
        // child 0: (p0, m1, m2, mid)  1: (m1, p1, mid, m2)  2: (m1, mid, p2, m2)  3: (mid, m1, m2, p3)
        const d3 n0 = child == 0 ? q.p[0] : child == 3 ? mid : m1;
        const d3 n1 = child == 0 ? m1 : child == 1 ? q.p[1] : child == 2 ? mid : m1;
        const d3 n2 = child == 0 ? m2 : child == 1 ? mid : child == 2 ? q.p[2] : m2;
        const d3 n3 = child == 0 ? mid : child == 3 ? q.p[3] : m2;
        q.p[0] = n0; q.p[1] = n1; q.p[2] = n2; q.p[3] = n3;
    }
    q.id = id;
    return q;
}

// id of leaf `leaf` (0 .. 6*4^depth - 1) in the reference's emission order: face-major,
// then depth-first with child 0..3 -- the first split is the most significant base-4
// digit of the in-face index, and QuadID stores level l's child at bits 2(l-1).
__host__ __device__ inline uint64_t uniform_leaf_id(int64_t leaf, int depth)
{
    const int64_t per_face = (int64_t)1 << (2 * depth);
    const uint64_t face = (uint64_t)(leaf / per_face);
    uint64_t j = (uint64_t)(leaf - (int64_t)face * per_face);
    uint64_t path = 0;
    for (int l = 0; l < depth; l++) {
        uint64_t digit = (j >> (2 * (depth - 1 - l))) & 3;
        path |= digit << (2 * l);
    }
    return (1ull << 63) | (face << 60) | ((uint64_t)depth << 55) | path;
}

// 104-byte record store (the quads are ~1 % of K1's traffic; the index buffer is the rest)
__device__ __forceinline__ void store_quad(Quad *dst, const Quad &q)
{
    *dst = q;
}

__global__ void __launch_bounds__(128)
k_quads_uniform(int depth, int64_t first, int64_t n, double radius, Quad *__restrict__ out)   // quads only
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        store_quad(out + i, quad_from_id(uniform_leaf_id(first + i, depth), radius));
}

__global__ void __launch_bounds__(128)
k_quads_from_ids(const uint64_t *__restrict__ ids, int64_t n, double radius, Quad *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t id = ids[i];
        Quad q;
        if ((id >> 63) && quad_root(id) < 6) q = quad_from_id(id, radius);
        else { for (int j = 0; j < 4; j++) q.p[j] = { 0.0, 0.0, 0.0 }; q.id = 0; }   // invalid id (main.cpp:35)
        store_quad(out + i, q);
    }
}

// ---- strip index buffer, closed form of main.cpp:427-474 ---------------------------------
// n = patch_verts.  Vertex numbering (main.cpp:406-422): n top-skirt verts, then n rows of
// (skirt, n verts, skirt), then n bottom-skirt verts.
__host__ __device__ inline uint32_t strip_index(int k, int n)
{
    const int w = n + 2;                       // vertices per full row
    if (k < 2 * n) {                           // top skirt: v0 = x, v1 = n+1+x   (:434-438)
        int x = k >> 1;
        return (k & 1) ? (uint32_t)(n + 1 + x) : (uint32_t)x;
    }
    k -= 2 * n;
    if (k < 2) return k == 0 ? (uint32_t)(2 * n) : (uint32_t)n;      // reset (:441-443)
    k -= 2;
    const int row_len = 2 * w + 2;             // 2(n+2) strip indices + 2 reset indices
    const int body = (n - 1) * 2 * w + (n - 2) * 2;
    if (k < body) {                            // rows y = 0 .. n-2       (:445-457)
        int y = k / row_len, r = k - y * row_len;
        int v0 = n + y * w, v1 = 2 * n + 2 + y * w;
        if (r < 2 * w) {
            int x = r >> 1;
            return (r & 1) ? (uint32_t)(v1 + x) : (uint32_t)(v0 + x);
        }
        return (r == 2 * w) ? (uint32_t)(v1 + w - 1) : (uint32_t)(v0 + w);
    }
    k -= body;
    const int v0e = n + (n - 1) * w, v1e = 2 * n + 2 + (n - 1) * w;  // cursors after the rows
    if (k < 2) return k == 0 ? (uint32_t)(v1e - 1) : (uint32_t)(v0e + 1);   // v0++, reset (:459-462)
    k -= 2;
    int x = k >> 1;                            // bottom skirt           (:465-469)
    return (k & 1) ? (uint32_t)(v1e + x) : (uint32_t)(v0e + 1 + x);
}

// merged index buffer: out[q*ni + k] = q*nv + strip[k].  Each thread keeps its slots of the
// strip in registers (shared memory for patches too large for that); a CTA then walks whole
// quads, each thread streaming VEC consecutive indices per store (16 bytes when ni % 4 == 0,
// i.e. even patch_verts; else 8 bytes -- ni is always even), so the index part is a pure
// coalesced HBM write: address arithmetic + stores.
// K1 as ONE launch: the first `quad_blocks` CTAs walk QuadIDs to corners (latency-bound fp64
// chains, few warps), all other CTAs stream the merged index buffer (bandwidth-bound).  The two
// outputs are independent, so running them side by side hides the corner chains completely.
// (forcing more resident CTAs with a register cap spills the fp64 corner chains and makes them the
// long pole: 31 -> 41 us at a 40-register cap)
template <int VEC>
__global__ void __launch_bounds__(256)
k_tessellate_fused(int depth, int64_t first, int64_t nquads, double radius, Quad *__restrict__ quads,
                   int quad_blocks, int n, int nv, int ni, uint32_t *__restrict__ indices,
                   const uint32_t *__restrict__ strip)
{
    extern __shared__ __align__(16) uint32_t s_strip[];
    if ((int)blockIdx.x < quad_blocks) {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nquads;
             i += (int64_t)quad_blocks * blockDim.x)
            store_quad(quads + i, quad_from_id(uniform_leaf_id(first + i, depth), radius));
        return;
    }
    // Each thread owns up to KV vectors of the strip (slots t, t + 256, ...) in REGISTERS and
    // walks the quads: the loop body is pure address arithmetic + streaming stores, no loads.
    constexpr int KV = 4;                                  // 4 x 256 threads x VEC indices >= ni for n <= ~44
    const int nvec = ni / VEC;
    const int64_t qstride = (int64_t)gridDim.x - quad_blocks;
    const int64_t q0 = (int64_t)blockIdx.x - quad_blocks;
    if (nvec <= KV * (int)blockDim.x) {
        uint32_t reg[KV][VEC];
#pragma unroll
        for (int j = 0; j < KV; j++) {
            const int v = threadIdx.x + j * blockDim.x;
#pragma unroll
            for (int e = 0; e < VEC; e++)                  // the library's cached strip if there is one, else the closed form
                reg[j][e] = v < nvec ? (strip ? __ldg(strip + v * VEC + e) : strip_index(v * VEC + e, n)) : 0u;
        }
        for (int64_t q = q0; q < nquads; q += qstride) {
            const uint32_t base = (uint32_t)q * (uint32_t)nv;
            uint32_t *dst = indices + q * ni;
#pragma unroll
            for (int j = 0; j < KV; j++) {
                const int v = threadIdx.x + j * blockDim.x;
                if (v < nvec) {
                    if (VEC == 4)
                        __stcs(reinterpret_cast<uint4 *>(dst) + v,
                               make_uint4(reg[j][0] + base, reg[j][1] + base, reg[j][2] + base, reg[j][3 % VEC] + base));
                    else
                        __stcs(reinterpret_cast<uint2 *>(dst) + v, make_uint2(reg[j][0] + base, reg[j][1] + base));
                }
            }
        }
        return;
    }
    // large patches: stage the strip in shared memory instead
    for (int k = threadIdx.x; k < ni; k += blockDim.x) s_strip[k] = strip ? __ldg(strip + k) : strip_index(k, n);
    __syncthreads();
    for (int64_t q = q0; q < nquads; q += qstride) {
        const uint32_t base = (uint32_t)q * (uint32_t)nv;
        if (VEC == 4) {
            uint4 *dst = reinterpret_cast<uint4 *>(indices + q * ni);
            for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
                uint4 s = reinterpret_cast<const uint4 *>(s_strip)[v];
                __stcs(dst + v, make_uint4(s.x + base, s.y + base, s.z + base, s.w + base));
            }
        } else {
            uint2 *dst = reinterpret_cast<uint2 *>(indices + q * ni);
            for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
                uint2 s = reinterpret_cast<const uint2 *>(s_strip)[v];
                __stcs(dst + v, make_uint2(s.x + base, s.y + base));
            }
        }
    }
}

// K1 with the index stream leaving each SM as bulk copies (TMA, SASS UBLKCP).  The first
// `quad_blocks` CTAs walk QuadIDs to corners as above.  Every other CTA keeps its share of the
// strip in registers, assembles the rebased strip of one quad (ni * 4 bytes: 8 144 at the
// reference's n = 30) in one of NB shared-memory buffers -- 16-byte STS, no global store
// instructions at all -- and thread 0 hands the buffer to the TMA unit as ONE
// cp.async.bulk shared -> global.  Up to NB - 1 quads per CTA are in flight while the next one is
// assembled, so HBM always has several 8 KB requests per SM queued whatever the warp occupancy is
// (the per-thread-store version had to keep 16-byte stores of many warps in flight for that and
// reached 0.66 of the copy peak).  One __syncthreads per quad.
constexpr int BULK_THREADS = 128;
constexpr int BULK_KV = 8;                      // strip vectors (uint4) a thread keeps in registers
constexpr int BULK_NB = 4;                      // staging buffers per CTA

__global__ void __launch_bounds__(BULK_THREADS)
k_tessellate_bulk(int depth, int64_t first, int64_t nquads, double radius, Quad *__restrict__ quads,
                  int quad_blocks, int n, int nv, int ni, uint32_t *__restrict__ indices,
                  const uint32_t *__restrict__ strip)
{
    extern __shared__ __align__(128) uint32_t s_buf[];                 // NB buffers of ni words (ni % 4 == 0)
    if ((int)blockIdx.x < quad_blocks) {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nquads;
             i += (int64_t)quad_blocks * blockDim.x)
            store_quad(quads + i, quad_from_id(uniform_leaf_id(first + i, depth), radius));
        return;
    }
    const int nvec = ni / 4;
    const bool in_regs = nvec <= BULK_KV * BULK_THREADS;               // n <= 44
    uint4 reg[BULK_KV];
#pragma unroll
    for (int j = 0; j < BULK_KV; j++) {
        const int v = threadIdx.x + j * BULK_THREADS;
        reg[j] = make_uint4(0u, 0u, 0u, 0u);
        if (in_regs && v < nvec)
            reg[j] = strip ? __ldg(reinterpret_cast<const uint4 *>(strip) + v)
                           : make_uint4(strip_index(4 * v, n), strip_index(4 * v + 1, n), strip_index(4 * v + 2, n), strip_index(4 * v + 3, n));
    }
    const int64_t qstride = (int64_t)gridDim.x - quad_blocks;
    int it = 0;
    for (int64_t q = (int64_t)blockIdx.x - quad_blocks; q < nquads; q += qstride, it++) {
        uint4 *buf = reinterpret_cast<uint4 *>(s_buf + (size_t)(it % BULK_NB) * ni);
        const uint32_t base = (uint32_t)q * (uint32_t)nv;              // idx[q][k] = q*nv + strip[k]
        if (in_regs) {
#pragma unroll
            for (int j = 0; j < BULK_KV; j++) {
                const int v = threadIdx.x + j * BULK_THREADS;
                if (v < nvec) buf[v] = make_uint4(reg[j].x + base, reg[j].y + base, reg[j].z + base, reg[j].w + base);
            }
        } else {                                                       // large patches: the strip comes from L1/L2
            for (int v = threadIdx.x; v < nvec; v += BULK_THREADS) {
                const uint4 sv = strip ? __ldg(reinterpret_cast<const uint4 *>(strip) + v)
                                       : make_uint4(strip_index(4 * v, n), strip_index(4 * v + 1, n), strip_index(4 * v + 2, n), strip_index(4 * v + 3, n));
                buf[v] = make_uint4(sv.x + base, sv.y + base, sv.z + base, sv.w + base);
            }
        }
        tma::fence_smem_writes();
        // the buffer the NEXT quad is assembled in was handed to the TMA unit NB - 1 quads ago:
        // all but the latest NB - 2 copies have been read out of shared memory after this wait
        if (threadIdx.x == 0) tma::wait_read<BULK_NB - 2>();
        __syncthreads();
        if (threadIdx.x == 0) {
            tma::store_bulk(indices + q * ni, buf, (uint32_t)ni * 4u);
            tma::commit();
        }
    }
    if (threadIdx.x == 0) tma::wait_all<0>();
}

// The merged index stream by a kernel small enough to live BESIDE the height-map kernel: K2 keeps
// an SM's arithmetic busy and never touches HBM, the index stream is nothing but HBM writes, so the
// two belong on the same SMs at the same time.  A resident 768-thread K2 CTA leaves 4 096
// registers and ~10 KB of shared memory per SM: this kernel takes 128 threads x <= 32 registers and
// no shared memory (the 8 KB strip of one patch is read through L1), one 16-byte store per thread
// per step, (quad, vector) advanced incrementally so there is no division in the loop.
__global__ void __launch_bounds__(128, 16)
k_index_stream_slim(int64_t nquads, int nv, int nvec, uint4 *__restrict__ out, const uint4 *__restrict__ strip)
{
    const uint32_t stride = gridDim.x * 128u;
    const uint32_t dq = stride / (uint32_t)nvec, dk = stride - dq * (uint32_t)nvec;
    const uint32_t t = blockIdx.x * 128u + threadIdx.x;
    uint32_t q = t / (uint32_t)nvec, k = t - q * (uint32_t)nvec;
    uint4 *dst = out + t;
    for (; q < (uint32_t)nquads; dst += stride) {
        const uint4 sv = __ldg(strip + k);
        const uint32_t base = q * (uint32_t)nv;                          // idx[q][k] = q*nv + strip[k]
        __stcs(dst, make_uint4(sv.x + base, sv.y + base, sv.z + base, sv.w + base));
        q += dq; k += dk;
        if (k >= (uint32_t)nvec) { k -= (uint32_t)nvec; q++; }
    }
}

// the reference's static patch: vertices (main.cpp:402-425) and indices (:427-474)
__global__ void k_patch_mesh(int n, float *__restrict__ verts, uint32_t *__restrict__ indices)
{
    const int nv = n * n + 4 * n, ni = 2 * n * n + 8 * n - 4, w = n + 2;
    const double div = __ddiv_rn(1.0, (double)(n - 1));               // main.cpp:404
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        if (!verts) break;
        float u, v, s;
        if (i < n) {                                      // top skirt (:406-407)
            u = __double2float_rn(__dmul_rn((double)i, div)); v = 0.0f; s = 1.0f;
        } else if (i < n + n * w) {
            int r = i - n, y = r / w, c = r - y * w;
            v = __double2float_rn(__dmul_rn((double)y, div));
            if (c == 0)          { u = 0.0f; s = 1.0f; }   // :411
            else if (c == w - 1) { u = 1.0f; s = 1.0f; }   // :416
            else { u = __double2float_rn(__dmul_rn((double)(c - 1), div)); s = 0.0f; }   // :414
        } else {                                          // bottom skirt (:419-420)
            u = __double2float_rn(__dmul_rn((double)(i - n - n * w), div)); v = 1.0f; s = 1.0f;
        }
        verts[3 * i] = u; verts[3 * i + 1] = v; verts[3 * i + 2] = s;
    }
    if (indices)
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ni; k += gridDim.x * blockDim.x)
            indices[k] = strip_index(k, n);
}

// =====================================================================================
static int sm_count_k1()
{
    static int n[64] = {};                                           // per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!n[dev]) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// The strip of one patch (main.cpp:427-474), kept in device memory per patch size: evaluating the
// closed form in every index-stream CTA cost ~3 us of a 30 us launch.  Built on the host (the closed
// form is __host__ __device__) and uploaded with a synchronous copy, so it is complete before any
// stream can read it; while a stream is being captured the kernels fall back to the closed form.
static std::mutex g_strip_mutex;
static struct StripCache { int n = 0; int dev = -1; uint32_t *d = nullptr; } g_strip;

void release_strip_cache()                                            // planet_gpu_shutdown
{
    std::lock_guard<std::mutex> lock(g_strip_mutex);
    if (g_strip.d) cudaFree(g_strip.d);
    g_strip = StripCache();
}

static const uint32_t *cached_strip(int n, int ni, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_strip_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_strip.d && g_strip.n == n && g_strip.dev == dev) return g_strip.d;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return nullptr;
    }
    if (g_strip.d) { cudaFree(g_strip.d); g_strip = StripCache(); }
    std::vector<uint32_t> h((size_t)ni);
    for (int k = 0; k < ni; k++) h[k] = strip_index(k, n);
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, (size_t)ni * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemcpy(d, h.data(), (size_t)ni * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        if (d) cudaFree(d);
        return nullptr;                                               // not fatal: closed form in the kernel
    }
    g_strip.n = n; g_strip.dev = dev; g_strip.d = d;
    return d;
}

int launch_index_stream_beside(const planet_gpu_params *p, int64_t nquads, uint32_t *d_indices, cudaStream_t stream);

int launch_tessellate_uniform(const planet_gpu_params *p, int depth, int64_t first, int64_t nquads,
                              Quad *d_quads, uint32_t *d_indices, cudaStream_t stream)
{
    if (nquads == 0) return 0;
    const int n = p->patch_verts;
    const int nv = n * n + 4 * n, ni = 2 * n * n + 8 * n - 4;
    if (d_indices) {
        if ((uint64_t)(nquads) * (uint64_t)nv > 0xFFFFFFFFull)
            return set_error(PLANET_E_INVALID, "merged vertex count %lld x %d exceeds uint32 indices",
                             (long long)nquads, nv);
        if ((reinterpret_cast<uintptr_t>(d_indices) & 7) != 0)
            return set_error(PLANET_E_INVALID, "index buffer must be 8-byte aligned");
    }
    if (d_quads && !d_indices) {
        int grid = (int)std::min<int64_t>((nquads + 31) / 32, (int64_t)sm_count_k1() * 16);
        k_quads_uniform<<<grid, 32, 0, stream>>>(depth, first, nquads, p->radius, d_quads);
    } else {
        const bool vec4 = (ni % 4 == 0) && (reinterpret_cast<uintptr_t>(d_indices) & 15) == 0;
        static const int idx_per_sm = [] {                               // tuning knob: index-stream CTAs per SM
            const char *e = getenv("PLANET_K1_IDX_PER_SM");
            return e ? std::max(1, std::min(8, atoi(e))) : 0;
        }();
        static const bool no_bulk = getenv("PLANET_K1_NO_BULK") != nullptr;   // test knob: the per-thread-store path
        const uint32_t *strip = cached_strip(n, ni, stream);
        const size_t bulk_smem = (size_t)BULK_NB * ni * sizeof(uint32_t);
        if (vec4 && !no_bulk && bulk_smem <= 200 * 1024) {
            // index stream as bulk copies: one 16-byte aligned request of ni*4 bytes per quad
            static size_t configured[64] = {};                           // per device
            int dev = 0;
            cudaGetDevice(&dev);
            if (bulk_smem > 48 * 1024 && bulk_smem > configured[dev & 63]) {
                PLANET_CUDA(cudaFuncSetAttribute(k_tessellate_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
                configured[dev & 63] = bulk_smem;
            }
            const int per_sm = idx_per_sm ? idx_per_sm : (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / bulk_smem));
            int quad_blocks = d_quads ? (int)std::min<int64_t>((nquads + BULK_THREADS - 1) / BULK_THREADS, sm_count_k1()) : 0;
            int idx_blocks = (int)std::min<int64_t>(nquads, (int64_t)sm_count_k1() * per_sm);
            k_tessellate_bulk<<<quad_blocks + idx_blocks, BULK_THREADS, bulk_smem, stream>>>(
                depth, first, nquads, p->radius, d_quads, quad_blocks, n, nv, ni, d_indices, strip);
        } else if ((size_t)ni * sizeof(uint32_t) > 200 * 1024) {
            // a patch whose strip alone outgrows shared memory (n > 158): the quads by their own kernel, the
            // indices by the kernel that reads the strip through L1 (it needs the 16-byte shape: even n)
            if (!vec4)
                return set_error(PLANET_E_UNSUPPORTED, "merged index buffer for patch_verts %d: the strip (%d indices) does not fit "
                                 "shared memory and is not a whole number of 16-byte vectors", n, ni);
            if (d_quads) {
                int grid = (int)std::min<int64_t>((nquads + 31) / 32, (int64_t)sm_count_k1() * 16);
                k_quads_uniform<<<grid, 32, 0, stream>>>(depth, first, nquads, p->radius, d_quads);
                count_launch();
            }
            return launch_index_stream_beside(p, nquads, d_indices, stream);
        } else {
            size_t smem = (size_t)ni * sizeof(uint32_t);
            if (smem > 48 * 1024) {
                PLANET_CUDA(cudaFuncSetAttribute(k_tessellate_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PLANET_CUDA(cudaFuncSetAttribute(k_tessellate_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            // indices are rebased to the caller's buffer: quad 0 of this call is vertex block 0
            int quad_blocks = d_quads ? (int)std::min<int64_t>((nquads + 255) / 256, sm_count_k1()) : 0;
            int idx_blocks = (int)std::min<int64_t>(nquads, (int64_t)sm_count_k1() * (idx_per_sm ? idx_per_sm : 6));
            if (vec4) k_tessellate_fused<4><<<quad_blocks + idx_blocks, 256, smem, stream>>>(
                          depth, first, nquads, p->radius, d_quads, quad_blocks, n, nv, ni, d_indices, strip);
            else      k_tessellate_fused<2><<<quad_blocks + idx_blocks, 256, smem, stream>>>(
                          depth, first, nquads, p->radius, d_quads, quad_blocks, n, nv, ni, d_indices, strip);
        }
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "tessellate launch");
}

// the merged strip indices of `nquads` patches by the slim kernel (see k_index_stream_slim)
int launch_index_stream_beside(const planet_gpu_params *p, int64_t nquads, uint32_t *d_indices, cudaStream_t stream)
{
    if (nquads == 0) return 0;
    const int n = p->patch_verts;
    const int nv = n * n + 4 * n, ni = 2 * n * n + 8 * n - 4;
    if ((uint64_t)nquads * (uint64_t)nv > 0xFFFFFFFFull)
        return set_error(PLANET_E_INVALID, "merged vertex count %lld x %d exceeds uint32 indices", (long long)nquads, nv);
    if ((ni & 3) || (reinterpret_cast<uintptr_t>(d_indices) & 15) || (uint64_t)nquads * (uint64_t)(ni / 4) > 0xFFFFFFFFull)
        return launch_tessellate_uniform(p, 0, 0, nquads, nullptr, d_indices, stream);   // shapes the 16-byte stream cannot take
    const uint32_t *strip = cached_strip(n, ni, stream);
    static bool carveout[64] = {};                                       // it must fit on an SM whose carve-out K2 has set
    int dev = 0;
    cudaGetDevice(&dev);
    if (!carveout[dev & 63]) {
        PLANET_CUDA(cudaFuncSetAttribute(k_index_stream_slim, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        carveout[dev & 63] = true;
    }
    k_index_stream_slim<<<sm_count_k1(), 128, 0, stream>>>(nquads, nv, ni / 4, reinterpret_cast<uint4 *>(d_indices),
                                                           reinterpret_cast<const uint4 *>(strip));
    count_launch();
    return check_cuda(cudaGetLastError(), "index stream launch");
}

int launch_quads_from_ids(const planet_gpu_params *p, const uint64_t *d_ids, int64_t n, Quad *d_quads,
                          cudaStream_t stream)
{
    if (n == 0) return 0;
    int grid = (int)std::min<int64_t>((n + 127) / 128, (int64_t)sm_count_k1() * 16);
    k_quads_from_ids<<<grid, 128, 0, stream>>>(d_ids, n, p->radius, d_quads);
    count_launch();
    return check_cuda(cudaGetLastError(), "quads_from_ids launch");
}

int launch_patch_mesh(int n, float *d_vertices, uint32_t *d_indices, cudaStream_t stream)
{
    k_patch_mesh<<<8, 256, 0, stream>>>(n, d_vertices, d_indices);
    count_launch();
    return check_cuda(cudaGetLastError(), "patch_mesh launch");
}

uint32_t host_strip_index(int k, int n) { return strip_index(k, n); }
uint64_t host_uniform_leaf_id(int64_t leaf, int depth) { return uniform_leaf_id(leaf, depth); }

} // namespace planet
