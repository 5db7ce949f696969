// k3_shade.cu -- K3: displacement, per-vertex normal and Lambert term.
//
// The reference does this per frame in a GLSL 1.40 vertex + fragment shader
// (main.cpp:286-380), fed by per-quad uniforms (main.cpp:662-677) and the quad's R32F
// height texture (render.cpp:415-435).  Here the same arithmetic runs once per vertex
// for whole batches of quads and the results stay in HBM as two float4 streams.
//
// HBM-bound by design: 4 B read + 32 B written per vertex, so everything else is kept
// off the critical path:
//   A. per quad, threads 0..3 form the uniforms P[j] = float(q.p[j] - cam),
//      N[j] = float(Normalize(q.p[j])) in fp64 exactly as main.cpp:666-672, while all
//      threads stage the (n+2)^2 height map into shared memory with 16-byte loads: each
//      height leaves HBM once, the 5-point stencil of compute_normal (main.cpp:338-346)
//      is served on chip;
//   B. the edge interpolants p = interpolate(a,b,UV.x), q = interpolate(c,d,UV.x)
//      (main.cpp:354-355) do not depend on UV.y: they are evaluated once per COLUMN (2n
//      per quad instead of 2 per vertex), together with q.p - p.p, xyscale (main.cpp:361)
//      and the column-constant terms of the third interpolate (acos, tan, 1/sin); q is
//      stored as deltas from p so the linear branch costs 6 FMAs.  Column data is kept
//      in shared memory as 80-byte records read with four LDS.128 (the stride puts the 8
//      records of a quarter-warp in 8 distinct bank groups);
//   C. one warp owns one quad and walks its vertex slots 32 at a time (for the reference's
//      n = 30 a step is one row incl. the two skirt vertices): height taps are consecutive
//      words and each float4 store instruction writes 512 contiguous bytes.
//      Warps never synchronise with each other, so A/B of one warp hide behind C of others.
// Normalisations use rsqrt (MUFU) + multiplies instead of the shader's sqrt + divide; the
// difference (<= 2 ulp) is far inside the 1e-4 rad parity bound.
//
// With the quad's own height map the sampler coordinate of main.cpp:358 is the centre of
// texel (vx+1, vy+1), so GL_LINEAR filtering (render.cpp:429-430) is a direct read.
#include "planet_common.cuh"
#include "planet_tma.cuh"

#include <algorithm>
#include <cstdlib>

namespace planet {

namespace shade {

constexpr int THREADS = 128;        // 4 warps per CTA, up to 6 CTAs per SM: best of the sweep in tools/k3_sweep.py
constexpr int WARPS = THREADS / 32;
constexpr int COL_STRIDE = 20;       // floats per column record (17 used): 80 B keeps LDS.128 aligned and
                                     // 8 consecutive records of a quarter-warp fall in 8 distinct bank groups

struct V { float3 p, n; };                                          // main.cpp:298

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
// MUFU.RSQ / MUFU.SQRT alone: rsqrtf()/sqrtf() wrap them in denormal scaling and a Newton step
// (4 and 12 instructions); every argument here is a squared length or >= 0.001
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float3 normalize(float3 a) { return a * rsqrt_fast(dot(a, a)); }
__device__ __forceinline__ float3 cross(float3 a, float3 b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// GLSL mix(x, y, a) = x*(1-a) + y*a
__device__ __forceinline__ float3 mix(float3 a, float3 b, float t, float omt)
{
    return f3(fmaf(b.x, t, a.x * omt), fmaf(b.y, t, a.y * omt), fmaf(b.z, t, a.z * omt));
}

// main.cpp:300-332, used for the two edge interpolants (phase B, 2n calls per quad)
__device__ V interpolate(const V &v0, const V &v1, float t)
{
    V r;
    float d = dot(v0.n, v1.n);
    if (1.0f - d < 0.001f) {                                        // interpolate_linear
        r.n = normalize(mix(v0.n, v1.n, t, 1.0f - t));
        r.p = mix(v0.p, v1.p, t, 1.0f - t);
        return r;
    }
    float theta2 = acosf(d);
    float k = 1.0f - t;
    float3 n = normalize(v0.n * sinf(k * theta2) + v1.n * sinf(t * theta2));
    float theta = theta2 * 0.5f;
    float gamma = theta - theta2 * t;
    float tan_theta = tanf(theta);
    float x = 1.0f - tanf(gamma) / tan_theta;
    float y = 1.0f / sinf(theta) - 1.0f / (cosf(gamma) * tan_theta);
    float3 v = (v1.p - v0.p) * 0.5f;
    r.p = v0.p + v * x + n * (y * sqrtf(dot(v, v)));
    r.n = n;
    return r;
}

// Column record: five float4 groups, read with LDS.128 (q is kept as the deltas q.p - p.p and
// q.n - p.n: the linear branch is then 6 FMAs).  The fifth group is only read on the slerp branch.
//   0: p.p.xyz, th2    1: p.n.xyz, 2*xyscale    2: (q.p - p.p).xyz, hlen    3: (q.n - p.n).xyz, itan    4: isin

// One WARP per quad: warps never wait for each other (no block barrier), so one warp's
// column phase overlaps the other warps' vertex phases.  Per-warp shared memory:
//   [4 corner uniforms | np column records of COL_STRIDE floats | (n+2)^2 staged heights]
// GL_LINEAR + GL_CLAMP_TO_EDGE fetch of a dim x dim R32F texture (render.cpp:429-433) at (u, v)
__device__ __forceinline__ float sample_bilinear(const float *tex, int dim, float u, float v)
{
    float x = u * (float)dim - 0.5f, y = v * (float)dim - 0.5f;
    float x0 = floorf(x), y0 = floorf(y);
    float fx = x - x0, fy = y - y0;
    int ix0 = min(max((int)x0, 0), dim - 1), ix1 = min(max((int)x0 + 1, 0), dim - 1);
    int iy0 = min(max((int)y0, 0), dim - 1), iy1 = min(max((int)y0 + 1, 0), dim - 1);
    float a = tex[iy0 * dim + ix0], b = tex[iy0 * dim + ix1], c = tex[iy1 * dim + ix0], d = tex[iy1 * dim + ix1];
    float top = fmaf(b - a, fx, a), bot = fmaf(d - c, fx, c);
    return fmaf(bot - top, fy, top);
}

// ---- the two inner bodies, shared by the warp-per-quad kernel and the CTA-per-chunk kernel ----
// phase B for one column: both edge interpolants at UV.x and the column-constant part of
// v = interpolate(p, q, UV.y) (main.cpp:310-326, 354-355), as one column record
__device__ __forceinline__ void column_record(const V *s_corner, float ux, int n, float4 *rec)
{
    V p = interpolate(s_corner[0], s_corner[1], ux);                 // main.cpp:354
    V q = interpolate(s_corner[2], s_corner[3], ux);                 // main.cpp:355
    float3 pq = q.p - p.p;
    float len = sqrtf(dot(pq, pq));
    float d = dot(p.n, q.n);
    float th2 = -1.0f, itan = 0.f, isin = 0.f;
    if (!(1.0f - d < 0.001f)) {
        th2 = acosf(d);
        float theta = th2 * 0.5f;
        itan = 1.0f / tanf(theta);
        isin = 1.0f / sinf(theta);
    }
    rec[0] = make_float4(p.p.x, p.p.y, p.p.z, th2);
    rec[1] = make_float4(p.n.x, p.n.y, p.n.z, 2.0f * (len / (float)(n - 1)));   // 2*xyscale, main.cpp:345,361 (29.0 == n-1)
    rec[2] = make_float4(pq.x, pq.y, pq.z, 0.5f * len);                         // w: length((q.p - p.p) * 0.5)
    rec[3] = make_float4(q.n.x - p.n.x, q.n.y - p.n.y, q.n.z - p.n.z, itan);
    rec[4] = make_float4(isin, 0.f, 0.f, 0.f);
}

// what phase C needs of its quad
struct QuadCtx {
    int n, w, dim; unsigned magic_w;
    const float *s_uv; const float4 *s_col; const float *s_h; const float *H;
    planet_gpu_texrect rect; float skirt_size;
    float4 *pos_q, *nrm_q;
};

// phase C for vertex slot i of the quad (main.cpp:348-367 + the fragment stage, :371-378)
template <bool STAGE, bool RECT>
__device__ __forceinline__ void shade_vertex(const QuadCtx &c, int i)
{
    const int n = c.n, w = c.w, dim = c.dim;
    // Slot -> (row, column).  Padding the top and bottom skirt rows (n slots each) to the
    // w slots of a body row makes the patch a (n+2) x w grid: virtual index = slot + 1,
    // + 1 more past the top row, + 1 more past the last body row.
    const int vi = i + 1 + (i >= n) + (i >= n + n * w);
    const int vrow = (int)__umulhi((unsigned)vi, c.magic_w);         // vi / w, exact for vi < 2^16, w <= 256
    const int col = vi - vrow * w;
    const int vx = min(max(col - 1, 0), n - 1), vy = min(max(vrow - 1, 0), n - 1);
    const float skirt = (vrow == 0 || vrow == n + 1 || col == 0 || col == w - 1) ? 1.0f : 0.0f;
    const float t = c.s_uv[vy], omt = 1.0f - t;
    const int row_off = (vy + 1) * dim;
    const float4 *rec = c.s_col + vx * (COL_STRIDE / 4);
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
    const float3 pp = f3(r0.x, r0.y, r0.z), pn = f3(r1.x, r1.y, r1.z);
    const float3 pq = f3(r2.x, r2.y, r2.z), dn = f3(r3.x, r3.y, r3.z);
    const float th2 = r0.w;
    // v = interpolate(p, q, UV.y)                                  // main.cpp:356
    float3 vp, vn;
    if (th2 < 0.0f) {                                                // interpolate_linear, main.cpp:300-308
        // mix(a, b, t) = a + (b - a) t with the column's precomputed b - a
        vn = normalize(f3(fmaf(dn.x, t, pn.x), fmaf(dn.y, t, pn.y), fmaf(dn.z, t, pn.z)));
        vp = f3(fmaf(pq.x, t, pp.x), fmaf(pq.y, t, pp.y), fmaf(pq.z, t, pp.z));
    } else {                                                         // main.cpp:314-331
        const float3 qn = pn + dn;
        vn = normalize(pn * sinf(omt * th2) + qn * sinf(t * th2));
        float gamma = th2 * 0.5f - th2 * t;
        float itan = r3.w;
        float x = 1.0f - tanf(gamma) * itan;
        float y = rec[4].x - itan / cosf(gamma);
        vp = pp + pq * (0.5f * x) + vn * (y * r2.w);
    }
    const int ti = row_off + vx + 1;                                 // texel (vx+1, vy+1)
    float hc, hl, hr, hu, hd;
    if (RECT) {
        // uv = mix(corners0, corners1, UV.xy), taps at +-pixel_size (main.cpp:339-344, 358)
        const float ux = c.s_uv[vx];                                 // UV.x (skirt literals 0.0f / 1.0f coincide)
        const float u = c.rect.corners[0] * (1.0f - ux) + c.rect.corners[2] * ux;
        const float v = c.rect.corners[1] * omt + c.rect.corners[3] * t;
        hc = sample_bilinear(c.s_h, dim, u, v);
        hl = sample_bilinear(c.s_h, dim, u - c.rect.pixel_size[0], v);
        hr = sample_bilinear(c.s_h, dim, u + c.rect.pixel_size[0], v);
        hu = sample_bilinear(c.s_h, dim, u, v - c.rect.pixel_size[1]);
        hd = sample_bilinear(c.s_h, dim, u, v + c.rect.pixel_size[1]);
    } else if (STAGE) { hc = c.s_h[ti]; hl = c.s_h[ti - 1]; hr = c.s_h[ti + 1]; hu = c.s_h[ti - dim]; hd = c.s_h[ti + dim]; }
    else { hc = __ldg(c.H + ti); hl = __ldg(c.H + ti - 1); hr = __ldg(c.H + ti + 1); hu = __ldg(c.H + ti - dim); hd = __ldg(c.H + ti + dim); }
    const float height = hc - c.skirt_size * skirt;                  // main.cpp:360
    float3 nt = normalize(f3(hl - hr, r1.w, hu - hd));               // main.cpp:339-345
    float3 tg = normalize(cross(vn, pq));                            // main.cpp:363
    // main.cpp:364-365 normalise bi = cross(t, n) and mat3(t, n, bi) * normal as well; t, n
    // are unit and orthogonal by construction and |normal| = 1, so both lengths are
    // 1 +- a few ulp and the two rsqrt/multiply groups are left out (<= 1e-6 rad)
    float3 bi = cross(tg, vn);
    float3 N = tg * nt.x + vn * nt.y + bi * nt.z;
    float3 pos = f3(fmaf(vn.x, height, vp.x), fmaf(vn.y, height, vp.y), fmaf(vn.z, height, vp.z));   // :366
    // fragment stage at the vertex: l = normalize(0,1,-1), main.cpp:374-378
    const float inv_sqrt2 = 0.70710678118654752f;
    float light = 0.001f + fmaxf((N.y - N.z) * inv_sqrt2, 0.0f);
    if (c.pos_q) __stcs(c.pos_q + i, make_float4(pos.x, pos.y, pos.z, height));
    if (c.nrm_q) __stcs(c.nrm_q + i, make_float4(N.x, N.y, N.z, sqrt_fast(light)));
}

// STAGE: height map staged in shared memory (else taps read global via L1).
// RECT:  each quad reads its map through a texrect (pool slot + corners + pixel size) with
//        bilinear filtering -- the cache / parent-fallback path (main.cpp:191-237, 334-346, 358).
// PUSH:  (fused multi-GPU gather, K4) the staged map of every k3_every-th quad is also sent to the
//        peers' gathered buffers, one bulk copy of the whole map per peer (planet_tma.cuh): the
//        share of the NVLink transfer the height-map kernel left to this one (PeerOut::k3_every).
template <bool STAGE, bool RECT, bool PUSH = false>
__global__ void __launch_bounds__(THREADS)
k_shade(const Quad *__restrict__ quads, int64_t nquads, int n, double cam_x, double cam_y, double cam_z,
        const float *__restrict__ heights, const planet_gpu_texrect *__restrict__ rects, float max_skirt,
        float4 *__restrict__ pos4, float4 *__restrict__ nrm4, int warp_smem_bytes, PeerOut peers = PeerOut())
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int dim = n + 2, dim2 = dim * dim, w = n + 2, nv = n * n + 4 * n;
    const int np = (n + 31) & ~31;                                   // padded column count
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float *s_uv = reinterpret_cast<float *>(smem);                   // np floats, shared by the block
    unsigned char *mine = smem + np * sizeof(float) + (size_t)warp * warp_smem_bytes;
    V *s_corner = reinterpret_cast<V *>(mine);                       // 4 x 24 B (in 128 B)
    float4 *s_col = reinterpret_cast<float4 *>(mine + 128);          // np records x COL_STRIDE floats
    float *s_h = reinterpret_cast<float *>(s_col) + COL_STRIDE * np; // dim2 floats (if staged)
    const double div = __ddiv_rn(1.0, (double)(n - 1));              // main.cpp:404
    const unsigned magic_w = (unsigned)(((1ull << 32) + (unsigned)w - 1) / (unsigned)w);   // ceil(2^32 / w)

    __shared__ float *s_peer[7];                                     // PUSH: the peers' buffers (indexed by lane below)
    if (PUSH && threadIdx.x < 7) s_peer[threadIdx.x] = peers.ptr[threadIdx.x];
    for (int i = threadIdx.x; i < n; i += blockDim.x)                // UV.x / UV.y values, main.cpp:406-420
        s_uv[i] = __double2float_rn(__dmul_rn((double)i, div));
    __syncthreads();

    const int64_t wstride = (int64_t)gridDim.x * nwarps;
    for (int64_t qi = (int64_t)blockIdx.x * nwarps + warp; qi < nquads; qi += wstride) {
        planet_gpu_texrect rect = {};
        if (RECT) rect = rects[qi];
        const float *H = heights + (RECT ? (int64_t)rect.slot : qi) * dim2;
        __syncwarp();                                                // previous quad fully shaded
        // ---- A: uniforms + height map staging ------------------------------------------------
        if (lane < 4) {
            const d3 p = quads[qi].p[lane];
            d3 rel = { p.x - cam_x, p.y - cam_y, p.z - cam_z };      // main.cpp:668
            d3 nd = exact::normalize(p);                             // main.cpp:669
            V c;
            c.p = f3((float)rel.x, (float)rel.y, (float)rel.z);
            c.n = f3((float)nd.x, (float)nd.y, (float)nd.z);
            s_corner[lane] = c;
        }
        if (PUSH) {                                                  // the previous pushed map has left shared memory
            if (lane < peers.n) tma::wait_read<0>();
            __syncwarp();
        }
        if (STAGE) {
            if ((dim2 & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0) {
                const float4 *H4 = reinterpret_cast<const float4 *>(H);
                for (int i = lane; i < dim2 / 4; i += 32)
                    reinterpret_cast<float4 *>(s_h)[i] = __ldcs(H4 + i);
            } else {
                for (int i = lane; i < dim2; i += 32) s_h[i] = __ldcs(H + i);
            }
        }
        if (PUSH) {
            if (shade_pushes_quad(peers.quad0 + qi, peers.k3_every)) {
                tma::fence_smem_writes();
                __syncwarp();
                if (lane < peers.n) {
                    tma::store_bulk(s_peer[lane] + qi * dim2, s_h, (uint32_t)dim2 * 4u);
                    tma::commit();
                }
            }
        }
        __syncwarp();
        // ---- B: per-column edge interpolants and column constants -------------------------------
        for (int x = lane; x < n; x += 32) column_record(s_corner, s_uv[x], n, s_col + x * (COL_STRIDE / 4));
        __syncwarp();
        // ---- C: the warp walks the patch rows ----------------------------------------------------
        float skirt_size = max_skirt;                                // main.cpp:674-677
        {
            int d = (int)quad_depth(quads[qi].id) - 1;
            if (d > 0) skirt_size /= (float)(2 << d);
        }
        QuadCtx ctx;
        ctx.n = n; ctx.w = w; ctx.dim = dim; ctx.magic_w = magic_w;
        ctx.s_uv = s_uv; ctx.s_col = s_col; ctx.s_h = s_h; ctx.H = H; ctx.rect = rect; ctx.skirt_size = skirt_size;
        ctx.pos_q = pos4 ? pos4 + qi * nv : nullptr;
        ctx.nrm_q = nrm4 ? nrm4 + qi * nv : nullptr;
        // The warp walks the quad's nv vertex slots 32 at a time (main.cpp:406-422 order: n top-skirt
        // slots, n rows of n+2, n bottom-skirt slots); a lane's slot gives its row and column.  For
        // the reference's n = 30 a 32-slot step is exactly one row; for other sizes lanes of one
        // step may sit on two rows, which costs nothing since everything is per lane.
        for (int i0 = 0; i0 < nv; i0 += 32)
            if (i0 + lane < nv) shade_vertex<STAGE, RECT>(ctx, i0 + lane);
    }
    if (PUSH) { if (lane < peers.n) tma::wait_all<0>(); }
}

// The same stage for FEW, LARGE patches (BASELINE config 5's end of the range: one 254 x 254 patch is
// 65 532 vertices): one warp per quad leaves the chip empty there (a single 256^2 patch took 1.2 ms),
// so here a whole CTA works on one (quad, chunk of vertex slots): the corner uniforms, the staged map
// and the column records are built by all its threads between block barriers, and the slots of the
// chunk are walked by all its warps.  Several CTAs share a quad; each rebuilds the column records
// (n interpolations -- small against a chunk of a thousand vertices or more).
constexpr int WIDE_THREADS = 256;
constexpr int WIDE_MIN_CHUNK = 1024;       // vertex slots per work item, at least (the host sizes chunks to ~4 items per SM)

template <bool STAGE>
__global__ void __launch_bounds__(WIDE_THREADS)
k_shade_wide(const Quad *__restrict__ quads, int64_t nquads, int n, double cam_x, double cam_y, double cam_z,
             const float *__restrict__ heights, float max_skirt, float4 *__restrict__ pos4, float4 *__restrict__ nrm4,
             int chunks, int chunk)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int dim = n + 2, dim2 = dim * dim, w = n + 2, nv = n * n + 4 * n;
    const int np = (n + 31) & ~31;
    float *s_uv = reinterpret_cast<float *>(smem);
    unsigned char *mine = smem + np * sizeof(float);
    V *s_corner = reinterpret_cast<V *>(mine);
    float4 *s_col = reinterpret_cast<float4 *>(mine + 128);
    float *s_h = reinterpret_cast<float *>(s_col) + COL_STRIDE * np;
    const double div = __ddiv_rn(1.0, (double)(n - 1));              // main.cpp:404
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_uv[i] = __double2float_rn(__dmul_rn((double)i, div));
    int64_t staged = -1;                                             // quad whose uniforms / map / columns are in shared memory
    for (int64_t item = blockIdx.x; item < nquads * chunks; item += gridDim.x) {
        const int64_t qi = item / chunks;
        const int ch = (int)(item - qi * chunks);
        const float *H = heights + qi * dim2;
        if (qi != staged) {                                          // (block-uniform)
            __syncthreads();                                         // the previous quad is fully shaded
            if (threadIdx.x < 4) {
                const d3 p = quads[qi].p[threadIdx.x];
                d3 rel = { p.x - cam_x, p.y - cam_y, p.z - cam_z };  // main.cpp:668
                d3 nd = exact::normalize(p);                         // main.cpp:669
                V c;
                c.p = f3((float)rel.x, (float)rel.y, (float)rel.z);
                c.n = f3((float)nd.x, (float)nd.y, (float)nd.z);
                s_corner[threadIdx.x] = c;
            }
            if (STAGE) for (int i = threadIdx.x; i < dim2; i += blockDim.x) s_h[i] = __ldcs(H + i);
            __syncthreads();
            for (int x = threadIdx.x; x < n; x += blockDim.x) column_record(s_corner, s_uv[x], n, s_col + x * (COL_STRIDE / 4));
            __syncthreads();
            staged = qi;
        }
        QuadCtx ctx;
        ctx.n = n; ctx.w = w; ctx.dim = dim; ctx.magic_w = (unsigned)(((1ull << 32) + (unsigned)w - 1) / (unsigned)w);
        ctx.s_uv = s_uv; ctx.s_col = s_col; ctx.s_h = s_h; ctx.H = H; ctx.rect = planet_gpu_texrect{};
        ctx.skirt_size = max_skirt;                                  // main.cpp:674-677
        {
            int d = (int)quad_depth(quads[qi].id) - 1;
            if (d > 0) ctx.skirt_size /= (float)(2 << d);
        }
        ctx.pos_q = pos4 ? pos4 + qi * nv : nullptr;
        ctx.nrm_q = nrm4 ? nrm4 + qi * nv : nullptr;
        const int hi = min(nv, (ch + 1) * chunk);
        for (int i = ch * chunk + threadIdx.x; i < hi; i += blockDim.x) shade_vertex<STAGE, false>(ctx, i);
    }
}

} // namespace shade

int launch_shade_push(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, const double *cam,
                      const float *d_heights, const planet_gpu_texrect *d_rects, float max_skirt, float *d_pos4,
                      float *d_nrm4, const PeerOut *peers, cudaStream_t stream);

int launch_shade(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, const double *cam,
                 const float *d_heights, const planet_gpu_texrect *d_rects, float max_skirt, float *d_pos4,
                 float *d_nrm4, cudaStream_t stream)
{
    return launch_shade_push(p, d_quads, nquads, cam, d_heights, d_rects, max_skirt, d_pos4, d_nrm4, nullptr, stream);
}

// can the shade kernel push maps for the fused gather (map staged in shared memory, 16-byte granular)?
bool shade_can_push(const planet_gpu_params *p, const float *d_heights)
{
    const int n = p->patch_verts, dim = n + 2, np = (n + 31) & ~31;
    if (n < 2 || n > 254) return false;
    const size_t col_bytes = 128 + (size_t)shade::COL_STRIDE * np * sizeof(float);
    return col_bytes + (size_t)dim * dim * sizeof(float) <= 200 * 1024 && (dim * dim) % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(d_heights) & 15) == 0;
}

int launch_shade_push(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, const double *cam,
                      const float *d_heights, const planet_gpu_texrect *d_rects, float max_skirt, float *d_pos4,
                      float *d_nrm4, const PeerOut *peers, cudaStream_t stream)
{
    if (nquads == 0) return 0;
    const int n = p->patch_verts;
    if (n < 2 || n > 254)
        return set_error(PLANET_E_UNSUPPORTED, "planet_gpu_shade: patch_verts %d outside [2, 254]", n);
    const int dim = n + 2, np = (n + 31) & ~31;
    const size_t col_bytes = 128 + (size_t)shade::COL_STRIDE * np * sizeof(float);
    const size_t hbytes = (size_t)dim * dim * sizeof(float);
    const size_t budget = 200 * 1024;
    const int stage = (col_bytes + hbytes) <= budget;               // else the stencil reads go to L1/L2
    if (d_rects && !stage)
        return set_error(PLANET_E_UNSUPPORTED, "planet_gpu_shade_cached: a %d x %d map does not fit shared memory", dim, dim);
    const size_t per_warp = (col_bytes + (stage ? hbytes : 0) + 15) & ~(size_t)15;   // records are read as float4
    int want_warps = shade::WARPS;
    if (const char *e = getenv("PLANET_K3_WARPS")) want_warps = std::max(1, std::min(shade::WARPS, atoi(e)));   // tuning knob
    int warps = (int)std::max<size_t>(1, std::min<size_t>(want_warps, (budget - np * sizeof(float)) / per_warp));
    size_t smem = np * sizeof(float) + (size_t)warps * per_warp;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static size_t configured[64] = {};                               // per device
    if (smem > 48 * 1024 && smem > configured[dev & 63]) {
        PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = smem;
    }
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(6, 2048 / (warps * 32)), budget / smem));
    if (const char *e = getenv("PLANET_K3_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // tuning knob
    int grid = (int)std::min<int64_t>((nquads + warps - 1) / warps, (int64_t)sms * per_sm);
    float4 *pos = reinterpret_cast<float4 *>(d_pos4), *nrm = reinterpret_cast<float4 *>(d_nrm4);
    // few quads, large patches: not enough warps at one per quad -- a CTA per (quad, chunk of slots) instead
    const int nv = n * n + 4 * n;
    const bool pushing = peers && peers->n > 0 && peers->k3_every != 0;
    if (!pushing && !d_rects && nv >= 2 * shade::WIDE_MIN_CHUNK && nquads < (int64_t)sms * 16 && !getenv("PLANET_K3_NO_WIDE")) {
        // chunks sized for about four work items per SM, a multiple of the CTA's 256 slots per round
        int64_t chunk = std::max<int64_t>(shade::WIDE_MIN_CHUNK, (nquads * nv + 4 * sms - 1) / (4 * sms));
        chunk = std::min<int64_t>((chunk + shade::WIDE_THREADS - 1) / shade::WIDE_THREADS * shade::WIDE_THREADS, nv);
        const int chunks = (int)((nv + chunk - 1) / chunk);
        // a CTA that shades a whole quad stages its map in shared memory; CTAs that share a quad read their
        // taps through L1 instead of each staging the whole map for a slice of it (dim 128, 256 patches:
        // 137 -> measured faster unstaged, profiles/r02zx / r02zy C5 K3 rows)
        const bool wstage = stage && chunks == 1;
        const size_t wsmem = np * sizeof(float) + ((col_bytes + (wstage ? hbytes : 0) + 15) & ~(size_t)15);
        static size_t wide_configured[64] = {};
        if (wsmem > 48 * 1024 && wsmem > wide_configured[dev & 63]) {
            PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade_wide<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade_wide<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            wide_configured[dev & 63] = wsmem;
        }
        const int wide_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, budget / wsmem));
        const int wgrid = (int)std::min<int64_t>(nquads * chunks, (int64_t)sms * wide_per_sm);
        if (wstage) shade::k_shade_wide<true><<<wgrid, shade::WIDE_THREADS, wsmem, stream>>>(d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, max_skirt, pos, nrm, chunks, (int)chunk);
        else       shade::k_shade_wide<false><<<wgrid, shade::WIDE_THREADS, wsmem, stream>>>(d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, max_skirt, pos, nrm, chunks, (int)chunk);
        count_launch();
        return check_cuda(cudaGetLastError(), "shade kernel launch");
    }
    if (peers && peers->n > 0 && peers->k3_every != 0) {
        if (d_rects || !stage) return set_error(PLANET_E_UNSUPPORTED, "shade kernel cannot push this map layout");
        shade::k_shade<true, false, true><<<grid, warps * 32, smem, stream>>>(
            d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, nullptr, max_skirt, pos, nrm, (int)per_warp, *peers);
    } else if (d_rects)
        shade::k_shade<true, true><<<grid, warps * 32, smem, stream>>>(
            d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, d_rects, max_skirt, pos, nrm, (int)per_warp);
    else if (stage)
        shade::k_shade<true, false><<<grid, warps * 32, smem, stream>>>(
            d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, nullptr, max_skirt, pos, nrm, (int)per_warp);
    else
        shade::k_shade<false, false><<<grid, warps * 32, smem, stream>>>(
            d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, nullptr, max_skirt, pos, nrm, (int)per_warp);
    count_launch();
    return check_cuda(cudaGetLastError(), "shade kernel launch");
}

} // namespace planet
