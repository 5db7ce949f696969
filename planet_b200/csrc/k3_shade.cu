// k3_shade.cu -- K3: displacement, per-vertex normal and Lambert term.
//
// The reference does this per frame in a GLSL 1.40 vertex + fragment shader
// (main.cpp:286-380), fed by per-quad uniforms (main.cpp:662-677) and the quad's R32F
// height texture (render.cpp:415-435).  Here the same arithmetic runs once per vertex
// for whole batches of quads and the results stay in HBM as two float4 streams.
//
// Per quad (one CTA iteration):
//   A. threads 0..3 form the uniforms P[j] = float(q.p[j] - cam), N[j] = float(Normalize(q.p[j]))
//      in fp64 exactly as main.cpp:666-672; all threads stage the (n+2)^2 height map into
//      shared memory with coalesced 16-byte loads (each height is read from HBM once; the
//      5-point stencil of compute_normal, main.cpp:338-346, is then served on chip);
//   B. 2n threads evaluate the two edge interpolants p = interpolate(a,b,UV.x),
//      q = interpolate(c,d,UV.x) (main.cpp:354-355) once per COLUMN -- they do not depend on
//      UV.y, so the shader's per-vertex acos/sin/tan work drops from 3 interpolate calls
//      to 1 -- together with q.p - p.p and xyscale (main.cpp:361);
//   C. every thread shades vertices: v = interpolate(p, q, UV.y), height, tangent frame,
//      Normal, position (main.cpp:356-366), Lambert (main.cpp:373-380); float4 stores.
//
// With the quad's own height map the sampler coordinate of main.cpp:358 is the centre of
// texel (vx+1, vy+1), so GL_LINEAR filtering (render.cpp:429-430) is a direct read.
#include "planet_common.cuh"

#include <algorithm>

namespace planet {

namespace shade {

constexpr int THREADS = 256;

struct V { float3 p, n; };                                          // main.cpp:298

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float length(float3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float3 normalize(float3 a) { float l = length(a); return f3(a.x / l, a.y / l, a.z / l); }
__device__ __forceinline__ float3 cross(float3 a, float3 b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 mix(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }

// main.cpp:300-332
__device__ __forceinline__ V interpolate(const V &v0, const V &v1, float t)
{
    V r;
    float d = dot(v0.n, v1.n);
    if (1.0f - d < 0.001f) {                                        // interpolate_linear
        r.n = normalize(mix(v0.n, v1.n, t));
        r.p = mix(v0.p, v1.p, t);
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
    r.p = v0.p + v * x + n * (y * length(v));
    r.n = n;
    return r;
}

// per-column data produced in phase B: 16 floats
struct Column { V p, q; float3 pq; float xyscale; };

__global__ void __launch_bounds__(THREADS)
k_shade(const Quad *__restrict__ quads, int64_t nquads, int n, double cam_x, double cam_y, double cam_z,
        const float *__restrict__ heights, float max_skirt, float4 *__restrict__ pos4,
        float4 *__restrict__ nrm4, int stage_heights)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int dim = n + 2, dim2 = dim * dim, w = n + 2, nv = n * n + 4 * n;
    V *s_corner = reinterpret_cast<V *>(smem);                       // 4 x 24 B
    Column *s_col = reinterpret_cast<Column *>(smem + 128);          // n x 64 B
    float *s_uv = reinterpret_cast<float *>(smem + 128 + (size_t)n * sizeof(Column));   // n floats
    float *s_h = s_uv + ((n + 3) & ~3);                              // dim2 floats (if staged)
    const double div = __ddiv_rn(1.0, (double)(n - 1));              // main.cpp:404

    for (int i = threadIdx.x; i < n; i += blockDim.x)                // UV.x / UV.y values (main.cpp:406-420)
        s_uv[i] = __double2float_rn(__dmul_rn((double)i, div));

    for (int64_t qi = blockIdx.x; qi < nquads; qi += gridDim.x) {
        const float *H = heights + qi * dim2;
        __syncthreads();                                             // previous quad fully shaded
        // ---- A: uniforms + height map staging ------------------------------------------------
        if (threadIdx.x < 4) {
            const d3 p = quads[qi].p[threadIdx.x];
            d3 rel = { p.x - cam_x, p.y - cam_y, p.z - cam_z };      // main.cpp:668
            d3 nd = exact::normalize(p);                             // main.cpp:669
            V c;
            c.p = f3((float)rel.x, (float)rel.y, (float)rel.z);
            c.n = f3((float)nd.x, (float)nd.y, (float)nd.z);
            s_corner[threadIdx.x] = c;
        }
        if (stage_heights) {
            if ((dim2 & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0) {
                const float4 *H4 = reinterpret_cast<const float4 *>(H);
                for (int i = threadIdx.x; i < dim2 / 4; i += blockDim.x)
                    reinterpret_cast<float4 *>(s_h)[i] = __ldg(H4 + i);
            } else {
                for (int i = threadIdx.x; i < dim2; i += blockDim.x) s_h[i] = __ldg(H + i);
            }
        }
        __syncthreads();
        // ---- B: per-column edge interpolants -------------------------------------------------
        for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
            int x = t >> 1, which = t & 1;
            V r = which ? interpolate(s_corner[2], s_corner[3], s_uv[x])     // main.cpp:355
                        : interpolate(s_corner[0], s_corner[1], s_uv[x]);    // main.cpp:354
            if (which) s_col[x].q = r; else s_col[x].p = r;
        }
        __syncthreads();
        for (int x = threadIdx.x; x < n; x += blockDim.x) {
            float3 pq = s_col[x].q.p - s_col[x].p.p;
            s_col[x].pq = pq;
            s_col[x].xyscale = length(pq) / (float)(n - 1);          // main.cpp:361 (29.0 == n-1)
        }
        __syncthreads();
        // ---- C: vertices -----------------------------------------------------------------------
        // main.cpp:674-677
        float skirt_size = max_skirt;
        {
            int d = (int)quad_depth(quads[qi].id) - 1;
            if (d > 0) skirt_size /= (float)(2 << d);
        }
        const float *Hs = stage_heights ? s_h : H;
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            int vx, vy; float skirt;
            if (i < n) { vx = i; vy = 0; skirt = 1.0f; }
            else if (i < n + n * w) {
                int r = i - n; vy = r / w; int c = r - vy * w;
                if (c == 0) { vx = 0; skirt = 1.0f; }
                else if (c == w - 1) { vx = n - 1; skirt = 1.0f; }
                else { vx = c - 1; skirt = 0.0f; }
            } else { vx = i - n - n * w; vy = n - 1; skirt = 1.0f; }

            const Column col = s_col[vx];
            V v = interpolate(col.p, col.q, s_uv[vy]);               // main.cpp:356
            const int tx = vx + 1, ty = vy + 1;
            const float hc = Hs[ty * dim + tx];
            const float height = hc - skirt_size * skirt;            // main.cpp:360
            const float x0 = Hs[ty * dim + tx - 1], x1 = Hs[ty * dim + tx + 1];
            const float y0 = Hs[(ty - 1) * dim + tx], y1 = Hs[(ty + 1) * dim + tx];
            float3 nt = normalize(f3(x0 - x1, 2.0f * col.xyscale, y0 - y1));     // main.cpp:345
            float3 nn = v.n;
            float3 t = normalize(cross(nn, col.pq));                 // main.cpp:363
            float3 bi = normalize(cross(t, nn));                     // main.cpp:364
            float3 N = normalize(t * nt.x + nn * nt.y + bi * nt.z);  // main.cpp:365
            float3 pos = v.p + v.n * height;                         // main.cpp:366
            // fragment stage at the vertex: l = normalize(0,1,-1)   // main.cpp:374-378
            const float inv_sqrt2 = 0.70710678118654752f;
            float lambert = N.y * inv_sqrt2 - N.z * inv_sqrt2;
            float light = 0.001f + fmaxf(lambert, 0.0f);
            int64_t o = qi * nv + i;
            if (pos4) pos4[o] = make_float4(pos.x, pos.y, pos.z, height);
            if (nrm4) nrm4[o] = make_float4(N.x, N.y, N.z, sqrtf(light));
        }
    }
}

} // namespace shade

int launch_shade(const planet_gpu_params *p, const Quad *d_quads, int64_t nquads, const double *cam,
                 const float *d_heights, float max_skirt, float *d_pos4, float *d_nrm4, cudaStream_t stream)
{
    if (nquads == 0) return 0;
    const int n = p->patch_verts;
    if (n < 2 || n > 254)
        return set_error(PLANET_E_UNSUPPORTED, "planet_gpu_shade: patch_verts %d outside [2, 254]", n);
    const int dim = n + 2;
    size_t base = 128 + (size_t)n * sizeof(shade::Column) + (size_t)((n + 3) & ~3) * sizeof(float);
    size_t hbytes = (size_t)dim * dim * sizeof(float);
    int stage = (base + hbytes) <= 200 * 1024;
    size_t smem = base + (stage ? hbytes : 0);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        PLANET_CUDA(cudaFuncSetAttribute(shade::k_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1024)));
    int grid = (int)std::min<int64_t>(nquads, (int64_t)sms * per_sm);
    shade::k_shade<<<grid, shade::THREADS, smem, stream>>>(
        d_quads, nquads, n, cam[0], cam[1], cam[2], d_heights, max_skirt,
        reinterpret_cast<float4 *>(d_pos4), reinterpret_cast<float4 *>(d_nrm4), stage);
    count_launch();
    return check_cuda(cudaGetLastError(), "shade kernel launch");
}

} // namespace planet
