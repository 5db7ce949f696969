// planet_call_surface.cuh -- the reference's call surface for this path, callable from CUDA.
//
// Same names, argument meaning and results as the reference, so terrain code written against
// perlin.h / vec3.h / math.h moves into a kernel unchanged:
//   float PerlinNoise3(double x, double y, double z)                              perlin.h:4, 50-88
//   float PerlinfBm   (double x, double y, double z, double lacunarity, float gain, int octaves)   main.cpp:689-707
//   float PerlinRidged(double x, double y, double z, double lacunarity, float gain, int octaves)   main.cpp:709-734
//   Vec3 / Vec3d with + - * / (vec3.h:25-44), Dot LengthSq Length Normalize SafeNormalize Cross Slerp
//   (vec3.h:46-72), V3 / V3d (math.h:47-70)
// The noise functions are bit-identical to the reference (unfused IEEE arithmetic in its
// evaluation order); they read the permutation and gradient tables from global memory through
// the read-only path, so they need no set-up.  Kernels on the throughput path stage the tables
// in shared memory instead (k2_heights.cu).
#pragma once

#include <cmath>

#include "planet_common.cuh"

#ifdef __CUDACC__

// ---- vec3.h / math.h -----------------------------------------------------------------------
// The reference instantiates one struct twice through macros ("we don't want templates"); here
// it is one template with the reference's two names as aliases.  Every product and sum is a
// separately rounded IEEE operation, as in the reference's x86-64 build: nvcc's default
// (-fmad=true) would contract a.x*b.x + a.y*b.y into an FMA and change the last bit of Dot, Cross
// and everything built on them, so on the device the arithmetic goes through the _rn intrinsics
// (which are never contracted) and the results are bit-identical to vec3.h on the host
// (tests/test_call_surface.py runs them in a kernel against the reference's own vec3.h).
namespace planet { namespace rn {
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float  add(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ float  sub(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ float  mul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ float  div(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ float  root(float a)           { return __fsqrt_rn(a); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double root(double a)          { return __dsqrt_rn(a); }
#else
template <class T> inline T add(T a, T b) { return a + b; }
template <class T> inline T sub(T a, T b) { return a - b; }
template <class T> inline T mul(T a, T b) { return a * b; }
template <class T> inline T div(T a, T b) { return a / b; }
template <class T> inline T root(T a)     { return std::sqrt(a); }
#endif
} }

template <class T> struct TVec3
{
    T x, y, z;
    __host__ __device__ TVec3 &operator+=(TVec3 v) { using namespace planet::rn; x = add(x, v.x); y = add(y, v.y); z = add(z, v.z); return *this; }
    __host__ __device__ TVec3 &operator-=(TVec3 v) { using namespace planet::rn; x = sub(x, v.x); y = sub(y, v.y); z = sub(z, v.z); return *this; }
    __host__ __device__ TVec3 &operator*=(T s) { using namespace planet::rn; x = mul(x, s); y = mul(y, s); z = mul(z, s); return *this; }
    __host__ __device__ TVec3 &operator/=(T s) { using namespace planet::rn; x = div(x, s); y = div(y, s); z = div(z, s); return *this; }
};
typedef TVec3<float> Vec3;
typedef TVec3<double> Vec3d;

template <class T> __host__ __device__ inline TVec3<T> operator+(TVec3<T> a, TVec3<T> b) { return a += b; }
template <class T> __host__ __device__ inline TVec3<T> operator-(TVec3<T> a, TVec3<T> b) { return a -= b; }
template <class T> __host__ __device__ inline TVec3<T> operator*(TVec3<T> a, T b) { return a *= b; }
template <class T> __host__ __device__ inline TVec3<T> operator*(T a, TVec3<T> b) { return b *= a; }
template <class T> __host__ __device__ inline TVec3<T> operator/(TVec3<T> a, T b) { return a /= b; }
template <class T> __host__ __device__ inline TVec3<T> operator-(TVec3<T> v) { return v *= T(-1.0); }

template <class T> __host__ __device__ inline T Dot(TVec3<T> a, TVec3<T> b)                         // vec3.h:46: (x + y) + z
{
    using namespace planet::rn;
    return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z));
}
template <class T> __host__ __device__ inline T LengthSq(TVec3<T> v) { return Dot(v, v); }
template <class T> __host__ __device__ inline T Length(TVec3<T> v) { return planet::rn::root(LengthSq(v)); }
template <class T> __host__ __device__ inline TVec3<T> Normalize(TVec3<T> v) { return v / Length(v); }
template <class T> __host__ __device__ inline TVec3<T> SafeNormalize(TVec3<T> v, T epsilon = T(0.0001))
{
    T len2 = LengthSq(v);
    if (len2 < epsilon) return TVec3<T>{ T(0), T(0), T(0) };
    return v / planet::rn::root(len2);
}
template <class T> __host__ __device__ inline TVec3<T> Cross(TVec3<T> a, TVec3<T> b)
{
    using namespace planet::rn;
    return TVec3<T>{ sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)), sub(mul(a.x, b.y), mul(a.y, b.x)) };
}
// vec3.h:68-72.  theta is a float whatever T is; the sines take T((1 - t) * theta).  acos and sin
// are the device math library's here (<= 2 ulp from glibc's), so Slerp agrees with the host to a
// few ulp, not bit for bit -- the one function of this surface with a tolerance.
template <class T> __host__ __device__ inline TVec3<T> Slerp(TVec3<T> a, TVec3<T> b, T t)
{
    using namespace planet::rn;
    float theta = std::acos(div(Dot(a, b), mul(Length(a), Length(b))));
    T th = T(theta);
    return (T(std::sin(mul(sub(T(1.0), t), th))) * a + T(std::sin(mul(t, th))) * b) / T(std::sin(theta));      // Sin(float): the divisor is a float sine in both instantiations
}
__host__ __device__ inline Vec3 V3(float x, float y, float z) { return Vec3{ x, y, z }; }
__host__ __device__ inline Vec3 V3(float s) { return Vec3{ s, s, s }; }
__host__ __device__ inline Vec3 V3(Vec3d v) { return Vec3{ (float)v.x, (float)v.y, (float)v.z }; }
__host__ __device__ inline Vec3d V3d(double x, double y, double z) { return Vec3d{ x, y, z }; }
__host__ __device__ inline Vec3d V3d(double s) { return Vec3d{ s, s, s }; }

// ---- perlin.h / main.cpp:686-734 -----------------------------------------------------------
__device__ __forceinline__ int PerlinRandom(int seed) { return __ldg(&planet::g_perm[seed & 255]); }   // perlin.h:38-41

__device__ __forceinline__ float PerlinGradient(float x, float y, float z, int ix, int iy, int iz)     // perlin.h:43-48
{
    int h = PerlinRandom(PerlinRandom(PerlinRandom(ix) + iy) + iz);
    const float *g = planet::g_grad[h & 15];
    float s = __fmul_rn(x, __ldg(g));
    s = __fadd_rn(s, __fmul_rn(y, __ldg(g + 1)));
    return __fadd_rn(s, __fmul_rn(z, __ldg(g + 2)));
}

__device__ inline float PerlinNoise3(double x, double y, double z)                                     // perlin.h:50-88
{
    using namespace planet::exact;
    int ix = cell(x), iy = cell(y), iz = cell(z);
    x = __dsub_rn(x, (double)ix); y = __dsub_rn(y, (double)iy); z = __dsub_rn(z, (double)iz);
    float u = fade(x), v = fade(y), w = fade(z);
    float x0 = __double2float_rn(x), x1 = __double2float_rn(__dadd_rn(x, -1.0));
    float y0 = __double2float_rn(y), y1 = __double2float_rn(__dadd_rn(y, -1.0));
    float z0 = __double2float_rn(z), z1 = __double2float_rn(__dadd_rn(z, -1.0));
    float g0 = PerlinGradient(x0, y0, z0, ix,     iy,     iz);
    float g1 = PerlinGradient(x1, y0, z0, ix + 1, iy,     iz);
    float g2 = PerlinGradient(x0, y1, z0, ix,     iy + 1, iz);
    float g3 = PerlinGradient(x1, y1, z0, ix + 1, iy + 1, iz);
    float g4 = PerlinGradient(x0, y0, z1, ix,     iy,     iz + 1);
    float g5 = PerlinGradient(x1, y0, z1, ix + 1, iy,     iz + 1);
    float g6 = PerlinGradient(x0, y1, z1, ix,     iy + 1, iz + 1);
    float g7 = PerlinGradient(x1, y1, z1, ix + 1, iy + 1, iz + 1);
    float a0 = lerp(g0, g1, u), a1 = lerp(g2, g3, u), a2 = lerp(g4, g5, u), a3 = lerp(g6, g7, u);
    return lerp(lerp(a0, a1, v), lerp(a2, a3, v), w);
}

__device__ inline float PerlinfBm(double x, double y, double z, double lacunarity, float gain, int octaves)   // main.cpp:689-707
{
    double frequency = 1.0;
    float amplitude = 1.0f, value = 0.0f;
    for (int i = 0; i < octaves; ++i) {
        float n = PerlinNoise3(__dmul_rn(x, frequency), __dmul_rn(y, frequency), __dmul_rn(z, frequency));
        value = __fadd_rn(value, __fmul_rn(n, amplitude));
        frequency = __dmul_rn(frequency, lacunarity);
        amplitude = __fmul_rn(amplitude, gain);
    }
    return value;
}

__device__ inline float PerlinRidged(double x, double y, double z, double lacunarity, float gain, int octaves)   // main.cpp:709-734
{
    double frequency = 1.0;
    float amplitude = 1.0f, weight = 1.0f, value = 0.0f;
    for (int i = 0; i < octaves; ++i) {
        float v = PerlinNoise3(__dmul_rn(x, frequency), __dmul_rn(y, frequency), __dmul_rn(z, frequency));
        v = (v < 0.0f) ? -v : v;
        v = __fsub_rn(1.0f, v);
        v = __fmul_rn(v, v);
        value = __fadd_rn(value, __fmul_rn(__fmul_rn(v, amplitude), weight));
        weight = v;
        frequency = __dmul_rn(frequency, lacunarity);
        amplitude = __fmul_rn(amplitude, gain);
    }
    return value;
}

#endif // __CUDACC__
