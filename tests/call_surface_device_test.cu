// The vector call surface (planet_b200/csrc/planet_call_surface.cuh) evaluated IN A KERNEL.
// stdin: int64 n, then a[n][3], b[n][3], t[n] as float64.  stdout: out64[n][33], out32[n][33] with
// the layout of ref_vec3_ops_* in oracle/ref_oracle.cpp (Dot, LengthSq, Length, Normalize,
// SafeNormalize, Cross, Slerp, a+b, a-b, a*t, t*a, a/t, -a), the f32 run on the inputs cast to
// float.  With the argument "host" the same functions run on the CPU.  tests/test_call_surface.py compares both with what the reference's vec3.h gave.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../planet_b200/csrc/planet_call_surface.cuh"

template <class T> __host__ __device__ void vec3_ops(const double *a, const double *b, const double *t, long i, T *out)
{
    TVec3<T> A = { T(a[3*i]), T(a[3*i+1]), T(a[3*i+2]) }, B = { T(b[3*i]), T(b[3*i+1]), T(b[3*i+2]) };
    T s = T(t[i]);
    T *o = out + 33 * i;
    TVec3<T> r[10] = { Normalize(A), SafeNormalize(A), Cross(A, B), Slerp(A, B, s), A + B, A - B, A * s, s * A, A / s, -A };
    o[0] = Dot(A, B); o[1] = LengthSq(A); o[2] = Length(A);
    for (int k = 0; k < 10; k++) { o[3 + 3*k] = r[k].x; o[4 + 3*k] = r[k].y; o[5 + 3*k] = r[k].z; }
}

template <class T> __global__ void k_vec3_ops(const double *a, const double *b, const double *t, long n, T *out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) vec3_ops<T>(a, b, t, i, out);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char **argv)
{
    const bool on_host = argc > 1 && argv[1][0] == 'h';      // "host": the same functions, compiled for the CPU
    int64_t n = 0;
    if (fread(&n, sizeof n, 1, stdin) != 1 || n <= 0 || n > (1 << 24)) return 1;
    std::vector<double> in((size_t)n * 7);
    if (fread(in.data(), sizeof(double), in.size(), stdin) != in.size()) return 1;
    std::vector<double> o64((size_t)n * 33); std::vector<float> o32((size_t)n * 33);
    if (on_host) {
        const double *a = in.data(), *b = a + 3 * n, *t = a + 6 * n;
        for (long i = 0; i < n; i++) { vec3_ops<double>(a, b, t, i, o64.data()); vec3_ops<float>(a, b, t, i, o32.data()); }
        fwrite(o64.data(), sizeof(double), o64.size(), stdout);
        fwrite(o32.data(), sizeof(float), o32.size(), stdout);
        return 0;
    }
    double *d_in, *d_o64; float *d_o32;
    CK(cudaMalloc(&d_in, in.size() * sizeof(double)));
    CK(cudaMalloc(&d_o64, (size_t)n * 33 * sizeof(double)));
    CK(cudaMalloc(&d_o32, (size_t)n * 33 * sizeof(float)));
    CK(cudaMemcpy(d_in, in.data(), in.size() * sizeof(double), cudaMemcpyHostToDevice));
    const double *a = d_in, *b = d_in + 3 * n, *t = d_in + 6 * n;
    int blocks = (int)((n + 127) / 128);
    k_vec3_ops<double><<<blocks, 128>>>(a, b, t, n, d_o64);
    k_vec3_ops<float><<<blocks, 128>>>(a, b, t, n, d_o32);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o64.data(), d_o64, o64.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o32.data(), d_o32, o32.size() * sizeof(float), cudaMemcpyDeviceToHost));
    fwrite(o64.data(), sizeof(double), o64.size(), stdout);
    fwrite(o32.data(), sizeof(float), o32.size(), stdout);
    return 0;
}
