// planet_host.h -- C++ host mirror of the reference's call surface for the terrain path.
//
// The reference is C++ (main.cpp, render.cpp); a maintainer switching its height generation to
// the GPU keeps its own types and only changes which two function pointers go into
// `HeightMapGenerator` (main.cpp:107-111).  This header restates exactly those types --
// layout-compatible with the reference's, so it can also be used stand-alone by a host program
// that does not include main.cpp -- and provides CreateGpuHeightMapGenerator(), the GPU
// counterpart of CreateHeightMapGenerator<F>() (main.cpp:113-158).
//
// Nothing here computes terrain on the CPU: both pointers forward to libplanet_gpu.so.
#ifndef PLANET_HOST_H
#define PLANET_HOST_H

#include <chrono>
#include <cstdint>
#include <cstdio>

#include "../../include/planet_gpu.h"

#ifndef PLANET_HOST_NO_TYPES   // define when compiling next to the reference's own math.h / main.cpp
struct Vec3d { double x, y, z; };                         // math.h:43-45 / vec3.h:25-33

struct QuadID { uint64_t value; };                        // main.cpp:19-22
inline uint64_t GetRoot(QuadID id)  { return (id.value >> 60) & 7; }            // main.cpp:26
inline uint64_t GetDepth(QuadID id) { return (id.value >> 55) & 31; }           // main.cpp:27
inline uint64_t GetIndex(QuadID id) { return id.value & ((1ull << 55) - 1); }   // main.cpp:28
inline QuadID MakeRootID(uint64_t root) { return { (1ull << 63) | (root << 60) }; }               // main.cpp:32-39
inline QuadID MakeChildID(QuadID id, uint64_t c)                                                   // main.cpp:41-49
{
    return { (id.value + (1ull << 55)) | (c << (2 * GetDepth(id))) };
}

struct Quad { Vec3d p[4]; QuadID id; };                   // main.cpp:68-72
static_assert(sizeof(Quad) == sizeof(planet_gpu_quad), "Quad must keep the reference's 104-byte layout");

struct HeightMapGenerator                                 // main.cpp:107-111
{
    float (*GetHeightAt)(const Vec3d &, int, int);
    void (*GenerateHeightMap)(float *, int, const Quad &, int);
};
#endif

// logging.h:6-7
#ifndef LOG_ERROR
#define LOG_WARNING(fmt, ...) fprintf(stdout, "[WARNING] " fmt "\n", ##__VA_ARGS__)
#define LOG_ERROR(fmt, ...) fprintf(stderr, "[ERROR] " fmt "\n", ##__VA_ARGS__)
#endif

// timing.h:4-30 over std::chrono instead of the SDL performance counter
#ifndef TIMING_H
inline uint64_t GetMicroTicks()
{
    using namespace std::chrono;
    return (uint64_t)duration_cast<microseconds>(steady_clock::now().time_since_epoch()).count();
}
static bool print_timings = false;
struct ScopeTimer
{
    const char *name; uint64_t start_time;
    ScopeTimer(const char *n) : name(n), start_time(GetMicroTicks()) {}
    ~ScopeTimer() { if (print_timings) printf("%-20s %10u us\n", name, uint32_t(GetMicroTicks() - start_time)); }
};
#define PLANET_JOIN__(a, b) a##b
#define PLANET_JOIN_(a, b) PLANET_JOIN__(a, b)
#define TIMED_FUNCTION() ScopeTimer PLANET_JOIN_(_timed_function_, __LINE__)(__FUNCTION__)
#define BEGIN_TIMED_BLOCK(id) uint64_t _timed_block_##id = GetMicroTicks()
#define END_TIMED_BLOCK(id) if (print_timings) printf("%-20s %10u us\n", #id, uint32_t(GetMicroTicks() - _timed_block_##id))
#endif

namespace planet_host {

// the two members of HeightMapGenerator, GPU-backed.  `const T &` and `const T *` are the same
// machine-level argument; these two-line shims make that explicit instead of casting pointers.
inline float GpuGetHeightAt(const Vec3d &p, int depth, int max_depth)                  // main.cpp:118-121
{
    return planet_gpu_get_height_at(&p.x, depth, max_depth);
}
inline void GpuGenerateHeightMap(float *data, int dim, const Quad &q, int max_depth)   // main.cpp:123-151
{
    TIMED_FUNCTION();                                                                   // main.cpp:126
    planet_gpu_generate_height_map(data, dim, &q, max_depth);
}

} // namespace planet_host

// GPU counterpart of CreateHeightMapGenerator<F>() (main.cpp:113-158).  `params` selects the
// functor constants (NULL = the reference's `Perlin`: ridged, gain 0.55, 6 + 12*depth/max_depth
// octaves, bit-exact arithmetic).  Returns a generator whose pointers are NULL if no GPU can be
// initialised -- the caller decides what to do; nothing falls back to the CPU.
inline HeightMapGenerator CreateGpuHeightMapGenerator(const planet_gpu_params *params = nullptr, int device = 0)
{
    HeightMapGenerator result = { nullptr, nullptr };
    if (planet_gpu_init(device) != PLANET_OK) {
        LOG_ERROR("planet_gpu_init: %s", planet_gpu_last_error());
        return result;
    }
    planet_gpu_params p;
    if (params) p = *params; else planet_gpu_default_params(&p);
    if (planet_gpu_set_params(&p) != PLANET_OK) {
        LOG_ERROR("planet_gpu_set_params: %s", planet_gpu_last_error());
        return result;
    }
    result.GetHeightAt = planet_host::GpuGetHeightAt;
    result.GenerateHeightMap = planet_host::GpuGenerateHeightMap;
    return result;
}

#endif // PLANET_HOST_H
