// planet_frame.cpp -- the reference's frame, headless and GPU-resident, in its own host language.
//
// Mirrors the structure of the reference's main.cpp with the terrain path moved behind the C-ABI:
//   InitPlanet   (main.cpp:280-516)  -> patch mesh + strip indices from K1, max_lod, skirt size
//   RenderPlanet (main.cpp:600-683)  -> K0 LOD selection, height-map cache (bookkeeping in one kernel on
//                                       device tables + one batched K2 launch per frame), K3 displacement
//                                       + normals through texrects; the leaf quads never visit the host
// No window, no GL: the vertex/normal/index buffers stay in device memory, which is where a
// CUDA-GL interop draw would read them.  Prints what the reference's title bar shows
// (main.cpp:1029-1037) plus timings through the reference's timing.h macro names.
//
// Build (from planet_b200/host):
//   g++ -O2 -std=c++17 planet_frame.cpp -I../../include -I/usr/local/cuda/include -L.. -lplanet_gpu
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/.. -o planet_frame
// Run: ./planet_frame [frames]
#include <cmath>
#include <cstdlib>
#include <vector>

#include <cuda_runtime_api.h>

#include "planet_host.h"
#include "planet_buffers.h"

struct Planet                                                        // main.cpp:161-181, GPU-resident
{
    double radius;
    int max_lod;
    float max_skirt_size;
    int patch_verts, vertex_count, index_count;
    GLuint vertex_buffer, index_buffer;                              // main.cpp:479-480, device memory here
    float *d_patch_vertices;                                         // (u, v, skirt) triples, main.cpp:402-425
    uint32_t *d_patch_indices;                                       // triangle strip, main.cpp:427-474
    planet_gpu_params params;
    void *cache;                                                     // HeightMapCache, main.cpp:78-84
    Quad *d_quads;                                                   // List<Quad> quads (main.cpp:178), device-resident
    planet_gpu_texrect *d_rects; float *d_pos4, *d_nrm4;
    int64_t capacity;
};

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { LOG_ERROR("%s: %s", #call, planet_gpu_last_error()); return false; } } while (0)
#define CUDA_OK(call) do { if ((call) != cudaSuccess) { LOG_ERROR("%s failed", #call); return false; } } while (0)

static bool InitPlanet(Planet &p, double radius)                     // main.cpp:280
{
    TIMED_FUNCTION();
    planet_gpu_default_params(&p.params);                            // the reference's Perlin functor, bit-exact
    p.params.radius = radius;
    p.radius = radius;
    p.patch_verts = p.params.patch_verts;                            // main.cpp:391
    p.vertex_count = planet_gpu_patch_vertex_count(p.patch_verts);   // main.cpp:393-394
    p.index_count = planet_gpu_patch_index_count(p.patch_verts);     // main.cpp:395-400
    CHECK(planet_gpu_init(0));
    // main.cpp:479-480 create the two buffers from CPU arrays; here K1 fills them in place
    p.vertex_buffer = CreateVertexBuffer(sizeof(float) * 3 * p.vertex_count, nullptr);
    p.index_buffer = CreateIndexBuffer(sizeof(uint32_t) * p.index_count, nullptr);
    if (!p.vertex_buffer || !p.index_buffer) { LOG_ERROR("CreateVertexBuffer/CreateIndexBuffer failed"); return false; }
    p.d_patch_vertices = (float *)MapBuffer(p.vertex_buffer);
    p.d_patch_indices = (uint32_t *)MapBuffer(p.index_buffer);
    CHECK(planet_gpu_patch_mesh(p.patch_verts, p.d_patch_vertices, p.d_patch_indices, nullptr));   // main.cpp:402-481
    p.max_lod = planet_gpu_max_lod(radius, p.patch_verts);           // main.cpp:497
    p.max_skirt_size = planet_gpu_max_skirt_size(radius, p.patch_verts);   // main.cpp:500
    p.cache = planet_gpu_cache_create(p.patch_verts + 2, 1024, 1499, 4096);   // main.cpp:75-76, 194
    if (!p.cache) { LOG_ERROR("cache: %s", planet_gpu_last_error()); return false; }
    p.capacity = 1 << 16;
    CUDA_OK(cudaMalloc((void **)&p.d_quads, sizeof(Quad) * p.capacity));
    CUDA_OK(cudaMalloc((void **)&p.d_rects, sizeof(planet_gpu_texrect) * p.capacity));
    CUDA_OK(cudaMalloc((void **)&p.d_pos4, sizeof(float) * 4 * p.vertex_count * 4096));
    CUDA_OK(cudaMalloc((void **)&p.d_nrm4, sizeof(float) * 4 * p.vertex_count * 4096));
    return true;
}

static bool RenderPlanet(Planet &planet, const Vec3d &cam_position)  // main.cpp:600
{
    TIMED_FUNCTION();
    const double cam[3] = { cam_position.x, cam_position.y, cam_position.z };

    BEGIN_TIMED_BLOCK(ProcessQuads);                                 // main.cpp:604-624 + ProcessQuad
    int64_t n = 0;
    CHECK(planet_gpu_select_lod(&planet.params, cam, planet.max_lod, (planet_gpu_quad *)planet.d_quads,
                                planet.capacity, &n, nullptr));
    if (n > 4096) { LOG_ERROR("%lld leaf quads exceed the demo's vertex buffers", (long long)n); return false; }
    END_TIMED_BLOCK(ProcessQuads);                                   // the leaves stay on the device

    BEGIN_TIMED_BLOCK(HeightMaps);                                   // main.cpp:652-660, generations_per_frame = 100
    int64_t generated = 0;                                           // bookkeeping on the device, one batched K2 launch for the misses
    CHECK(planet_gpu_cache_frame_device(planet.cache, &planet.params, (const planet_gpu_quad *)planet.d_quads, n,
                                        planet.max_lod, 100, planet.d_rects, &generated, nullptr));
    END_TIMED_BLOCK(HeightMaps);

    BEGIN_TIMED_BLOCK(Draw);                                         // main.cpp:662-679 + the GLSL stage
    CHECK(planet_gpu_shade_cached(&planet.params, (const planet_gpu_quad *)planet.d_quads, n, cam,
                                  planet_gpu_cache_pool(planet.cache), planet.d_rects, planet.max_skirt_size,
                                  planet.d_pos4, planet.d_nrm4, nullptr));
    CUDA_OK(cudaDeviceSynchronize());
    END_TIMED_BLOCK(Draw);

    // for the statistics line only (the reference prints tri_count, main.cpp:1030): which quads borrowed a parent's map
    std::vector<planet_gpu_texrect> rects((size_t)n);
    CUDA_OK(cudaMemcpy(rects.data(), planet.d_rects, sizeof(planet_gpu_texrect) * n, cudaMemcpyDeviceToHost));
    int fallback = 0;
    for (const auto &r : rects) fallback += r.flags == PLANET_TEXRECT_PARENT;
    const int tri_count = (int)n * (planet.patch_verts - 1) * (planet.patch_verts - 1) * 2;   // main.cpp:1030
    printf("tris: %d, quads: %d, generated: %d, parent fallback: %d, cached: %d\n", tri_count, (int)n, (int)generated,
           fallback, planet_gpu_cache_count(planet.cache));
    return true;
}

int main(int argc, char **argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 6;
    print_timings = true;                                            // the reference toggles this with `T` (main.cpp:996)
    const double radius = 6371000.0;                                 // main.cpp:821
    Planet planet = {};
    if (!InitPlanet(planet, radius)) return 1;
    printf("max_lod: %d, max_skirt_size: %.3f, patch: %d verts / %d indices\n", planet.max_lod, planet.max_skirt_size,
           planet.vertex_count, planet.index_count);

    Vec3d cam = { 0.0, 0.0, -radius - 10.0 };                        // main.cpp:864
    for (int f = 0; f < frames; f++) {
        if (!RenderPlanet(planet, cam)) return 1;
        // fly along the surface: 20 km per frame eastwards, staying 10 m above the sphere
        double a = 20000.0 * (f + 1) / radius;
        cam = { std::sin(a) * (radius + 10.0), 0.0, -std::cos(a) * (radius + 10.0) };
    }
    // first displaced vertex and normal of the last frame, to show the buffers are real
    float v[8];
    cudaMemcpy(v, planet.d_pos4, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(v + 4, planet.d_nrm4, 16, cudaMemcpyDeviceToHost);
    printf("vertex 0: pos (%.3f %.3f %.3f) height %.3f  normal (%.4f %.4f %.4f) light %.4f\n",
           v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    planet_gpu_cache_destroy(planet.cache);
    DeleteBuffers();
    planet_gpu_shutdown();
    return 0;
}
