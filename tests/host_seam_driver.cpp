// Test driver for the C++ host mirror: fills a reference-layout HeightMapGenerator with the GPU
// entry points and calls it the way the reference's two callers do --
//   GetHeightMapForQuad: float data[dim*dim]; hmap_gen.GenerateHeightMap(data, dim, q, max_lod)   main.cpp:243-244
//   ProcessQuad:         hmap_gen.GetHeightAt(q.p[i], 0, 1)                                       main.cpp:552
// usage: host_seam_driver <quads.bin> <max_lod> ; writes raw floats to stdout:
//   for every quad: 32*32 heights, then 4 corner heights from GetHeightAt(p, 0, 1).
#include <cstdlib>
#include <vector>

#include "../planet_b200/host/planet_host.h"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<Quad> quads;
    Quad q;
    while (fread(&q, sizeof q, 1, f) == 1) quads.push_back(q);
    fclose(f);
    const int max_lod = atoi(argv[2]);

    HeightMapGenerator hmap_gen = CreateGpuHeightMapGenerator();        // main.cpp:843
    if (!hmap_gen.GenerateHeightMap) return 3;

    const int dim = 32;                                                  // main.cpp:194
    for (const Quad &quad : quads) {
        float data[dim * dim];                                           // main.cpp:243
        hmap_gen.GenerateHeightMap(data, dim, quad, max_lod);            // main.cpp:244
        fwrite(data, sizeof(float), dim * dim, stdout);
        float corner[4];
        for (int i = 0; i < 4; i++) corner[i] = hmap_gen.GetHeightAt(quad.p[i], 0, 1);   // main.cpp:552
        fwrite(corner, sizeof(float), 4, stdout);
    }
    planet_gpu_shutdown();
    return 0;
}
