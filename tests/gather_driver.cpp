// Multi-process test driver for K4 (planet_gpu_gather_*): one process per GPU, no Python, no torch.
//
//   gather_driver <world> <depth> <mode>      mode: fast | exact | ragged | nccl | ce | smpush | split | split5of8 | concurrent | concurrent_ragged
//
// The parent forks `world` ranks before any CUDA call.  Rank 0 writes the NCCL unique id to a file
// in a scratch directory, the others read it (the C-ABI leaves the transport to the caller).  The
// quads are the 6 * 4^depth leaves of the uniform tree in the reference's emission order
// (main.cpp:589-592, 604-624), split by patch range [rank * Q / world, (rank + 1) * Q / world)
// (`ragged`: an uneven split).  Every rank then
//   1. runs K1 + the fused K2/K4 kernel for its shard, three steps in a row on two buffers (so the
//      release flags and the buffer rotation are exercised), and waits for the peers' shards;
//   2. computes ALL quads' height maps on its own GPU with the plain K2 call;
//   3. checks that the gathered buffer equals that buffer byte for byte (sharding and gathering
//      are invisible in the result).
// `nccl` does step 1 with the plain K2 call into the local buffer + planet_gpu_gather_nccl, `ce` with
// planet_gpu_gather_begin / _push / _publish (chunks handed to the copy engines).
// Exit code 0 and one line "gather ok ..." per rank on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include <cuda_runtime_api.h>

#include "../include/planet_gpu.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "[rank %d] %s: %s\n", rank, #call, planet_gpu_last_error()); return 10; } } while (0)
#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "[rank %d] %s: %s\n", rank, #call, cudaGetErrorString(e_)); return 11; } } while (0)

static int run_rank(int rank, int world, int depth, const std::string &mode, const std::string &dir)
{
    CHECK(planet_gpu_init(rank));
    unsigned char id[PLANET_GATHER_ID_BYTES] = {};
    const std::string path = dir + "/nccl_id", tmp = path + ".tmp";
    if (world > 1) {
        if (rank == 0) {
            CHECK(planet_gpu_gather_unique_id(id));
            FILE *f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) return 12;
            fclose(f);
            rename(tmp.c_str(), path.c_str());
        } else {
            FILE *f = nullptr;
            for (int tries = 0; tries < 3000 && !(f = fopen(path.c_str(), "rb")); tries++) usleep(10000);
            if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "[rank %d] no unique id\n", rank); return 12; }
            fclose(f);
        }
    }

    planet_gpu_params params;
    planet_gpu_default_params(&params);
    params.noise_kind = PLANET_NOISE_FBM; params.gain = 0.5f; params.fixed_octaves = 8;
    params.precision = mode == "exact" ? PLANET_PRECISION_EXACT : PLANET_PRECISION_FAST;
    const int dim = 32, max_lod = 18;
    const int64_t Q = 6ll << (2 * depth), texels = (int64_t)dim * dim;
    // patch-range partition; `ragged` moves the boundaries off the even split
    std::vector<int64_t> lo(world + 1);
    for (int r = 0; r <= world; r++) lo[r] = Q * r / world;
    if (mode == "ragged" || mode == "concurrent_ragged") for (int r = 1; r < world; r++) lo[r] += 37 * r + 1;
    const int64_t first = lo[rank], n = lo[rank + 1] - lo[rank];

    void *g = planet_gpu_gather_create(world > 1 ? id : nullptr, rank, world, Q * texels * sizeof(float), 2);
    if (!g) { fprintf(stderr, "[rank %d] gather_create: %s\n", rank, planet_gpu_last_error()); return 13; }

    planet_gpu_quad *d_quads = nullptr, *d_all = nullptr;
    float *d_want = nullptr;
    CUDA_OK(cudaMalloc((void **)&d_quads, sizeof(planet_gpu_quad) * n));
    CUDA_OK(cudaMalloc((void **)&d_all, sizeof(planet_gpu_quad) * Q));
    CUDA_OK(cudaMalloc((void **)&d_want, sizeof(float) * Q * texels));
    cudaStream_t stream;
    CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CHECK(planet_gpu_tessellate_uniform(&params, depth, first, n, d_quads, nullptr, stream));

    int which = 0;
    if (mode == "nccl") {
        float *buf = planet_gpu_gather_buffer(g, 0);
        CHECK(planet_gpu_generate_height_maps(&params, d_quads, n, dim, max_lod, buf + first * texels, stream));
        std::vector<int64_t> off(world), size(world);
        for (int r = 0; r < world; r++) { off[r] = lo[r] * texels * 4; size[r] = (lo[r + 1] - lo[r]) * texels * 4; }
        CHECK(planet_gpu_gather_nccl(g, 0, off.data(), size.data(), stream));
    } else if (mode == "split" || mode == "split5of8") {
        // the NVLink transfer spread over K2 and K3: every 4th map travels from the shade kernel
        float *d_pos = nullptr, *d_nrm = nullptr;
        const int nv = planet_gpu_patch_vertex_count(params.patch_verts);
        CUDA_OK(cudaMalloc((void **)&d_pos, sizeof(float) * 4 * nv * n));
        CUDA_OK(cudaMalloc((void **)&d_nrm, sizeof(float) * 4 * nv * n));
        const double cam[3] = { 0.0, 0.0, -6371010.0 };
        CHECK(planet_gpu_gather_set_shade_share(g, mode == "split" ? 4 : -5));      // every 4th map / 5 maps in 8
        for (int step = 0; step < 3; step++) {
            params.seed_offset[1] = 0.25 * step;
            CHECK(planet_gpu_gather_height_maps(g, &params, d_quads, n, first, dim, max_lod, stream));
            CHECK(planet_gpu_gather_shade(g, &params, d_quads, n, first, cam, -1.0f, d_pos, d_nrm, stream));
            CHECK(planet_gpu_gather_wait(g, /*release*/ 1, stream));
        }
        which = planet_gpu_gather_last_buffer(g);
        CUDA_OK(cudaStreamSynchronize(stream));
        cudaFree(d_pos); cudaFree(d_nrm);
    } else if (mode == "ce" || mode == "smpush") {
        // plain K2 in four chunks, every finished chunk pushed to the peers by the copy engines (ce) or by
        // the library's pusher kernel on its side stream (smpush)
        if (mode == "smpush") CHECK(planet_gpu_gather_set_push_mode(g, PLANET_GATHER_PUSH_SM_KERNEL));
        for (int step = 0; step < 3; step++) {
            params.seed_offset[1] = 0.25 * step;
            CHECK(planet_gpu_gather_begin(g, stream));
            float *buf = planet_gpu_gather_buffer(g, planet_gpu_gather_last_buffer(g));
            for (int c = 0; c < 4; c++) {
                const int64_t a = n * c / 4, b = n * (c + 1) / 4;
                CHECK(planet_gpu_generate_height_maps(&params, d_quads + a, b - a, dim, max_lod, buf + (first + a) * texels, stream));
                CHECK(planet_gpu_gather_push(g, (first + a) * texels * 4, (b - a) * texels * 4, stream));
            }
            CHECK(planet_gpu_gather_publish(g));
            CHECK(planet_gpu_gather_wait(g, /*release*/ 1, stream));
        }
        which = planet_gpu_gather_last_buffer(g);
    } else {
        // concurrent: one K2 launch publishing its progress, the pusher kernel following it on the side stream
        if (mode == "concurrent" || mode == "concurrent_ragged") CHECK(planet_gpu_gather_set_push_mode(g, PLANET_GATHER_PUSH_CONCURRENT));
        for (int step = 0; step < 3; step++) {
            // every step computes a different terrain (seed offset), so a step satisfied by an earlier
            // step's data -- a missed release wait, a wrong buffer rotation -- fails the comparison below
            params.seed_offset[1] = 0.25 * step;
            CHECK(planet_gpu_gather_height_maps(g, &params, d_quads, n, first, dim, max_lod, stream));
            CHECK(planet_gpu_gather_wait(g, /*release*/ 1, stream));
        }
        which = planet_gpu_gather_last_buffer(g);
    }
    CUDA_OK(cudaStreamSynchronize(stream));
    CHECK(planet_gpu_gather_error(g));

    // the unsharded answer, computed here
    CHECK(planet_gpu_tessellate_uniform(&params, depth, 0, Q, d_all, nullptr, stream));
    CHECK(planet_gpu_generate_height_maps(&params, d_all, Q, dim, max_lod, d_want, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    std::vector<float> got((size_t)(Q * texels)), want((size_t)(Q * texels));
    CUDA_OK(cudaMemcpy(got.data(), planet_gpu_gather_buffer(g, which), got.size() * 4, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(want.data(), d_want, want.size() * 4, cudaMemcpyDeviceToHost));
    int64_t bad = 0, first_bad = -1;
    for (size_t i = 0; i < got.size(); i++)
        if (memcmp(&got[i], &want[i], 4) != 0) { if (first_bad < 0) first_bad = (int64_t)i; bad++; }
    CHECK(planet_gpu_gather_barrier(g, stream));
    planet_gpu_gather_destroy(g);
    cudaFree(d_quads); cudaFree(d_all); cudaFree(d_want);
    planet_gpu_shutdown();
    if (bad) { fprintf(stderr, "[rank %d] %lld of %zu floats differ, first at %lld (quad %lld)\n", rank, (long long)bad, got.size(), (long long)first_bad, (long long)(first_bad / texels)); return 20; }
    printf("gather ok rank %d/%d mode %s: %lld quads gathered, my shard [%lld, %lld), %lld kernel launches\n", rank, world,
           mode.c_str(), (long long)Q, (long long)first, (long long)(first + n), (long long)planet_gpu_launch_count());
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: gather_driver <world> <depth> fast|exact|ragged|nccl|ce|split\n"); return 2; }
    const int world = atoi(argv[1]), depth = atoi(argv[2]);
    const std::string mode = argv[3];
    char dir[] = "/tmp/planet_gather_XXXXXX";
    if (!mkdtemp(dir)) return 2;
    std::vector<pid_t> kids;
    for (int r = 0; r < world; r++) {
        pid_t pid = fork();                                   // before any CUDA call: every rank gets a fresh context
        if (pid == 0) { const int code = run_rank(r, world, depth, mode, dir); fflush(stdout); fflush(stderr); _exit(code); }
        kids.push_back(pid);
    }
    int worst = 0;
    for (pid_t pid : kids) {
        int status = 0;
        waitpid(pid, &status, 0);
        const int code = WIFEXITED(status) ? WEXITSTATUS(status) : 99;
        if (code > worst) worst = code;
    }
    unlink((std::string(dir) + "/nccl_id").c_str());
    rmdir(dir);
    return worst;
}
