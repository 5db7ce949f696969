/* planet_gpu.h -- C-ABI of the B200 terrain hot path of pgcomp/planet.
 *
 * Drop-in boundary: the reference's only indirection on this
 * path is
 *
 *     struct HeightMapGenerator {                       // main.cpp:107-111
 *         float (*GetHeightAt)(const Vec3d &, int, int);
 *         void  (*GenerateHeightMap)(float *, int, const Quad &, int);
 *     };
 *
 * installed by InitPlanet (main.cpp:280, 495) and called from GetHeightMapForQuad
 * (main.cpp:244) and ProcessQuad (main.cpp:552, 555).  On the SysV/Itanium ABI a
 * `const T &` parameter is passed as `const T *`, so planet_gpu_get_height_at and
 * planet_gpu_generate_height_map below are assignable to those two members (see
 * INTEGRATION.md and planet_b200/host/planet_host.h for the binding).
 *
 * Everything else here is the batched, stream-ordered form of the same
 * computations (device pointers, caller-owned memory, nothing retained), which
 * the single-threaded reference lacks.  Plain C types only: no CUDA, torch or
 * C++ types appear in any signature; `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).
 *
 * Errors: every entry point that can fail returns 0 on success or a negative
 * PLANET_E_* code and records a message retrievable with
 * planet_gpu_last_error().  The two legacy-shaped entry points cannot return a
 * status (reference signatures); on failure they log "[ERROR] ..." to stderr
 * (logging.h:7 convention) and produce NaN.  There is NO CPU fallback anywhere:
 * without a usable CUDA device every compute call fails.
 */
#ifndef PLANET_GPU_H
#define PLANET_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLANET_GPU_ABI_VERSION 1

enum {
    PLANET_OK = 0,
    PLANET_E_NO_DEVICE = -1,      /* no CUDA device / driver */
    PLANET_E_INVALID = -2,        /* precondition violated (dim <= 3, max_depth == 0, depth >= 32, ...) */
    PLANET_E_CUDA = -3,           /* a CUDA runtime call failed; see planet_gpu_last_error() */
    PLANET_E_UNSUPPORTED = -4     /* valid in the reference but outside what this build handles */
};

/* noise kinds: main.cpp:829 (ridged, the reference default), :830 (fBm), :835-841 (ConstantZero) */
enum { PLANET_NOISE_RIDGED = 0, PLANET_NOISE_FBM = 1, PLANET_NOISE_ZERO = 2 };

/* arithmetic modes of the noise kernel */
enum {
    /* bit-exact: fp64 lattice split, fp64 fade, unfused fp32 -- every rounding of
     * perlin.h:50-88 / main.cpp:689-734 reproduced (IEEE _rn intrinsics, no FMA) */
    PLANET_PRECISION_EXACT = 0,
    /* throughput: 64-bit fixed-point lattice split (exact for lacunarity 2.0), fp32
     * fade with FMA; |dh| <= 1e-5 * height_scale * sum(gain^k) (tests state it) */
    PLANET_PRECISION_FAST = 1
};

/* The reference's compile-time constants on this path, as data. */
typedef struct planet_gpu_params {
    double radius;          /* main.cpp:821  6371000.0 */
    int32_t patch_verts;    /* main.cpp:391  30 (height map dim = patch_verts + 2, main.cpp:194) */
    int32_t noise_kind;     /* PLANET_NOISE_* */
    double lacunarity;      /* main.cpp:829  2.0f widened to double */
    float gain;             /* main.cpp:829  0.55f */
    int32_t fixed_octaves;  /* <= 0: main.cpp:827, 6 + 12*depth/max_depth */
    double coord_scale;     /* main.cpp:828  0.00001 */
    float height_scale;     /* main.cpp:831  8848.0f */
    int32_t precision;      /* PLANET_PRECISION_* */
    double seed_offset[3];  /* added to the scaled coordinate; {0,0,0} == the reference */
} planet_gpu_params;

/* main.cpp:68-72: struct Quad { Vec3d p[4]; QuadID id; } -- 104 bytes, same layout */
typedef struct planet_gpu_quad {
    double p[4][3];
    uint64_t id;
} planet_gpu_quad;

/* ---- lifecycle ------------------------------------------------------------------- */
int  planet_gpu_abi_version(void);
void planet_gpu_default_params(planet_gpu_params *out);      /* the reference's defaults, EXACT mode */
/* Selects the device (must be Blackwell: the kernels are sm_100a only).  One device per process:
 * the staging buffers of the host-pointer calls and the LOD scratch belong to it, matching the
 * one-process-per-GPU model the multi-GPU path uses.  Calling any other entry point first
 * adopts the caller's current CUDA device. */
int  planet_gpu_init(int device);
void planet_gpu_shutdown(void);
const char *planet_gpu_last_error(void);
/* name, SM count, boost clock (kHz) and FP32 lanes/SM of the active device: the
 * denominators bench.py prints beside every roofline fraction */
int  planet_gpu_device_info(char *name, int name_cap, int *sm_count, int *clock_khz, int *fp32_lanes_per_sm);

/* ---- the reference-shaped seam (main.cpp:107-111) ----------------------------------- */
/* parameters used by the two legacy-shaped calls below (they have no params argument) */
int   planet_gpu_set_params(const planet_gpu_params *p);
/* replaces Gen::GetHeightAt, main.cpp:118-121.  Host pointer; synchronous. */
float planet_gpu_get_height_at(const double *p, int depth, int max_depth);
/* replaces Gen::GenerateHeightMap, main.cpp:123-151.  `data` is a caller-owned HOST
 * buffer of dim*dim floats, fully overwritten, row-major data[y*dim+x]; `quad` points
 * to a reference-layout Quad (104 bytes).  Synchronous: H2D quad, kernel, D2H heights. */
void  planet_gpu_generate_height_map(float *data, int dim, const void *quad, int max_depth);

/* ---- batched device API (all pointers are DEVICE pointers, work is stream-ordered) --- */
/* K2: nquads height maps of dim*dim floats each, out[q*dim*dim + y*dim + x]
 * (main.cpp:123-151 applied to every quad). */
int planet_gpu_generate_height_maps(const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                    int64_t nquads, int dim, int max_depth, float *d_out,
                                    void *stream);
/* K2 with the multi-GPU gather fused in, for callers that map the peers' buffers themselves
 * (planet_gpu_gather_* below does it for them): identical to planet_gpu_generate_height_maps, and
 * every finished tile is additionally stored to the same offset of up to 7 peer buffers (peer_out:
 * HOST array of n_peers DEVICE pointers, typically this rank's shard inside each peer GPU's
 * gathered buffer, mapped with CUDA IPC / peer access).  The stores travel over NVLink while the
 * kernel computes; the buffers are complete once every rank's kernel has finished. */
int planet_gpu_generate_height_maps_gathered(const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                             int64_t nquads, int dim, int max_depth, float *d_out,
                                             float *const *peer_out, int n_peers, void *stream);
/* ---- K4: multi-GPU gather of finished patches (one process per GPU) -------------- */
/* The reference is a single process; here the leaf quads are split by patch range over the GPUs of
 * one box and the ONE exchange on the path is gathering every rank's finished height maps into one
 * buffer, in the reference's emission order (main.cpp:589-592, 604-624), on every rank.  The
 * gather object is C++ behind this ABI: NCCL (ncclCommInitRank from `id`) for bootstrap, barriers
 * and the plain collective; cudaIpcGetMemHandle / cudaIpcOpenMemHandle for the peer mappings the
 * fused kernel stores through.  All ranks must call create / destroy together. */
#define PLANET_GATHER_ID_BYTES 128
/* rank 0 fills `id` (ncclGetUniqueId) and hands the 128 bytes to the other ranks by any means */
int   planet_gpu_gather_unique_id(void *id);
/* `bytes`: size of ONE gathered buffer (all ranks' shards); n_buffers 1 or 2 (2: a step never waits
 * for the peers to finish reading the previous one).  world 1 needs no id.  NULL on error. */
void *planet_gpu_gather_create(const void *id, int rank, int world, int64_t bytes, int n_buffers);
void  planet_gpu_gather_destroy(void *gather);
float *planet_gpu_gather_buffer(void *gather, int which);          /* DEVICE pointer, this rank's buffer `which` */
int   planet_gpu_gather_last_buffer(const void *gather);           /* buffer the last gather_height_maps filled */
/* K2 + K4 in one kernel: planet_gpu_generate_height_maps for quads [first_quad, first_quad + nquads)
 * of the gathered buffer's quad order, written to that range of the next buffer on THIS rank and,
 * as 512-byte bulk copies over NVLink while the kernel computes, on every peer; then signals the
 * peers.  Stream-ordered, no host synchronisation. */
int   planet_gpu_gather_height_maps(void *gather, const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                    int64_t nquads, int64_t first_quad, int dim, int max_depth, void *stream);
/* Spreading the NVLink transfer over K2 AND K3: with every >= 2, planet_gpu_gather_height_maps leaves
 * every `every`-th map to the shade kernel, which stages each map in shared memory anyway and sends
 * those as one 4 KB bulk copy per peer (0 = off: K2 pushes everything; -1 .. -7: that many maps in every 8 go to the
 * shade kernel, for shares other than 1/every).  planet_gpu_gather_shade is
 * planet_gpu_shade for the quads of the preceding gather_height_maps, reading their maps from the
 * gathered buffer; when a share was left to it, it pushes it and signals the peers (so it MUST
 * follow every gather_height_maps while a share is set). */
int   planet_gpu_gather_set_shade_share(void *gather, int every);
int   planet_gpu_gather_shade(void *gather, const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t nquads,
                              int64_t first_quad, const double *cam_pos, float max_skirt, float *d_pos4, float *d_nrm4,
                              void *stream);
/* The same exchange by the copy engines, for callers that want the transfer to run under the
 * kernels that FOLLOW K2 as well: begin (next buffer; waits, stream-ordered, for the peers' release
 * of it), then any kernels that write bytes of planet_gpu_gather_buffer(g, last_buffer) on
 * `stream`, each followed by push (those bytes go to the same offset of every peer's buffer, one
 * cudaMemcpyAsync per peer on the gather's own streams, ordered behind `stream` by an event), then
 * publish (each peer is signalled as soon as its copies are done), then planet_gpu_gather_wait. */
int   planet_gpu_gather_begin(void *gather, void *stream);
/* who moves the bytes of planet_gpu_gather_push: the copy engines (default), or a small kernel of the
 * library's own on one high-priority side stream (16-byte loads from the local buffer, 16-byte
 * stores to every peer; 128 threads x 32 registers per CTA so that it is resident on the SMs BESIDE
 * the K2 CTAs computing the next chunk -- the SMs reach ~700 GB/s all-to-all where the copy engines
 * reach ~480, and no compute warp waits for the link as in the fused kernel) */
/* PLANET_GATHER_PUSH_CONCURRENT changes planet_gpu_gather_height_maps instead: ONE un-chunked K2 launch
 * that stores to this rank's buffer only and publishes, per warp, how many 512-byte tiles it has
 * finished, and the pusher kernel on the side stream following those counters and sending the
 * finished tiles to every peer while K2 (and then K3) compute on the same SMs; falls back to the
 * fused kernel for batches the big FAST kernel does not take. */
enum { PLANET_GATHER_PUSH_COPY_ENGINES = 0, PLANET_GATHER_PUSH_SM_KERNEL = 1, PLANET_GATHER_PUSH_CONCURRENT = 2 };
int   planet_gpu_gather_set_push_mode(void *gather, int mode);
int   planet_gpu_gather_push(void *gather, int64_t offset_bytes, int64_t size_bytes, void *stream);
int   planet_gpu_gather_publish(void *gather);
/* stream-ordered wait until every peer's shard of the last step has landed in this rank's buffer;
 * release != 0 also tells the peers this rank is done reading it (they may overwrite it n_buffers
 * steps later).  A peer that does not signal within 2 s sets an error (planet_gpu_gather_error). */
int   planet_gpu_gather_wait(void *gather, int release, void *stream);
/* the plain collective, for comparison and for data already in the local buffer: every rank r's
 * bytes [offset_bytes[r], +size_bytes[r]) of buffer `which` are sent to all ranks (ncclAllGather in
 * place for equal shards in rank order, grouped ncclBroadcast otherwise).  HOST arrays of world entries. */
int   planet_gpu_gather_nccl(void *gather, int which, const int64_t *offset_bytes, const int64_t *size_bytes, void *stream);
int   planet_gpu_gather_barrier(void *gather, void *stream);        /* all ranks; synchronises `stream` */
int   planet_gpu_gather_error(void *gather);                        /* 0, or PLANET_E_CUDA after a timed-out wait */

/* batched Gen::GetHeightAt (main.cpp:118-121): n points (xyz doubles), one (depth, max_depth) */
int planet_gpu_heights_at(const planet_gpu_params *p, const double *d_xyz, int64_t n, int depth,
                          int max_depth, float *d_out, void *stream);
/* perlin.h:50-88 / main.cpp:689-734 on n points; kind PLANET_NOISE_FBM or _RIDGED, or
 * octaves == 0 for a single PerlinNoise3 evaluation.  precision as in params. */
int planet_gpu_noise(const double *d_xyz, int64_t n, int kind, double lacunarity, float gain,
                     int octaves, int precision, float *d_out, void *stream);

/* K1: subdivision geometry.  Quads of `nquads` leaves starting at leaf `first` of the
 * uniform depth-`depth` tree over all six root faces, in the order the reference's
 * recursion emits them (face-major, then child 0..3 depth-first; main.cpp:589-592,
 * 619-624).  Corners are bit-identical to the reference's midpoint rule
 * (main.cpp:546-547, 581-594).  d_indices (may be NULL) receives the triangle-strip
 * index buffer of main.cpp:427-474 for every quad, rebased to a merged vertex buffer:
 * d_indices[q*ni + k] = q*nv + strip[k], nv = n*n + 4n, ni = 2n*n + 8n - 4. */
int planet_gpu_tessellate_uniform(const planet_gpu_params *p, int depth, int64_t first,
                                  int64_t nquads, planet_gpu_quad *d_quads, uint32_t *d_indices,
                                  void *stream);
/* The merged strip index buffer of `nquads` patches alone (d_indices[q*ni + k] = q*nv + strip[k],
 * main.cpp:427-474 rebased) by a kernel small enough to be resident BESIDE the height-map kernel:
 * issue it on a second stream next to planet_gpu_generate_height_maps -- K2 keeps the SMs'
 * arithmetic busy and never touches HBM, the index stream is nothing but HBM writes.  Same bytes as
 * planet_gpu_tessellate_uniform(..., NULL, d_indices, ...). */
int planet_gpu_merged_indices_beside(const planet_gpu_params *p, int64_t nquads, uint32_t *d_indices, void *stream);
/* corners of arbitrary quads from their QuadIDs (main.cpp:19-65 encoding) */
int planet_gpu_quads_from_ids(const planet_gpu_params *p, const uint64_t *d_ids, int64_t n,
                              planet_gpu_quad *d_quads, void *stream);
/* the reference's static patch mesh (main.cpp:402-474): nv (u,v,skirt) float triples and
 * ni uint32 strip indices; either pointer may be NULL */
int planet_gpu_patch_mesh(int patch_verts, float *d_vertices, uint32_t *d_indices, void *stream);
int planet_gpu_patch_vertex_count(int patch_verts);   /* n*n + 4n          (main.cpp:393-394) */
int planet_gpu_patch_index_count(int patch_verts);    /* 2n*n + 8n - 4     (main.cpp:395-400) */
/* host-side integer closed forms the kernels use (no device needed):
 *   strip index k of the patch index buffer, main.cpp:427-474;
 *   QuadID of leaf `leaf` of the uniform depth-`depth` tree in emission order */
uint32_t planet_gpu_strip_index(int k, int patch_verts);
uint64_t planet_gpu_uniform_leaf_id(int64_t leaf, int depth);
int   planet_gpu_max_lod(double radius, int patch_verts);          /* main.cpp:497 */
float planet_gpu_max_skirt_size(double radius, int patch_verts);   /* main.cpp:500 */

/* K0 (the caller of the path): camera-driven LOD selection, i.e. the leaf
 * quads RenderPlanet/ProcessQuad (main.cpp:537-624) would put in planet.quads for a camera at
 * cam_pos (HOST pointer, 3 doubles), in the same order, with bit-identical corners and ids.
 * Split decisions evaluate GetHeightAt(p, 0, 1) in EXACT arithmetic whatever p->precision says.
 * d_quads (DEVICE) must hold `capacity` quads; *count (HOST) receives the number of leaves.
 * One cooperative launch walks all levels; `stream` is synchronised once, to read the leaf count. */
int planet_gpu_select_lod(const planet_gpu_params *p, const double *cam_pos, int max_lod,
                          planet_gpu_quad *d_quads, int64_t capacity, int64_t *count, void *stream);

/* K3: the GLSL stage (main.cpp:286-380) for every vertex of every quad, in the patch
 * vertex order of main.cpp:406-422.  d_heights holds the quads' own height maps
 * (dim = patch_verts + 2).  Outputs, nv float4 per quad:
 *   d_pos4 = (v.p + v.n*height, height)           position relative to cam_pos
 *   d_nrm4 = (Normal, sqrt(0.001 + max(0, N.l)))  world normal + Lambert term
 * skirt size per quad follows main.cpp:674-677 from max_skirt (pass a negative value
 * to use planet_gpu_max_skirt_size).  Either output may be NULL. */
int planet_gpu_shade(const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t nquads,
                     const double *cam_pos /* host, 3 doubles */, const float *d_heights,
                     float max_skirt, float *d_pos4, float *d_nrm4, void *stream);

/* ---- height-map residency (main.cpp:75-102, 191-278) ------------------ */
/* What GetHeightMapForQuad returns per quad (TextureRect, main.cpp:184-189), with the GL texture
 * name replaced by a slot of the cache's device pool. */
enum { PLANET_TEXRECT_HIT = 0, PLANET_TEXRECT_GENERATED = 1, PLANET_TEXRECT_PARENT = 2 };
typedef struct planet_gpu_texrect {
    int32_t slot;            /* pool slot holding the map to sample                         */
    int32_t flags;           /* PLANET_TEXRECT_*                                            */
    float corners[4];        /* corners[0].xy, corners[1].xy  (main.cpp:197-198, 232-233)   */
    float pixel_size[2];     /* main.cpp:199, 234                                           */
} planet_gpu_texrect;

/* cache of `cache_max` maps of dim x dim floats in a hash table of `map_max` slots (the reference:
 * 32, 1024, 1499 -- main.cpp:75-76, 194); `extra_slots` bounds the generations of one frame beyond
 * the cache size (evicted slots are recycled one frame later).  Returns NULL on error. */
void *planet_gpu_cache_create(int dim, int cache_max, int map_max, int extra_slots);
void  planet_gpu_cache_destroy(void *cache);
int   planet_gpu_cache_count(const void *cache);
/* One frame of the reference's per-leaf loop (main.cpp:655-660 calling GetHeightMapForQuad) with
 * the leaf quads ALREADY ON THE DEVICE (planet_gpu_select_lod's output): the bookkeeping -- the
 * reference's probe order, LRU choice, fallback rule and budget accounting (generations_per_frame
 * = 100 at main.cpp:653) -- runs in one kernel on device-resident tables, every miss of the frame
 * is generated by ONE batched K2 launch into its pool slot, and d_rects (DEVICE, n entries)
 * receives the texrect each quad draws with, for planet_gpu_shade_cached.  The host reads back two
 * integers (*n_generated, may be NULL, and an error flag); `stream` is synchronised once for that.
 * A frame that fails (pool exhausted, allocation or launch error) leaves the cache unchanged. */
int   planet_gpu_cache_frame_device(void *cache, const planet_gpu_params *p, const planet_gpu_quad *d_quads,
                                    int64_t n, int max_lod, int generations_per_frame,
                                    planet_gpu_texrect *d_rects, int64_t *n_generated, void *stream);
/* the same frame for a caller whose leaf list is in HOST memory: uploads the quads, runs the
 * device frame, downloads the texrects to h_rects; d_rects (DEVICE, may be NULL) keeps a copy */
int   planet_gpu_cache_frame(void *cache, const planet_gpu_params *p, const planet_gpu_quad *h_quads,
                             int64_t n, int max_lod, int generations_per_frame,
                             planet_gpu_texrect *h_rects, planet_gpu_texrect *d_rects, void *stream);
/* Planning only, on the HOST (no device needed, nothing is generated): the same bookkeeping code
 * run with one lane on a host copy of the tables, for callers that want the decisions without a
 * GPU (capacity planning, tests).  A cache object is either planned here or driven by the two
 * calls above, not both.  *n_generate = maps the frame would generate. */
int   planet_gpu_cache_plan_frame(void *cache, const planet_gpu_quad *h_quads, int64_t n,
                                  int generations_per_frame, planet_gpu_texrect *h_rects, int64_t *n_generate);
const float *planet_gpu_cache_pool(void *cache);      /* DEVICE pointer to slot 0 */
/* copies the maps in the given pool slots to HOST memory (n x dim x dim floats); synchronous */
int   planet_gpu_cache_read_slots(void *cache, const int32_t *slots, int64_t n, float *h_out);
/* K3 reading each quad's map through its texrect: bilinear GL_LINEAR / CLAMP_TO_EDGE sampling
 * (render.cpp:429-433) at mix(corners0, corners1, UV.xy) and +-pixel_size (main.cpp:334-346, 358),
 * i.e. what the shader does when the map is the parent's.  dim = patch_verts + 2. */
int planet_gpu_shade_cached(const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t nquads,
                            const double *cam_pos, const float *d_pool, const planet_gpu_texrect *d_rects,
                            float max_skirt, float *d_pos4, float *d_nrm4, void *stream);

/* ---- host-buffer convenience (what a reference-side caller with host memory uses) ---- */
/* batched GenerateHeightMap with HOST quads and HOST output: H2D quads, K2, D2H heights;
 * returns after the data is in h_out.  d_mirror (DEVICE pointer, may be NULL) additionally
 * keeps the height maps resident on the device -- the role the GL texture plays after
 * main.cpp:245 -- so K3 can consume them without a second upload. */
int planet_gpu_generate_height_maps_host(const planet_gpu_params *p, const planet_gpu_quad *h_quads,
                                         int64_t nquads, int dim, int max_depth, float *h_out,
                                         float *d_mirror);

/* One frame's worth of the whole path for a caller with HOST quads: H2D quads, K2 (dim =
 * patch_verts + 2), the height maps to h_heights (what the reference uploads with glTexImage2D,
 * main.cpp:245) AND resident in d_heights, then K3 (main.cpp:348-380) into d_pos4 / d_nrm4 on the
 * device while the last height maps are still crossing PCIe.  Returns when everything is done.
 * cam_pos: 3 host doubles; max_skirt < 0 = planet_gpu_max_skirt_size(radius, patch_verts). */
int planet_gpu_terrain_host(const planet_gpu_params *p, const planet_gpu_quad *h_quads, int64_t nquads,
                            int max_depth, const double *cam_pos, float max_skirt, float *h_heights,
                            float *d_heights, float *d_pos4, float *d_nrm4);

/* ---- measurement helpers ------------------------------------------------------------- */
/* dependent-free FFMA loop on every SM for about `ms` milliseconds; returns achieved
 * FP32 TFLOP/s (2 flop per FFMA) and the elapsed device time; the FP32 roofline
 * denominator measured on this device under this kernel's own power/clock regime */
int planet_gpu_measure_fp32_peak(double ms, double *tflops, double *elapsed_ms);
/* number of kernels this library has launched since planet_gpu_init (bench.py's gpu_launches) */
int64_t planet_gpu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PLANET_GPU_H */
