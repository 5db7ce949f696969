// height_map_cache.cu -- the step after the path: residency of generated
// height maps, with the decisions taken ON THE DEVICE.
//
// Replaces HeightMapCache / MapFind (main.cpp:75-102) and GetHeightMapForQuad
// (main.cpp:191-278).  The reference keeps one GL texture per cached quad and, for every leaf of
// every frame, looks the quad up, generates on a miss (32x32 map on the CPU + glTexImage2D),
// evicts the least recently used entry beyond 1 024, and -- once the per-frame budget of 100
// generations (main.cpp:652-653) is spent -- borrows the parent's map with remapped texture
// corners.  Here the textures are slots of ONE device-resident pool, the bookkeeping (id table,
// slot table, last-used ticks, free-slot stack) lives in device memory, and a whole frame is:
//
//   k_plan_frame (one CTA)  leaf quads (device, straight from K0) -> texrects + the miss list
//     all threads          an exact index over the id table is built for the frame (the reference
//                          answers "is this id cached?" with a scan of all 1 499 slots); one thread per
//                          leaf then finds where the leaf and its parent sit as the frame starts;
//     all threads          in a frame that cannot evict (count + misses <= cache size) every hit is
//                          resolved at once: nothing can take its entry away before its turn;
//     one warp             walks what is left -- the misses, or every leaf if evictions are possible
//                          -- in the reference's order, against what the frame has changed so far:
//                          a miss spends budget, falls back to the parent or generates; an
//                          eviction is ONE packed-key maximum over the table (age, then lowest
//                          slot) by the warp; the empty slot an insertion takes is the first of the
//                          reference's own probe sequence;
//     all threads          texrects from the packed per-leaf results; the missing quads compacted
//                          into the K2 batch; the new state written back.
//   K2                     ONE batched launch generates every miss (main.cpp:244 once per quad);
//   k_scatter_maps         moves the finished maps to their pool slots;
//   K3 (shade_cached)      reads each quad's map through its texrect.
//
// The host reads back two integers per frame (misses, error), no quads, no texrects.  Every
// decision -- probe order (lo32 ^ hi32, linear, first match), who is evicted (oldest tick, lowest
// table slot on ties), the fallback rule, the budget -- is the reference's; tests/test_cache.py
// checks them against the reference's own GetHeightMapForQuad on a 30-frame flight with hits,
// fallbacks and evictions.  The bookkeeping is a template over the number of cooperating lanes:
// 32 inside the kernel, 1 for planet_gpu_cache_plan_frame, the planning-only entry that runs the
// SAME code on a host copy of the state (no device needed; no terrain is computed there).
//
// A frame works on a copy of the state (two state buffers, ping-pong) and becomes current only
// when the plan succeeded and the generation was launched: a failed frame -- pool exhausted,
// allocation or launch error -- leaves the cache exactly as it was (no entry ever points at a
// slot that was not generated).  Slots freed by eviction are reused from the next frame on: the
// reference had already drawn with a texture before deleting it, here shading follows generation.
#include "planet_common.cuh"

#include <cstddef>
#include <cstring>
#include <vector>

namespace planet {

int launch_height_maps(const planet_gpu_params *, const Quad *, int64_t, int, int, float *, cudaStream_t);

namespace cache {

// ---- the bookkeeping state: one blob of 32-bit words, in host or device memory ---------------
enum { H_COUNT, H_TICK, H_NFREE, H_NDEFERRED, H_NGEN, H_ERROR, H_WORDS = 8 };
enum { ERR_POOL_EXHAUSTED = 1 };

struct Shape { int dim, cache_max, map_max, pool_slots; };

struct State {                 // views into one blob
    int32_t *hdr;              // H_*
    uint64_t *ids;             // [map_max] QuadID value, 0 = empty          (main.cpp:81)
    int32_t *slot_of;          // [map_max] pool slot instead of a GL name   (main.cpp:82)
    uint32_t *last_tick;       // [map_max]                                  (main.cpp:83)
    int32_t *free_slots;       // [pool_slots] stack of unused pool slots
    int32_t *deferred;         // [pool_slots] freed this frame, reusable next frame
};
__host__ __device__ inline size_t state_words(const Shape &s) { return H_WORDS + (size_t)s.map_max * 4 + (size_t)s.pool_slots * 2; }
__host__ __device__ inline State state_at(void *blob, const Shape &s)
{
    int32_t *w = (int32_t *)blob;
    State v;
    v.hdr = w;
    v.ids = (uint64_t *)(w + H_WORDS);                                   // 8-byte aligned: H_WORDS is even
    v.slot_of = w + H_WORDS + 2 * s.map_max;
    v.last_tick = (uint32_t *)(w + H_WORDS + 3 * s.map_max);
    v.free_slots = w + H_WORDS + 4 * s.map_max;
    v.deferred = v.free_slots + s.pool_slots;
    return v;
}

// per-leaf scratch of one frame
struct Scratch {
    int32_t *own, *par;        // [n] phase-1 probe results (table index or -1)
    int32_t *index2;           // [index2_size] the frame's exact index over the id table, see Index2
    int32_t index2_size;       // a power of two >= 2 * (map_max + n)
    int32_t *miss_src;         // [n] leaf index of the k-th generated map
    int32_t *miss_slot;        // [n] its pool slot
};

// ---- lanes: 32 on the device (one warp), 1 on the host --------------------------------------
template <int LANES> struct Lanes;
template <> struct Lanes<1> {
    static __host__ __device__ int lane() { return 0; }
    static __host__ __device__ unsigned ballot(bool p) { return p ? 1u : 0u; }
    static __host__ __device__ unsigned long long max64(unsigned long long v) { return v; }
    static __host__ __device__ void sync() {}
};
#ifdef __CUDACC__
template <> struct Lanes<32> {
    static __device__ int lane() { return threadIdx.x & 31; }
    static __device__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    static __device__ unsigned long long max64(unsigned long long v)
    {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
            v = o > v ? o : v;
        }
        return v;
    }
    static __device__ void sync() { __syncwarp(); }
};
#endif
__host__ __device__ inline int first_bit(unsigned b)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)b) - 1;
#else
    return __builtin_ctz(b);
#endif
}

__host__ __device__ inline uint32_t hash_of(uint64_t key) { return (uint32_t)key ^ (uint32_t)(key >> 32); }   // main.cpp:91

// MapFind (main.cpp:86-102): the first table index in probe order from hash(key) whose id equals
// `want` (want == 0: the first empty slot of the key's probe sequence); LANES slots per step.
template <int LANES>
__host__ __device__ inline int find(const uint64_t *ids, int map_max, uint64_t key, uint64_t want)
{
    const uint32_t h = hash_of(key) % (uint32_t)map_max;
    for (int base = 0; base < map_max; base += LANES) {
        const int i = base + Lanes<LANES>::lane();
        uint32_t index = h + (uint32_t)i;                                // < 2 * map_max
        if (index >= (uint32_t)map_max) index -= (uint32_t)map_max;
        const unsigned hit = Lanes<LANES>::ballot(i < map_max && ids[index] == want);
        if (hit) {
            uint32_t r = h + (uint32_t)(base + first_bit(hit));
            return (int)(r >= (uint32_t)map_max ? r - (uint32_t)map_max : r);
        }
    }
    return -1;
}

// "Is this id anywhere in the table, and where?"  The reference answers with a scan of all map_max
// slots (an eviction leaves no tombstone, so a probe cannot stop at an empty slot).  Here the frame
// builds, once and in parallel, an exact index over the table -- open addressing on a multiplicative
// hash, entry = table slot + 1 -- and every such question afterwards is one or two probes, in the
// parallel pre-pass and on the sequential path alike.  Insertions of the frame are added to it; an
// eviction needs nothing: a stale entry points at a slot that no longer holds the id, fails the
// compare and the probe moves on.  (Which EMPTY slot an insertion takes is still the reference's own
// probe sequence, find(key, 0): that order decides later LRU ties.)
__host__ __device__ inline uint32_t hash2(uint64_t id, int size)
{
    // the TOP bits of a multiplicative hash: a QuadID keeps its path in the low bits and root / depth in
    // bits 55..63, and a shallow quad's low bits are all zero -- the low bits of a product only see the
    // low bits of its input (taking those made the probe chains hundreds of entries long)
#ifdef __CUDA_ARCH__
    const int shift = __clz(size) + 1;
#else
    const int shift = __builtin_clz((unsigned)size) + 1;
#endif
    const uint32_t x = (uint32_t)id * 0x9E3779B1u ^ (uint32_t)(id >> 32) * 0x85EBCA77u;
    return (x * 2654435761u) >> shift;
}
__host__ __device__ inline int lookup2(const State &st, const Scratch &sc, uint64_t id)
{
    for (uint32_t h = hash2(id, sc.index2_size);; h = (h + 1) & (uint32_t)(sc.index2_size - 1)) {
        const int e = sc.index2[h];
        if (e == 0) return -1;
        if (st.ids[e - 1] == id) return e - 1;
    }
}
// single writer (the sequential path, or the host)
__host__ __device__ inline void insert2(const Scratch &sc, uint64_t id, int slot)
{
    uint32_t h = hash2(id, sc.index2_size);
    while (sc.index2[h] != 0) h = (h + 1) & (uint32_t)(sc.index2_size - 1);
    sc.index2[h] = slot + 1;
}
#ifdef __CUDACC__
// many writers (the parallel build)
__device__ inline void insert2_atomic(const Scratch &sc, uint64_t id, int slot)
{
    uint32_t h = hash2(id, sc.index2_size);
    while (atomicCAS(&sc.index2[h], 0, slot + 1) != 0) h = (h + 1) & (uint32_t)(sc.index2_size - 1);
}
#endif

// the entry main.cpp:249-261 evicts: the largest tick age (signed, as there), the lowest table
// index among equals -- one packed key (age << 32 | ~index), one maximum
template <int LANES>
__host__ __device__ inline int oldest(const State &st, int map_max, uint32_t tick)
{
    unsigned long long best = 0;
    for (int base = 0; base < map_max; base += LANES) {
        const int i = base + Lanes<LANES>::lane();
        if (i < map_max && st.ids[i] != 0) {
            const int age = (int)(tick - st.last_tick[i]);
            if (age >= 0) {
                const unsigned long long key = ((unsigned long long)(uint32_t)age << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
                best = key > best ? key : best;
            }
        }
    }
    best = Lanes<LANES>::max64(best);
    return best ? (int)(0xFFFFFFFFu - (uint32_t)best) : 0;                // nothing qualified: slot 0, as main.cpp:249
}

__host__ __device__ inline uint64_t parent_of(uint64_t id)                // main.cpp:57-65
{
    return (id - (1ull << 55)) & ~(3ull << (2 * (quad_depth(id) - 1)));
}

// the texture window a quad samples: its own map (border texel excluded), or its quadrant of the
// parent's map (main.cpp:197-199, 212-236).  Bit 0 of the child index selects the x half, bit 1
// the y half.
__host__ __device__ inline void window(planet_gpu_texrect &r, int dim_i, int quadrant)
{
    const float dim = (float)dim_i;
    const float lo = 1.5f, hi = dim - 1.5f;
    if (quadrant < 0) {
        r.corners[0] = r.corners[1] = lo / dim;
        r.corners[2] = r.corners[3] = hi / dim;
        r.pixel_size[0] = r.pixel_size[1] = 1.0f / dim;
        return;
    }
    const float below = dim / 2.0f - 0.5f, above = dim / 2.0f + 0.5f;
    r.corners[0] = ((quadrant & 1) ? above : lo) / dim;
    r.corners[1] = ((quadrant & 2) ? above : lo) / dim;
    r.corners[2] = ((quadrant & 1) ? hi : below) / dim;
    r.corners[3] = ((quadrant & 2) ? hi : below) / dim;
    r.pixel_size[0] = r.pixel_size[1] = ((dim / 2.0f - 1.0f) / (float)(dim_i - 3)) / dim;
}

// the ids of the frame's leaves: inside the 104-byte quads (stride 104) or a packed copy (stride 8)
struct LeafIds {
    const unsigned char *base; size_t stride;
    __host__ __device__ uint64_t operator[](int64_t i) const { return *reinterpret_cast<const uint64_t *>(base + (size_t)i * stride); }
};
__host__ __device__ inline LeafIds ids_in(const planet_gpu_quad *quads)
{
    return LeafIds{ reinterpret_cast<const unsigned char *>(quads) + offsetof(planet_gpu_quad, id), sizeof(planet_gpu_quad) };
}

// phase 1 for one leaf: where the leaf and its parent sit in the table as the frame starts
__host__ __device__ inline void probe_leaf(const State &st, const LeafIds &leaf, int64_t i, const Scratch &sc)
{
    const uint64_t id = leaf[i];
    sc.own[i] = lookup2(st, sc, id);
    // the parent is probed even when the leaf itself is cached: the leaf's entry may be evicted earlier in
    // this frame, and its turn then needs the parent (found by the randomised lists of tests/test_cache.py)
    sc.par[i] = quad_depth(id) > 0 ? lookup2(st, sc, parent_of(id)) : -1;
}

// What the bookkeeping decides per leaf, packed into one word: the pool slot to sample (24 bits), the
// PLANET_TEXRECT_* kind (2 bits) and, for a parent fallback, the child's quadrant (2 bits).  The texrect
// itself (six float divisions) is built from it afterwards, in parallel, off the sequential path.
__host__ __device__ inline int32_t pack_result(int slot, int kind, int quadrant) { return slot | (kind << 24) | (quadrant << 26); }
__host__ __device__ inline planet_gpu_texrect rect_of(int32_t res, int dim)
{
    planet_gpu_texrect r;
    r.slot = res & 0xFFFFFF;
    r.flags = (res >> 24) & 3;
    window(r, dim, r.flags == PLANET_TEXRECT_PARENT ? (res >> 26) & 3 : -1);
    return r;
}

// the walking state of one frame (identical in every lane of the cooperating group)
struct Walk { int budget, count, n_free, n_gen, n_evicted, error; uint32_t tick; };

template <int LANES>
__host__ __device__ inline Walk begin_frame(const State &st, int budget)
{
    // textures deleted last frame are gone now: their slots return to the stack
    Walk w;
    w.n_free = st.hdr[H_NFREE];
    const int n_deferred = st.hdr[H_NDEFERRED];
    for (int j = Lanes<LANES>::lane(); j < n_deferred; j += LANES) st.free_slots[w.n_free + j] = st.deferred[j];
    w.n_free += n_deferred;
    Lanes<LANES>::sync();
    w.budget = budget; w.count = st.hdr[H_COUNT]; w.n_gen = w.n_evicted = w.error = 0;
    w.tick = (uint32_t)st.hdr[H_TICK];
    return w;
}

// GetHeightMapForQuad (main.cpp:191-278) for leaf i against what the frame has changed so far.  All
// lanes of the cooperating group execute this with identical control flow; lane 0 writes.
template <int LANES>
__host__ __device__ inline void resolve_leaf(const State &st, const Shape &sh, const LeafIds &leaf, int64_t i, const Scratch &sc,
                                             Walk &w, int32_t *res)
{
    const bool writer = Lanes<LANES>::lane() == 0;
    const uint64_t id = leaf[i];
    int kind = PLANET_TEXRECT_HIT, quadrant = 0;
    // the leaf itself: the start-of-frame probe if that slot still holds it, else the table as it is now
    // (an entry evicted earlier in the frame, or one inserted by an earlier leaf of the frame)
    int index = sc.own[i];
    if (index < 0 || st.ids[index] != id) index = w.n_gen || w.n_evicted ? lookup2(st, sc, id) : -1;
    if (index < 0) {
        if (w.budget <= 0 && quad_depth(id) > 0) {                           // main.cpp:208: budget spent -> try the parent's map
            const uint64_t pid = parent_of(id);
            int p = sc.par[i];
            if (p < 0 || st.ids[p] != pid) p = w.n_gen || w.n_evicted ? lookup2(st, sc, pid) : -1;
            if (p >= 0) {
                index = p;
                kind = PLANET_TEXRECT_PARENT;
                quadrant = (int)((id >> (2 * (quad_depth(id) - 1))) & 3);   // GetChildIndex, main.cpp:51-55
            }
        }
        if (w.budget > 0 || index < 0) {                                     // main.cpp:239: generate
            w.budget--;
            if (w.count == sh.cache_max) {                                   // main.cpp:247-266
                const int victim = oldest<LANES>(st, sh.map_max, w.tick);
                if (writer) { st.deferred[w.n_evicted] = st.slot_of[victim]; st.ids[victim] = 0; }
                w.n_evicted++;
                w.count--;
                Lanes<LANES>::sync();
            }
            if (w.n_free == 0) { w.error = ERR_POOL_EXHAUSTED; return; }
            const int slot = st.free_slots[--w.n_free];
            index = find<LANES>(st.ids, sh.map_max, id, 0);                  // main.cpp:268: first empty slot of the probe sequence
            if (writer) {
                st.ids[index] = id; st.slot_of[index] = slot;
                insert2(sc, id, index);
                sc.miss_src[w.n_gen] = (int32_t)i; sc.miss_slot[w.n_gen] = slot;
            }
            w.n_gen++; w.count++;
            kind = PLANET_TEXRECT_GENERATED;                                 // a generated map is the quad's own
            Lanes<LANES>::sync();
        }
    }
    if (writer) {
        st.last_tick[index] = w.tick;                                        // main.cpp:275
        res[i] = pack_result(st.slot_of[index], kind, quadrant);
    }
    Lanes<LANES>::sync();
}

// A hit in a frame that cannot evict (count + start-of-frame misses <= cache_max): nothing the frame
// does can take the entry away before the leaf's turn, and a hit changes nothing but its own tick --
// so all hits of such a frame are resolved at once, in any order, and only the misses are walked.
__host__ __device__ inline void touch_hit(const State &st, int index, uint32_t tick, int32_t *res_i)
{
    st.last_tick[index] = tick;
    *res_i = pack_result(st.slot_of[index], PLANET_TEXRECT_HIT, 0);
}

template <int LANES>
__host__ __device__ inline void end_frame(const State &st, const Walk &w)
{
    if (Lanes<LANES>::lane() == 0) {
        st.hdr[H_COUNT] = w.count; st.hdr[H_TICK] = (int32_t)(w.tick + 1);   // main.cpp:682
        st.hdr[H_NFREE] = w.n_free; st.hdr[H_NDEFERRED] = w.n_evicted;
        st.hdr[H_NGEN] = w.n_gen; st.hdr[H_ERROR] = w.error;
    }
}

// ---- the device frame -----------------------------------------------------------------
constexpr int PLAN_THREADS = 1024;

// With `in_smem` the whole frame runs out of shared memory: the state blob (40 KB at the reference's
// 1 024 / 1 499), the frame's index over it, the per-leaf scratch and a packed copy of the leaf ids are
// staged there and the new state is written to `next` in one coalesced pass at the end.  States too
// large for shared memory take the same code path on the global copy.
//   phase 0, all threads the state copy; the exact index over the id table, built with atomicCAS
//   phase 1, all threads one thread per leaf: where the leaf and its parent sit as the frame starts
//   phase 2, warp 0      counts the start-of-frame misses; if the frame cannot evict, lists them in order
//   phase 3, all threads (only then) every hit resolved at once
//   phase 4, warp 0      walks the misses in order -- or, if evictions are possible, every leaf.  One warp
//                        running dependent code retires an instruction every ~5 cycles, so what stays
//                        on this path is only what depends on the frame's earlier decisions (the first
//                        version walked all leaves, scanned the table for every absent id and built the
//                        texrects here: 90-160 us for 141 leaves)
//   phase 5, all threads texrects from the packed results, the K2 batch, the state write-back
__global__ void __launch_bounds__(PLAN_THREADS)
k_plan_frame(const int32_t *__restrict__ cur, int32_t *__restrict__ next, Shape sh, const planet_gpu_quad *__restrict__ quads,
             int64_t n, int budget, Scratch sc_global, int32_t *res_global, planet_gpu_texrect *__restrict__ rects,
             Quad *__restrict__ miss_quads, int in_smem)
{
    extern __shared__ __align__(16) int32_t s_plan[];
    __shared__ int s_fast, s_nmiss;
    const size_t words = state_words(sh);
    int32_t *work = in_smem ? s_plan : next;
    for (size_t w = threadIdx.x; w < words; w += PLAN_THREADS) work[w] = cur[w];
    Scratch sc = sc_global;
    int32_t *res = res_global, *miss_pos = res_global + n;
    LeafIds leaf = ids_in(quads);
    if (in_smem) {
        int32_t *at = s_plan + ((words + 1) & ~(size_t)1);                     // keep the id copy 8-byte aligned
        uint64_t *s_ids = reinterpret_cast<uint64_t *>(at);
        for (int64_t i = threadIdx.x; i < n; i += PLAN_THREADS) s_ids[i] = quads[i].id;
        leaf = LeafIds{ reinterpret_cast<const unsigned char *>(s_ids), sizeof(uint64_t) };
        at += 2 * n;
        sc = Scratch{ at, at + n, at + 6 * n, sc_global.index2_size, at + 2 * n, at + 3 * n };
        res = at + 4 * n; miss_pos = at + 5 * n;
    }
    for (int h = threadIdx.x; h < sc.index2_size; h += PLAN_THREADS) sc.index2[h] = 0;
    __syncthreads();
    const State st = state_at(work, sh);
    for (int k = threadIdx.x; k < sh.map_max; k += PLAN_THREADS)
        if (st.ids[k] != 0) insert2_atomic(sc, st.ids[k], k);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = threadIdx.x; i < n; i += PLAN_THREADS) probe_leaf(st, leaf, i, sc);
    __syncthreads();
    if (warp == 0) {                                                         // the misses, in order
        int nm = 0;
        for (int64_t base = 0; base < n; base += 32) {
            const int64_t i = base + lane;
            const bool miss = i < n && sc.own[i] < 0;
            const unsigned m = __ballot_sync(0xffffffffu, miss);
            if (miss) miss_pos[nm + __popc(m & ((1u << lane) - 1))] = (int32_t)i;
            nm += __popc(m);
        }
        if (lane == 0) { s_nmiss = nm; s_fast = st.hdr[H_COUNT] + nm <= sh.cache_max; }
    }
    __syncthreads();
    const bool fast = s_fast != 0;
    if (fast) {
        const uint32_t tick = (uint32_t)st.hdr[H_TICK];
        for (int64_t i = threadIdx.x; i < n; i += PLAN_THREADS)
            if (sc.own[i] >= 0) touch_hit(st, sc.own[i], tick, res + i);
    }
    __syncthreads();
    if (warp == 0) {
        Walk w = begin_frame<32>(st, budget);
        if (fast) for (int k = 0; k < s_nmiss && !w.error; k++) resolve_leaf<32>(st, sh, leaf, miss_pos[k], sc, w, res);
        else      for (int64_t i = 0; i < n && !w.error; i++) resolve_leaf<32>(st, sh, leaf, i, sc, w, res);
        end_frame<32>(st, w);
    }
    __syncthreads();
    const bool failed = st.hdr[H_ERROR] != 0;
    if (!failed)
        for (int64_t i = threadIdx.x; i < n; i += PLAN_THREADS) rects[i] = rect_of(res[i], sh.dim);
    // the K2 batch: the missing quads, compacted (13 eight-byte words each), and their pool slots
    const int n_gen = failed ? 0 : st.hdr[H_NGEN];
    const uint64_t *src = reinterpret_cast<const uint64_t *>(quads);
    uint64_t *dst = reinterpret_cast<uint64_t *>(miss_quads);
    for (int w = threadIdx.x; w < n_gen * 13; w += PLAN_THREADS) {
        const int k = w / 13, part = w - k * 13;
        dst[w] = src[(size_t)sc.miss_src[k] * 13 + part];
    }
    if (in_smem) {
        for (int k = threadIdx.x; k < n_gen; k += PLAN_THREADS) sc_global.miss_slot[k] = sc.miss_slot[k];   // k_scatter_maps reads them
        for (size_t w = threadIdx.x; w < words; w += PLAN_THREADS) next[w] = work[w];
    }
}

__global__ void k_scatter_maps(const float *__restrict__ maps, const int *__restrict__ slots, int n, int texels,
                               float *__restrict__ pool)
{
    // one CTA per generated map: contiguous 16-byte copies into its pool slot
    for (int m = blockIdx.x; m < n; m += gridDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(maps + (size_t)m * texels);
        float4 *dst = reinterpret_cast<float4 *>(pool + (size_t)slots[m] * texels);
        for (int i = threadIdx.x; i < texels / 4; i += blockDim.x) dst[i] = src[i];
        for (int i = (texels & ~3) + threadIdx.x; i < texels; i += blockDim.x)
            pool[(size_t)slots[m] * texels + i] = maps[(size_t)m * texels + i];
    }
}

struct Cache {
    Shape sh;
    enum { UNUSED, HOST_PLANNED, DEVICE_PLANNED } mode = UNUSED;
    int count = 0;                           // host copy of H_COUNT after the last successful frame
    // planning-only mode: the state in host memory (two copies, like the device)
    std::vector<int32_t> h_state[2];
    std::vector<int32_t> h_scratch;
    // device mode
    int32_t *d_state[2] = { nullptr, nullptr };
    int current = 0;
    float *d_pool = nullptr;
    int32_t *d_scratch = nullptr; Quad *d_miss_quads = nullptr; planet_gpu_quad *d_quads = nullptr; planet_gpu_texrect *d_rects = nullptr;
    size_t leaf_cap = 0;
    float *d_miss_maps = nullptr; size_t miss_cap = 0;
    int32_t *h_hdr = nullptr;                // pinned: the two integers a frame reads back
    size_t plan_smem_set = 0;
    int32_t *d_index2 = nullptr; size_t index2_cap = 0;
    std::vector<int32_t> h_index2;
};

static void fresh_state(const Shape &sh, int32_t *blob)
{
    memset(blob, 0, state_words(sh) * sizeof(int32_t));
    State st = state_at(blob, sh);
    for (int k = 0; k < sh.map_max; k++) st.slot_of[k] = -1;
    for (int k = 0; k < sh.pool_slots; k++) st.free_slots[k] = sh.pool_slots - 1 - k;   // slot 0 is handed out first
    st.hdr[H_NFREE] = sh.pool_slots;
}

// size of the frame's index: a power of two, at most half full after every leaf of the frame was inserted
static int index2_size_for(const Shape &sh, int64_t n)
{
    int size = 1024;
    while ((int64_t)size < 2 * ((int64_t)sh.map_max + n)) size <<= 1;
    return size;
}
// own | par | miss_src | miss_slot | (res | miss_pos)   -- six arrays of n; the index lives elsewhere
static Scratch scratch_at(int32_t *base, size_t n, int32_t *index2, int index2_size)
{
    return Scratch{ base, base + n, index2, index2_size, base + 2 * n, base + 3 * n };
}

static int latch(Cache *c, int mode)
{
    if (c->mode == Cache::UNUSED) c->mode = (decltype(c->mode))mode;
    if (c->mode != mode)
        return set_error(PLANET_E_INVALID, "this cache is already %s: planet_gpu_cache_plan_frame (host bookkeeping only) and "
                         "planet_gpu_cache_frame / _frame_device (device bookkeeping) keep separate states",
                         c->mode == Cache::HOST_PLANNED ? "planned on the host" : "planned on the device");
    return 0;
}

static int pool_exhausted(const Cache *c)
{
    return set_error(PLANET_E_INVALID, "height-map pool exhausted: more than %d generations in one frame beyond the cache size "
                     "(the frame was dropped, the cache is unchanged)", c->sh.pool_slots - c->sh.cache_max);
}

static int ensure_device(Cache *c, int64_t n)
{
    if (!ensure_init()) return PLANET_E_NO_DEVICE;
    if (!c->d_state[0]) {
        const size_t bytes = state_words(c->sh) * sizeof(int32_t);
        std::vector<int32_t> init(state_words(c->sh));
        fresh_state(c->sh, init.data());
        PLANET_CUDA(cudaMalloc(&c->d_state[0], bytes));
        PLANET_CUDA(cudaMalloc(&c->d_state[1], bytes));
        PLANET_CUDA(cudaMemcpy(c->d_state[0], init.data(), bytes, cudaMemcpyHostToDevice));
        PLANET_CUDA(cudaMallocHost(&c->h_hdr, H_WORDS * sizeof(int32_t)));
        c->current = 0;
    }
    if (!c->d_pool) PLANET_CUDA(cudaMalloc(&c->d_pool, (size_t)c->sh.pool_slots * c->sh.dim * c->sh.dim * sizeof(float)));
    if ((size_t)n > c->leaf_cap) {
        cudaFree(c->d_scratch); cudaFree(c->d_miss_quads); cudaFree(c->d_quads); cudaFree(c->d_rects);   // cudaFree(nullptr) is a no-op
        c->d_scratch = nullptr; c->d_miss_quads = nullptr; c->d_quads = nullptr; c->d_rects = nullptr; c->leaf_cap = 0;
        const size_t want = (size_t)n + 1024;
        PLANET_CUDA(cudaMalloc(&c->d_scratch, want * 6 * sizeof(int32_t)));
        PLANET_CUDA(cudaMalloc(&c->d_miss_quads, want * sizeof(Quad)));
        PLANET_CUDA(cudaMalloc(&c->d_quads, want * sizeof(planet_gpu_quad)));
        PLANET_CUDA(cudaMalloc(&c->d_rects, want * sizeof(planet_gpu_texrect)));
        c->leaf_cap = want;
    }
    return 0;
}

// plan on the device, generate the misses, make the frame current.  d_quads / d_rects: device.
static int frame_on_device(Cache *c, const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t n, int max_lod,
                           int budget, planet_gpu_texrect *d_rects, int64_t *n_generated, cudaStream_t stream)
{
    int rc = ensure_device(c, n);
    if (rc) return rc;
    const int index2_size = index2_size_for(c->sh, n);
    if ((size_t)index2_size > c->index2_cap) {
        cudaFree(c->d_index2);
        c->d_index2 = nullptr; c->index2_cap = 0;
        PLANET_CUDA(cudaMalloc(&c->d_index2, (size_t)index2_size * sizeof(int32_t)));
        c->index2_cap = (size_t)index2_size;
    }
    const Scratch sc = scratch_at(c->d_scratch, c->leaf_cap, c->d_index2, index2_size);
    int32_t *cur = c->d_state[c->current], *next = c->d_state[c->current ^ 1];
    // the frame out of shared memory when state + ids + scratch + index fit (they do unless the cache or the frame is huge)
    const size_t smem = (((state_words(c->sh) + 1) & ~(size_t)1) + 8 * (size_t)n + (size_t)index2_size) * sizeof(int32_t);
    const int in_smem = smem <= 200 * 1024;
    if (in_smem && smem > 48 * 1024 && smem > c->plan_smem_set) {
        PLANET_CUDA(cudaFuncSetAttribute(k_plan_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        c->plan_smem_set = 200 * 1024;
    }
    k_plan_frame<<<1, PLAN_THREADS, in_smem ? smem : 0, stream>>>(cur, next, c->sh, d_quads, n, budget, sc, c->d_scratch + 4 * c->leaf_cap,
                                                                  d_rects, c->d_miss_quads, in_smem);
    count_launch();
    PLANET_CUDA(cudaGetLastError());
    PLANET_CUDA(cudaMemcpyAsync(c->h_hdr, next, H_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PLANET_CUDA(cudaStreamSynchronize(stream));                          // the frame's one read-back: misses, error
    if (c->h_hdr[H_ERROR]) return pool_exhausted(c);
    const int64_t n_gen = c->h_hdr[H_NGEN];
    const int texels = c->sh.dim * c->sh.dim;
    if (n_gen > 0) {
        if ((size_t)n_gen > c->miss_cap) {
            cudaFree(c->d_miss_maps);
            c->d_miss_maps = nullptr; c->miss_cap = 0;
            PLANET_CUDA(cudaMalloc(&c->d_miss_maps, ((size_t)n_gen + 256) * texels * sizeof(float)));
            c->miss_cap = (size_t)n_gen + 256;
        }
        rc = launch_height_maps(p, c->d_miss_quads, n_gen, c->sh.dim, max_lod, c->d_miss_maps, stream);
        if (rc) return rc;
        k_scatter_maps<<<(int)std::min<int64_t>(n_gen, 1184), 128, 0, stream>>>(c->d_miss_maps, sc.miss_slot, (int)n_gen, texels, c->d_pool);
        count_launch();
        PLANET_CUDA(cudaGetLastError());
    }
    c->current ^= 1;                                                     // the frame went through
    c->count = c->h_hdr[H_COUNT];
    if (n_generated) *n_generated = n_gen;
    return 0;
}

} // namespace cache
} // namespace planet

using namespace planet;
using namespace planet::cache;

extern "C" {

void *planet_gpu_cache_create(int dim, int cache_max, int map_max, int extra_slots)
{
    if (dim <= 3 || cache_max <= 0 || map_max <= cache_max || extra_slots < 0 || map_max > (1 << 24) || cache_max + (int64_t)extra_slots > (1 << 24)) {
        set_error(PLANET_E_INVALID, "cache_create(dim=%d, cache_max=%d, map_max=%d, extra=%d): need dim > 3, 0 < cache_max < map_max", dim, cache_max, map_max, extra_slots);
        return nullptr;
    }
    Cache *c = new Cache();
    c->sh = Shape{ dim, cache_max, map_max, cache_max + extra_slots };
    return c;
}

void planet_gpu_cache_destroy(void *cache)
{
    Cache *c = (Cache *)cache;
    if (!c) return;
    cudaFree(c->d_state[0]); cudaFree(c->d_state[1]); cudaFree(c->d_pool); cudaFree(c->d_scratch);
    cudaFree(c->d_miss_quads); cudaFree(c->d_quads); cudaFree(c->d_rects); cudaFree(c->d_miss_maps); cudaFree(c->d_index2);
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    delete c;
}

int planet_gpu_cache_count(const void *cache) { return cache ? ((const Cache *)cache)->count : 0; }

int planet_gpu_cache_plan_frame(void *cache, const planet_gpu_quad *h_quads, int64_t n, int generations_per_frame,
                                planet_gpu_texrect *h_rects, int64_t *n_generate)
{
    Cache *c = (Cache *)cache;
    if (!c || n < 0 || (n > 0 && (!h_quads || !h_rects))) return set_error(PLANET_E_INVALID, "NULL argument");
    int rc = latch(c, Cache::HOST_PLANNED);
    if (rc) return rc;
    const size_t words = state_words(c->sh);
    if (c->h_state[0].empty()) {
        c->h_state[0].resize(words); c->h_state[1].resize(words);
        fresh_state(c->sh, c->h_state[0].data());
        c->current = 0;
    }
    // the same three phases as k_plan_frame, one lane
    int32_t *cur = c->h_state[c->current].data(), *next = c->h_state[c->current ^ 1].data();
    memcpy(next, cur, words * sizeof(int32_t));
    c->h_scratch.resize((size_t)n * 5 + 1);
    c->h_index2.assign((size_t)index2_size_for(c->sh, n), 0);
    const Scratch sc = scratch_at(c->h_scratch.data(), (size_t)n, c->h_index2.data(), (int)c->h_index2.size());
    const State st = state_at(next, c->sh);
    const LeafIds leaf = ids_in(h_quads);
    int32_t *res = c->h_scratch.data() + 4 * (size_t)n;
    for (int k = 0; k < c->sh.map_max; k++)
        if (st.ids[k] != 0) insert2(sc, st.ids[k], k);
    for (int64_t i = 0; i < n; i++) probe_leaf(st, leaf, i, sc);
    int n_miss = 0;
    for (int64_t i = 0; i < n; i++) n_miss += sc.own[i] < 0;
    const bool fast = st.hdr[H_COUNT] + n_miss <= c->sh.cache_max;           // the frame cannot evict
    if (fast)
        for (int64_t i = 0; i < n; i++)
            if (sc.own[i] >= 0) touch_hit(st, sc.own[i], (uint32_t)st.hdr[H_TICK], res + i);
    Walk w = begin_frame<1>(st, generations_per_frame);
    for (int64_t i = 0; i < n && !w.error; i++)
        if (!fast || sc.own[i] < 0) resolve_leaf<1>(st, c->sh, leaf, i, sc, w, res);
    end_frame<1>(st, w);
    if (!w.error)
        for (int64_t i = 0; i < n; i++) h_rects[i] = rect_of(res[i], c->sh.dim);
    if (st.hdr[H_ERROR]) return pool_exhausted(c);
    c->current ^= 1;
    c->count = st.hdr[H_COUNT];
    if (n_generate) *n_generate = st.hdr[H_NGEN];
    return 0;
}

const float *planet_gpu_cache_pool(void *cache)
{
    Cache *c = (Cache *)cache;
    if (!c || ensure_device(c, 0)) return nullptr;
    return c->d_pool;
}

int planet_gpu_cache_read_slots(void *cache, const int32_t *slots, int64_t n, float *h_out)
{
    Cache *c = (Cache *)cache;
    if (!c || !slots || !h_out) return set_error(PLANET_E_INVALID, "NULL argument");
    if (!planet_gpu_cache_pool(cache)) return PLANET_E_CUDA;
    const size_t texels = (size_t)c->sh.dim * c->sh.dim;
    PLANET_CUDA(cudaDeviceSynchronize());
    for (int64_t i = 0; i < n; i++) {
        if (slots[i] < 0 || slots[i] >= c->sh.pool_slots) return set_error(PLANET_E_INVALID, "slot %d out of range", slots[i]);
        PLANET_CUDA(cudaMemcpy(h_out + i * texels, c->d_pool + (size_t)slots[i] * texels, texels * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int planet_gpu_cache_frame_device(void *cache, const planet_gpu_params *p, const planet_gpu_quad *d_quads, int64_t n,
                                  int max_lod, int generations_per_frame, planet_gpu_texrect *d_rects,
                                  int64_t *n_generated, void *stream)
{
    Cache *c = (Cache *)cache;
    if (!c || n < 0 || n > (1 << 24) || (n > 0 && (!d_quads || !d_rects))) return set_error(PLANET_E_INVALID, "cache_frame_device: bad argument");
    int rc = validate_params(p);
    if (rc) return rc;
    if ((rc = latch(c, Cache::DEVICE_PLANNED))) return rc;
    return frame_on_device(c, p, d_quads, n, max_lod, generations_per_frame, d_rects, n_generated, (cudaStream_t)stream);
}

int planet_gpu_cache_frame(void *cache, const planet_gpu_params *p, const planet_gpu_quad *h_quads, int64_t n,
                           int max_lod, int generations_per_frame, planet_gpu_texrect *h_rects,
                           planet_gpu_texrect *d_rects, void *stream_)
{
    Cache *c = (Cache *)cache;
    if (!c || n < 0 || n > (1 << 24) || (n > 0 && (!h_quads || !h_rects))) return set_error(PLANET_E_INVALID, "cache_frame: bad argument");
    int rc = validate_params(p);
    if (rc) return rc;
    if ((rc = latch(c, Cache::DEVICE_PLANNED))) return rc;
    if ((rc = ensure_device(c, n))) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    // host quads in, host texrects out: the same device frame between one upload and one download
    if (n > 0) PLANET_CUDA(cudaMemcpyAsync(c->d_quads, h_quads, n * sizeof(planet_gpu_quad), cudaMemcpyHostToDevice, stream));
    planet_gpu_texrect *rects = d_rects ? d_rects : c->d_rects;
    rc = frame_on_device(c, p, c->d_quads, n, max_lod, generations_per_frame, rects, nullptr, stream);
    if (rc) return rc;
    if (n > 0) PLANET_CUDA(cudaMemcpyAsync(h_rects, rects, n * sizeof(planet_gpu_texrect), cudaMemcpyDeviceToHost, stream));
    PLANET_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

} // extern "C"
