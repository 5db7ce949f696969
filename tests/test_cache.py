"""the height-map cache (main.cpp:75-102, 191-278).

CPU half: the host bookkeeping (planet_gpu_cache_plan_frame) against the reference's own
GetHeightMapForQuad run over a camera path that exercises hits, budget-exhausted parent
fallback and LRU eviction.  GPU half: the generated pool contents and the texrect shading."""
import numpy as np
import pytest

import planet_b200 as pb
from oracle.bindings import RADIUS, height_params


def camera_path():
    """Start far out (few coarse leaves), jump down to the surface (more than 100 new leaves in one
    frame -> parent fallback), then travel round the planet (cache fills beyond 1 024 -> LRU)."""
    R = RADIUS
    cams = [np.array([0.0, 0.0, -3.0 * R]), np.array([0.0, 0.0, -R - 200000.0]), np.array([0.0, 0.0, -R - 10.0]),
            np.array([0.0, 0.0, -R - 10.0]), np.array([0.0, 0.0, -R - 3000.0])]
    for k in range(1, 26):                                  # a great-circle tour at 10 m altitude
        a = 2 * np.pi * k / 25
        cams.append(np.array([np.sin(a) * (R + 10.0), 0.3 * R * np.sin(3 * a), -np.cos(a) * (R + 10.0)]))
        cams[-1] *= (R + 10.0) / np.linalg.norm(cams[-1])
    return cams


def reference_frames(ref):
    """Per frame: leaf quads, the id of the quad whose map each draw sampled, corners, pixel size."""
    ref.reset_cache()
    owner = {}                                              # GL texture name -> QuadID it was generated for
    frames = []
    for cam in camera_path():
        quads, maps, draws, tex = ref.render_next_frame(cam, height_params())
        # a texture generated this frame belongs to the first quad of the frame drawn with it
        for q, t, d in zip(quads, tex, draws):
            t = int(t)
            if t in maps and t not in owner:
                owner[t] = int(q["id"])
        frames.append(dict(cam=cam, quads=quads, owners=np.array([owner[int(t)] for t in tex], np.uint64),
                           corners=draws[:, 25:29].copy(), pixel=draws[:, 29:31].copy(), generated=len(maps),
                           maps=maps, tex=tex, pn=draws[:, 0:24].copy(), skirt=draws[:, 24].copy()))
    return frames


@pytest.fixture(scope="module")
def frames(ref):
    return reference_frames(ref)


def test_path_exercises_fallback_and_eviction(frames, ref):
    fallback = sum(int((f["pixel"][:, 0] != np.float32(1.0 / 32)).sum()) for f in frames)
    total_generated = sum(f["generated"] for f in frames)
    assert fallback > 0, "camera path never hit the budget-exhausted parent fallback"
    assert total_generated > 1024 + 100, "camera path never filled the cache (no LRU eviction)"
    assert ref.L.ref_cache_count() == 1024


def test_plan_frame_makes_the_reference_decisions(frames):
    """Same map owner (own / parent / which cached quad), same corners and pixel size, bit for bit."""
    cache = pb.HeightMapCache(32, 1024, 1499, extra_slots=1024)
    slot_owner = {}
    for k, f in enumerate(frames):
        rects, n_gen = cache.plan_frame(f["quads"], 100)
        assert n_gen == f["generated"], (k, n_gen, f["generated"])
        for q, r in zip(f["quads"], rects):
            if r["flags"] == pb.TEXRECT_GENERATED:
                slot_owner[int(r["slot"])] = int(q["id"])
        owners = np.array([slot_owner[int(s)] for s in rects["slot"]], np.uint64)
        assert (owners == f["owners"]).all(), k
        assert rects["corners"].tobytes() == f["corners"].tobytes(), k
        assert rects["pixel_size"].tobytes() == f["pixel"].tobytes(), k
        want_parent = f["pixel"][:, 0] != np.float32(1.0 / 32)
        assert ((rects["flags"] == pb.TEXRECT_PARENT) == want_parent).all(), k
    assert cache.count == 1024
    cache.close()


def awkward_lists(port):
    """(quad list, budget) per call: lists that are NOT one frame's leaf set -- a quad twice, a quad
    together with its parent, cached quads placed behind enough new ones that they are evicted before
    their own turn comes, quads with no cached parent and no budget.  These are the cases where a
    probe of the table as it stood at the start of the frame is not the answer."""
    f0d4, f0d5, f1d5 = port.uniform_quads(0, 4), port.uniform_quads(0, 5), port.uniform_quads(1, 5)
    roots = np.concatenate([port.uniform_quads(f, 0) for f in range(6)])
    return [
        (f0d4, 2000),                                                  # 256 parents, all generated
        (f0d5[:900], 100),                                             # 100 generated, 800 borrow the parent's quadrant
        (np.concatenate([f0d5[900:902], f0d5[900:902], f0d4[:2], f0d5[:3], f0d5[901:903]]), 1),   # duplicates, parent + child
        (f1d5[:800], 5000),                                            # fills the cache: evictions start mid-list
        (np.concatenate([f1d5[800:1000], f0d4[:64], f0d5[:40]]), 5000),    # the old entries at the back are evicted before their turn
        (np.concatenate([roots, port.uniform_quads(2, 3)[:20], f1d5[1000:1024]]), 0),   # no budget: roots and orphans generate anyway
        (np.concatenate([f1d5[::7], f0d5[::5], f0d4[::3]]), 37),
    ]


def run_lists(ref, port, plan):
    """Every list through the reference's GetHeightMapForQuad and through `plan`; same decisions."""
    ref.reset_cache()
    tex_owner, slot_owner = {}, {}
    for k, (quads, budget) in enumerate(awkward_lists(port)):
        want, want_count = ref.cache_lookup(quads, budget)
        rects, n_gen, count = plan(quads, budget)
        gen = want[:, 7] != 0
        assert n_gen == int(gen.sum()) and count == want_count, (k, n_gen, int(gen.sum()), count, want_count)
        assert ((rects["flags"] == pb.TEXRECT_GENERATED) == gen).all(), k
        for q, t, r, g in zip(quads, want[:, 0], rects, gen):
            if g:
                tex_owner[int(t)] = int(q["id"]); slot_owner[int(r["slot"])] = int(q["id"])
        assert [tex_owner[int(t)] for t in want[:, 0]] == [slot_owner[int(s)] for s in rects["slot"]], k
        assert rects["corners"].tobytes() == want[:, 1:5].tobytes(), k
        assert rects["pixel_size"].tobytes() == want[:, 5:7].tobytes(), k
        assert ((rects["flags"] == pb.TEXRECT_PARENT) == (want[:, 5] != np.float32(1.0 / 32))).all(), k


def test_plan_frame_on_lists_that_are_not_a_leaf_set(ref, port):
    cache = pb.HeightMapCache(32, 1024, 1499, extra_slots=6000)

    def plan(quads, budget):
        rects, n_gen = cache.plan_frame(quads, budget)
        return rects, n_gen, cache.count
    run_lists(ref, port, plan)
    cache.close()


def random_lists(port, seed, frames=24):
    """Seeded random quad lists over a universe of 2 304 ids (two faces at depth 5 plus their depth-4 parents and the
    roots): sizes 50..700, budgets 0..300, duplicates and parent/child mixes included, enough distinct ids to keep the
    1 024-entry cache evicting."""
    rng = np.random.default_rng(seed)
    universe = np.concatenate([port.uniform_quads(0, 5), port.uniform_quads(1, 5), port.uniform_quads(0, 4),
                               np.concatenate([port.uniform_quads(f, 0) for f in range(6)])])
    out = []
    for _ in range(frames):
        n = int(rng.integers(50, 700))
        hot = rng.integers(0, len(universe) - n)                       # a window (locality, like a camera) + scattered ids
        idx = np.concatenate([np.arange(hot, hot + n // 2), rng.integers(0, len(universe), n - n // 2)])
        rng.shuffle(idx)
        out.append((universe[idx], int(rng.choice([0, 1, 7, 100, 300]))))
    return out


def run_random(ref, port, plan, seed):
    ref.reset_cache()
    tex_owner, slot_owner = {}, {}
    for k, (quads, budget) in enumerate(random_lists(port, seed)):
        want, want_count = ref.cache_lookup(quads, budget)
        rects, n_gen, count = plan(quads, budget)
        gen = want[:, 7] != 0
        assert n_gen == int(gen.sum()) and count == want_count, (seed, k)
        assert ((rects["flags"] == pb.TEXRECT_GENERATED) == gen).all(), (seed, k)
        for q, t, r, g in zip(quads, want[:, 0], rects, gen):
            if g:
                tex_owner[int(t)] = int(q["id"]); slot_owner[int(r["slot"])] = int(q["id"])
        assert [tex_owner[int(t)] for t in want[:, 0]] == [slot_owner[int(s)] for s in rects["slot"]], (seed, k)
        assert rects["corners"].tobytes() == want[:, 1:5].tobytes() and rects["pixel_size"].tobytes() == want[:, 5:7].tobytes(), (seed, k)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_plan_frame_on_random_lists(ref, port, seed):
    cache = pb.HeightMapCache(32, 1024, 1499, extra_slots=4000)

    def plan(quads, budget):
        rects, n_gen = cache.plan_frame(quads, budget)
        return rects, n_gen, cache.count
    run_random(ref, port, plan, seed)
    cache.close()


def test_a_failed_frame_leaves_the_cache_unchanged(frames):
    cache = pb.HeightMapCache(32, 8, 11, extra_slots=4)
    few = frames[2]["quads"][:6]
    r0, n0 = cache.plan_frame(few, 100)
    assert n0 == 6 and cache.count == 6
    with pytest.raises(pb.PlanetGpuError, match="pool exhausted"):
        cache.plan_frame(frames[2]["quads"], 100)
    assert cache.count == 6
    r1, n1 = cache.plan_frame(few, 100)                         # still all hits, on the same slots
    assert n1 == 0 and (r1["slot"] == r0["slot"]).all() and (r1["flags"] == pb.TEXRECT_HIT).all()
    cache.close()


def test_one_cache_is_planned_on_the_host_or_on_the_device_not_both(frames):
    cache = pb.HeightMapCache(32, 1024, 1499, extra_slots=64)
    cache.plan_frame(frames[0]["quads"], 100)
    import ctypes as C
    q = np.ascontiguousarray(frames[0]["quads"])
    rects = np.zeros(len(q), pb.TEXRECT_DTYPE)
    p = pb.default_params()
    rc = pb.lib().planet_gpu_cache_frame(cache.handle, C.byref(p), q.ctypes.data, len(q), 18, 100, rects.ctypes.data, None, None)
    assert rc == -2 and b"already planned on the host" in pb.lib().planet_gpu_last_error()
    cache.close()


def test_draw_uniforms_of_the_glsl_stage_are_pinned(frames, port):
    """The GLSL stage itself cannot be executed here, but everything it is fed can be pinned: the
    per-quad uniforms P[4], N[4] and SkirtSize of the oracle (and of K3, which computes them the
    same way) equal, bit for bit, what the reference passed to glUniform on a 30-frame flight."""
    max_skirt = port.max_skirt_size()
    for k, f in enumerate(frames):
        got = port.quad_uniforms(f["quads"], f["cam"])
        assert got.tobytes() == f["pn"].tobytes(), k
        skirt = np.array([port.skirt_size_for_quad(max_skirt, q["id"]) for q in f["quads"]], np.float32)
        assert skirt.tobytes() == f["skirt"].tobytes(), k


def test_pool_exhaustion_is_an_error_not_a_corruption(frames):
    cache = pb.HeightMapCache(32, 8, 11, extra_slots=4)
    with pytest.raises(pb.PlanetGpuError, match="pool exhausted"):
        cache.plan_frame(frames[2]["quads"], 100)
    cache.close()


@pytest.mark.gpu
def test_cache_frames_generate_the_reference_maps_and_shade_through_texrects(frames, gpu, port):
    cache = gpu.HeightMapCache(32, 1024, 1499, extra_slots=1024)
    p = gpu.default_params()                                 # EXACT: maps must be the reference's bits
    for k, f in enumerate(frames[:8]):
        rects, d_rects = cache.frame(f["quads"], 18, p, 100)
        gen = rects["flags"] == gpu.TEXRECT_GENERATED
        if gen.any():
            got = cache.read_slots(rects["slot"][gen])
            want = np.stack([f["maps"][int(t)] for t in f["tex"][gen]])
            assert got.tobytes() == want.tobytes(), k        # one batched K2 launch == 1 GenerateHeightMap per quad
        # shading through the texrects (own maps and parent fallbacks) against the GLSL restatement
        dq = gpu.quads_to_device(f["quads"])
        pos, nrm = gpu.shade_cached(dq, cache, d_rects, f["cam"], p)
        pool = cache.read_slots(np.arange(cache.pool_slots, dtype=np.int32))
        r6 = np.concatenate([rects["corners"], rects["pixel_size"]], axis=1)
        wpos, wnrm = port.shade_patches_rect(f["quads"], f["cam"], pool, rects["slot"], r6)
        a, b = nrm.cpu().numpy()[..., :3].astype(np.float64), wnrm[..., :3].astype(np.float64)
        ang = np.arctan2(np.linalg.norm(np.cross(a, b), axis=-1), (a * b).sum(-1))
        assert ang.max() <= 1e-4, (k, ang.max())
        # bilinear weights carry ~1 ulp of fp32 texture-coordinate rounding (FMA vs not): the height
        # error scales with the map's slope, so bound it by 1e-5 of each map's height range
        dh = np.abs(pos.cpu().numpy()[..., 3].astype(np.float64) - wpos[..., 3]).max(axis=1)
        rng = np.ptp(pool[rects["slot"]].reshape(len(rects), -1), axis=1)
        assert (dh <= 1e-5 * rng + 1e-3).all(), (k, dh.max())
    cache.close()


@pytest.mark.gpu
def test_device_bookkeeping_takes_the_reference_decisions(frames, ref, port, gpu):
    """k_plan_frame (32 lanes, tables in device memory) against the reference's GetHeightMapForQuad:
    the 30-frame flight (owners, corners, pixel sizes, flags, counts) and the awkward lists."""
    cache = gpu.HeightMapCache(32, 1024, 1499, extra_slots=1024)
    host = gpu.HeightMapCache(32, 1024, 1499, extra_slots=1024)
    p = gpu.fbm_params(4, 0.5, gpu.FAST)                        # what is generated does not matter here
    slot_owner = {}
    for k, f in enumerate(frames):
        d_rects, n_gen = cache.frame_device(gpu.quads_to_device(f["quads"]), 18, p, 100)
        rects = d_rects.cpu().numpy().view(gpu.TEXRECT_DTYPE).reshape(-1)
        assert n_gen == f["generated"], (k, n_gen, f["generated"])
        for q, r in zip(f["quads"], rects):
            if r["flags"] == gpu.TEXRECT_GENERATED:
                slot_owner[int(r["slot"])] = int(q["id"])
        assert (np.array([slot_owner[int(s)] for s in rects["slot"]], np.uint64) == f["owners"]).all(), k
        assert rects["corners"].tobytes() == f["corners"].tobytes() and rects["pixel_size"].tobytes() == f["pixel"].tobytes(), k
        h_rects, h_gen = host.plan_frame(f["quads"], 100)       # one lane on the host: the same bytes
        assert h_rects.tobytes() == rects.tobytes() and h_gen == n_gen and host.count == cache.count, k
    assert cache.count == 1024
    cache.close(); host.close()

    cache = gpu.HeightMapCache(32, 1024, 1499, extra_slots=6000)

    def plan(quads, budget):
        d_rects, n_gen = cache.frame_device(gpu.quads_to_device(quads), 18, p, budget)
        return d_rects.cpu().numpy().view(gpu.TEXRECT_DTYPE).reshape(-1), n_gen, cache.count
    run_lists(ref, port, plan)
    cache.close()


@pytest.mark.gpu
def test_a_failed_device_frame_leaves_the_cache_unchanged(frames, gpu):
    cache = gpu.HeightMapCache(32, 8, 11, extra_slots=4)
    p = gpu.default_params()
    few = gpu.quads_to_device(frames[2]["quads"][:6])
    r0, n0 = cache.frame_device(few, 18, p, 100)
    with pytest.raises(gpu.PlanetGpuError, match="pool exhausted"):
        cache.frame_device(gpu.quads_to_device(frames[2]["quads"]), 18, p, 100)
    assert cache.count == 6
    r1, n1 = cache.frame_device(few, 18, p, 100)
    assert n0 == 6 and n1 == 0 and (r1[:, 0] == r0[:, 0]).all()
    cache.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 4])
def test_device_bookkeeping_on_random_lists(ref, port, gpu, seed):
    cache = gpu.HeightMapCache(32, 1024, 1499, extra_slots=4000)
    p = gpu.fbm_params(1, 0.5, gpu.FAST)

    def plan(quads, budget):
        d_rects, n_gen = cache.frame_device(gpu.quads_to_device(quads), 18, p, budget)
        return d_rects.cpu().numpy().view(gpu.TEXRECT_DTYPE).reshape(-1), n_gen, cache.count
    run_random(ref, port, plan, seed)
    cache.close()
