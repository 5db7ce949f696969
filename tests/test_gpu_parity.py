"""Parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against
 (a) the committed golden vectors produced by the reference's own code, and
 (b) the CPU oracle on the same seeded inputs,
plus size-independent properties at BASELINE.json's full size.

Bars (north_star): index buffers, QuadIDs, quad corners and permutation hashing bit-exact;
EXACT-mode heights bit-exact; FAST-mode heights |dh| <= 1e-5 * height_scale * sum(gain^k);
normals <= 1e-4 rad."""
import ctypes as C

import numpy as np
import pytest

from conftest import as_bits, quads_from_bytes
from oracle.bindings import FBM, RIDGED, fnv1a32, height_params

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # north_star: <= 1e-5 relative on height (relative to the fractal's amplitude sum)


def amp_sum(gain, octaves):
    return float(sum(np.float32(gain) ** k for k in range(max(octaves, 1))))


def dev_points(pb, pts):
    import torch
    return torch.from_numpy(np.ascontiguousarray(pts, np.float64)).cuda()


def to_np(t):
    return t.detach().cpu().numpy()


def orc_params(p):
    """planet_gpu_params -> oracle height_params (same constants)."""
    return height_params(kind=p.noise_kind, lacunarity=p.lacunarity, gain=p.gain,
                         fixed_octaves=p.fixed_octaves, coord_scale=p.coord_scale,
                         height_scale=p.height_scale)


# ---------------------------------------------------------------------------------------------
# perlin.h: hashing + PerlinNoise3, main.cpp:689-734 fractals
# ---------------------------------------------------------------------------------------------
def test_noise3_exact_is_bit_identical_to_reference(gpu, golden):
    got = to_np(gpu.noise(dev_points(gpu, golden["noise_points"]), octaves=0, precision=gpu.EXACT))
    assert (as_bits(got) == golden["noise_bits"]).all()


def test_noise3_fast_within_tolerance(gpu, golden):
    pts = golden["noise_points"]
    got = to_np(gpu.noise(dev_points(gpu, pts), octaves=0, precision=gpu.FAST))
    want = golden["noise_bits"].view(np.float32)
    err = np.abs(got.astype(np.float64) - want)
    assert err.max() <= REL_TOL, (err.max(), pts[err.argmax()])


@pytest.mark.parametrize("name,kind,gain,octs", [("fbm", FBM, 0.5, (1, 2, 8, 12, 16)),
                                                 ("ridged", RIDGED, 0.55, (1, 6, 7, 12, 18))])
def test_fractals_exact_bits_and_fast_tolerance(gpu, golden, name, kind, gain, octs):
    pts = dev_points(gpu, golden["fractal_points"])
    for o in octs:
        want_bits = golden[f"{name}_{o}_bits"]
        ex = to_np(gpu.noise(pts, kind=kind, lacunarity=2.0, gain=gain, octaves=o, precision=gpu.EXACT))
        assert (as_bits(ex) == want_bits).all(), (name, o)
        fa = to_np(gpu.noise(pts, kind=kind, lacunarity=2.0, gain=gain, octaves=o, precision=gpu.FAST))
        err = np.abs(fa.astype(np.float64) - want_bits.view(np.float32)).max()
        assert err <= REL_TOL * amp_sum(gain, o), (name, o, err)


def test_general_lacunarity_takes_the_exact_path(gpu, golden):
    pts = dev_points(gpu, golden["fractal_points"])
    for prec in (gpu.EXACT, gpu.FAST):      # FAST is only defined for lacunarity 2.0; falls to EXACT on the GPU
        got = to_np(gpu.noise(pts, kind=FBM, lacunarity=2.17, gain=0.47, octaves=6, precision=prec))
        assert (as_bits(got) == golden["fbm_lac217_gain047_6_bits"]).all()


def test_noise_against_oracle_on_seeded_points(gpu, port):
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-70, 70, (20000, 3)), rng.uniform(-1e5, 1e5, (5000, 3)),
                          np.floor(rng.uniform(-300, 300, (2000, 3)))])        # incl. lattice points
    d = dev_points(gpu, pts)
    for kind, gain, o in ((FBM, 0.5, 8), (RIDGED, 0.55, 12), (FBM, 0.5, 0)):
        want = port.noise3(pts) if o == 0 else port.fractal(pts, kind, 2.0, gain, o)
        ex = to_np(gpu.noise(d, kind=kind, gain=gain, octaves=o, precision=gpu.EXACT))
        assert (as_bits(ex) == as_bits(want)).all()
        fa = to_np(gpu.noise(d, kind=kind, gain=gain, octaves=o, precision=gpu.FAST))
        assert np.abs(fa.astype(np.float64) - want).max() <= REL_TOL * amp_sum(gain, o)


def test_points_kernels_at_scale_fast_vs_exact(gpu):
    """1.3 M points take the replicated-table points kernel; FAST must track EXACT everywhere."""
    import torch
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    pts = (torch.rand((1_300_000, 3), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 140.0
    for kind, gain, o in ((FBM, 0.5, 8), (RIDGED, 0.55, 12)):
        ex = gpu.noise(pts, kind=kind, gain=gain, octaves=o, precision=gpu.EXACT)
        fa = gpu.noise(pts, kind=kind, gain=gain, octaves=o, precision=gpu.FAST)
        assert (ex.double() - fa.double()).abs().max().item() <= REL_TOL * amp_sum(gain, o)
    sphere = pts / pts.norm(dim=1, keepdim=True) * 6371000.0
    p_ex, p_fa = gpu.default_params(), gpu.default_params(precision=gpu.FAST)
    ex = gpu.heights_at(sphere, 9, 18, p_ex); fa = gpu.heights_at(sphere, 9, 18, p_fa)
    assert (ex.double() - fa.double()).abs().max().item() <= REL_TOL * 8848.0 * amp_sum(0.55, 12)


# ---------------------------------------------------------------------------------------------
# K1: QuadIDs, corners, patch mesh, merged index buffer -- all bit-exact
# ---------------------------------------------------------------------------------------------
def test_tessellate_uniform_corners_and_ids_bit_exact(gpu, golden, port):
    roots = gpu.quads_to_host(gpu.tessellate_uniform(0))
    assert roots.tobytes() == golden["root_quads"].tobytes()
    d2 = gpu.quads_to_host(gpu.tessellate_uniform(2))
    assert d2.tobytes() == golden["depth2_quads"].tobytes()
    d5 = gpu.quads_to_host(gpu.tessellate_uniform(5, first=0, nquads=1024))
    assert fnv1a32(d5) == int(golden["depth5_face0_fnv"])
    d7 = gpu.quads_to_host(gpu.tessellate_uniform(7, first=3 * 16384, nquads=16384))
    assert fnv1a32(d7) == int(golden["depth7_face3_fnv"])
    # a ragged sub-range that straddles two faces, against the oracle
    sub = gpu.quads_to_host(gpu.tessellate_uniform(4, first=250, nquads=13))
    want = np.concatenate([port.uniform_quads(0, 4), port.uniform_quads(1, 4)])[250:263]
    assert sub.tobytes() == want.tobytes()


def test_quads_from_ids_reproduces_processquad_leaves(gpu, golden):
    quads = quads_from_bytes(golden["frame_quads"])               # depths 0..10, from ProcessQuad itself
    got = gpu.quads_to_host(gpu.quads_from_ids(quads["id"].copy()))
    assert got.tobytes() == quads.tobytes()


def test_quads_from_ids_deepest_path_and_invalid_id(gpu, golden, port):
    ids = golden["quad_ids"]
    got = gpu.quads_to_host(gpu.quads_from_ids(np.concatenate([ids, np.array([0], np.uint64)])))
    for k, i in enumerate(ids):
        assert got[k].tobytes() == port.quad_from_id(i).tobytes(), hex(int(i))
    assert int(got[-1]["id"]) == 0                                # invalid id (main.cpp:35) -> zero quad


def test_quads_from_ids_on_random_paths(gpu, port):
    """400 seeded QuadIDs, every root, depths 0..27, random child paths (all four children at every level within a
    warp: the walk selects its two midpoints without branching) against the restatement of main.cpp:546-547, 581-594."""
    rng = np.random.default_rng(7)
    ids = []
    for _ in range(400):
        qid = port.make_root_id(int(rng.integers(0, 6)))
        for _ in range(int(rng.integers(0, 28))):
            qid = port.make_child_id(qid, int(rng.integers(0, 4)))
        ids.append(qid)
    ids = np.array(ids, np.uint64)
    got = gpu.quads_to_host(gpu.quads_from_ids(ids))
    for k, i in enumerate(ids):
        assert got[k].tobytes() == port.quad_from_id(i).tobytes(), hex(int(i))


def test_patch_mesh_bit_exact(gpu, golden, port):
    v, i = gpu.patch_mesh(30)
    assert to_np(v).tobytes() == golden["patch_vertex_buffer"].tobytes()
    assert to_np(i).view(np.uint32).tobytes() == golden["patch_index_buffer"].tobytes()
    for n in (2, 5, 31, 50):
        v, i = gpu.patch_mesh(n)
        assert to_np(v).tobytes() == port.patch_vertices(n).tobytes(), n
        assert (to_np(i).view(np.uint32) == port.patch_indices(n)).all(), n


def test_merged_index_buffer(gpu, golden):
    strip = golden["patch_index_buffer"].view(np.uint32)
    for nq in (1, 3, 96):
        _, idx = gpu.tessellate_uniform(2, first=0, nquads=nq, with_indices=True)
        idx = to_np(idx).view(np.uint32).reshape(nq, 2036)
        want = strip[None, :] + (np.arange(nq, dtype=np.uint32) * 1020)[:, None]
        assert (idx == want).all(), nq
    p = gpu.default_params(patch_verts=5)                         # total % 4 == 2 tail
    _, idx = gpu.tessellate_uniform(1, first=0, nquads=3, params=p, with_indices=True)
    ni, nv = gpu.patch_index_count(5), gpu.patch_vertex_count(5)
    L = gpu.lib()
    strip5 = np.array([L.planet_gpu_strip_index(k, 5) for k in range(ni)], np.uint32)
    want = strip5[None, :] + (np.arange(3, dtype=np.uint32) * nv)[:, None]
    assert (to_np(idx).view(np.uint32).reshape(3, ni) == want).all()


# ---------------------------------------------------------------------------------------------
# K0: camera-driven LOD selection (ProcessQuad / RenderPlanet, main.cpp:537-624)
# ---------------------------------------------------------------------------------------------
def test_select_lod_reproduces_the_reference_default_frame(gpu, golden):
    """planet.quads of the reference's first frame: 117 leaves, depths 0..10, same order, same bytes."""
    got = gpu.quads_to_host(gpu.select_lod(golden["frame_cam"], int(golden["max_lod"])))
    assert len(got) == 117
    assert got.tobytes() == golden["frame_quads"].tobytes()


def test_select_lod_other_cameras_against_the_reference(gpu, ref):
    for cam in ([3.0e6, 5.0e6, -4.5e6], [0.0, 6371000.0 + 2000.0, 0.0], [-9.0e6, 1.0e5, 2.0e5], [0.0, 0.0, 0.0]):
        want, _, _ = ref.render_frame(np.array(cam, np.float64))
        got = gpu.quads_to_host(gpu.select_lod(cam))
        assert len(got) == len(want), (cam, len(got), len(want))
        assert got.tobytes() == want.tobytes(), cam


def test_select_lod_on_random_cameras(gpu, ref):
    """24 seeded cameras from 1 m above the sphere to 3 radii out, anywhere around the planet (and a few inside
    it): the leaves of the reference's RenderPlanet, in its order, byte for byte."""
    rng = np.random.default_rng(20261019)
    R = 6371000.0
    for k in range(24):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        alt = float(np.exp(rng.uniform(0.0, np.log(3.0 * R)))) if k % 6 else -float(rng.uniform(1.0, 0.5 * R))
        cam = d * (R + alt)
        want, _, _ = ref.render_frame(cam)
        got = gpu.quads_to_host(gpu.select_lod(cam))
        assert len(got) == len(want), (k, cam, len(got), len(want))
        assert got.tobytes() == want.tobytes(), (k, cam)


def test_select_lod_fallback_paths(gpu, golden):
    """The device-wide sort (more leaves than one CTA sorts) and the per-level launch path."""
    import os, subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, planet_b200 as pb; pb.init(0);"
            "g = np.load(%r); q = pb.quads_to_host(pb.select_lod(g['frame_cam'], 18));"
            "print('OK' if q.tobytes() == g['frame_quads'].tobytes() else 'BAD')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = code % (root, os.path.join(root, "tests", "golden", "planet_golden.npz"))
    for knob in ({"PLANET_K0_BLOCK_SORT_MAX": "64"}, {"PLANET_K0_LEVEL_LAUNCHES": "1"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **knob), capture_output=True, text=True)
        assert out.stdout.strip().endswith("OK"), (knob, out.stdout, out.stderr)


def test_select_lod_capacity_error(gpu, golden):
    with pytest.raises(gpu.PlanetGpuError, match="capacity"):
        gpu.select_lod(golden["frame_cam"], capacity=40)


# ---------------------------------------------------------------------------------------------
# K2: GenerateHeightMap / GetHeightAt
# ---------------------------------------------------------------------------------------------
def test_default_frame_height_maps_exact_bits(gpu, golden):
    """117 height maps the reference's real main() produced on its first frame (ridged, 6..12 octaves)."""
    quads = quads_from_bytes(golden["frame_quads"])
    got = to_np(gpu.generate_height_maps(gpu.quads_to_device(quads), 32, int(golden["max_lod"]),
                                         gpu.default_params()))
    assert got.tobytes() == golden["frame_height_maps"].tobytes()


def test_default_frame_height_maps_fast_tolerance(gpu, golden, port):
    quads = quads_from_bytes(golden["frame_quads"])
    p = gpu.default_params(precision=gpu.FAST)
    got = to_np(gpu.generate_height_maps(gpu.quads_to_device(quads), 32, int(golden["max_lod"]), p))
    want = golden["frame_height_maps"]
    for k, q in enumerate(quads):
        o = 6 + 12 * port.get_depth(int(q["id"])) // 18
        err = np.abs(got[k].astype(np.float64) - want[k]).max()
        assert err <= REL_TOL * 8848.0 * amp_sum(0.55, o), (k, o, err)


@pytest.mark.parametrize("precision", ["EXACT", "FAST"])
def test_fbm_height_maps_golden_incl_ragged_dims(gpu, golden, port, precision):
    prec = getattr(gpu, precision)
    ml = int(golden["max_lod"])

    def check(quads, dim, octaves, want, kind=gpu.FBM, gain=0.5):
        p = gpu.default_params(noise_kind=kind, gain=gain, fixed_octaves=octaves, precision=prec)
        got = to_np(gpu.generate_height_maps(gpu.quads_to_device(quads), dim, ml, p))
        if prec == gpu.EXACT:
            assert got.tobytes() == want.tobytes(), (dim, octaves)
        else:
            o = octaves if octaves > 0 else 6
            err = np.abs(got.astype(np.float64) - want).max()
            assert err <= REL_TOL * 8848.0 * amp_sum(gain, o), (dim, octaves, err)

    d5 = port.uniform_quads(0, 5)
    check(d5[golden["fbm8_depth5_pick"]], 32, 8, golden["fbm8_depth5_maps"])
    q7 = quads_from_bytes(golden["fbm8_depth7_quads"])
    check(q7, 32, 8, golden["fbm8_depth7_maps"])
    for dim in (4, 5, 7, 33, 52, 64):                              # minimum (dim > 3), odd and ragged sizes
        check(q7[:2], dim, 12, golden[f"fbm12_dim{dim}_maps"])
    roots = quads_from_bytes(golden["root_quads"])[[1, 5]]
    check(roots, 128, 0, golden["ridged_roots15_dim128_maps"], kind=gpu.RIDGED, gain=0.55)


def test_fbm8_depth5_checksum_of_checksums(gpu, golden):
    """All 1 024 face-0 depth-5 maps, EXACT: FNV of the whole buffer == the reference's."""
    quads = gpu.tessellate_uniform(5, first=0, nquads=1024)
    p = gpu.default_params(noise_kind=gpu.FBM, gain=0.5, fixed_octaves=8)
    maps = to_np(gpu.generate_height_maps(quads, 32, 18, p))
    assert fnv1a32(maps) == int(golden["fbm8_depth5_fnv"]) == 0x62660FBE


def test_ridged_18_octaves_and_zero_functor(gpu, port):
    quads = port.uniform_quads(2, 3)[::7]
    dq = gpu.quads_to_device(quads)
    for prec in (gpu.EXACT, gpu.FAST):
        p = gpu.default_params(fixed_octaves=18, precision=prec)
        got = to_np(gpu.generate_height_maps(dq, 32, 18, p))
        want = port.generate_height_maps(quads, 32, 18, orc_params(p))
        if prec == gpu.EXACT:
            assert got.tobytes() == want.tobytes()
        else:
            assert np.abs(got.astype(np.float64) - want).max() <= REL_TOL * 8848.0 * amp_sum(0.55, 18)
        z = gpu.default_params(noise_kind=gpu.ZERO, precision=prec)           # ConstantZero, main.cpp:835-841
        assert not to_np(gpu.generate_height_maps(dq, 32, 18, z)).any()


def test_heights_at_golden_and_legacy_scalar_shim(gpu, golden):
    pts, depth = golden["height_at_points"], golden["height_at_depth"]
    want = golden["height_at_bits"]
    for d in np.unique(depth):
        sel = depth == d
        ex = to_np(gpu.heights_at(dev_points(gpu, pts[sel]), int(d), 18, gpu.default_params()))
        assert (as_bits(ex) == want[sel]).all(), d
        fa = to_np(gpu.heights_at(dev_points(gpu, pts[sel]), int(d), 18, gpu.default_params(precision=gpu.FAST)))
        o = 6 + 12 * int(d) // 18
        assert np.abs(fa.astype(np.float64) - want[sel].view(np.float32)).max() <= REL_TOL * 8848.0 * amp_sum(0.55, o)
    # ProcessQuad's call shape: GetHeightAt(p, 0, 1) through the reference-shaped pointer
    gpu.set_params(gpu.default_params())
    lod = np.array([gpu.get_height_at(p, 0, 1) for p in pts[:16]], np.float32)
    assert (as_bits(lod) == golden["height_at_lod_bits"][:16]).all()


def test_reference_shaped_generate_height_map(gpu, golden):
    """void GenerateHeightMap(float*, int, const Quad&, int) with host buffers, as GetHeightMapForQuad calls it."""
    quads = quads_from_bytes(golden["frame_quads"])
    gpu.set_params(gpu.default_params())
    for k in (0, 57, 116):
        got = gpu.generate_height_map(quads[k], 32, int(golden["max_lod"]))
        assert got.tobytes() == golden["frame_height_maps"][k].tobytes()


def test_reference_shaped_calls_from_eight_host_threads(gpu, golden):
    """SURVEY 8b threading: the reference's GenerateHeightMap is a pure function, callable from 8
    threads at once with identical output; the drop-in serialises on a mutex and must give the same
    bits whichever thread calls (ctypes drops the GIL for the duration of each call)."""
    from concurrent.futures import ThreadPoolExecutor
    quads = quads_from_bytes(golden["frame_quads"])
    gpu.set_params(gpu.default_params())
    ks = list(range(0, len(quads), 3))
    max_lod, want_maps = int(golden["max_lod"]), golden["frame_height_maps"]   # NpzFile is not thread-safe: read first

    def one(k):
        m = gpu.generate_height_map(quads[k], 32, max_lod)
        h = gpu.get_height_at(np.array(quads[k]["p"][0]), 0, 1)
        return k, m.tobytes(), h

    with ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(one, ks * 3))
    want_h = {}
    for k, m, h in results:
        assert m == want_maps[k].tobytes(), k
        assert want_h.setdefault(k, h) == h and np.isfinite(h)


def test_seam_calls_fall_back_to_the_batched_path_when_the_fast_lane_does_not_apply(gpu, golden):
    """FAST params, a map too large for the fast lane, and the ZERO functor through the two
    reference-shaped calls: same results as the batched device API."""
    import torch
    quads = quads_from_bytes(golden["frame_quads"])
    pt = np.array(quads[5]["p"][2])
    try:
        fast = gpu.default_params(precision=gpu.FAST)
        gpu.set_params(fast)
        got = gpu.generate_height_map(quads[40], 32, 18)
        want = to_np(gpu.generate_height_maps(gpu.quads_to_device(quads[40:41]), 32, 18, fast))[0]
        assert got.tobytes() == want.tobytes()
        h = gpu.get_height_at(pt, 0, 1)
        want_h = to_np(gpu.heights_at(dev_points(gpu, pt[None, :]), 0, 1, fast))[0]
        assert np.float32(h).tobytes() == np.float32(want_h).tobytes()
        exact = gpu.default_params()
        gpu.set_params(exact)
        big = gpu.generate_height_map(quads[3], 2304, 18)          # > 2048: batch of one through the host path
        want_big = to_np(gpu.generate_height_maps(gpu.quads_to_device(quads[3:4]), 2304, 18, exact))[0]
        assert big.tobytes() == want_big.tobytes()
        gpu.set_params(gpu.default_params(noise_kind=gpu.ZERO))
        assert not gpu.generate_height_map(quads[0], 32, 18).any()
        assert gpu.get_height_at(pt, 0, 1) == 0.0
    finally:
        gpu.set_params(gpu.default_params())


def test_shutdown_releases_and_the_library_comes_back(gpu, golden):
    """planet_gpu_shutdown frees the staging buffers, the LOD scratch and the cached strip; the next
    call re-initialises lazily and every path that owned one of them still gives the same bytes."""
    quads = quads_from_bytes(golden["frame_quads"])
    gpu.set_params(gpu.default_params())
    before = gpu.generate_height_map(quads[7], 32, 18)
    _, idx_before = gpu.tessellate_uniform(2, first=0, nquads=5, with_indices=True)
    lod_before = gpu.quads_to_host(gpu.select_lod((0.0, 0.0, -6371010.0), 18, gpu.default_params()))
    gpu.lib().planet_gpu_shutdown()
    gpu.set_params(gpu.default_params())
    assert gpu.generate_height_map(quads[7], 32, 18).tobytes() == before.tobytes()
    _, idx_after = gpu.tessellate_uniform(2, first=0, nquads=5, with_indices=True)
    assert to_np(idx_after).tobytes() == to_np(idx_before).tobytes()
    lod_after = gpu.quads_to_host(gpu.select_lod((0.0, 0.0, -6371010.0), 18, gpu.default_params()))
    assert lod_after.tobytes() == lod_before.tobytes()


def test_large_maps_dim_1024_and_4096(gpu, golden):
    """C5-sized single patches: the reference's own dim-1024 probe (first leaf of the
    default frame, default ridged functor, FNV-1a 0x57fc9fba) through both EXACT kernels, and FAST
    against EXACT at 1024^2 and 4096^2."""
    from oracle.bindings import fnv1a32
    quads = quads_from_bytes(golden["frame_quads"])
    dq = gpu.quads_to_device(quads[:2])
    exact, fast = gpu.default_params(), gpu.default_params(precision=gpu.FAST)
    one = to_np(gpu.generate_height_maps(dq[:1], 1024, 18, exact))            # 2^20 samples: small-batch kernel
    assert fnv1a32(one) == int(golden["ridged_leaf0_dim1024_fnv"])
    two = to_np(gpu.generate_height_maps(dq, 1024, 18, exact))                # 2^21 samples: table kernel
    assert two[0].tobytes() == one[0].tobytes()
    tol = REL_TOL * 8848.0 * amp_sum(0.55, 6)
    f1 = to_np(gpu.generate_height_maps(dq, 1024, 18, fast))
    assert np.abs(f1.astype(np.float64) - two).max() <= tol
    e4 = gpu.generate_height_maps(dq[:1], 4096, 18, exact)
    f4 = gpu.generate_height_maps(dq[:1], 4096, 18, fast)
    assert bool((e4.double() - f4.double()).abs().max() <= tol)


@pytest.mark.parametrize("n,nq", [(30, 16384), (30, 7), (12, 333), (5, 40), (44, 100)])
def test_index_stream_beside_k2_writes_the_same_indices(gpu, n, nq):
    """planet_gpu_merged_indices_beside (the slim kernel that shares the SMs with K2; n = 5 has ni % 4 != 0 and
    takes the general path) against K1's own index stream, alone and while K2 runs on another stream."""
    import torch
    p = gpu.fbm_params(8, 0.5, gpu.FAST, patch_verts=n)
    _, want = gpu.tessellate_uniform(7, first=0, nquads=nq, params=p, with_indices=True)
    assert torch.equal(gpu.merged_indices_beside(nq, p), want)
    quads = gpu.tessellate_uniform(7, first=0, nquads=max(nq, 2048), params=p)
    side = torch.cuda.Stream()
    out = torch.full_like(want, -1)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        gpu.merged_indices_beside(nq, p, out=out, stream=side)
    maps = gpu.generate_height_maps(quads, n + 2, 18, p)                    # main stream, at the same time
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    assert torch.equal(maps, gpu.generate_height_maps(quads, n + 2, 18, p))


def test_k1_indices_at_c5_patch_sizes(gpu, port):
    """Merged strip indices for the large patches of config 5: n = 126 (the strip still fits shared memory) and
    n = 254 (it does not: the indices come from the kernel that reads the strip through L1); an odd n past
    that limit is refused with a message, not a launch failure."""
    import torch
    for n in (126, 254):
        p = gpu.default_params(patch_verts=n)
        nv, ni = gpu.patch_vertex_count(n), gpu.patch_index_count(n)
        quads, idx = gpu.tessellate_uniform(2, first=5, nquads=3, params=p, with_indices=True)
        strip = port.patch_indices(n).astype(np.int64)
        want = (np.arange(3, dtype=np.int64)[:, None] * nv + strip[None, :]).astype(np.uint32)
        assert to_np(idx).view(np.uint32).reshape(3, ni).tobytes() == want.tobytes(), n
        assert gpu.quads_to_host(quads).tobytes() == np.concatenate([port.uniform_quads(f, 2) for f in range(6)])[5:8].tobytes()
    with pytest.raises(gpu.PlanetGpuError, match="does not fit"):
        gpu.tessellate_uniform(1, first=0, nquads=1, params=gpu.default_params(patch_verts=201), with_indices=True)
    torch.cuda.synchronize()


@pytest.mark.parametrize("dim", [128, 512, 2048])
def test_c5_dims_between_the_probes(gpu, port, golden, dim):
    """BASELINE config 5 sweeps dim 64..4096; 1024 and 4096 are covered above, these are the sizes
    between: EXACT bit-identical to the oracle on whole maps (the small-batch kernel at 128 and 512,
    the table kernel at 2048), FAST within tolerance, for a shallow and a deep leaf of the frame."""
    quads = quads_from_bytes(golden["frame_quads"])[[0, 116 if dim < 2048 else 0]][: 2 if dim < 2048 else 1]
    exact, fast = gpu.default_params(), gpu.default_params(precision=gpu.FAST)
    dq = gpu.quads_to_device(quads)
    got = to_np(gpu.generate_height_maps(dq, dim, 18, exact))
    want = port.generate_height_maps(quads, dim, 18, orc_params(exact), nthreads=8)
    assert as_bits(got).tobytes() == as_bits(want).tobytes()
    f = to_np(gpu.generate_height_maps(dq, dim, 18, fast))
    octs = 6 + 12 * int((int(quads["id"].max()) >> 55) & 31) // 18
    assert np.abs(f.astype(np.float64) - want).max() <= REL_TOL * 8848.0 * amp_sum(0.55, octs)


@pytest.mark.parametrize("n", [62, 126, 254])
def test_c5_shade_at_larger_patches(gpu, port, golden, n):
    """K3 at the C5 patch sizes it supports (dim = n + 2 = 64, 128, 256: the map of a patch must fit
    the warp's shared-memory staging; larger patches return PLANET_E_UNSUPPORTED, tested below)."""
    quads = quads_from_bytes(golden["frame_quads"])[[3, 60]]
    maps = port.generate_height_maps(quads, n + 2, 18, height_params(), nthreads=4)
    check_shade(gpu, port, quads, maps, golden["frame_cam"], n=n)
    if n == 254:
        import torch
        with pytest.raises(gpu.PlanetGpuError, match="patch_verts"):
            gpu.shade(gpu.quads_to_device(quads), torch.zeros((2, 258, 258), device="cuda"), golden["frame_cam"], gpu.default_params(patch_verts=256))


@pytest.mark.parametrize("n,nq", [(44, 3), (62, 1), (126, 5), (254, 2)])
def test_shade_cta_per_chunk_kernel_gives_the_warp_per_quad_bytes(gpu, golden, monkeypatch, n, nq):
    """Few large patches are shaded by a CTA per (quad, chunk of slots) instead of a warp per quad: same
    per-vertex code, same column records -- the outputs must be the same bytes."""
    import torch
    quads = gpu.quads_to_device(quads_from_bytes(golden["frame_quads"])[[1, 30, 77, 100, 116][:nq]])
    p = gpu.fbm_params(5, 0.5, gpu.FAST, patch_verts=n)
    maps = gpu.generate_height_maps(quads, n + 2, 18, p)
    cam = golden["frame_cam"]
    pos_w, nrm_w = gpu.shade(quads, maps, cam, p)
    monkeypatch.setenv("PLANET_K3_NO_WIDE", "1")
    pos_q, nrm_q = gpu.shade(quads, maps, cam, p)
    assert torch.equal(pos_w, pos_q) and torch.equal(nrm_w, nrm_q)
    assert bool(torch.isfinite(pos_w).all()) and bool((nrm_w[..., :3].norm(dim=-1) - 1).abs().max() < 1e-5)


def test_batched_api_from_four_threads_on_four_streams(gpu):
    """The batched entry points are stream-ordered and re-entrant: four host threads, each on its
    own CUDA stream, run K1 + K2 + K3 on different quad ranges and get the single-threaded bytes."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    p = gpu.fbm_params(octaves=8, gain=0.5, precision=gpu.FAST)
    cam = (0.0, 0.0, -6371010.0)

    def run(k, stream=None):
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
        with ctx:
            quads, idx = gpu.tessellate_uniform(6, first=1000 * k, nquads=700 + 100 * k, params=p, with_indices=True, stream=stream)
            maps = gpu.generate_height_maps(quads, 32, 18, p, stream=stream)
            pos, nrm = gpu.shade(quads, maps, cam, p, stream=stream)
            if stream is not None:
                stream.synchronize()
            else:
                torch.cuda.synchronize()
            return [t.cpu() for t in (quads, idx, maps, pos, nrm)]

    want = [run(k) for k in range(4)]
    streams = [torch.cuda.Stream() for _ in range(4)]
    for _ in range(3):
        with ThreadPoolExecutor(max_workers=4) as pool:
            got = list(pool.map(lambda k: run(k, streams[k]), range(4)))
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                assert torch.equal(a, b)


def test_the_step_can_be_captured_in_a_cuda_graph(gpu):
    """K1 + K2 + K3 make no host synchronisation and no allocation of their own, so a frame-sized
    step captures into a CUDA graph (the way to take the launch latency out of small frames) and
    replays to the same bytes -- also when K1 meets a patch size whose strip it has not cached yet
    (it then evaluates the closed form in the kernel instead of uploading during capture)."""
    import torch
    cam = (0.0, 0.0, -6371010.0)
    for patch in (30, 22):                                         # 22: strip not cached before the capture
        p = gpu.fbm_params(octaves=8, gain=0.5, precision=gpu.FAST, patch_verts=patch)
        dim, nq = patch + 2, 117
        nv, ni = gpu.patch_vertex_count(patch), gpu.patch_index_count(patch)
        quads = torch.empty((nq, 13), dtype=torch.int64, device="cuda")
        idx = torch.empty(nq * ni, dtype=torch.int32, device="cuda")
        maps = torch.empty((nq, dim, dim), dtype=torch.float32, device="cuda")
        pos = torch.empty((nq, nv, 4), dtype=torch.float32, device="cuda")
        nrm = torch.empty_like(pos)
        L, C = gpu.lib(), gpu.C

        def step(stream):
            sp = C.c_void_p(stream.cuda_stream)
            gpu._check(L.planet_gpu_tessellate_uniform(C.byref(p), 5, 300, nq, quads.data_ptr(), idx.data_ptr(), sp))
            gpu._check(L.planet_gpu_generate_height_maps(C.byref(p), quads.data_ptr(), nq, dim, 18, maps.data_ptr(), sp))
            camv = (C.c_double * 3)(*cam)
            gpu._check(L.planet_gpu_shade(C.byref(p), quads.data_ptr(), nq, camv, maps.data_ptr(), -1.0,
                                          pos.data_ptr(), nrm.data_ptr(), sp))

        if patch == 30:
            step(torch.cuda.current_stream())                      # warm: function attributes, strip cache
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step(torch.cuda.current_stream())
        for t in (quads, idx, maps, pos, nrm):
            t.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got = [t.clone() for t in (quads, idx, maps, pos, nrm)]
        step(torch.cuda.current_stream())
        torch.cuda.synchronize()
        for a, b in zip(got, (quads, idx, maps, pos, nrm)):
            assert torch.equal(a, b)
        assert bool(maps.abs().max() > 0)


def test_host_batch_path_equals_device_path(gpu, golden):
    quads = quads_from_bytes(golden["frame_quads"])
    p = gpu.default_params(precision=gpu.FAST)
    host = gpu.generate_height_maps_host(quads, 32, 18, p)
    dev = to_np(gpu.generate_height_maps(gpu.quads_to_device(quads), 32, 18, p))
    assert host.tobytes() == dev.tobytes()


@pytest.mark.parametrize("nquads", [117, 1024, 6144, 16384])    # 1, 2, 4 and 7 pipeline chunks (1, 2, 4, 8, 8 ... waves)
def test_terrain_host_equals_separate_calls(gpu, nquads):
    """planet_gpu_terrain_host = generate_height_maps_host + shade, byte for byte (K3 only
    overlaps the PCIe drain; it must still see every finished map)."""
    torch = gpu._torch()
    p = gpu.fbm_params(octaves=8, gain=0.5, precision=gpu.FAST)
    depth = 6
    first = 4 ** depth + 11
    dq = gpu.tessellate_uniform(depth, first, nquads, p)
    quads = to_np(dq).view(gpu.QUAD_DTYPE).reshape(-1)
    cam = (0.0, 0.0, 3.0 * p.radius)
    out, heights, pos, nrm = gpu.terrain_host(quads, 18, cam, p)
    want_h = gpu.generate_height_maps_host(quads, 32, 18, p)
    assert out.tobytes() == want_h.tobytes()
    assert to_np(heights).tobytes() == want_h.tobytes()
    want_pos, want_nrm = gpu.shade(dq, heights, cam, p)
    assert torch.equal(pos, want_pos) and torch.equal(nrm, want_nrm)


def test_seed_offset_is_a_coordinate_shift(gpu, port):
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(4096, 3)); pts = pts / np.linalg.norm(pts, axis=1, keepdims=True) * 6371000.0
    seed = (17.25, -3.5, 101.125)
    p = gpu.default_params(noise_kind=gpu.FBM, gain=0.5, fixed_octaves=8, seed_offset=seed)
    got = to_np(gpu.heights_at(dev_points(gpu, pts), 0, 18, p))
    shifted = pts * 0.00001 + np.array(seed)                       # same two roundings as the kernel
    want = port.fractal(shifted, FBM, 2.0, 0.5, 8) * np.float32(8848.0)
    assert (as_bits(got) == as_bits(want)).all()


def test_wide_quads_take_the_per_sample_reduction(gpu, port):
    """FAST: a quad spanning more than ~100 lattice cells cannot be reduced mod 256 per quad."""
    roots = port.root_quads()
    p = gpu.default_params(noise_kind=gpu.FBM, gain=0.5, fixed_octaves=6, coord_scale=1e-4, precision=gpu.FAST)
    got = to_np(gpu.generate_height_maps(gpu.quads_to_device(roots), 48, 18, p))
    want = port.generate_height_maps(roots, 48, 18, orc_params(p))
    assert np.abs(got.astype(np.float64) - want).max() <= REL_TOL * 8848.0 * amp_sum(0.5, 6)


def test_fast_mode_steps_aside_where_its_shifts_would_wrap(gpu, port, golden):
    """FAST is the lattice split by 32-bit funnel shifts and a 40-bit magic division: neither holds
    for more than 32 octaves or dim > 8192.  The library must then run the EXACT arithmetic on the
    GPU (never wrong heights): with max_depth 9 a depth-10 quad asks for 6 + 12*10/9 = 19 octaves, but
    the bound over every depth a QuadID can carry (31 -> 47) is over 32; dim 8200 is past the
    division's range.  (Beyond ~30 octaves the reference itself casts out-of-range doubles to int.)"""
    import torch
    deep = quads_from_bytes(golden["frame_quads"])
    deep = deep[np.argsort((deep["id"] >> np.uint64(55)) & np.uint64(31))[-3:]].copy()     # the frame's deepest leaves (depth 10)
    fast = gpu.default_params(precision=gpu.FAST)                   # ridged 0.55, 6 + 12*depth/max_depth octaves
    exact = gpu.default_params(precision=gpu.EXACT)
    got = to_np(gpu.generate_height_maps(gpu.quads_to_device(deep), 32, 9, fast))
    want = port.generate_height_maps(deep, 32, 9, orc_params(exact))
    assert np.isfinite(want).all()
    assert as_bits(got).tobytes() == as_bits(want).tobytes()        # EXACT path taken: the reference's bits
    one = gpu.quads_to_device(port.root_quads()[:1])
    p8 = gpu.fbm_params(2, 0.5, gpu.FAST)
    a = gpu.generate_height_maps(one, 8200, 18, p8)
    b = gpu.generate_height_maps(one, 8200, 18, gpu.fbm_params(2, 0.5, gpu.EXACT))
    assert torch.equal(a, b)
    rows = [0, 1, 4099, 8199]
    sub = to_np(a[0, rows][:, ::911])
    xs = np.arange(0, 8200, 911)
    q = port.root_quads()[0]["p"]
    for ri, y in enumerate(rows):                                   # main.cpp:132-146 by hand for a few texels
        for ci, x in enumerate(xs):
            fx, fy = (x - 1) / (8200 - 3), (y - 1) / (8200 - 3)
            v0, v1 = q[1] - q[0], q[3] - q[2]
            top, bot = q[0] + v0 * fx, q[2] + v1 * fx
            pt = top + (bot - top) * fy
            h = port.get_height_at(pt, 0, 18, orc_params(gpu.fbm_params(2, 0.5, gpu.EXACT)))
            assert abs(float(sub[ri, ci]) - h) <= 1e-3 * 8848.0, (x, y)


def test_argument_validation_returns_errors(gpu, port):
    import torch
    q = gpu.quads_to_device(port.root_quads())
    with pytest.raises(gpu.PlanetGpuError, match="dim"):
        gpu.generate_height_maps(q, 3, 18)                          # main.cpp:128 assert(dim > 3)
    with pytest.raises(gpu.PlanetGpuError, match="max_depth"):
        gpu.generate_height_maps(q, 32, 0)                          # main.cpp:827 divides by max_depth
    with pytest.raises(gpu.PlanetGpuError, match="depth"):
        gpu.tessellate_uniform(28, first=0, nquads=1)               # depth + 1 < 32 and 55 path bits
    with pytest.raises(gpu.PlanetGpuError, match="leaf range"):
        gpu.tessellate_uniform(2, first=90, nquads=10)              # only 96 leaves at depth 2
    with pytest.raises(gpu.PlanetGpuError):
        gpu.generate_height_maps(q, 32, 18, gpu.default_params(noise_kind=9))
    empty = torch.empty((0, 13), dtype=torch.int64, device="cuda")
    assert gpu.generate_height_maps(empty, 32, 18).shape == (0, 32, 32)     # empty input is a no-op


# ---------------------------------------------------------------------------------------------
# K3: displacement + normals + Lambert vs the GLSL restatement
# ---------------------------------------------------------------------------------------------
def angle(a, b):
    c = np.cross(a.astype(np.float64), b.astype(np.float64))
    return np.arctan2(np.linalg.norm(c, axis=-1), (a.astype(np.float64) * b).sum(-1))


def check_shade(gpu, port, quads, maps, cam, n=30):
    p = gpu.default_params(patch_verts=n)
    import torch
    pos, nrm = gpu.shade(gpu.quads_to_device(quads), torch.from_numpy(np.ascontiguousarray(maps)).cuda(), cam, p)
    pos, nrm = to_np(pos), to_np(nrm)
    wpos, wnrm = port.shade_patches(quads, cam, maps, n=n)
    ang = angle(nrm[..., :3], wnrm[..., :3])
    assert ang.max() <= 1e-4, ang.max()                             # north_star: <= 1e-4 angular on normals
    assert np.abs(nrm[..., 3] - wnrm[..., 3]).max() <= 1e-4        # Lambert colour
    assert (pos[..., 3] == wpos[..., 3]).all()                      # height - skirt: one fp32 op, exact
    for k, q in enumerate(quads):
        rel = np.abs(q["p"] - cam).max()
        ext = np.linalg.norm(q["p"][3] - q["p"][0])
        tol = 8 * 2.0 ** -23 * rel + 1e-5 * ext                     # few fp32 ulps of the camera-relative corner
        assert np.abs(pos[k, :, :3].astype(np.float64) - wpos[k, :, :3]).max() <= tol, k


def test_shade_default_frame(gpu, port, golden):
    """All 117 leaf quads (depths 0..10: both the slerp and the linear branch of interpolate)."""
    check_shade(gpu, port, quads_from_bytes(golden["frame_quads"]), golden["frame_height_maps"], golden["frame_cam"])


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_shade_on_random_cameras_and_patch_sizes(gpu, port, ref, seed):
    """Seeded sweep over what K3 is given in use: the leaf set the reference's RenderPlanet selects for a random camera
    (1 m above the surface to two radii out; so a quad's size stays in proportion to its distance, which is what keeps
    the shader's fp32 camera-relative arithmetic well conditioned -- a 600 m quad seen from 10 000 km is not a case the
    reference ever draws), patch sizes 2..61, noisy maps.  Depths 0..18: both branches of interpolate."""
    rng = np.random.default_rng(seed)
    n = int(rng.choice([2, 5, 12, 30, 31, 44, 61]))
    d = rng.normal(size=3); d /= np.linalg.norm(d)
    cam = d * (6371000.0 + float(np.exp(rng.uniform(0.0, np.log(2.0 * 6371000.0)))))
    leaves, _, _ = ref.render_frame(cam)
    quads = leaves[rng.choice(len(leaves), min(len(leaves), 24), replace=False)]
    maps = port.generate_height_maps(quads, n + 2, 18, height_params(kind=FBM, gain=0.6, fixed_octaves=int(rng.integers(1, 7))), nthreads=4)
    check_shade(gpu, port, quads, maps, cam, n=n)


def test_shade_other_patch_size_and_camera(gpu, port):
    quads = port.uniform_quads(4, 2)
    n = 12
    maps = port.generate_height_maps(quads, n + 2, 18, height_params(kind=FBM, gain=0.5, fixed_octaves=5))
    check_shade(gpu, port, quads, maps, np.array([1.0e6, -2.5e6, 7.1e6]), n=n)


# ---------------------------------------------------------------------------------------------
# BASELINE.json config 2 at full size: properties that need no 16M-sample oracle run
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2(gpu):
    """root face 0, uniform depth 7: 16 384 quads x 32^2 = 16 777 216 samples, fBm 8 octaves."""
    quads = gpu.tessellate_uniform(7, first=0, nquads=16384)
    p = gpu.fbm_params(octaves=8, gain=0.5, precision=gpu.FAST)
    maps = gpu.generate_height_maps(quads, 32, 18, p)
    return quads, p, maps


def test_c2_fast_vs_oracle_on_a_strided_subset(gpu, port, c2):
    quads, p, maps = c2
    sel = np.arange(0, 16384, 61)
    hq = gpu.quads_to_host(quads)[sel]
    want = port.generate_height_maps(hq, 32, 18, orc_params(p), nthreads=4)
    got = to_np(maps)[sel]
    assert np.abs(got.astype(np.float64) - want).max() <= REL_TOL * 8848.0 * amp_sum(0.5, 8)


def test_c2_fast_vs_exact_everywhere(gpu, c2):
    import torch
    quads, p, maps = c2
    ex = gpu.generate_height_maps(quads, 32, 18, gpu.fbm_params(octaves=8, gain=0.5, precision=gpu.EXACT))
    err = (maps.double() - ex.double()).abs().max().item()
    assert err <= REL_TOL * 8848.0 * amp_sum(0.5, 8), err
    assert torch.isfinite(maps).all()


def test_c2_sharding_is_byte_invisible(gpu, c2):
    """Any partition of the quad range gives the bytes of the unpartitioned run."""
    import torch
    quads, p, maps = c2
    parts = []
    for g in range(8):                                              # 8 shards as on 8 GPUs
        lo, hi = g * 2048, (g + 1) * 2048
        q = gpu.tessellate_uniform(7, first=lo, nquads=2048)
        parts.append(gpu.generate_height_maps(q, 32, 18, p))
    assert torch.equal(torch.cat(parts), maps)
    ragged = torch.cat([gpu.generate_height_maps(quads[a:b], 32, 18, p) for a, b in ((0, 1), (1, 1000), (1000, 16384))])
    assert torch.equal(ragged, maps)


def test_c2_sibling_seams_are_continuous(gpu, c2):
    """Child 0 and child 1 of one parent share an edge: last interior column == first interior column."""
    quads, p, maps = c2
    m = to_np(maps).reshape(-1, 4, 32, 32)                          # consecutive leaves = the 4 children
    right_of_0, left_of_1 = m[:, 0, 1:31, 30], m[:, 1, 1:31, 1]
    tol = 2 * REL_TOL * 8848.0 * amp_sum(0.5, 8)
    assert np.abs(right_of_0.astype(np.float64) - left_of_1).max() <= tol
    bottom_of_0, top_of_2 = m[:, 0, 30, 1:31], m[:, 2, 1, 1:31]
    assert np.abs(bottom_of_0.astype(np.float64) - top_of_2).max() <= tol


def test_c2_shade_normals_are_unit_and_positions_displace_radially(gpu, c2):
    import torch
    quads, p, maps = c2
    cam = (0.0, 0.0, -6371000.0 - 10.0)
    pos, nrm = gpu.shade(quads[:4096], maps[:4096], cam)
    n3 = nrm[..., :3]
    assert (n3.norm(dim=-1) - 1).abs().max().item() <= 1e-5
    assert (nrm[..., 3] >= 0.0316).all() and (nrm[..., 3] <= 1.0005).all()      # sqrt(0.001) .. sqrt(1.001)
    # |pos + cam| - R == height (to fp32 resolution at 6.4e6 m) for non-skirt vertices
    interior = torch.ones(1020, dtype=torch.bool, device="cuda")
    interior[:30] = False; interior[-30:] = False
    interior[30::32][:30] = False; interior[61::32][:30] = False
    w = pos[:, interior, :3].double() + torch.tensor(cam, dtype=torch.float64, device="cuda")
    r = w.norm(dim=-1) - 6371000.0
    # the shader lifts the flat quad to the arc (linear branch below depth 6 keeps the chord):
    # at depth 7 the sagitta of a 78 km quad is ~120 m, so compare with a loose bound and the exact w channel
    assert (r - pos[:, interior, 3].double()).abs().max().item() < 250.0


# ---------------------------------------------------------------------------------------------
# out-of-bounds canaries (compute-sanitizer is not available on this pool): every output is a
# window inside a sentinel-filled buffer, at ragged sizes, and the sentinel must survive
# ---------------------------------------------------------------------------------------------
def _window(torch, numel, dtype, pad=4099, offset_elems=0):
    big = torch.full((numel + 2 * pad + offset_elems,), -12345.0 if dtype.is_floating_point else -12345,
                     dtype=dtype, device="cuda")
    return big, big[pad + offset_elems: pad + offset_elems + numel]


def _intact(big, win_start, numel):
    sentinel = big[0].item()
    return bool((big[:win_start] == sentinel).all()) and bool((big[win_start + numel:] == sentinel).all())


@pytest.mark.parametrize("dim,nq,octaves,offset", [(4, 37, 3, 0), (5, 129, 2, 1), (33, 7, 8, 0), (32, 1025, 1, 3), (52, 300, 12, 0), (64, 257, 4, 1)])
@pytest.mark.parametrize("small_path", [True, False])
def test_k2_writes_exactly_its_output(gpu, port, dim, nq, octaves, offset, small_path, monkeypatch):
    import torch
    quads = gpu.tessellate_uniform(5, first=11, nquads=nq)
    numel = nq * dim * dim
    big, win = _window(torch, numel, torch.float32, offset_elems=offset)       # offset 1/3: not 8-byte aligned
    p = gpu.fbm_params(octaves, 0.5, gpu.FAST)
    # small_path False exercises the replicated-table kernel even for small batches
    if not small_path:
        import subprocess, sys, json, os
        code = ("import sys; sys.path.insert(0, %r); import torch, planet_b200 as pb; pb.init(0);"
                "q = pb.tessellate_uniform(5, first=11, nquads=%d);"
                "big = torch.full((%d,), -12345.0, device='cuda'); win = big[%d:%d];"
                "pb.generate_height_maps(q, %d, 18, pb.fbm_params(%d, 0.5, pb.FAST), out=win.view(%d, %d, %d));"
                "torch.cuda.synchronize();"
                "ok = bool((big[:%d] == -12345.0).all()) and bool((big[%d:] == -12345.0).all()) and bool((win != -12345.0).all());"
                "print('OK' if ok else 'BAD')") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), nq,
                                                   numel + 8198 + offset, 4099 + offset, 4099 + offset + numel,
                                                   dim, octaves, nq, dim, dim, 4099 + offset, 4099 + offset + numel)
        env = dict(os.environ, PLANET_K2_SMALL_MAX="0")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert out.stdout.strip().endswith("OK"), out.stdout + out.stderr
        return
    gpu.generate_height_maps(quads, dim, 18, p, out=win.view(nq, dim, dim))
    torch.cuda.synchronize()
    assert _intact(big, 4099 + offset, numel)
    assert bool((win != -12345.0).all())                                        # and every texel was written
    sel = [0, nq // 2, nq - 1]
    want = port.generate_height_maps(gpu.quads_to_host(quads)[sel], dim, 18, orc_params(p))
    got = to_np(win.view(nq, dim, dim))[sel]
    assert np.abs(got.astype(np.float64) - want).max() <= REL_TOL * 8848.0 * amp_sum(0.5, octaves)


@pytest.mark.parametrize("kind", ["fbm", "ridged"])
@pytest.mark.parametrize("nq", [700, 2100])                                     # compact- and replicated-table kernels
def test_k2_tile_paths_agree_bit_for_bit(gpu, kind, nq):
    """A height is a function of (quad, x, y) only.  With an 8-byte aligned output every tile of a
    32 x 32 map takes the run-of-regular-tiles path of K2; with the output one float off it, every
    tile takes the ragged path (scalar stores, per-sample guards).  Same bits either way."""
    import torch
    p = gpu.fbm_params(8, 0.5, gpu.FAST) if kind == "fbm" else gpu.default_params(precision=gpu.FAST)
    quads = gpu.tessellate_uniform(6, first=5, nquads=nq, params=p)
    aligned = gpu.generate_height_maps(quads, 32, 18, p)
    big, win = _window(torch, nq * 1024, torch.float32, pad=4100, offset_elems=1)
    assert win.data_ptr() % 8 == 4
    gpu.generate_height_maps(quads, 32, 18, p, out=win.view(nq, 32, 32))
    torch.cuda.synchronize()
    assert torch.equal(win.view(nq, 32, 32), aligned)
    assert _intact(big, 4101, nq * 1024)


@pytest.mark.parametrize("n,nq", [(30, 5), (5, 3), (31, 17), (50, 9), (12, 1)])
def test_k1_k3_write_exactly_their_outputs(gpu, n, nq):
    import torch
    p = gpu.default_params(patch_verts=n)
    nv, ni, dim = gpu.patch_vertex_count(n), gpu.patch_index_count(n), n + 2
    bq, wq = _window(torch, nq * 13, torch.int64)
    bi, wi = _window(torch, nq * ni, torch.int32, pad=4100)                     # keeps the window 8-byte aligned
    L, C = gpu.lib(), gpu.C
    gpu._check(L.planet_gpu_tessellate_uniform(C.byref(p), 3, 20, nq, wq.data_ptr(), wi.data_ptr(), None))
    torch.cuda.synchronize()
    assert _intact(bq, 4099, nq * 13) and _intact(bi, 4100, nq * ni)
    strip = np.array([L.planet_gpu_strip_index(k, n) for k in range(ni)], np.uint32)
    want = strip[None, :] + (np.arange(nq, dtype=np.uint32) * nv)[:, None]
    assert (to_np(wi).view(np.uint32).reshape(nq, ni) == want).all()
    # the index stream alone, by the kernel that runs beside K2: same bytes, nothing outside its window
    bs, ws = _window(torch, nq * ni, torch.int32, pad=4100)
    gpu._check(L.planet_gpu_merged_indices_beside(C.byref(p), nq, ws.data_ptr(), None))
    torch.cuda.synchronize()
    assert _intact(bs, 4100, nq * ni) and torch.equal(ws, wi)
    quads = wq.view(nq, 13)
    heights = gpu.generate_height_maps(quads, dim, 18, gpu.fbm_params(4, 0.5, gpu.FAST, patch_verts=n))
    bp, wp = _window(torch, nq * nv * 4, torch.float32, pad=4100)
    bn, wn = _window(torch, nq * nv * 4, torch.float32, pad=4100)
    gpu.shade(quads, heights, (1.0e6, 2.0e6, -7.0e6), p, pos=wp.view(nq, nv, 4), nrm=wn.view(nq, nv, 4))
    torch.cuda.synchronize()
    assert _intact(bp, 4100, nq * nv * 4) and _intact(bn, 4100, nq * nv * 4)
    assert bool((wp != -12345.0).all()) and bool((wn != -12345.0).all())


def test_c4_shape_against_the_oracle(gpu, port):
    """BASELINE config 4's shape: patch 50 (dim 52), depth 8, fBm 12 octaves -- a strided subset."""
    p = gpu.fbm_params(12, 0.5, gpu.FAST, patch_verts=50)
    quads = gpu.tessellate_uniform(8, first=5 * 65536 + 1000, nquads=4096, params=p)
    maps = gpu.generate_height_maps(quads, 52, 18, p)
    sel = np.arange(0, 4096, 409)
    hq = gpu.quads_to_host(quads)[sel]
    assert hq.tobytes() == port.uniform_quads(5, 8)[1000:1000 + 4096][sel].tobytes()
    want = port.generate_height_maps(hq, 52, 18, orc_params(p), nthreads=4)
    assert np.abs(to_np(maps)[sel].astype(np.float64) - want).max() <= REL_TOL * 8848.0 * amp_sum(0.5, 12)
    check_shade(gpu, port, hq[:3], to_np(maps)[sel][:3], np.array([0.0, 0.0, -6371010.0]), n=50)


def test_k2_fused_gather_stores_identical_bytes_to_every_peer(gpu):
    """planet_gpu_generate_height_maps_gathered: the peer buffers (here: 7 more buffers on this GPU,
    on an 8-GPU box: CUDA-IPC-mapped buffers of the other ranks) get exactly the local bytes."""
    import torch
    for prec, nq, dim in ((gpu.FAST, 3000, 32), (gpu.FAST, 40, 33), (gpu.EXACT, 64, 32), (gpu.EXACT, 1100, 32)):
        p = gpu.fbm_params(8, 0.5, prec)
        quads = gpu.tessellate_uniform(6, first=100, nquads=nq)
        want = gpu.generate_height_maps(quads, dim, 18, p)
        for n_peers in (1, 7):
            out = torch.zeros_like(want)
            peers = [torch.full_like(want, -1.0) for _ in range(n_peers)]
            gpu.generate_height_maps_gathered(quads, dim, 18, out, peers, p)
            torch.cuda.synchronize()
            assert torch.equal(out, want)
            for t_ in peers:
                assert torch.equal(t_, want)
    with pytest.raises(gpu.PlanetGpuError, match="n_peers"):
        gpu.generate_height_maps_gathered(quads, 32, 18, want, [want] * 8, p)


def test_exact_large_batches_use_the_table_kernel_and_keep_every_bit(gpu, port):
    """Above PLANET_K2_SMALL_MAX samples EXACT mode runs k_height_maps_exact_tab (replicated tables,
    FMA dots on exact products, doubled gradient codes).  It must stay bit-identical to the
    reference: checked against the oracle and, for what the oracle cannot express (seed offsets),
    against the small-batch kernel, which is the line-by-line exact:: code."""
    import torch
    cases = [
        dict(kind=gpu.FBM, gain=0.5, fixed_octaves=8, dim=32, depth=6, first=777, nq=1500),
        dict(kind=gpu.RIDGED, gain=0.55, fixed_octaves=0, dim=33, depth=6, first=3, nq=1100),      # default functor, odd dim
        dict(kind=gpu.FBM, gain=0.7, fixed_octaves=5, dim=110, depth=2, first=0, nq=96, lacunarity=2.5, precision=gpu.FAST),
    ]
    for c in cases:
        p = gpu.default_params(noise_kind=c["kind"], gain=c["gain"], fixed_octaves=c["fixed_octaves"],
                               lacunarity=c.get("lacunarity", 2.0), precision=c.get("precision", gpu.EXACT))
        quads = gpu.tessellate_uniform(c["depth"], first=c["first"], nquads=c["nq"])
        assert c["nq"] * c["dim"] ** 2 > 1 << 20
        got = to_np(gpu.generate_height_maps(quads, c["dim"], 18, p))
        want = port.generate_height_maps(gpu.quads_to_host(quads), c["dim"], 18, orc_params(p), nthreads=8)
        assert (as_bits(got) == as_bits(want)).all(), c
    # seed offset + a scale that puts samples on both sides of the coordinate planes
    p = gpu.default_params(noise_kind=gpu.RIDGED, gain=0.55, fixed_octaves=12, precision=gpu.EXACT,
                           coord_scale=3e-6, seed_offset=(-7.25, 0.5, 1e3))
    quads = gpu.tessellate_uniform(5, first=0, nquads=2048)
    whole = gpu.generate_height_maps(quads, 32, 18, p)
    parts = torch.cat([gpu.generate_height_maps(quads[i:i + 512], 32, 18, p) for i in range(0, 2048, 512)])
    assert torch.equal(whole, parts)


def test_randomised_configurations_against_the_oracle(gpu, port):
    """40 random (dim, octaves, gain, kind, seed offset, depth, batch) draws: EXACT bit-exact,
    FAST within the stated tolerance -- through whichever table layout the batch size selects."""
    rng = np.random.default_rng(20261018)
    for trial in range(40):
        dim = int(rng.choice([4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 64, 70]))
        octaves = int(rng.integers(1, 21))
        gain = float(rng.choice([0.3, 0.5, 0.55, 0.7, 0.9]))
        kind = int(rng.choice([gpu.FBM, gpu.RIDGED]))
        depth = int(rng.integers(0, 9))
        nq = int(rng.integers(1, 40))
        first = int(rng.integers(0, 6 * 4 ** depth - min(nq, 6 * 4 ** depth) + 1))
        nq = min(nq, 6 * 4 ** depth - first)
        seed = tuple(rng.uniform(-500, 500, 3)) if trial % 3 == 0 else (0.0, 0.0, 0.0)
        scale = float(rng.choice([1e-5, 1e-5, 3e-6, 4e-5]))
        quads = gpu.tessellate_uniform(depth, first=first, nquads=nq)
        hq = gpu.quads_to_host(quads)
        for prec in (gpu.EXACT, gpu.FAST):
            p = gpu.default_params(noise_kind=kind, gain=gain, fixed_octaves=octaves, precision=prec,
                                   coord_scale=scale, seed_offset=seed)
            got = to_np(gpu.generate_height_maps(quads, dim, 18, p))
            if seed == (0.0, 0.0, 0.0):
                want = port.generate_height_maps(hq, dim, 18, orc_params(p))
            else:
                # the oracle has no seed: evaluate the fractal on the kernel's own (scaled + shifted) points
                want = np.empty_like(got)
                div = 1.0 / (dim - 3)
                u = (np.arange(dim) - 1) * div
                for k, q in enumerate(hq):
                    p0 = q["p"][0][None, :] + (q["p"][1] - q["p"][0])[None, :] * u[:, None]
                    p1 = q["p"][2][None, :] + (q["p"][3] - q["p"][2])[None, :] * u[:, None]
                    pts = p0[None, :, :] + (p1 - p0)[None, :, :] * u[:, None, None]        # [y, x, 3]
                    pts = pts * scale + np.array(seed)
                    want[k] = (port.fractal(pts.reshape(-1, 3), kind, 2.0, gain, octaves) * np.float32(8848.0)).reshape(dim, dim)
            tol = REL_TOL * 8848.0 * amp_sum(gain, octaves)
            if prec == gpu.EXACT:
                assert got.tobytes() == want.tobytes(), (trial, dim, octaves, gain, kind, depth, seed)
            else:
                err = np.abs(got.astype(np.float64) - want).max()
                assert err <= tol, (trial, dim, octaves, gain, kind, depth, seed, err, tol)
