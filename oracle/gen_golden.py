"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/planet_golden.npz.

Run in the build container (needs /root/reference): every array in the fixture is
produced by the reference's OWN code through oracle/_ref/libplanet_ref.so
(oracle/ref_oracle.cpp #includes /root/reference/main.cpp unmodified), including one
headless frame of the reference's real main().  While writing, the plain-C
restatement (oracle/planet_oracle.c) is asserted bit-identical on every case, so a
stale fixture and a wrong restatement are both caught here.

    python oracle/gen_golden.py

The GPU box has no /root/reference; there the tests read only the committed .npz.
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bindings import (FBM, RIDGED, RADIUS, PortOracle, RefOracle,  # noqa: E402
                             fnv1a32, height_params)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "planet_golden.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b, what):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape and a.tobytes() == b.tobytes(), "port != reference: " + what


def main():
    ref, port = RefOracle(), PortOracle()
    g = {}
    rng = np.random.default_rng(20261018)

    # --- perlin.h tables and hash chain -------------------------------------------------
    t, v = ref.tables()
    same(t, port.tables()[0], "table"); same(v, port.tables()[1], "vectors")
    g["table"], g["vectors"] = t, v
    seeds = np.concatenate([np.arange(-520, 520), rng.integers(-2**31, 2**31 - 1, 256)]).astype(np.int64)
    g["random_seeds"] = seeds
    g["random_values"] = np.array([ref.random(s) for s in seeds], np.int32)
    same(g["random_values"], np.array([port.random(s) for s in seeds], np.int32), "PerlinRandom")
    cells = rng.integers(-70000, 70000, (512, 3)).astype(np.int32)
    cells[:4] = [[1, 2, 3], [-1, -2, -3], [0, 0, 0], [255, 256, -256]]
    frac = rng.random((512, 3)).astype(np.float32)
    g["gradient_cells"], g["gradient_frac"] = cells, frac
    gr = np.array([ref.L.ref_perlin_gradient(*map(float, f), *map(int, c)) for f, c in zip(frac, cells)], np.float32)
    gp = np.array([port.L.orc_perlin_gradient(*map(float, f), *map(int, c)) for f, c in zip(frac, cells)], np.float32)
    same(gr, gp, "PerlinGradient")
    g["gradient_bits"] = bits(gr)

    # --- PerlinNoise3 -------------------------------------------------------------------
    special = np.array([
        [0.5, 0.5, 0.5], [0.1, 0.2, 0.3], [-1.25, 3.75, -0.5], [63.71, 0, 0],
        [-36.78, 36.78, -36.78], [255.9, 256.1, -0.0001], [-2.0, 0.5, 0.5],
        [0, 0, 0], [-0.0, -0.0, -0.0], [1, 2, 3], [-1, -2, -3], [-1e-300, 1e-300, -1e-17],
        [8350000.25, -8350000.75, 4194304.5], [-255.0, -256.0, -257.0],
        [0.9999999999999999, -0.9999999999999999, 1.0000000000000002],
    ], np.float64)
    pts = np.concatenate([special,
                          rng.uniform(-300, 300, (3000, 3)),
                          rng.uniform(-8.4e6, 8.4e6, (500, 3)),          # 63.7 * 2^17
                          np.round(rng.uniform(-50, 50, (200, 3))),      # lattice points
                          rng.uniform(-1, 1, (300, 3))])
    g["noise_points"] = pts
    nr = ref.noise3(pts); same(nr, port.noise3(pts), "PerlinNoise3")
    g["noise_bits"] = bits(nr)

    # --- PerlinfBm / PerlinRidged -------------------------------------------------------
    d = rng.normal(size=(768, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    sph = d * 63.71
    g["fractal_points"] = sph
    for kind, name, gain, octs in ((FBM, "fbm", 0.5, (1, 2, 8, 12, 16)), (RIDGED, "ridged", 0.55, (1, 6, 7, 12, 18))):
        for o in octs:
            fr = ref.fractal(sph, kind, 2.0, gain, o)
            same(fr, port.fractal(sph, kind, 2.0, gain, o), f"{name} {o}")
            g[f"{name}_{o}_bits"] = bits(fr)
    # general lacunarity (not a power of two): exercises the double path
    fr = ref.fractal(sph, FBM, 2.17, 0.47, 6); same(fr, port.fractal(sph, FBM, 2.17, 0.47, 6), "fbm lac 2.17")
    g["fbm_lac217_gain047_6_bits"] = bits(fr)

    # --- InitPlanet: patch mesh, max_lod, skirt ----------------------------------------
    max_lod, max_skirt = ref.init_planet()
    assert max_lod == port.max_lod() and np.float32(max_skirt) == np.float32(port.max_skirt_size())
    g["max_lod"], g["max_skirt_size"] = np.int32(max_lod), np.float32(max_skirt)
    vb, ib = ref.patch_buffer(0), ref.patch_buffer(1)
    same(vb.view(np.float32).reshape(-1, 3), port.patch_vertices(30), "patch vertices")
    same(ib.view(np.uint32), port.patch_indices(30), "patch indices")
    g["patch_vertex_buffer"], g["patch_index_buffer"] = vb, ib
    assert fnv1a32(vb) == 0xA9622A0D and fnv1a32(ib) == 0xFA54AAA4      # SURVEY.md 8c probes

    # --- QuadID + subdivision geometry --------------------------------------------------
    roots = ref.root_quads(); same(roots, port.root_quads(), "root quads")
    g["root_quads"] = roots.view(np.uint8).reshape(6, 104)
    d2 = np.concatenate([ref.uniform_quads(f, 2) for f in range(6)])
    same(d2, np.concatenate([port.uniform_quads(f, 2) for f in range(6)]), "depth-2 quads")
    g["depth2_quads"] = d2.view(np.uint8).reshape(-1, 104)
    d5 = ref.uniform_quads(0, 5); same(d5, port.uniform_quads(0, 5), "depth-5 face-0 quads")
    assert fnv1a32(d5) == 0x83EED4C1
    g["depth5_face0_fnv"] = np.uint32(fnv1a32(d5))
    g["depth5_face0_first_last"] = d5[[0, 1, 2, 3, 341, 682, 1022, 1023]].view(np.uint8).reshape(-1, 104)
    d7 = ref.uniform_quads(3, 7); same(d7, port.uniform_quads(3, 7), "depth-7 face-3 quads")
    g["depth7_face3_fnv"] = np.uint32(fnv1a32(d7))
    ids = [ref.make_root_id(r) for r in range(6)]
    for _ in range(10):
        ids.append(ref.make_child_id(ids[int(rng.integers(len(ids)))], int(rng.integers(4))))
    cur = ref.make_root_id(4)
    for lvl in range(27):                                 # deepest legal path (depth+1 < 32, 55 path bits)
        cur = ref.make_child_id(cur, (lvl * 7 + 3) % 4); ids.append(cur)
    ids = np.array(ids, np.uint64)
    g["quad_ids"] = ids
    for fn in ("get_root", "get_depth", "get_index"):
        r = np.array([getattr(ref, fn)(int(i)) for i in ids], np.uint64)
        same(r, np.array([getattr(port, fn)(int(i)) for i in ids], np.uint64), fn)
        g["quad_" + fn] = r
    deep = ids[6:]
    g["quad_parent"] = np.array([ref.get_parent_id(int(i)) for i in deep], np.uint64)
    g["quad_child_index"] = np.array([ref.get_child_index(int(i)) for i in deep], np.uint64)
    same(g["quad_parent"], np.array([port.get_parent_id(int(i)) for i in deep], np.uint64), "parent")
    same(g["quad_child_index"], np.array([port.get_child_index(int(i)) for i in deep], np.uint64), "child idx")
    g["quad_children"] = np.array([[ref.make_child_id(int(i), c) for c in range(4)] for i in ids[:-1]], np.uint64)
    same(g["quad_children"], np.array([[port.make_child_id(int(i), c) for c in range(4)] for i in ids[:-1]], np.uint64), "children")

    # --- the reference's real main(): one headless frame with its LOCAL Perlin functor ---
    rc, main_maps, main_draws = ref.run_reference_main(tempfile.mkdtemp())
    assert rc == 0 and len(main_maps) == 117 and len(main_draws) == 117
    cam = np.array([0.0, 0.0, -RADIUS - 10.0])            # main.cpp:864
    quads, maps, draws = ref.render_frame(cam, height_params())
    assert len(quads) == 117
    # the restated functor (ref_oracle.cpp OracleHeight) == the unreachable local one
    same(np.stack(main_maps), np.stack(maps), "restated Perlin functor vs main()'s local functor")
    same(main_draws, draws, "per-draw uniforms")
    pm = port.generate_height_maps(quads, 32, max_lod, height_params())
    same(np.stack(maps), pm, "default-frame height maps")
    for k, q in enumerate(quads):                         # ProcessQuad's leaves == descent by id
        same(q, port.quad_from_id(q["id"]), f"quad_from_id {k}")
    g["frame_cam"] = cam
    g["frame_quads"] = quads.view(np.uint8).reshape(-1, 104)
    g["frame_height_maps"] = np.stack(main_maps)
    g["frame_draws"] = main_draws
    big = ref.generate_height_maps(quads[:1], 1024, max_lod, height_params())
    assert fnv1a32(big) == 0x57FC9FBA and big[0, 0, 0] == np.float32(10731.9414)   # SURVEY.md 8c probe
    g["ridged_leaf0_dim1024_fnv"] = np.uint32(0x57FC9FBA)

    # --- GenerateHeightMap / GetHeightAt, fBm configs and ragged dims -------------------
    fbm8 = height_params(kind=FBM, gain=0.5, fixed_octaves=8)
    all5 = ref.generate_height_maps(d5, 32, max_lod, fbm8, nthreads=8)
    assert fnv1a32(all5) == 0x62660FBE                     # SURVEY.md 8c probe
    same(all5, port.generate_height_maps(d5, 32, max_lod, fbm8, nthreads=8), "fbm8 depth-5 maps")
    pick = [0, 1, 2, 3, 341, 682, 1022, 1023]
    g["fbm8_depth5_fnv"] = np.uint32(0x62660FBE)
    g["fbm8_depth5_pick"] = np.array(pick, np.int32)
    g["fbm8_depth5_maps"] = all5[pick]
    d7s = d7[[0, 5461, 10922, 16383]]
    g["fbm8_depth7_quads"] = d7s.view(np.uint8).reshape(-1, 104)
    m7 = ref.generate_height_maps(d7s, 32, max_lod, fbm8); same(m7, port.generate_height_maps(d7s, 32, max_lod, fbm8), "d7")
    g["fbm8_depth7_maps"] = m7
    fbm12 = height_params(kind=FBM, gain=0.5, fixed_octaves=12)
    for dim in (4, 5, 7, 33, 52, 64):
        m = ref.generate_height_maps(d7s[:2], dim, max_lod, fbm12)
        same(m, port.generate_height_maps(d7s[:2], dim, max_lod, fbm12), f"dim {dim}")
        g[f"fbm12_dim{dim}_maps"] = m
    m = ref.generate_height_maps(roots[[1, 5]], 128, max_lod, height_params())
    same(m, port.generate_height_maps(roots[[1, 5]], 128, max_lod, height_params()), "roots 128")
    g["ridged_roots15_dim128_maps"] = m
    hp = rng.normal(size=(256, 3)); hp = hp / np.linalg.norm(hp, axis=1, keepdims=True) * RADIUS
    dd = rng.integers(0, 19, 256)
    g["height_at_points"], g["height_at_depth"] = hp, dd.astype(np.int32)
    hr = np.array([ref.get_height_at(p, int(k), 18) for p, k in zip(hp, dd)], np.float32)
    same(hr, np.array([port.get_height_at(p, int(k), 18) for p, k in zip(hp, dd)], np.float32), "GetHeightAt")
    g["height_at_bits"] = bits(hr)
    lod = np.array([ref.get_height_at(p, 0, 1) for p in hp], np.float32)      # ProcessQuad's call (main.cpp:552)
    g["height_at_lod_bits"] = bits(lod)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
