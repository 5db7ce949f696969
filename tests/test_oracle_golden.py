"""The CPU oracle pinned against the reference's own outputs (tests/golden, written by
oracle/gen_golden.py from the reference compiled here).  Runs without a GPU."""
import numpy as np
import pytest

from conftest import as_bits, quads_from_bytes
from oracle.bindings import FBM, RIDGED, fnv1a32, height_params


@pytest.fixture(scope="module", params=["port", "ref"])
def orc(request):
    return request.getfixturevalue(request.param)


def test_tables(orc, golden):
    t, v = orc.tables()
    assert (t == golden["table"]).all() and (v == golden["vectors"]).all()
    assert sorted(t.tolist()) == list(range(256))                    # a permutation (perlin.h:10-28)


def test_perlin_random_masks_negative_seeds(orc, golden):
    got = np.array([orc.random(int(s)) for s in golden["random_seeds"]], np.int32)
    assert (got == golden["random_values"]).all()
    # SURVEY 8c known answers: R(R(R(1)+2)+3) = 23, R(R(R(-1)-2)-3) = 156
    assert orc.random(orc.random(orc.random(1) + 2) + 3) == 23
    assert orc.random(orc.random(orc.random(-1) - 2) - 3) == 156


def test_noise3_bits(orc, golden):
    assert (as_bits(orc.noise3(golden["noise_points"])) == golden["noise_bits"]).all()


def test_noise3_known_answers(orc):
    pts = np.array([[0.5, 0.5, 0.5], [0.1, 0.2, 0.3], [-1.25, 3.75, -0.5], [63.71, 0, 0],
                    [-36.78, 36.78, -36.78], [255.9, 256.1, -0.0001], [-2.0, 0.5, 0.5]])
    want = np.array([0xBEC00000, 0xBE808C14, 0xBE5B12E0, 0x3DDA4398, 0x3D86715A, 0xBE484E40, 0x3E000000],
                    np.uint32)
    assert (as_bits(orc.noise3(pts)) == want).all()


@pytest.mark.parametrize("name,kind,gain,octs", [("fbm", FBM, 0.5, (1, 2, 8, 12, 16)),
                                                 ("ridged", RIDGED, 0.55, (1, 6, 7, 12, 18))])
def test_fractal_bits(orc, golden, name, kind, gain, octs):
    for o in octs:
        got = orc.fractal(golden["fractal_points"], kind, 2.0, gain, o)
        assert (as_bits(got) == golden[f"{name}_{o}_bits"]).all(), (name, o)


def test_fractal_general_lacunarity(orc, golden):
    got = orc.fractal(golden["fractal_points"], FBM, 2.17, 0.47, 6)
    assert (as_bits(got) == golden["fbm_lac217_gain047_6_bits"]).all()


def test_quad_id_helpers(orc, golden):
    ids = golden["quad_ids"]
    for fn in ("get_root", "get_depth", "get_index"):
        assert [getattr(orc, fn)(int(i)) for i in ids] == golden["quad_" + fn].tolist()
    deep = ids[6:]
    assert [orc.get_parent_id(int(i)) for i in deep] == golden["quad_parent"].tolist()
    assert [orc.get_child_index(int(i)) for i in deep] == golden["quad_child_index"].tolist()
    kids = [[orc.make_child_id(int(i), c) for c in range(4)] for i in ids[:-1]]
    assert kids == golden["quad_children"].tolist()
    assert orc.make_root_id(0) == 0x8000000000000000


def test_subdivision_geometry(orc, golden):
    assert orc.root_quads().tobytes() == golden["root_quads"].tobytes()
    d2 = np.concatenate([orc.uniform_quads(f, 2) for f in range(6)])
    assert d2.tobytes() == golden["depth2_quads"].tobytes()
    d5 = orc.uniform_quads(0, 5)
    assert fnv1a32(d5) == int(golden["depth5_face0_fnv"]) == 0x83EED4C1
    assert d5[[0, 1, 2, 3, 341, 682, 1022, 1023]].tobytes() == golden["depth5_face0_first_last"].tobytes()
    assert int(d5["id"][0]) == 0x8280000000000000 and int(d5["id"][-1]) == 0x82800000000003FF


def test_patch_mesh(orc, golden):
    assert orc.patch_vertices(30).tobytes() == golden["patch_vertex_buffer"].tobytes()
    assert orc.patch_indices(30).tobytes() == golden["patch_index_buffer"].tobytes()
    ib = orc.patch_indices(30)
    assert len(ib) == 2036 and ib[:8].tolist() == [0, 31, 1, 32, 2, 33, 3, 34] and ib[-1] == 1019


def test_port_patch_mesh_other_sizes(port):
    for n in (2, 3, 5, 31, 50):
        ib, vb = port.patch_indices(n), port.patch_vertices(n)
        assert len(ib) == 2 * n * n + 8 * n - 4 and len(vb) == n * n + 4 * n
        assert ib.max() == len(vb) - 1 and set(ib.tolist()) == set(range(len(vb)))


def test_default_frame_height_maps(orc, golden):
    """117 leaf quads of the reference's default camera; maps captured from its real main()."""
    quads = quads_from_bytes(golden["frame_quads"])
    assert len(quads) == 117
    got = orc.generate_height_maps(quads, 32, int(golden["max_lod"]), height_params())
    assert got.tobytes() == golden["frame_height_maps"].tobytes()


def test_quad_from_id_matches_processquad_leaves(port, golden):
    quads = quads_from_bytes(golden["frame_quads"])
    for q in quads:
        assert port.quad_from_id(q["id"]).tobytes() == q.tobytes()
    depths = sorted(set(port.get_depth(int(i)) for i in quads["id"]))
    assert depths[0] == 0 and depths[-1] == 10


def test_fbm_height_maps(orc, golden):
    fbm8 = height_params(kind=FBM, gain=0.5, fixed_octaves=8)
    ml = int(golden["max_lod"])
    d5 = orc.uniform_quads(0, 5)
    pick = golden["fbm8_depth5_pick"]
    assert orc.generate_height_maps(d5[pick], 32, ml, fbm8).tobytes() == golden["fbm8_depth5_maps"].tobytes()
    q7 = quads_from_bytes(golden["fbm8_depth7_quads"])
    assert orc.generate_height_maps(q7, 32, ml, fbm8).tobytes() == golden["fbm8_depth7_maps"].tobytes()
    fbm12 = height_params(kind=FBM, gain=0.5, fixed_octaves=12)
    for dim in (4, 5, 7, 33, 52, 64):                               # ragged / minimum sizes (dim > 3)
        got = orc.generate_height_maps(q7[:2], dim, ml, fbm12)
        assert got.tobytes() == golden[f"fbm12_dim{dim}_maps"].tobytes(), dim


def test_fbm8_depth5_checksum_of_all_maps(orc, golden):
    fbm8 = height_params(kind=FBM, gain=0.5, fixed_octaves=8)
    maps = orc.generate_height_maps(orc.uniform_quads(0, 5), 32, int(golden["max_lod"]), fbm8, nthreads=4)
    assert fnv1a32(maps) == int(golden["fbm8_depth5_fnv"]) == 0x62660FBE


def test_ridged_roots_dim128(orc, golden):
    roots = quads_from_bytes(golden["root_quads"])[[1, 5]]
    got = orc.generate_height_maps(roots, 128, int(golden["max_lod"]), height_params())
    assert got.tobytes() == golden["ridged_roots15_dim128_maps"].tobytes()


def test_get_height_at(orc, golden):
    pts, depth = golden["height_at_points"], golden["height_at_depth"]
    got = np.array([orc.get_height_at(p, int(d), 18) for p, d in zip(pts, depth)], np.float32)
    assert (as_bits(got) == golden["height_at_bits"]).all()
    lod = np.array([orc.get_height_at(p, 0, 1) for p in pts], np.float32)   # ProcessQuad's call
    assert (as_bits(lod) == golden["height_at_lod_bits"]).all()


def test_init_planet_scalars(port, golden):
    assert port.max_lod() == int(golden["max_lod"]) == 18
    assert np.float32(port.max_skirt_size()) == golden["max_skirt_size"]
    assert abs(float(golden["max_skirt_size"]) - 244267.0) < 1.0


def test_shade_restatement_sanity(port, golden):
    """Unpinned part (GLSL): structural checks of the restatement on the default frame."""
    quads = quads_from_bytes(golden["frame_quads"])[[0, 5, 40, 116]]
    maps = golden["frame_height_maps"][[0, 5, 40, 116]]
    cam = golden["frame_cam"]
    pos, nrm = port.shade_patches(quads, cam, maps)
    assert pos.shape == (4, 1020, 4) and np.isfinite(pos).all() and np.isfinite(nrm).all()
    assert np.allclose(np.linalg.norm(nrm[..., :3], axis=-1), 1.0, atol=1e-5)
    # corner vertex (0,0) is slot 31 (after 30 top-skirt verts and the left skirt vert)
    for k, q in enumerate(quads):
        p0 = q["p"][0] - cam
        up = q["p"][0] / np.linalg.norm(q["p"][0])
        want = p0 + up * maps[k][1, 1]
        assert np.allclose(pos[k, 31, :3], want, rtol=0, atol=2.0), k      # fp32 at |p| ~ 6e6: ulp 0.5 m
        assert pos[k, 31, 3] == maps[k][1, 1]
    # skirt vertices sit skirt_size lower than the edge vertex under them
    sk = port.skirt_size_for_quad(port.max_skirt_size(), quads["id"][2])
    assert np.isclose(pos[2, 30, 3], pos[2, 31, 3] - sk)


def test_glsl_text_the_restatement_follows(ref):
    """The shader cannot be executed here, so the restatement (oracle/planet_oracle.c: glsl_*) is a
    transcription.  This states WHICH text it transcribes: the strings the reference hands to
    glShaderSource (render.cpp:111), captured by the recording GL, hashed, and checked for the lines
    whose arithmetic the restatement and K3 reproduce.  If the reference's shader changes, this fails."""
    from oracle.bindings import fnv1a32
    ref.render_frame(np.array([0.0, 0.0, -6371010.0]))       # InitPlanet compiles the shader
    vs, fs = ref.captured_shaders()
    assert (fnv1a32(vs.encode()), fnv1a32(fs.encode())) == (0x06e4320c, 0xd545d6af)
    assert "#define VERTEX_SHADER 1" in vs and "#define FRAGMENT_SHADER 1" in fs
    for line in ("if (1.0 - dot(v0.n, v1.n) < 0.001)",
                 "vec3 n = normalize(mix(v0.n, v1.n, t));",
                 "float theta2 = acos(dot(v0.n, v1.n));",
                 "vec3 n = normalize(sin(k*theta2)*v0.n + sin(t*theta2)*v1.n);",
                 "float gamma = theta - theta2*t;",
                 "float x = 1.0 - tan(gamma)/tan_theta;",
                 "float y = 1.0/sin(theta) - 1.0/(cos(gamma)*tan_theta);",
                 "vec3 p = v0.p + x*v + y*n*length(v);",
                 "return normalize(vec3(x0 - x1, 2.0*xyscale, y0 - y1));",
                 "vec2 uv = mix(HeightMap_corners[0], HeightMap_corners[1], UV.xy);",
                 "float height = sample_height(uv) - SkirtSize*UV.z;",
                 "vec3 normal = compute_normal(uv, length(q.p - p.p) / 29.0);",
                 "vec3 t = normalize(cross(n, q.p - p.p));",
                 "vec3 bi = normalize(cross(t, n));",
                 "Normal = normalize(mat3(t, n, bi) * normal);",
                 "gl_Position = Projection * View * vec4(v.p + v.n*height, 1.0);",
                 "vec3 l = normalize(vec3(0.0, 1.0, -1.0));",
                 "float light = 0.001 + max(0.0, dot(n, l));",
                 "FragColor = vec4(vec3(sqrt(light)), 1.0);"):
        assert line in vs, line
