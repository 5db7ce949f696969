"""The GLSL stage (main.cpp:286-380) has no CPU implementation in the reference and no GL runs on
either box, so K3 and the C restatement (oracle/planet_oracle.c) could share a transcription
error.  tests/glsl_eval.py is a third, independent evaluation -- float64 numpy written from the
shader text, fed with the attribute stream, uniforms and textures the reference itself handed to
GL (golden patch_vertex_buffer / frame_draws / frame_height_maps).  Here:
  * the restatement and K3 against that evaluator on the reference's whole first frame
    (117 draws, depths 0..10, both branches of interpolate);
  * invariants the shader's own geometry implies, checked on the evaluator AND on K3:
    interpolate(v0, v1, 0/1) returns its endpoints; a flat height map gives Normal == v.n;
    the slerp branch keeps v.p on the sphere through the corners (|v.p + cam| == R) with v.n
    radial; the linear branch stays within the chord's sagitta R*theta^2/8 below it."""
import numpy as np
import pytest

from conftest import quads_from_bytes
from glsl_eval import draw_uniforms, interpolate, shade_draw

R = 6371000.0


def angle(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=-1), (a * b).sum(-1))


def frame(golden):
    uv3 = golden["patch_vertex_buffer"].view(np.float32).reshape(-1, 3)
    return uv3, quads_from_bytes(golden["frame_quads"]), golden["frame_height_maps"], golden["frame_cam"], golden["frame_draws"]


def check_against_evaluator(golden, pos, nrm):
    """pos, nrm: (117, 1020, 4) float32 from the restatement or from K3."""
    uv3, quads, maps, cam, draws = frame(golden)
    worst = 0.0
    for k in range(len(quads)):
        r = shade_draw(uv3, draws[k], maps[k])
        assert r["margin"] > 1e-5, "a branch test of interpolate within fp32 rounding of its threshold"
        worst = max(worst, angle(r["normal"], nrm[k, :, :3]).max())
        assert np.abs(r["light"] - nrm[k, :, 3]).max() <= 1e-4
        assert np.abs(r["height"] - pos[k, :, 3]).max() <= 0.05                    # fp32 texel weights vs float64
        rel = np.abs(quads[k]["p"] - cam).max()
        ext = np.linalg.norm(quads[k]["p"][3] - quads[k]["p"][0])
        tol = 8 * 2.0 ** -23 * rel + 1e-5 * ext                                    # few fp32 ulps of the camera-relative corner
        assert np.abs(r["pos"] - pos[k, :, :3]).max() <= tol, k
    assert worst <= 1e-4, worst                                                    # north_star: <= 1e-4 angular
    return worst


def test_restatement_agrees_with_the_independent_float64_evaluation(port, golden):
    uv3, quads, maps, cam, draws = frame(golden)
    pos, nrm = port.shade_patches(quads, cam, maps)
    assert check_against_evaluator(golden, pos, nrm) <= 5e-6                       # measured 9.1e-7 rad


def test_draw_uniforms_helper_reproduces_the_captured_draws(port, golden):
    uv3, quads, maps, cam, draws = frame(golden)
    for k in (0, 7, 50, 116):
        u = draw_uniforms(quads[k], cam, float(golden["max_skirt_size"]))
        assert u[:31].tobytes() == draws[k][:31].tobytes()


def test_interpolate_returns_its_endpoints(golden):
    uv3, quads, maps, cam, draws = frame(golden)
    for k in (0, 20, 116):                                                         # slerp, linear, slerp (root)
        P = draws[k][:12].reshape(4, 3).astype(np.float64); N = draws[k][12:24].reshape(4, 3).astype(np.float64)
        for t, want in ((0.0, 0), (1.0, 1)):
            p, n, _ = interpolate(P[0:1], N[0:1], P[1:2], N[1:2], np.array([[t]]))
            assert np.abs(p - P[want]).max() <= 1e-9 * np.abs(P).max() and angle(n, N[want:want + 1]).max() <= 1e-7


def sphere_invariants(r, draw, cam, flat_height):
    """What the shader's geometry implies for one draw with a constant height map."""
    assert angle(r["normal"], r["vn"]).max() <= 2e-6                               # compute_normal == (0,1,0): TBN * e_y == n
    world = r["vp"] + cam
    radial = world / np.linalg.norm(world, axis=-1, keepdims=True)
    N = draw[12:24].reshape(4, 3).astype(np.float64)
    theta = max(angle(N[0], N[1]), angle(N[2], N[3]), angle(N[0], N[2]))
    lin_ab, lin_cd, lin_pq = (b.any() for b in r["branch"])
    dr = np.linalg.norm(world, axis=-1) - R
    if not (lin_ab or lin_cd or lin_pq):                                           # all three interpolations slerp: on the sphere
        assert np.abs(dr).max() <= 4.0, np.abs(dr).max()                           # fp32 uniforms at 6.4e6 m: 0.5 m ulp
        assert angle(r["vn"], radial).max() <= 1e-6
    else:                                                                          # chords: below the sphere by at most the sagitta(s)
        assert dr.max() <= 4.0 and dr.min() >= -2.0 * R * theta * theta / 8 - 4.0
    shift = r["pos"] - r["vp"]
    assert np.abs(np.linalg.norm(shift, axis=-1) - np.abs(r["height"])).max() <= 1e-6 * max(1.0, abs(flat_height))
    return theta


def test_flat_height_map_invariants_of_the_shader(golden):
    uv3, quads, maps, cam, draws = frame(golden)
    flat = np.full((32, 32), 1234.5, np.float32)
    seen = set()
    for k in range(0, 117, 3):
        r = shade_draw(uv3, draws[k], flat)
        sphere_invariants(r, draws[k], cam, 1234.5)
        seen.add(tuple(bool(b.any()) for b in r["branch"]))
        skirtless = uv3[:, 2] == 0
        assert np.abs(r["height"][skirtless] - 1234.5).max() <= 1e-3
        assert np.abs(r["height"][~skirtless] - (1234.5 - draws[k][24])).max() <= 1e-3
    assert (False, False, False) in seen and (True, True, True) in seen            # both branches exercised


@pytest.mark.gpu
def test_k3_agrees_with_the_independent_float64_evaluation(gpu, golden):
    import torch
    uv3, quads, maps, cam, draws = frame(golden)
    dq = gpu.quads_to_device(quads)
    pos, nrm = gpu.shade(dq, torch.from_numpy(np.ascontiguousarray(maps)).cuda(), cam)
    worst = check_against_evaluator(golden, pos.cpu().numpy(), nrm.cpu().numpy())
    assert worst <= 1e-5, worst


@pytest.mark.gpu
def test_k3_flat_height_map_normal_is_the_interpolated_normal_and_positions_lie_on_the_arc(gpu, golden):
    import torch
    uv3, quads, maps, cam, draws = frame(golden)
    flat = torch.full((len(quads), 32, 32), 1234.5, dtype=torch.float32, device="cuda")
    pos, nrm = gpu.shade(gpu.quads_to_device(quads), flat, cam)
    pos, nrm = pos.cpu().numpy().astype(np.float64), nrm.cpu().numpy().astype(np.float64)
    flat_np = np.full((32, 32), 1234.5, np.float32)
    skirtless = uv3[:, 2] == 0
    for k in range(len(quads)):
        r = shade_draw(uv3, draws[k], flat_np)
        assert angle(nrm[k, :, :3], r["vn"]).max() <= 1e-5                         # Normal == v.n
        world = pos[k, skirtless, :3] + cam                                        # = (v.p + cam) + v.n * h
        ev = r["pos"][skirtless] + cam
        all_slerp = not any(b.any() for b in r["branch"])
        if all_slerp:                                                              # on the sphere of radius R + h
            assert np.abs(np.linalg.norm(world, axis=-1) - (R + 1234.5)).max() <= 8.0, k
        tol = 8 * 2.0 ** -23 * np.abs(quads[k]["p"] - cam).max() + 1e-5 * np.linalg.norm(quads[k]["p"][3] - quads[k]["p"][0])
        assert np.abs(world - ev).max() <= tol
        # the four corner vertices are the endpoints of interpolate: P[i] + N[i] * h (slots of UV = (0,0),(1,0),(0,1),(1,1))
        for i, (ux, uy) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
            slot = int(np.flatnonzero((uv3[:, 0] == ux) & (uv3[:, 1] == uy) & skirtless)[0])
            want = draws[k][3 * i:3 * i + 3].astype(np.float64) + draws[k][12 + 3 * i:15 + 3 * i].astype(np.float64) * 1234.5
            # (in the slerp branch y = 1/sin(theta) - 1/(cos(gamma) tan(theta)) cancels to ~1 fp32 ulp of 1/sin(theta),
            #  times half the chord: the shader's own rounding, a fraction of a metre on a 1000 km quad)
            assert np.abs(pos[k, slot, :3] - want).max() <= tol, (k, i)


def test_restatement_of_the_parent_fallback_sampling_agrees_with_the_evaluator(ref, port, golden):
    """The other way the shader reads a map: through texrect corners that select the child's quadrant of the
    PARENT's texture, with a larger pixel size (main.cpp:212-236, 334-346, 358) and GL_LINEAR filtering between
    texels.  Draws of that kind are taken from the reference on a flight that exhausts the generation budget; the
    restatement's rect path is compared with the float64 evaluation fed with the reference's own uniforms."""
    from test_cache import camera_path
    from oracle.bindings import height_params
    uv3 = golden["patch_vertex_buffer"].view(np.float32).reshape(-1, 3)
    ref.reset_cache()
    textures, checked = {}, 0
    for cam in camera_path()[:6]:
        quads, maps, draws, tex = ref.render_next_frame(cam, height_params())
        textures.update(maps)
        fallback = np.flatnonzero(draws[:, 29] != np.float32(1.0 / 32))
        for k in fallback[:: max(1, len(fallback) // 12)]:
            tmap = textures[int(tex[k])]
            r = shade_draw(uv3, draws[k], tmap)
            pos, nrm = port.shade_patches_rect(quads[k:k + 1], cam, tmap[None], np.zeros(1, np.int32), draws[k:k + 1, 25:31])
            assert angle(r["normal"], nrm[0, :, :3]).max() <= 1e-5
            hrange = float(np.ptp(tmap))
            assert np.abs(r["height"] - pos[0, :, 3]).max() <= 1e-5 * hrange + 0.05      # fp32 bilinear weights vs float64
            rel = np.abs(quads[k]["p"] - cam).max()
            ext = np.linalg.norm(quads[k]["p"][3] - quads[k]["p"][0])
            assert np.abs(r["pos"] - pos[0, :, :3]).max() <= 8 * 2.0 ** -23 * rel + 1e-5 * ext + 1e-5 * hrange
            checked += 1
    assert checked >= 10, "the flight produced no parent-fallback draws to check"
