"""TEST INFRASTRUCTURE.  The reference's vertex + fragment shader (main.cpp:286-380) evaluated in
float64 numpy, written from the GLSL text alone -- it shares no code with the C restatement
(oracle/planet_oracle.c: glsl_*) nor with K3, so a transcription error common to those two shows
up here.  Inputs are what the reference itself hands to GL, all pinned by the recording GL of
oracle/ref_oracle.cpp: the UV attribute stream (golden patch_vertex_buffer), the per-draw
uniforms P[4], N[4], SkirtSize, HeightMap_corners, HeightMap_pixel_size (golden frame_draws) and
the GL_R32F height map (golden frame_height_maps; GL_LINEAR, GL_CLAMP_TO_EDGE, render.cpp:429-433).

Output per vertex: v.p + v.n*height (what the shader multiplies by Projection*View), the Normal
varying, and the fragment stage's sqrt(light) evaluated at the vertex."""
import numpy as np


def _dot(a, b):
    return (a * b).sum(-1, keepdims=True)


def _normalize(a):
    return a / np.sqrt(_dot(a, a))


def _mix(a, b, t):
    return a * (1.0 - t) + b * t                                   # GLSL mix: x*(1-a) + y*a


def interpolate_linear(p0, n0, p1, n1, t):                         # main.cpp:300-308
    return _mix(p0, p1, t), _normalize(_mix(n0, n1, t))


def interpolate(p0, n0, p1, n1, t):                                # main.cpp:310-331
    """p*, n*: (..., 3); t: (..., 1).  Both branches are evaluated, the shader's test selects."""
    d = _dot(n0, n1)
    linear = (1.0 - d) < 0.001
    lp, ln = interpolate_linear(p0, n0, p1, n1, t)
    with np.errstate(all="ignore"):
        theta2 = np.arccos(np.clip(d, -1.0, 1.0))
        k = 1.0 - t
        n = _normalize(np.sin(k * theta2) * n0 + np.sin(t * theta2) * n1)
        theta = theta2 * 0.5
        gamma = theta - theta2 * t
        tan_theta = np.tan(theta)
        x = 1.0 - np.tan(gamma) / tan_theta
        y = 1.0 / np.sin(theta) - 1.0 / (np.cos(gamma) * tan_theta)
        v = (p1 - p0) * 0.5
        p = p0 + x * v + y * n * np.sqrt(_dot(v, v))
    return np.where(linear, lp, p), np.where(linear, ln, n), linear


def texture_linear(tex, uv):
    """GL_LINEAR / GL_CLAMP_TO_EDGE fetch of a (H, W) single-channel texture at uv (..., 2) in [0,1]."""
    h, w = tex.shape
    x = uv[..., 0] * w - 0.5
    y = uv[..., 1] * h - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = x - x0, y - y0
    xi0 = np.clip(x0, 0, w - 1).astype(int); xi1 = np.clip(x0 + 1, 0, w - 1).astype(int)
    yi0 = np.clip(y0, 0, h - 1).astype(int); yi1 = np.clip(y0 + 1, 0, h - 1).astype(int)
    t = tex.astype(np.float64)
    return (t[yi0, xi0] * (1 - fx) + t[yi0, xi1] * fx) * (1 - fy) + (t[yi1, xi0] * (1 - fx) + t[yi1, xi1] * fx) * fy


def shade_draw(uv3, draw32, height_map):
    """One Draw(planet.patch) (main.cpp:655-680 sets the uniforms, :345-365 is main()).
    uv3: (nv, 3) UV attribute; draw32: one golden frame_draws row; height_map: (H, W) float32.
    Returns dict(pos=(nv,3), normal=(nv,3), light=(nv,), height=(nv,), vn=(nv,3), vp=(nv,3), branch=...)."""
    uv3 = np.asarray(uv3, np.float64)
    d = np.asarray(draw32, np.float64)
    P = d[0:12].reshape(4, 3); N = d[12:24].reshape(4, 3)
    skirt = d[24]
    c0, c1 = d[25:27], d[27:29]
    pix = d[29:31]
    ux, uy, uz = uv3[:, 0:1], uv3[:, 1:2], uv3[:, 2]
    nv = len(uv3)
    b = lambda a: np.broadcast_to(a, (nv, 3))
    pp, pn, lin_ab = interpolate(b(P[0]), b(N[0]), b(P[1]), b(N[1]), ux)
    qp, qn, lin_cd = interpolate(b(P[2]), b(N[2]), b(P[3]), b(N[3]), ux)
    vp, vn, lin_pq = interpolate(pp, pn, qp, qn, uy)
    uv = _mix(c0[None, :], c1[None, :], uv3[:, :2])
    height = texture_linear(height_map, uv) - skirt * uz
    # compute_normal (main.cpp:337-343): offs = (px, 0, py); offs.xy = (px, 0), offs.yz = (0, py)
    ox = np.array([pix[0], 0.0]); oy = np.array([0.0, pix[1]])
    x0 = texture_linear(height_map, uv - ox); x1 = texture_linear(height_map, uv + ox)
    y0 = texture_linear(height_map, uv - oy); y1 = texture_linear(height_map, uv + oy)
    xyscale = np.sqrt(_dot(qp - pp, qp - pp))[:, 0] / 29.0
    normal = _normalize(np.stack([x0 - x1, 2.0 * xyscale, y0 - y1], -1))
    t = _normalize(np.cross(vn, qp - pp))
    bi = _normalize(np.cross(t, vn))
    # mat3(t, n, bi) has columns t, n, bi
    world = _normalize(t * normal[:, 0:1] + vn * normal[:, 1:2] + bi * normal[:, 2:3])
    l = np.array([0.0, 1.0, -1.0]) / np.sqrt(2.0)                  # fragment stage, main.cpp:371-374
    light = 0.001 + np.maximum(0.0, (world * l).sum(-1))
    return dict(pos=vp + vn * height[:, None], normal=world, light=np.sqrt(light), height=height,
                vp=vp, vn=vn, branch=(lin_ab[:, 0], lin_cd[:, 0], lin_pq[:, 0]),
                margin=min(abs(1.0 - float((N[0] * N[1]).sum()) - 0.001), abs(1.0 - float((N[2] * N[3]).sum()) - 0.001),
                           float(np.abs(1.0 - (pn * qn).sum(-1) - 0.001).min())))


def draw_uniforms(quad, cam, max_skirt, corners=(1.5 / 32, 1.5 / 32, 30.5 / 32, 30.5 / 32), pixel=(1 / 32, 1 / 32)):
    """The uniforms main.cpp:655-680 would set for `quad` (used where no captured draw exists):
    P = float(q.p - cam), N = float(Normalize(q.p)), SkirtSize = max_skirt / (2 << (depth-1)) for depth > 1."""
    p = np.asarray(quad["p"], np.float64)
    P = (p - np.asarray(cam, np.float64)).astype(np.float32)
    N = (p / np.linalg.norm(p, axis=1, keepdims=True)).astype(np.float32)
    depth = int((int(quad["id"]) >> 55) & 31) - 1
    s = np.float32(max_skirt)
    if depth > 0:
        s = np.float32(s / np.float32(2 << depth))
    return np.concatenate([P.ravel(), N.ravel(), [s], corners, pixel, [0]]).astype(np.float32)
