"""planet_b200 -- B200-native terrain hot path of pgcomp/planet behind a C-ABI.

The product is ``libplanet_gpu.so`` (hand-written sm_100a CUDA, include/planet_gpu.h).
This module is the thin Python host layer the tests and bench.py use: a ctypes binding
with the reference's vocabulary (quads, height maps, patches).  PyTorch appears only as
plumbing -- device memory, streams, torch.distributed -- never as the compute path, and
there is no CPU fallback: if the library or a GPU is missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

RIDGED, FBM, ZERO = 0, 1, 2            # PLANET_NOISE_*
EXACT, FAST = 0, 1                     # PLANET_PRECISION_*
RADIUS = 6371000.0                     # main.cpp:821

QUAD_DTYPE = np.dtype([("p", np.float64, (4, 3)), ("id", np.uint64)])   # main.cpp:68-72, 104 B
QUAD_WORDS = 13                        # a Quad as 13 x 8-byte words (torch side: int64[n, 13])


class PlanetGpuError(RuntimeError):
    pass


class TexRect(C.Structure):
    """planet_gpu_texrect: what GetHeightMapForQuad returns per quad (main.cpp:184-189)."""
    _fields_ = [("slot", C.c_int32), ("flags", C.c_int32), ("corners", C.c_float * 4), ("pixel_size", C.c_float * 2)]


TEXRECT_DTYPE = np.dtype([("slot", np.int32), ("flags", np.int32), ("corners", np.float32, 4), ("pixel_size", np.float32, 2)])
TEXRECT_HIT, TEXRECT_GENERATED, TEXRECT_PARENT = 0, 1, 2


class Params(C.Structure):
    """planet_gpu_params (include/planet_gpu.h) -- the reference's constants, SURVEY App. D."""
    _fields_ = [("radius", C.c_double), ("patch_verts", C.c_int32), ("noise_kind", C.c_int32),
                ("lacunarity", C.c_double), ("gain", C.c_float), ("fixed_octaves", C.c_int32),
                ("coord_scale", C.c_double), ("height_scale", C.c_float), ("precision", C.c_int32),
                ("seed_offset", C.c_double * 3)]


_lib = None


def lib():
    """Load (building if needed) the C-ABI library.  Raises if it cannot be built/loaded."""
    global _lib
    if _lib is None:
        so = _build.build()
        L = C.CDLL(so)
        i, i64, f, d, vp = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_void_p
        pp = C.POINTER(Params)
        sig = {
            "planet_gpu_abi_version": (i, []),
            "planet_gpu_default_params": (None, [pp]),
            "planet_gpu_init": (i, [i]),
            "planet_gpu_shutdown": (None, []),
            "planet_gpu_last_error": (C.c_char_p, []),
            "planet_gpu_device_info": (i, [C.c_char_p, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
            "planet_gpu_set_params": (i, [pp]),
            "planet_gpu_get_height_at": (f, [vp, i, i]),
            "planet_gpu_generate_height_map": (None, [vp, i, vp, i]),
            "planet_gpu_generate_height_maps": (i, [pp, vp, i64, i, i, vp, vp]),
            "planet_gpu_generate_height_maps_gathered": (i, [pp, vp, i64, i, i, vp, vp, i, vp]),
            "planet_gpu_gather_unique_id": (i, [vp]),
            "planet_gpu_gather_create": (vp, [vp, i, i, i64, i]),
            "planet_gpu_gather_destroy": (None, [vp]),
            "planet_gpu_gather_buffer": (vp, [vp, i]),
            "planet_gpu_gather_last_buffer": (i, [vp]),
            "planet_gpu_gather_height_maps": (i, [vp, pp, vp, i64, i64, i, i, vp]),
            "planet_gpu_gather_wait": (i, [vp, i, vp]),
            "planet_gpu_gather_set_shade_share": (i, [vp, i]),
            "planet_gpu_gather_shade": (i, [vp, pp, vp, i64, i64, vp, f, vp, vp, vp]),
            "planet_gpu_gather_begin": (i, [vp, vp]),
            "planet_gpu_gather_push": (i, [vp, i64, i64, vp]),
            "planet_gpu_gather_publish": (i, [vp]),
            "planet_gpu_gather_set_push_mode": (i, [vp, i]),
            "planet_gpu_gather_nccl": (i, [vp, i, vp, vp, vp]),
            "planet_gpu_gather_barrier": (i, [vp, vp]),
            "planet_gpu_gather_error": (i, [vp]),
            "planet_gpu_heights_at": (i, [pp, vp, i64, i, i, vp, vp]),
            "planet_gpu_noise": (i, [vp, i64, i, d, f, i, i, vp, vp]),
            "planet_gpu_tessellate_uniform": (i, [pp, i, i64, i64, vp, vp, vp]),
            "planet_gpu_merged_indices_beside": (i, [pp, i64, vp, vp]),
            "planet_gpu_quads_from_ids": (i, [pp, vp, i64, vp, vp]),
            "planet_gpu_patch_mesh": (i, [i, vp, vp, vp]),
            "planet_gpu_patch_vertex_count": (i, [i]),
            "planet_gpu_patch_index_count": (i, [i]),
            "planet_gpu_strip_index": (C.c_uint32, [i, i]),
            "planet_gpu_uniform_leaf_id": (C.c_uint64, [i64, i]),
            "planet_gpu_max_lod": (i, [d, i]),
            "planet_gpu_max_skirt_size": (f, [d, i]),
            "planet_gpu_select_lod": (i, [pp, vp, i, vp, i64, C.POINTER(i64), vp]),
            "planet_gpu_shade": (i, [pp, vp, i64, vp, vp, f, vp, vp, vp]),
            "planet_gpu_cache_create": (vp, [i, i, i, i]),
            "planet_gpu_cache_destroy": (None, [vp]),
            "planet_gpu_cache_count": (i, [vp]),
            "planet_gpu_cache_plan_frame": (i, [vp, vp, i64, i, vp, C.POINTER(i64)]),
            "planet_gpu_cache_frame": (i, [vp, pp, vp, i64, i, i, vp, vp, vp]),
            "planet_gpu_cache_frame_device": (i, [vp, pp, vp, i64, i, i, vp, C.POINTER(i64), vp]),
            "planet_gpu_cache_pool": (vp, [vp]),
            "planet_gpu_cache_read_slots": (i, [vp, vp, i64, vp]),
            "planet_gpu_shade_cached": (i, [pp, vp, i64, vp, vp, vp, f, vp, vp, vp]),
            "planet_gpu_generate_height_maps_host": (i, [pp, vp, i64, i, i, vp, vp]),
            "planet_gpu_terrain_host": (i, [pp, vp, i64, i, vp, f, vp, vp, vp, vp]),
            "planet_gpu_measure_fp32_peak": (i, [d, C.POINTER(d), C.POINTER(d)]),
            "planet_gpu_launch_count": (i64, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)            # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "planet_gpu_abi_version", "planet_gpu_default_params", "planet_gpu_init", "planet_gpu_shutdown",
    "planet_gpu_last_error", "planet_gpu_device_info", "planet_gpu_set_params",
    "planet_gpu_get_height_at", "planet_gpu_generate_height_map", "planet_gpu_generate_height_maps",
    "planet_gpu_generate_height_maps_gathered", "planet_gpu_gather_unique_id", "planet_gpu_gather_create",
    "planet_gpu_gather_destroy", "planet_gpu_gather_buffer", "planet_gpu_gather_last_buffer",
    "planet_gpu_gather_height_maps", "planet_gpu_gather_wait", "planet_gpu_gather_begin", "planet_gpu_gather_push",
    "planet_gpu_gather_publish", "planet_gpu_gather_set_push_mode", "planet_gpu_gather_set_shade_share", "planet_gpu_gather_shade", "planet_gpu_gather_nccl",
    "planet_gpu_gather_barrier", "planet_gpu_gather_error", "planet_gpu_heights_at", "planet_gpu_noise", "planet_gpu_tessellate_uniform", "planet_gpu_merged_indices_beside",
    "planet_gpu_quads_from_ids", "planet_gpu_patch_mesh", "planet_gpu_patch_vertex_count",
    "planet_gpu_patch_index_count", "planet_gpu_strip_index", "planet_gpu_uniform_leaf_id",
    "planet_gpu_max_lod", "planet_gpu_max_skirt_size",
    "planet_gpu_select_lod", "planet_gpu_shade", "planet_gpu_generate_height_maps_host",
    "planet_gpu_terrain_host",
    "planet_gpu_cache_create", "planet_gpu_cache_destroy", "planet_gpu_cache_count",
    "planet_gpu_cache_plan_frame", "planet_gpu_cache_frame", "planet_gpu_cache_frame_device", "planet_gpu_cache_pool",
    "planet_gpu_cache_read_slots", "planet_gpu_shade_cached", "planet_gpu_measure_fp32_peak",
    "planet_gpu_launch_count",
]


def _check(rc):
    if rc != 0:
        raise PlanetGpuError(f"planet_gpu error {rc}: {lib().planet_gpu_last_error().decode()}")


def default_params(**over):
    """The reference's defaults (ridged, gain 0.55, octaves 6+12*depth/max_depth, EXACT)."""
    p = Params()
    lib().planet_gpu_default_params(C.byref(p))
    for k, v in over.items():
        if k == "seed_offset":
            p.seed_offset = (C.c_double * 3)(*v)
        else:
            setattr(p, k, v)
    return p


def fbm_params(octaves=8, gain=0.5, precision=FAST, **over):
    """BASELINE.json configs 2-5: fBm, fixed octave count, gain 0.5."""
    return default_params(noise_kind=FBM, gain=gain, fixed_octaves=octaves, precision=precision, **over)


def init(device=0):
    _check(lib().planet_gpu_init(int(device)))


def patch_vertex_count(n=30): return lib().planet_gpu_patch_vertex_count(n)
def patch_index_count(n=30): return lib().planet_gpu_patch_index_count(n)
def max_lod(radius=RADIUS, n=30): return lib().planet_gpu_max_lod(radius, n)
def max_skirt_size(radius=RADIUS, n=30): return lib().planet_gpu_max_skirt_size(radius, n)
def launch_count(): return lib().planet_gpu_launch_count()


def device_info():
    name = C.create_string_buffer(128)
    sm, khz, lanes = C.c_int(), C.c_int(), C.c_int()
    _check(lib().planet_gpu_device_info(name, 128, C.byref(sm), C.byref(khz), C.byref(lanes)))
    return {"name": name.value.decode(), "sm_count": sm.value, "clock_khz": khz.value,
            "fp32_lanes_per_sm": lanes.value}


def measure_fp32_peak(ms=200.0):
    tf, el = C.c_double(), C.c_double()
    _check(lib().planet_gpu_measure_fp32_peak(ms, C.byref(tf), C.byref(el)))
    return tf.value, el.value


# ---- torch plumbing -----------------------------------------------------------------------
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise PlanetGpuError("no CUDA device: planet_b200 has no CPU path")
    return torch


def _stream(stream=None):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def quads_to_device(quads, device="cuda"):
    """numpy structured quads (QUAD_DTYPE) -> int64[n, 13] device tensor (same bytes)."""
    torch = _torch()
    q = np.ascontiguousarray(quads, QUAD_DTYPE)
    return torch.from_numpy(q.view(np.int64).reshape(-1, QUAD_WORDS).copy()).to(device)


def quads_to_host(t):
    return t.detach().cpu().numpy().reshape(-1).view(QUAD_DTYPE)


def tessellate_uniform(depth, first=0, nquads=None, params=None, with_indices=False, stream=None):
    """K1: quads [first, first+nquads) of the uniform depth-`depth` tree (+ merged strip indices)."""
    torch = _torch()
    params = params or default_params()
    if nquads is None:
        nquads = 6 * 4 ** depth - first
    quads = torch.empty((nquads, QUAD_WORDS), dtype=torch.int64, device="cuda")
    idx = None
    if with_indices:
        idx = torch.empty(nquads * patch_index_count(params.patch_verts), dtype=torch.int32, device="cuda")
    _check(lib().planet_gpu_tessellate_uniform(C.byref(params), depth, first, nquads, quads.data_ptr(),
                                               idx.data_ptr() if idx is not None else None, _stream(stream)))
    return (quads, idx) if with_indices else quads


def merged_indices_beside(nquads, params=None, out=None, stream=None):
    """The merged strip index buffer alone, by the kernel small enough to run beside K2 (issue it on a side stream)."""
    torch = _torch()
    params = params or default_params()
    if out is None:
        out = torch.empty(nquads * patch_index_count(params.patch_verts), dtype=torch.int32, device="cuda")
    _check(lib().planet_gpu_merged_indices_beside(C.byref(params), nquads, out.data_ptr(), _stream(stream)))
    return out


def quads_from_ids(ids, params=None, stream=None):
    torch = _torch()
    params = params or default_params()
    ids = ids if hasattr(ids, "data_ptr") else torch.from_numpy(np.ascontiguousarray(ids, np.uint64).view(np.int64)).cuda()
    out = torch.empty((ids.numel(), QUAD_WORDS), dtype=torch.int64, device="cuda")
    _check(lib().planet_gpu_quads_from_ids(C.byref(params), ids.data_ptr(), ids.numel(), out.data_ptr(), _stream(stream)))
    return out


def patch_mesh(n=30, stream=None):
    torch = _torch()
    v = torch.empty((patch_vertex_count(n), 3), dtype=torch.float32, device="cuda")
    i = torch.empty(patch_index_count(n), dtype=torch.int32, device="cuda")
    _check(lib().planet_gpu_patch_mesh(n, v.data_ptr(), i.data_ptr(), _stream(stream)))
    return v, i


def select_lod(cam_pos, max_lod=None, params=None, capacity=65536, stream=None):
    """K0: the leaf quads ProcessQuad/RenderPlanet (main.cpp:537-624) select for a camera, in the
    reference's order.  Returns an int64[n,13] device tensor."""
    torch = _torch()
    params = params or default_params()
    if max_lod is None:
        max_lod = lib().planet_gpu_max_lod(params.radius, params.patch_verts)
    buf = torch.empty((capacity, QUAD_WORDS), dtype=torch.int64, device="cuda")
    cam = (C.c_double * 3)(*[float(c) for c in cam_pos])
    n = C.c_int64(0)
    _check(lib().planet_gpu_select_lod(C.byref(params), cam, max_lod, buf.data_ptr(), capacity, C.byref(n), _stream(stream)))
    return buf[:n.value]


def generate_height_maps(quads, dim, max_depth, params=None, out=None, stream=None):
    """K2: batched GenerateHeightMap (main.cpp:123-151).  quads: int64[n,13] device tensor."""
    torch = _torch()
    params = params or default_params()
    n = quads.shape[0]
    if out is None:
        out = torch.empty((n, dim, dim), dtype=torch.float32, device="cuda")
    _check(lib().planet_gpu_generate_height_maps(C.byref(params), quads.data_ptr(), n, dim, max_depth,
                                                 out.data_ptr(), _stream(stream)))
    return out


def generate_height_maps_gathered(quads, dim, max_depth, out, peer_outs, params=None, stream=None):
    """K2 with the all-gather fused in: `out` is this rank's shard inside its own gathered buffer,
    `peer_outs` the same shard inside each peer's (IPC-mapped) gathered buffer."""
    params = params or default_params()
    n = quads.shape[0]
    arr = (C.c_void_p * max(len(peer_outs), 1))(*[t.data_ptr() for t in peer_outs])
    _check(lib().planet_gpu_generate_height_maps_gathered(C.byref(params), quads.data_ptr(), n, dim, max_depth,
                                                          out.data_ptr(), arr, len(peer_outs), _stream(stream)))
    return out


def heights_at(points, depth, max_depth, params=None, stream=None):
    """Batched GetHeightAt (main.cpp:118-121).  points: float64[n,3] device tensor."""
    torch = _torch()
    params = params or default_params()
    n = points.shape[0]
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _check(lib().planet_gpu_heights_at(C.byref(params), points.data_ptr(), n, depth, max_depth,
                                       out.data_ptr(), _stream(stream)))
    return out


def noise(points, kind=FBM, lacunarity=2.0, gain=0.5, octaves=0, precision=EXACT, stream=None):
    """PerlinNoise3 (octaves=0) / PerlinfBm / PerlinRidged on float64[n,3] device points."""
    torch = _torch()
    n = points.shape[0]
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _check(lib().planet_gpu_noise(points.data_ptr(), n, kind, lacunarity, gain, octaves, precision,
                                  out.data_ptr(), _stream(stream)))
    return out


def shade(quads, heights, cam_pos, params=None, max_skirt=-1.0, want_pos=True, want_nrm=True,
          pos=None, nrm=None, stream=None):
    """K3: displaced positions + normals/Lambert for every patch vertex (main.cpp:286-380)."""
    torch = _torch()
    params = params or default_params()
    n = quads.shape[0]
    nv = patch_vertex_count(params.patch_verts)
    if want_pos and pos is None:
        pos = torch.empty((n, nv, 4), dtype=torch.float32, device="cuda")
    if want_nrm and nrm is None:
        nrm = torch.empty((n, nv, 4), dtype=torch.float32, device="cuda")
    cam = (C.c_double * 3)(*[float(c) for c in cam_pos])
    _check(lib().planet_gpu_shade(C.byref(params), quads.data_ptr(), n, cam, heights.data_ptr(), max_skirt,
                                  pos.data_ptr() if pos is not None else None,
                                  nrm.data_ptr() if nrm is not None else None, _stream(stream)))
    return pos, nrm


# ---- host-buffer entry points (the reference-facing calls) ----------------------------------
def generate_height_maps_host(quads_np, dim, max_depth, params=None, out=None, mirror=None):
    """Host quads in, host height maps out: H2D + K2 + D2H inside the call.  `mirror`: optional
    device tensor that also keeps the maps resident (the GL texture's role, main.cpp:245)."""
    params = params or default_params()
    q = np.ascontiguousarray(quads_np, QUAD_DTYPE)
    if out is None:
        out = np.empty((len(q), dim, dim), np.float32)
    _check(lib().planet_gpu_generate_height_maps_host(C.byref(params), q.ctypes.data, len(q), dim, max_depth,
                                                      out.ctypes.data,
                                                      mirror.data_ptr() if mirror is not None else None))
    return out


def terrain_host(quads_np, max_depth, cam_pos, params=None, max_skirt=-1.0, out=None, heights=None,
                 pos=None, nrm=None):
    """Host quads in; height maps out on the host AND resident on the device, displaced positions +
    normals on the device (K3 runs while the last maps cross PCIe).  Returns (out, heights, pos, nrm)."""
    torch = _torch()
    params = params or default_params()
    q = np.ascontiguousarray(quads_np, QUAD_DTYPE)
    n, dim = len(q), params.patch_verts + 2
    nv = patch_vertex_count(params.patch_verts)
    if out is None:
        out = np.empty((n, dim, dim), np.float32)
    if heights is None:
        heights = torch.empty((n, dim, dim), dtype=torch.float32, device="cuda")
    if pos is None:
        pos = torch.empty((n, nv, 4), dtype=torch.float32, device="cuda")
    if nrm is None:
        nrm = torch.empty((n, nv, 4), dtype=torch.float32, device="cuda")
    cam = (C.c_double * 3)(*[float(c) for c in cam_pos])
    _check(lib().planet_gpu_terrain_host(C.byref(params), q.ctypes.data, n, max_depth, cam, max_skirt,
                                         out.ctypes.data, heights.data_ptr(), pos.data_ptr(), nrm.data_ptr()))
    return out, heights, pos, nrm


def set_params(params):
    _check(lib().planet_gpu_set_params(C.byref(params)))


def generate_height_map(quad_np, dim, max_depth):
    """The reference-shaped call: void GenerateHeightMap(float*, int, const Quad&, int)."""
    q = np.ascontiguousarray(quad_np, QUAD_DTYPE).reshape(1)
    out = np.empty((dim, dim), np.float32)
    lib().planet_gpu_generate_height_map(out.ctypes.data, dim, q.ctypes.data, max_depth)
    return out


def get_height_at(p, depth, max_depth):
    """The reference-shaped call: float GetHeightAt(const Vec3d&, int, int)."""
    v = np.ascontiguousarray(p, np.float64)
    return lib().planet_gpu_get_height_at(v.ctypes.data, depth, max_depth)


class HeightMapCache:
    """Device-resident pool with the reference's cache semantics (main.cpp:75-102, 191-278)."""

    def __init__(self, dim=32, cache_max=1024, map_max=1499, extra_slots=1024):
        self.dim, self.extra = dim, extra_slots
        self.handle = lib().planet_gpu_cache_create(dim, cache_max, map_max, extra_slots)
        if not self.handle:
            raise PlanetGpuError(lib().planet_gpu_last_error().decode())
        self.pool_slots = cache_max + extra_slots

    def close(self):
        if self.handle:
            lib().planet_gpu_cache_destroy(self.handle)
            self.handle = None

    __del__ = close

    @property
    def count(self):
        return lib().planet_gpu_cache_count(self.handle)

    def plan_frame(self, quads_np, generations_per_frame=100):
        """Host bookkeeping only: (texrects, number of maps to generate)."""
        q = np.ascontiguousarray(quads_np, QUAD_DTYPE)
        rects = np.zeros(len(q), TEXRECT_DTYPE)
        n_gen = C.c_int64(0)
        _check(lib().planet_gpu_cache_plan_frame(self.handle, q.ctypes.data, len(q), generations_per_frame,
                                                 rects.ctypes.data, C.byref(n_gen)))
        return rects, n_gen.value

    def frame(self, quads_np, max_lod, params=None, generations_per_frame=100, stream=None):
        """Plan + one batched K2 launch for the misses.  Returns (host texrects, device texrects)."""
        torch = _torch()
        params = params or default_params()
        q = np.ascontiguousarray(quads_np, QUAD_DTYPE)
        rects = np.zeros(len(q), TEXRECT_DTYPE)
        d_rects = torch.empty((len(q), 8), dtype=torch.int32, device="cuda")
        _check(lib().planet_gpu_cache_frame(self.handle, C.byref(params), q.ctypes.data, len(q), max_lod,
                                            generations_per_frame, rects.ctypes.data, d_rects.data_ptr(), _stream(stream)))
        return rects, d_rects

    def frame_device(self, d_quads, max_lod, params=None, generations_per_frame=100, stream=None):
        """The frame with the leaf quads already on the device (K0's output): bookkeeping in one
        kernel on device-resident tables + one batched K2 launch.  Returns (device texrects, maps generated)."""
        torch = _torch()
        params = params or default_params()
        n = d_quads.shape[0]
        d_rects = torch.empty((n, 8), dtype=torch.int32, device="cuda")
        n_gen = C.c_int64(0)
        _check(lib().planet_gpu_cache_frame_device(self.handle, C.byref(params), d_quads.data_ptr(), n, max_lod,
                                                   generations_per_frame, d_rects.data_ptr(), C.byref(n_gen), _stream(stream)))
        return d_rects, n_gen.value

    def pool_ptr(self):
        return lib().planet_gpu_cache_pool(self.handle)

    def read_slots(self, slots):
        """Host copy of the maps in the given pool slots: float32[n, dim, dim]."""
        s = np.ascontiguousarray(slots, np.int32)
        out = np.empty((len(s), self.dim, self.dim), np.float32)
        _check(lib().planet_gpu_cache_read_slots(self.handle, s.ctypes.data, len(s), out.ctypes.data))
        return out


def shade_cached(quads, cache, d_rects, cam_pos, params=None, max_skirt=-1.0, stream=None):
    """K3 through the cache's texrects (bilinear sampling, parent fallback)."""
    torch = _torch()
    params = params or default_params()
    n = quads.shape[0]
    nv = patch_vertex_count(params.patch_verts)
    pos = torch.empty((n, nv, 4), dtype=torch.float32, device="cuda")
    nrm = torch.empty((n, nv, 4), dtype=torch.float32, device="cuda")
    cam = (C.c_double * 3)(*[float(c) for c in cam_pos])
    _check(lib().planet_gpu_shade_cached(C.byref(params), quads.data_ptr(), n, cam, cache.pool_ptr(), d_rects.data_ptr(),
                                         max_skirt, pos.data_ptr(), nrm.data_ptr(), _stream(stream)))
    return pos, nrm
