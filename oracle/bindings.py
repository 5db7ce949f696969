"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

``PortOracle``  wraps oracle/libplanet_oracle.so (plain-C restatement, planet_oracle.c).
``RefOracle``   wraps oracle/_ref/libplanet_ref.so (the reference's own sources compiled
                with headless stubs, ref_oracle.cpp); present only if it was built in a
                container that has /root/reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  planet_b200/ never does: the product has no CPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libplanet_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libplanet_ref.so")
REF_O3_SO = os.path.join(HERE, "_ref", "libplanet_ref_o3.so")     # -O3 -march=x86-64-v3, same bits

RIDGED, FBM, ZERO = 0, 1, 2
RADIUS = 6371000.0                      # main.cpp:821

QUAD_DTYPE = np.dtype([("p", np.float64, (4, 3)), ("id", np.uint64)])   # main.cpp:68-72
assert QUAD_DTYPE.itemsize == 104


def build(force=False):
    """Compile the checkers (the port always; the reference when /root/reference exists)."""
    if force or not os.path.exists(PORT_SO) or \
            os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "planet_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if os.path.exists("/root/reference/main.cpp") and (
            force or not os.path.exists(REF_SO) or not os.path.exists(REF_O3_SO) or
            os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(HERE, "ref_oracle.cpp"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def fnv1a32(data) -> int:
    """FNV-1a-32 of a bytes-like / ndarray; the hash SURVEY.md quotes for buffers."""
    b = np.frombuffer(memoryview(np.ascontiguousarray(data)).cast("B"), dtype=np.uint8)
    h = 0x811C9DC5
    # chunked pure-python would be slow; do it in numpy-free C-ish loop only for small inputs
    for x in b.tobytes():
        h = ((h ^ x) * 0x01000193) & 0xFFFFFFFF
    return h


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _HeightParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("lacunarity", C.c_double), ("gain", C.c_float),
                ("fixed_octaves", C.c_int), ("coord_scale", C.c_double),
                ("height_scale", C.c_float)]


def height_params(kind=RIDGED, lacunarity=2.0, gain=0.55, fixed_octaves=0,
                  coord_scale=0.00001, height_scale=8848.0):
    return dict(kind=kind, lacunarity=lacunarity, gain=gain, fixed_octaves=fixed_octaves,
                coord_scale=coord_scale, height_scale=height_scale)


class PortOracle:
    """Plain-C restatement (oracle/planet_oracle.c)."""
    kind = "port"

    def __init__(self):
        build()
        L = self.L = C.CDLL(PORT_SO)
        f, d, i, l, u64, vp = C.c_float, C.c_double, C.c_int, C.c_long, C.c_uint64, C.c_void_p
        hp = C.POINTER(_HeightParams)
        sig = {
            "orc_perlin_random": (i, [i]),
            "orc_perlin_gradient": (f, [f, f, f, i, i, i]),
            "orc_perlin_noise3": (f, [d, d, d]),
            "orc_noise3_batch": (None, [vp, l, vp]),
            "orc_perlin_fbm": (f, [d, d, d, d, f, i]),
            "orc_perlin_ridged": (f, [d, d, d, d, f, i]),
            "orc_fractal_batch": (None, [vp, l, i, d, f, i, vp]),
            "orc_get_height_at": (f, [hp, vp, i, i]),
            "orc_generate_height_map": (None, [hp, vp, i, vp, i]),
            "orc_generate_height_maps": (None, [hp, vp, l, i, i, vp, i]),
            "orc_make_root_id": (u64, [u64]), "orc_make_child_id": (u64, [u64, u64]),
            "orc_get_parent_id": (u64, [u64]), "orc_get_root": (u64, [u64]),
            "orc_get_depth": (u64, [u64]), "orc_get_index": (u64, [u64]),
            "orc_get_child_index": (u64, [u64]),
            "orc_root_quads": (None, [d, vp]), "orc_split_quad": (None, [d, vp, vp]),
            "orc_uniform_quads": (l, [d, i, i, vp]), "orc_quad_from_id": (i, [d, u64, vp]),
            "orc_patch_vertex_count": (i, [i]), "orc_patch_index_count": (i, [i]),
            "orc_patch_vertices": (None, [i, vp]), "orc_patch_indices": (None, [i, vp]),
            "orc_max_lod": (i, [d, i]), "orc_max_skirt_size": (f, [d, i]),
            "orc_skirt_size_for_quad": (f, [f, u64]),
            "orc_quad_uniforms": (None, [vp, vp, vp]),
            "orc_shade_patches": (None, [vp, l, vp, vp, i, f, vp, vp]),
            "orc_shade_patches_rect": (None, [vp, l, vp, vp, vp, vp, i, f, vp, vp]),
            "orc_perlin_tables": (None, [vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    @staticmethod
    def _hp(params):
        return C.byref(_HeightParams(**(params or height_params())))

    def tables(self):
        t = np.zeros(256, np.uint8); v = np.zeros((16, 3), np.float32)
        self.L.orc_perlin_tables(_p(t), _p(v))
        return t, v

    def random(self, seed): return self.L.orc_perlin_random(int(seed))

    def noise3(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        out = np.empty(len(xyz), np.float32)
        self.L.orc_noise3_batch(_p(xyz), len(xyz), _p(out))
        return out

    def fractal(self, xyz, kind, lacunarity, gain, octaves):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        out = np.empty(len(xyz), np.float32)
        self.L.orc_fractal_batch(_p(xyz), len(xyz), kind, lacunarity, gain, octaves, _p(out))
        return out

    def get_height_at(self, p, depth, max_depth, params=None):
        p = np.ascontiguousarray(p, np.float64)
        return self.L.orc_get_height_at(self._hp(params), _p(p), depth, max_depth)

    def generate_height_maps(self, quads, dim, max_depth, params=None, nthreads=1):
        quads = np.ascontiguousarray(quads, QUAD_DTYPE)
        out = np.empty((len(quads), dim, dim), np.float32)
        self.L.orc_generate_height_maps(self._hp(params), _p(quads), len(quads), dim, max_depth,
                                        _p(out), nthreads)
        return out

    def make_root_id(self, r): return self.L.orc_make_root_id(r)
    def make_child_id(self, i, c): return self.L.orc_make_child_id(i, c)
    def get_parent_id(self, i): return self.L.orc_get_parent_id(i)
    def get_root(self, i): return self.L.orc_get_root(i)
    def get_depth(self, i): return self.L.orc_get_depth(i)
    def get_index(self, i): return self.L.orc_get_index(i)
    def get_child_index(self, i): return self.L.orc_get_child_index(i)

    def root_quads(self, radius=RADIUS):
        q = np.zeros(6, QUAD_DTYPE); self.L.orc_root_quads(radius, _p(q)); return q

    def split_quad(self, quad, radius=RADIUS):
        quad = np.ascontiguousarray(quad, QUAD_DTYPE).reshape(1)
        out = np.zeros(4, QUAD_DTYPE); self.L.orc_split_quad(radius, _p(quad), _p(out)); return out

    def uniform_quads(self, face, depth, radius=RADIUS):
        out = np.zeros(4 ** depth, QUAD_DTYPE)
        n = self.L.orc_uniform_quads(radius, face, depth, _p(out))
        assert n == len(out)
        return out

    def quad_from_id(self, qid, radius=RADIUS):
        out = np.zeros(1, QUAD_DTYPE)
        ok = self.L.orc_quad_from_id(radius, int(qid), _p(out))
        return out[0] if ok else None

    def patch_vertices(self, n=30):
        out = np.empty((self.L.orc_patch_vertex_count(n), 3), np.float32)
        self.L.orc_patch_vertices(n, _p(out)); return out

    def patch_indices(self, n=30):
        out = np.empty(self.L.orc_patch_index_count(n), np.uint32)
        self.L.orc_patch_indices(n, _p(out)); return out

    def max_lod(self, radius=RADIUS, n=30): return self.L.orc_max_lod(radius, n)
    def max_skirt_size(self, radius=RADIUS, n=30): return self.L.orc_max_skirt_size(radius, n)
    def skirt_size_for_quad(self, max_skirt, qid): return self.L.orc_skirt_size_for_quad(max_skirt, int(qid))

    def quad_uniforms(self, quads, cam_pos):
        """P[4], N[4] of every quad's draw (main.cpp:666-672) as float32[n, 24]."""
        quads = np.ascontiguousarray(quads, QUAD_DTYPE)
        cam = np.ascontiguousarray(cam_pos, np.float64)
        out = np.empty((len(quads), 24), np.float32)
        for k in range(len(quads)):
            self.L.orc_quad_uniforms(_p(quads[k:k + 1]), _p(cam), _p(out[k]))
        return out

    def shade_patches(self, quads, cam_pos, heights, n=30, max_skirt=None, radius=RADIUS):
        quads = np.ascontiguousarray(quads, QUAD_DTYPE)
        heights = np.ascontiguousarray(heights, np.float32)
        cam = np.ascontiguousarray(cam_pos, np.float64)
        if max_skirt is None:
            max_skirt = self.max_skirt_size(radius, n)
        nv = self.L.orc_patch_vertex_count(n)
        pos = np.empty((len(quads), nv, 4), np.float32); nrm = np.empty_like(pos)
        self.L.orc_shade_patches(_p(quads), len(quads), _p(cam), _p(heights), n, max_skirt,
                                 _p(pos), _p(nrm))
        return pos, nrm


    def shade_patches_rect(self, quads, cam_pos, pool, slots, rects, n=30, max_skirt=None, radius=RADIUS):
        """GLSL stage through texture rects (cache / parent fallback): rects float32[nq, 6]."""
        quads = np.ascontiguousarray(quads, QUAD_DTYPE)
        pool = np.ascontiguousarray(pool, np.float32)
        slots = np.ascontiguousarray(slots, np.int32)
        rects = np.ascontiguousarray(rects, np.float32).reshape(len(quads), 6)
        cam = np.ascontiguousarray(cam_pos, np.float64)
        if max_skirt is None:
            max_skirt = self.max_skirt_size(radius, n)
        nv = self.L.orc_patch_vertex_count(n)
        pos = np.empty((len(quads), nv, 4), np.float32); nrm = np.empty_like(pos)
        self.L.orc_shade_patches_rect(_p(quads), len(quads), _p(cam), _p(pool), _p(slots), _p(rects), n,
                                      max_skirt, _p(pos), _p(nrm))
        return pos, nrm


class RefOracle:
    """The reference's own code (oracle/ref_oracle.cpp includes /root/reference/main.cpp)."""
    kind = "reference"

    @staticmethod
    def available():
        if os.path.exists("/root/reference/main.cpp"):
            build()
        return os.path.exists(REF_SO)

    @staticmethod
    def o3_available():
        """The -O3 / AVX2 build exists and this host can run it."""
        if not RefOracle.available() or not os.path.exists(REF_O3_SO):
            return False
        try:
            flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
        except (OSError, StopIteration):
            return False
        return all(f in flags for f in ("avx2", "bmi2", "fma"))

    def __init__(self, o3=False):
        if not self.available():
            raise RuntimeError("oracle/_ref/libplanet_ref.so not built (needs /root/reference)")
        if o3 and not self.o3_available():
            raise RuntimeError("oracle/_ref/libplanet_ref_o3.so not built or not runnable on this CPU")
        self.flags = "-O3 -march=x86-64-v3 -ffp-contract=off" if o3 else "-O2 -ffp-contract=off"
        L = self.L = C.CDLL(REF_O3_SO if o3 else REF_SO)
        f, d, i, l, u64, vp = C.c_float, C.c_double, C.c_int, C.c_long, C.c_uint64, C.c_void_p
        sig = {
            "ref_perlin_tables": (None, [vp, vp]), "ref_perlin_random": (i, [i]),
            "ref_perlin_gradient": (f, [f, f, f, i, i, i]),
            "ref_perlin_noise3": (f, [d, d, d]),
            "ref_perlin_fbm": (f, [d, d, d, d, f, i]), "ref_perlin_ridged": (f, [d, d, d, d, f, i]),
            "ref_noise3_batch": (None, [vp, l, vp]),
            "ref_fractal_batch": (None, [vp, l, i, d, f, i, vp]),
            "ref_make_root_id": (u64, [u64]), "ref_make_child_id": (u64, [u64, u64]),
            "ref_get_parent_id": (u64, [u64]), "ref_get_root": (u64, [u64]),
            "ref_get_depth": (u64, [u64]), "ref_get_index": (u64, [u64]),
            "ref_get_child_index": (u64, [u64]), "ref_sizeof_quad": (i, []),
            "ref_set_functor": (None, [i, d, f, i, d, f]),
            "ref_get_height_at": (f, [vp, i, i]),
            "ref_generate_height_map": (None, [vp, i, vp, i]),
            "ref_generate_height_maps": (None, [vp, l, i, i, vp, i]),
            "ref_init_planet": (i, [d, vp, vp]), "ref_patch_buffer": (l, [i, vp, l]),
            "ref_root_quads": (i, [d, vp]), "ref_split_quad": (i, [d, vp, vp]),
            "ref_uniform_quads": (l, [d, i, i, vp]),
            "ref_render_frame": (l, [d, vp]), "ref_frame_quads": (l, [vp, l]),
            "ref_render_next_frame": (l, [d, vp]), "ref_reset_cache": (None, []),
            "ref_captured_draw_texture": (l, [l]), "ref_captured_height_map_id": (l, [l]),
            "ref_cache_count": (i, []),
            "ref_run_reference_main": (i, [C.c_char_p]),
            "ref_captured_height_map_count": (l, []),
            "ref_captured_height_map": (i, [l, vp, vp, vp]),
            "ref_captured_draw_count": (l, []), "ref_captured_draw": (i, [l, vp]),
            "ref_captured_shader_count": (l, []), "ref_captured_shader_source": (l, [l, vp, l]),
            "ref_captured_buffer_count": (l, []), "ref_captured_buffer": (l, [l, vp, l]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        assert L.ref_sizeof_quad() == 104

    def _set(self, params):
        p = params or height_params()
        self.L.ref_set_functor(p["kind"], p["lacunarity"], p["gain"], p["fixed_octaves"],
                               p["coord_scale"], p["height_scale"])

    def tables(self):
        t = np.zeros(256, np.uint8); v = np.zeros((16, 3), np.float32)
        self.L.ref_perlin_tables(_p(t), _p(v))
        return t, v

    def random(self, seed): return self.L.ref_perlin_random(int(seed))

    def noise3(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        out = np.empty(len(xyz), np.float32)
        self.L.ref_noise3_batch(_p(xyz), len(xyz), _p(out))
        return out

    def fractal(self, xyz, kind, lacunarity, gain, octaves):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        out = np.empty(len(xyz), np.float32)
        self.L.ref_fractal_batch(_p(xyz), len(xyz), kind, lacunarity, gain, octaves, _p(out))
        return out

    def get_height_at(self, p, depth, max_depth, params=None):
        self._set(params)
        p = np.ascontiguousarray(p, np.float64)
        return self.L.ref_get_height_at(_p(p), depth, max_depth)

    def generate_height_maps(self, quads, dim, max_depth, params=None, nthreads=1):
        self._set(params)
        quads = np.ascontiguousarray(quads, QUAD_DTYPE)
        out = np.empty((len(quads), dim, dim), np.float32)
        self.L.ref_generate_height_maps(_p(quads), len(quads), dim, max_depth, _p(out), nthreads)
        return out

    def make_root_id(self, r): return self.L.ref_make_root_id(r)
    def make_child_id(self, i, c): return self.L.ref_make_child_id(i, c)
    def get_parent_id(self, i): return self.L.ref_get_parent_id(i)
    def get_root(self, i): return self.L.ref_get_root(i)
    def get_depth(self, i): return self.L.ref_get_depth(i)
    def get_index(self, i): return self.L.ref_get_index(i)
    def get_child_index(self, i): return self.L.ref_get_child_index(i)

    def init_planet(self, radius=RADIUS):
        ml = C.c_int(); sk = C.c_float()
        assert self.L.ref_init_planet(radius, C.byref(ml), C.byref(sk))
        return ml.value, sk.value

    def patch_buffer(self, which):
        self.init_planet()
        n = self.L.ref_patch_buffer(which, None, 0)
        buf = np.zeros(n, np.uint8); self.L.ref_patch_buffer(which, _p(buf), n)
        return buf

    def patch_vertices(self, n=30):
        assert n == 30, "the reference's patch size is a compile-time constant (main.cpp:391)"
        return self.patch_buffer(0).view(np.float32).reshape(-1, 3)

    def patch_indices(self, n=30):
        assert n == 30
        return self.patch_buffer(1).view(np.uint32)

    def root_quads(self, radius=RADIUS):
        q = np.zeros(6, QUAD_DTYPE); assert self.L.ref_root_quads(radius, _p(q)) == 6; return q

    def split_quad(self, quad, radius=RADIUS):
        quad = np.ascontiguousarray(quad, QUAD_DTYPE).reshape(1)
        out = np.zeros(4, QUAD_DTYPE); assert self.L.ref_split_quad(radius, _p(quad), _p(out)) == 4
        return out

    def uniform_quads(self, face, depth, radius=RADIUS):
        out = np.zeros(4 ** depth, QUAD_DTYPE)
        assert self.L.ref_uniform_quads(radius, face, depth, _p(out)) == len(out)
        return out

    def render_frame(self, cam_pos, params=None, radius=RADIUS):
        """One RenderPlanet() with a cold cache: (leaf quads, their height maps, draw uniforms)."""
        self._set(params)
        cam = np.ascontiguousarray(cam_pos, np.float64)
        n = self.L.ref_render_frame(radius, _p(cam))
        quads = np.zeros(n, QUAD_DTYPE); self.L.ref_frame_quads(_p(quads), n)
        return quads, self.captured_height_maps(), self.captured_draws()

    def render_next_frame(self, cam_pos, params=None, radius=RADIUS):
        """The next frame of a sequence (cache, LRU ticks and render_tick carry over).  Returns
        (leaf quads, {texture id: height map} generated this frame, draw uniforms, texture id per draw)."""
        self._set(params)
        cam = np.ascontiguousarray(cam_pos, np.float64)
        n = self.L.ref_render_next_frame(radius, _p(cam))
        quads = np.zeros(n, QUAD_DTYPE); self.L.ref_frame_quads(_p(quads), n)
        maps = self.captured_height_maps()
        ids = [self.L.ref_captured_height_map_id(k) for k in range(len(maps))]
        draws = self.captured_draws()
        tex = np.array([self.L.ref_captured_draw_texture(k) for k in range(len(draws))], np.int64)
        return quads, dict(zip(ids, maps)), draws, tex

    def reset_cache(self):
        self.L.ref_reset_cache()

    def cache_lookup(self, quads, budget=100, radius=RADIUS):
        """GetHeightMapForQuad (main.cpp:191-278) over an arbitrary quad list as one frame.
        Returns (float32[n, 8] = texture name, corners[4], pixel_size[2], generated flag; cache.count)."""
        q = np.ascontiguousarray(quads, QUAD_DTYPE)
        out = np.zeros((len(q), 8), np.float32)
        self.L.ref_cache_lookup.restype = C.c_long
        count = self.L.ref_cache_lookup(C.c_double(radius), _p(q), C.c_long(len(q)), C.c_int(budget), _p(out))
        return out, int(count)

    def run_reference_main(self, scratch_dir):
        """The reference's real main() for one headless frame (its local `Perlin` functor)."""
        rc = self.L.ref_run_reference_main(scratch_dir.encode())
        return rc, self.captured_height_maps(), self.captured_draws()

    def captured_shaders(self):
        """The GLSL texts InitPlanet handed to glShaderSource (vertex stage, fragment stage)."""
        out = []
        for k in range(self.L.ref_captured_shader_count()):
            n = self.L.ref_captured_shader_source(k, None, 0)
            buf = C.create_string_buffer(n + 1)
            self.L.ref_captured_shader_source(k, buf, n + 1)
            out.append(buf.value.decode())
        return out

    def captured_height_maps(self):
        n = self.L.ref_captured_height_map_count()
        maps = []
        for k in range(n):
            w = C.c_int(); h = C.c_int()
            self.L.ref_captured_height_map(k, None, C.byref(w), C.byref(h))
            a = np.empty((h.value, w.value), np.float32)
            self.L.ref_captured_height_map(k, _p(a), None, None)
            maps.append(a)
        return maps

    def captured_draws(self):
        n = self.L.ref_captured_draw_count()
        out = np.zeros((n, 32), np.float32)
        for k in range(n):
            self.L.ref_captured_draw(k, _p(out[k]))
        return out


def best_oracle():
    """The reference itself when its build travelled here, else the port."""
    return RefOracle() if RefOracle.available() else PortOracle()


def fastest_oracle():
    """For timing the CPU side without sandbagging it: the reference at -O3 / AVX2 where that build
    runs, else best_oracle().  Same bits either way (tests/test_oracle_golden.py)."""
    return RefOracle(o3=True) if RefOracle.o3_available() else best_oracle()
