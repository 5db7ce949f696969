"""The drop-in boundary without a GPU: the library builds, loads, exports every symbol the
header declares, validates arguments, and refuses to compute when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import planet_b200 as pb
from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "planet_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(planet_gpu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = pb.lib()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/planet_gpu.h but not exported"
    assert sorted(pb.EXPORTED_SYMBOLS) == names


def test_header_is_plain_c_and_cxx():
    """include/planet_gpu.h is the boundary a reference-side binding (C, C++, cgo, ctypes ...) compiles against:
    plain C99, no CUDA, torch or C++ types in any signature."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "planet_gpu.h")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    import re
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)       # comments may name what the signatures avoid
    for banned in ("cudaStream_t", "torch", "std::", "#include <cuda", "at::", "c10::"):
        assert banned not in code, banned


def test_abi_version_and_struct_layout():
    assert pb.lib().planet_gpu_abi_version() == 1
    assert C.sizeof(pb.Params) == 72                       # planet_gpu_params
    assert pb.QUAD_DTYPE.itemsize == 104                   # main.cpp:68-72


def test_default_params_are_the_reference_constants():
    p = pb.default_params()
    assert (p.radius, p.patch_verts, p.noise_kind, p.lacunarity) == (6371000.0, 30, pb.RIDGED, 2.0)
    assert p.gain == np.float32(0.55) and p.fixed_octaves == 0
    assert p.coord_scale == 0.00001 and p.height_scale == 8848.0 and p.precision == pb.EXACT
    assert list(p.seed_offset) == [0.0, 0.0, 0.0]


def test_host_scalars_match_reference(golden):
    assert pb.max_lod() == int(golden["max_lod"])
    assert np.float32(pb.max_skirt_size()) == golden["max_skirt_size"]
    assert pb.patch_vertex_count(30) == 1020 and pb.patch_index_count(30) == 2036


def test_strip_index_closed_form_matches_reference_buffer(golden, port):
    L = pb.lib()
    want = golden["patch_index_buffer"].view(np.uint32)
    got = np.array([L.planet_gpu_strip_index(k, 30) for k in range(2036)], np.uint32)
    assert (got == want).all()
    for n in (2, 3, 4, 5, 7, 31, 50, 64):                  # closed form for N != 30 too
        got = np.array([L.planet_gpu_strip_index(k, n) for k in range(pb.patch_index_count(n))], np.uint32)
        assert (got == port.patch_indices(n)).all(), n


def test_uniform_leaf_ids_follow_recursion_order(port):
    L = pb.lib()
    for depth in (0, 1, 2, 4):
        want = np.concatenate([port.uniform_quads(f, depth)["id"] for f in range(6)])
        got = np.array([L.planet_gpu_uniform_leaf_id(k, depth) for k in range(len(want))], np.uint64)
        assert (got == want).all(), depth


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert pb.lib().planet_gpu_init(0) == -1                # PLANET_E_NO_DEVICE
    assert b"no CPU path" in pb.lib().planet_gpu_last_error()
    with pytest.raises(pb.PlanetGpuError):
        pb.generate_height_maps_host(np.zeros(1, pb.QUAD_DTYPE), 32, 18)
    # the reference-shaped call cannot return a status: it logs and fills NaN
    out = pb.generate_height_map(np.zeros(1, pb.QUAD_DTYPE), 8, 18)
    assert np.isnan(out).all()
    assert np.isnan(pb.get_height_at([1.0, 2.0, 3.0], 0, 1))
    with pytest.raises(pb.PlanetGpuError):
        pb.tessellate_uniform(2)


def test_product_never_touches_the_oracle():
    """planet_b200/ and include/ must not import, link or load anything under oracle/."""
    bad = []
    for base in ("planet_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle|libplanet_ref|planet_oracle", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_vector_call_surface_on_the_host(tmp_path):
    """vec3.h / math.h names as __host__ __device__ functions (planet_call_surface.cuh), host half."""
    import subprocess
    exe = str(tmp_path / "call_surface_host_test")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "call_surface_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert "call surface ok" in out


def test_k2_octave_loop_has_not_grown():
    """K2 is bound by the registers its octave loop writes, so its run time tracks the instruction
    count of that loop.  The count is read from the built library with cuobjdump (no GPU): 182
    instructions per two samples for fBm, 191 for ridged, 188 / 199 for the mixed-octave-count
    variants of the ragged-tile path when this bound was set (first float-table version: 192 /
    201 / 198 / 206).  A change that adds registers to the kernel (e.g. live pointers across the
    loop) or defeats the absolute / uniform-base addressing of the table reads shows up here
    first."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sass_loops import octave_loops
    found = octave_loops()
    assert len(found) == 4, sorted(found)                    # (plain, fused gather) x (fBm, ridged)
    for name, loops in found.items():
        fbm = "ELi1EEEv" in name                             # KIND = PLANET_NOISE_FBM
        octave = sorted(n for n, _ in loops if n < 215)      # the run-of-tiles loop also lands in the window
        assert len(octave) >= 3, (name, octave)
        assert octave[0] <= (184 if fbm else 193), (name, octave)
        assert octave[-1] <= (190 if fbm else 201), (name, octave)


def _sass_by_function():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "planet_b200", "libplanet_gpu.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    code = {c.split("\n", 1)[0].strip(): c for c in sass.split("Function : ")[1:]}
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        usage[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)))
    return code, usage


def test_kernels_that_must_use_the_tma_unit_do_and_the_plain_k2_does_not():
    """Read from the built library, no GPU: the index stream of K1, the fused gather of K2, K3's share of it
    and the TMA pusher move their bytes as bulk copies (SASS UBLKCP); the single-GPU K2 carries none of it."""
    code, _ = _sass_by_function()
    def one(*parts):
        names = [n for n in code if all(p in n for p in parts)]
        assert len(names) == 1, (parts, names)
        return code[names[0]]
    assert "UBLKCP" in one("k_tessellate_bulk")
    assert "UBLKCP" in one("k_height_maps_fast", "Li768ELi32ELb1ELi1E")          # <768, 32, gather, fBm>
    assert "UBLKCP" not in one("k_height_maps_fast", "Li768ELi32ELb0ELi1E")      # <768, 32, no gather, fBm>
    assert "UBLKCP" in one("k_shade", "Lb1ELb0ELb1E")                            # <STAGE, !RECT, PUSH>
    assert "UBLKCP" not in one("k_shade", "Lb1ELb0ELb0E")
    assert "UBLKCP" in one("k_push_progress_tma")


def test_kernels_meant_to_share_an_sm_with_k2_fit_beside_it():
    """A resident K2 CTA (768 threads x 80 registers, ~217 KB of shared memory) leaves an SM 4 096 registers and
    ~10 KB: the pusher kernels and the slim index stream are built to fit there (128 threads x <= 32 registers,
    <= 10 KB incl. the 1 KB the driver reserves per CTA), and their polling loads carry no fence."""
    code, usage = _sass_by_function()
    k2 = [u for n, u in usage.items() if "k_height_maps_fast" in n and "Li768ELi32ELb0ELi1E" in n]
    assert len(k2) == 1 and k2[0]["reg"] <= 80
    for key in ("k_push_progress_tma", "k_push_progressE", "k_push_range", "k_index_stream_slim"):
        hits = [(n, u) for n, u in usage.items() if key in n]
        assert len(hits) == 1, (key, [n for n, _ in hits])
        name, u = hits[0]
        assert u["reg"] <= 32 and u["shared"] <= 10 * 1024, (name, u)
        assert 768 * 80 + 128 * max(u["reg"], 24) <= 65536
    # the progress counters are polled with a relaxed load; the acquire fence (MEMBAR.GPU + CCTL.IVALL, an L1
    # invalidation K2 would feel) is paid once per observed advance, not once per poll
    for key, fences in (("k_push_progress_tma", 3), ("k_push_progressE", 1)):
        name = [n for n in code if key in n][0]
        lines = code[name].splitlines()
        polls = [i for i, l in enumerate(lines) if "LDG.E.64.STRONG.GPU" in l]
        assert len(polls) == 1, name
        assert not any("CCTL" in l or "MEMBAR" in l for l in lines[polls[0]:polls[0] + 3]), name
        assert sum("CCTL.IVALL" in l for l in lines) <= fences, name
