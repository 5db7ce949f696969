"""The "switch unchanged" claim, executed.  oracle/_ref/ref_gpu_caller is the reference's own
main.cpp (#include'd unmodified by oracle/ref_oracle.cpp, -DPLANET_REAL_CALLER) built as a program:
InitPlanet (main.cpp:280, :495) receives the GPU HeightMapGenerator of planet_b200/host/planet_host.h
instead of CreateHeightMapGenerator<Perlin>() (main.cpp:843), then RenderPlanet (main.cpp:600) runs
the reference's ProcessQuad / GetHeightMapForQuad, whose calls through the two function pointers
(main.cpp:244, :552, :555) land in libplanet_gpu.so.  What the reference hands to glTexImage2D
(render.cpp:426) must equal what its CPU generator produced in the reference's real main().

The program is built from /root/reference by oracle/Makefile (build() does it) and travels to the
GPU box prebuilt; nothing here reads /root/reference at run time."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, as_bits, quads_from_bytes

EXE = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_caller")


def parse_frames(blob, frames):
    """int64 leaf count, the leaf Quads (104 B), int64 map count, the 32 x 32 maps -- per frame."""
    out, at = [], 0
    for _ in range(frames):
        n = int(np.frombuffer(blob, np.int64, 1, at)[0]); at += 8
        quads = np.frombuffer(blob, np.uint8, n * 104, at).reshape(n, 104); at += n * 104
        m = int(np.frombuffer(blob, np.int64, 1, at)[0]); at += 8
        maps = np.frombuffer(blob, np.float32, m * 1024, at).reshape(m, 32, 32); at += m * 4096
        out.append((quads, maps))
    assert at == len(blob)
    return out


def need_exe():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/ref_gpu_caller not built (no /root/reference here)")


def test_reference_caller_with_its_own_cpu_generator_reproduces_the_golden_frame(golden):
    """The harness itself, on CPU: with main.cpp:843's generator left in place the program emits
    the frame the reference's real main() produced (so any difference in GPU mode is the library's)."""
    need_exe()
    r = subprocess.run([EXE, "1", "cpu"], capture_output=True, check=True)
    (quads, maps), = parse_frames(r.stdout, 1)
    assert quads.tobytes() == np.ascontiguousarray(golden["frame_quads"]).tobytes()
    assert (as_bits(maps) == as_bits(golden["frame_height_maps"])).all()


def test_reference_caller_fails_loudly_without_a_gpu():
    need_exe()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "1"], capture_output=True)
    assert r.returncode == 3 and b"no CPU path" in r.stderr and r.stdout == b""


@pytest.mark.gpu
def test_reference_init_and_render_planet_on_the_gpu_generator(golden, gpu):
    """InitPlanet(planet, R, <GPU generator>) + one RenderPlanet at the default camera: the same
    117 leaves (every ProcessQuad split decision went through planet_gpu_get_height_at) and the 117
    glTexImage2D-captured maps bit-identical to golden["frame_height_maps"]."""
    need_exe()
    r = subprocess.run([EXE, "1"], capture_output=True, check=True)
    (quads, maps), = parse_frames(r.stdout, 1)
    assert len(quads) == 117 and len(maps) == 117
    assert quads.tobytes() == np.ascontiguousarray(golden["frame_quads"]).tobytes()
    assert (as_bits(maps) == as_bits(golden["frame_height_maps"])).all()
    q = quads_from_bytes(quads)
    assert int(((q["id"] >> np.uint64(55)) & np.uint64(31)).max()) == 10       # depths 0..10, as in the reference's frame


@pytest.mark.gpu
def test_reference_flight_of_eight_frames_gpu_generator_vs_cpu_generator(gpu):
    """Eight frames, 20 km apart, through the reference's own height-map cache (hits, the 100-per-frame
    generation budget, parent fallback): leaves and uploaded maps of every frame are byte-identical
    whichever generator InitPlanet was given."""
    need_exe()
    a = subprocess.run([EXE, "8"], capture_output=True, check=True).stdout
    b = subprocess.run([EXE, "8", "cpu"], capture_output=True, check=True).stdout
    fa, fb = parse_frames(a, 8), parse_frames(b, 8)
    assert sum(len(m) for _, m in fb[1:]) > 0, "the flight should generate new maps after frame 0"
    for (qa, ma), (qb, mb) in zip(fa, fb):
        assert qa.tobytes() == qb.tobytes()
        assert (as_bits(ma) == as_bits(mb)).all()
    assert a == b


@pytest.mark.gpu
def test_reference_cache_loop_behind_gpu_lod_selection(gpu):
    """Both halves of INTEGRATION.md's minimum edit: the GPU generator installed through InitPlanet AND
    RenderPlanet's ProcessQuad recursion (main.cpp:604-624) replaced by one planet_gpu_select_lod, the
    reference's own GetHeightMapForQuad loop (main.cpp:652-660) running behind it.  Leaves and uploaded maps
    of an 8-frame flight are byte-identical to the untouched reference with its CPU generator."""
    need_exe()
    a = subprocess.run([EXE, "8", "lod"], capture_output=True, check=True).stdout
    b = subprocess.run([EXE, "8", "cpu"], capture_output=True, check=True).stdout
    for k, ((qa, ma), (qb, mb)) in enumerate(zip(parse_frames(a, 8), parse_frames(b, 8))):
        assert qa.tobytes() == qb.tobytes(), k
        assert (as_bits(ma) == as_bits(mb)).all(), k
    assert a == b
