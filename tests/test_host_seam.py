"""The C++ host mirror (planet_b200/host/planet_host.h): a reference-layout HeightMapGenerator
filled with the GPU entry points, driven exactly like the reference's two call sites."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, as_bits, quads_from_bytes


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    import planet_b200 as pb
    pb.lib()                                                   # make sure the .so is built
    exe = str(tmp_path_factory.mktemp("host") / "host_seam_driver")
    lib_dir = os.path.join(ROOT, "planet_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe,
                           os.path.join(ROOT, "tests", "host_seam_driver.cpp"),
                           "-L" + lib_dir, "-lplanet_gpu", "-Wl,-rpath," + lib_dir])
    return exe


def test_host_mirror_compiles_and_fails_loudly_without_gpu(driver, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([driver, os.devnull, "18"], capture_output=True)
    assert r.returncode == 3 and b"[ERROR] planet_gpu_init" in r.stderr and r.stdout == b""


@pytest.mark.gpu
def test_reference_call_sites_through_function_pointers(driver, golden, tmp_path, gpu):
    quads = quads_from_bytes(golden["frame_quads"])[[0, 3, 20, 64, 116]]
    path = tmp_path / "quads.bin"
    path.write_bytes(quads.tobytes())
    r = subprocess.run([driver, str(path), str(int(golden["max_lod"]))], capture_output=True, check=True)
    out = np.frombuffer(r.stdout, np.float32).reshape(len(quads), 32 * 32 + 4)
    want = golden["frame_height_maps"][[0, 3, 20, 64, 116]].reshape(len(quads), -1)
    assert (as_bits(out[:, :1024]) == as_bits(want)).all()     # GenerateHeightMap, bit-exact
    from oracle.bindings import PortOracle, height_params
    orc = PortOracle()
    corner = np.array([[orc.get_height_at(p, 0, 1, height_params()) for p in q["p"]] for q in quads], np.float32)
    assert (as_bits(out[:, 1024:]) == as_bits(corner)).all()   # GetHeightAt(p, 0, 1), bit-exact


@pytest.mark.gpu
def test_headless_frame_program_matches_the_reference_frame_counts(tmp_path, gpu):
    """planet_b200/host/planet_frame.cpp: InitPlanet + RenderPlanet on the GPU, in C++."""
    exe = str(tmp_path / "planet_frame")
    host = os.path.join(ROOT, "planet_b200", "host")
    lib_dir = os.path.join(ROOT, "planet_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(host, "planet_frame.cpp"),
                           "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           "-L" + lib_dir, "-lplanet_gpu", "-L/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    out = subprocess.run([exe, "3"], capture_output=True, text=True, check=True).stdout
    assert "max_lod: 18" in out and "patch: 1020 verts / 2036 indices" in out
    lines = [l for l in out.splitlines() if l.startswith("tris:")]
    # the reference's first frame: 117 leaf quads, 117 * 29 * 29 * 2 triangles, all generated
    assert lines[0].startswith("tris: 196794, quads: 117, generated: 117, parent fallback: 0, cached: 117")
    assert len(lines) == 3 and "vertex 0: pos" in out
