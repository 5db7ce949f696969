"""K4 behind the C-ABI (planet_gpu_gather_*, csrc/k4_gather.cu), driven by a C++ multi-process
program (tests/gather_driver.cpp: one process per GPU, NCCL unique id handed over through a file,
no Python or torch on the path).  On a one-GPU box only the world-1 object runs; with two or more
GPUs the fused K2/K4 kernel, the release-flag rotation over two buffers, ragged shards, EXACT
arithmetic and the plain NCCL collective are each compared byte for byte with the unsharded
buffer computed on the same GPU."""
import ctypes as C
import os
import signal
import subprocess

import pytest

import planet_b200 as pb
from conftest import ROOT


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    pb.lib()
    exe = str(tmp_path_factory.mktemp("gather") / "gather_driver")
    lib_dir = os.path.join(ROOT, "planet_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "gather_driver.cpp"),
                           "-I/usr/local/cuda/include", "-L" + lib_dir, "-lplanet_gpu", "-L/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath," + lib_dir])
    return exe


def run_driver(exe, world, depth, mode, timeout=240):
    """Runs the driver in its own process group and kills the whole group on a timeout (a rank that
    dies early must not leave its peers spinning on the box)."""
    p = subprocess.Popen([exe, str(world), str(depth), mode], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        pytest.fail(f"gather_driver {world} {depth} {mode} timed out\n{out}\n{err[-2000:]}")
    return p.returncode, out, err


def test_gather_object_validates_and_fails_loudly_without_gpu(driver):
    import torch
    L = pb.lib()
    if torch.cuda.is_available():
        assert not L.planet_gpu_gather_create(None, 0, 9, 1024, 2) and b"world in [1, 8]" in L.planet_gpu_last_error()
        assert not L.planet_gpu_gather_create(None, 0, 2, 1024, 2)          # world > 1 needs the unique id
        assert not L.planet_gpu_gather_create(None, 0, 1, 1024, 3)          # one or two buffers
        return
    assert not L.planet_gpu_gather_create(None, 0, 1, 1024, 2)
    assert b"no CPU path" in L.planet_gpu_last_error()
    rc, out, err = run_driver(driver, 1, 2, "fast")
    assert rc == 10 and "planet_gpu_init" in err and out == ""


def test_unique_id_comes_from_nccl():
    buf = C.create_string_buffer(128)
    rc = pb.lib().planet_gpu_gather_unique_id(buf)
    if rc == -4:                                                            # PLANET_E_UNSUPPORTED: no libnccl.so.2 on this box
        pytest.skip(pb.lib().planet_gpu_last_error().decode())
    assert rc == 0 and any(buf.raw)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "exact", "nccl", "ce", "smpush", "split", "concurrent"])
def test_world_of_one_is_the_plain_path(driver, gpu, mode):
    rc, out, err = run_driver(driver, 1, 5, mode)
    assert rc == 0, err[-2000:]
    assert "gather ok rank 0/1" in out


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "ragged", "exact", "nccl", "ce", "smpush", "split", "split5of8", "concurrent", "concurrent_ragged"])
def test_two_ranks_gather_the_unsharded_bytes(driver, gpu, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 8)
    rc, out, err = run_driver(driver, world, 5, mode)
    assert rc == 0, err[-2000:]
    for r in range(world):
        assert f"gather ok rank {r}/{world}" in out
