"""The reference's vector call surface (vec3.h:25-72, math.h:47-70) as __host__ __device__
functions (planet_b200/csrc/planet_call_surface.cuh), run in a kernel and on the host against
what the reference's own vec3.h produced (tests/golden/vec3_golden.npz, written by
oracle/gen_vec3_golden.py through oracle/_ref).  Bar: bit-exact for everything built from
+ - * / sqrt (Dot, LengthSq, Length, Normalize, SafeNormalize, Cross, the operators) -- which
needs the unfused _rn arithmetic on the device; Slerp goes through acos/sin of the device math
library and is held to 1e-6 of the operands' length."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

COLS = {"Dot": (0, 1), "LengthSq": (1, 2), "Length": (2, 3), "Normalize": (3, 6), "SafeNormalize": (6, 9), "Cross": (9, 12),
        "Slerp": (12, 15), "a+b": (15, 18), "a-b": (18, 21), "a*t": (21, 24), "t*a": (24, 27), "a/t": (27, 30), "-a": (30, 33)}


@pytest.fixture(scope="module")
def vec3_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "vec3_golden.npz"))


@pytest.fixture(scope="module")
def program(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("vec3") / "call_surface_device_test")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "--expt-relaxed-constexpr", "-o", exe, os.path.join(ROOT, "tests", "call_surface_device_test.cu")])
    return exe


def run(program, g, *args):
    n = len(g["t"])
    blob = np.int64(n).tobytes() + g["a"].tobytes() + g["b"].tobytes() + g["t"].tobytes()
    out = subprocess.run([program, *args], input=blob, capture_output=True, check=True).stdout
    o64 = np.frombuffer(out, np.float64, n * 33).reshape(n, 33)
    o32 = np.frombuffer(out, np.float32, n * 33, n * 33 * 8).reshape(n, 33)
    return o64, o32


def check(o64, o32, g, slerp_exact):
    scale = np.maximum(np.linalg.norm(g["a"], axis=1), np.linalg.norm(g["b"], axis=1))[:, None]
    for name, (lo, hi) in COLS.items():
        for got, want in ((o64, g["out64"]), (o32, g["out32"])):
            a, b = got[:, lo:hi], want[:, lo:hi]
            if name == "Slerp" and not slerp_exact:
                # float operands a hair apart give Dot/(|a||b|) >= 1: acos is 0 or NaN and vec3.h:70-71 divides
                # by sin(0) -- the reference's own hazard, reproduced: the same rows are non-finite
                ok = np.isfinite(b).all(axis=1)
                assert (np.isfinite(a).all(axis=1) == ok).all() and ok.sum() > 0.7 * len(ok), name
                assert (np.abs(a[ok].astype(np.float64) - b[ok]) <= 1e-6 * scale[ok]).all(), name
            else:
                assert a.tobytes() == b.tobytes(), f"{name} ({got.dtype}) is not bit-identical to vec3.h"


def test_vector_call_surface_host_half_is_bit_identical_to_vec3_h(program, vec3_golden):
    o64, o32 = run(program, vec3_golden, "host")
    check(o64, o32, vec3_golden, slerp_exact=True)           # same libm as the reference: Slerp too


@pytest.mark.gpu
def test_vector_call_surface_in_a_kernel_is_bit_identical_to_vec3_h(program, vec3_golden, gpu):
    o64, o32 = run(program, vec3_golden)
    check(o64, o32, vec3_golden, slerp_exact=False)
    # the small-angle rows are neighbouring quad corners: Slerp must stay on the chord's arc there
    s = o64[6:300, 12:15]
    la, lb = np.linalg.norm(vec3_golden["a"][6:300], axis=1), np.linalg.norm(vec3_golden["b"][6:300], axis=1)
    ls = np.linalg.norm(s, axis=1)
    assert (ls >= np.minimum(la, lb) * (1 - 1e-5)).all() and (ls <= np.maximum(la, lb) * (1 + 1e-5)).all()
