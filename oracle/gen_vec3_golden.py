"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/vec3_golden.npz: the reference's own vec3.h
functions (vec3.h:25-72, both instantiations), run through oracle/_ref/libplanet_ref.so
(ref_vec3_ops_f32 / _f64 in oracle/ref_oracle.cpp) on seeded inputs.

    python oracle/gen_vec3_golden.py        (build container only: needs oracle/_ref)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import REF_SO  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "vec3_golden.npz")


def main():
    L = C.CDLL(REF_SO)
    rng = np.random.default_rng(20261019)
    n = 1024
    a = rng.normal(size=(n, 3)) * np.exp(rng.uniform(-3, 16, (n, 1)))          # lengths 0.05 .. 9e6 (planet radius scale)
    b = rng.normal(size=(n, 3)) * np.exp(rng.uniform(-3, 16, (n, 1)))
    a[:6] = [[1, -2, 3.5], [6371000, 0, 0], [0, 0, -6371010], [1e-3, 0, 0], [3678.3, -3678.3, 3678.3], [0.1, 0.2, 0.3]]
    b[:6] = [[-0.25, 4, 2], [0, 6371000, 0], [20000, 0, -6371010], [0, 1e-3, 0], [3678.3, 3678.3, -3678.3], [0.3, 0.2, 0.1]]
    b[6:300] = a[6:300] + rng.normal(size=(294, 3)) * np.linalg.norm(a[6:300], axis=1, keepdims=True) * 1e-3   # small angles: neighbouring quad corners
    t = rng.uniform(0, 1, n)
    t[:4] = [0.5, 0.25, 1.0 / 31, 30.0 / 31]
    out64 = np.zeros((n, 33), np.float64)
    L.ref_vec3_ops_f64(a.ctypes, b.ctypes, t.ctypes, C.c_long(n), out64.ctypes)
    a32, b32, t32 = a.astype(np.float32), b.astype(np.float32), t.astype(np.float32)
    out32 = np.zeros((n, 33), np.float32)
    L.ref_vec3_ops_f32(a32.ctypes, b32.ctypes, t32.ctypes, C.c_long(n), out32.ctypes)
    np.savez_compressed(OUT, a=a, b=b, t=t, out64=out64, out32=out32)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
