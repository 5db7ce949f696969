"""Builds planet_b200/libplanet_gpu.so: the C-ABI library (include/planet_gpu.h) with the
hand-written sm_100a kernels.  nvcc cross-compiles without a GPU; the .so is built in-tree
so it travels to the GPU box with the repository snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libplanet_gpu.so")
SOURCES = ["planet_api.cu", "k0_lod.cu", "k1_tessellate.cu", "k2_heights.cu", "k3_shade.cu", "k4_gather.cu", "height_map_cache.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _newest_source():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "planet_gpu.h"))
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("PLANET_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
