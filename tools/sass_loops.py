"""Instruction counts of the octave loops of the K2 FAST kernel, read from the built library with
cuobjdump (no GPU needed).  K2 is issue-bound, so its run time tracks these numbers; a change that
lengthens the loop shows up here before it shows up on the GPU (tests/test_abi.py keeps a bound).

    python tools/sass_loops.py [path/to/libplanet_gpu.so]
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "planet_b200", "libplanet_gpu.so")
_ADDR = re.compile(r"/\*([0-9a-f]{4,5})\*/")


def octave_loops(lib=LIB, kernel="k_height_maps_fast", variant="Li768ELi32"):
    """{mangled name: [(instruction count, opcode histogram), ...]} for every backward branch spanning
    150-260 instructions (the octave loops; prologue and tile loops are far outside that range)."""
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    result = {}
    for chunk in out.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if kernel not in name or variant not in name:
            continue
        lines = [l for l in chunk.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
        addr = lambda l: int(_ADDR.search(l).group(1), 16)
        loops = []
        for l in lines:
            if "BRA" not in l:
                continue
            m = re.search(r"0x([0-9a-f]+)", l.split("BRA", 1)[1])
            if not m:
                continue
            target, here = int(m.group(1), 16), addr(l)
            if target < here and 150 < (here - target) // 16 < 260:
                hist = collections.Counter()
                for k in lines:
                    if target <= addr(k) <= here:
                        op = _ADDR.sub("", k, count=1).strip()
                        op = re.sub(r"^@!?U?P\d+\s+", "", op)
                        hist[op.split()[0].rstrip(";").split(".")[0]] += 1
                loops.append((sum(hist.values()), dict(hist)))
        result[name] = loops
    return result


if __name__ == "__main__":
    for name, loops in octave_loops(*(sys.argv[1:2] or [LIB])).items():
        print(name)
        for n, hist in loops:
            print(f"  {n:4d}  " + " ".join(f"{k}:{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])))
