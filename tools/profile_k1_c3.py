"""K1 at C3 size (98 304 quads: 10 MB of quads + 801 MB of merged strip indices, six times the L2) -- the command ncu wraps
to put dram__bytes_write beside bench.py's roofline_k1_c3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import planet_b200 as pb

pb.init(0)
p = pb.fbm_params(8, 0.5, pb.FAST)
for _ in range(3):
    quads, idx = pb.tessellate_uniform(7, first=0, nquads=98304, params=p, with_indices=True)
torch.cuda.synchronize()
print("ok", int(idx[-1]))
