"""C2-sized K2 batch in EXACT arithmetic (the bit-identical mode) -- the command ncu wraps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import planet_b200 as pb

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pb.init(0)
p = pb.fbm_params(octaves=8, gain=0.5, precision=pb.EXACT)
quads = pb.tessellate_uniform(7, first=0, nquads=16384, params=p)
for _ in range(reps):
    h = pb.generate_height_maps(quads, 32, 18, p)
torch.cuda.synchronize()
print("ok", float(h.abs().max()))
