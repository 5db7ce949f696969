"""One C2-sized pass of K1/K2/K3 a few times -- the command ncu wraps (profiles/ recipes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import planet_b200 as pb

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pb.init(0)
p = pb.fbm_params(octaves=8, gain=0.5, precision=pb.FAST)
for _ in range(reps):
    quads, idx = pb.tessellate_uniform(7, first=0, nquads=16384, params=p, with_indices=True)
    h = pb.generate_height_maps(quads, 32, 18, p)
    pos, nrm = pb.shade(quads, h, (0.0, 0.0, -6371010.0), p)
torch.cuda.synchronize()
print("ok", float(h.abs().max()), float(nrm[..., 3].mean()))
