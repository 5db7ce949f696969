"""Height-map batches that walk every K2 code path once (regular runs, ragged tiles, odd and tiny
dims, compact and replicated tables, FAST and EXACT, fBm and ridged) -- the command
compute-sanitizer wraps:   compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import planet_b200 as pb

pb.init(0)
cases = []
for prec in (pb.FAST, pb.EXACT):
    for mk in (lambda **k: pb.fbm_params(octaves=8, gain=0.5, **k), lambda **k: pb.default_params(**k)):
        p = mk(precision=prec)
        for depth, first, nq, dim in ((5, 0, 1024, 32),      # 1 M samples: compact-table path, regular runs
                                      (6, 7, 2051, 32),      # 2.1 M: replicated tables, run boundaries mid-warp
                                      (4, 3, 97, 5),         # tiny maps: several quads per warp tile
                                      (5, 0, 1100, 33),      # odd dim: second texel on the next row / quad
                                      (5, 1, 450, 52),       # C4's patch: tiles straddle quads
                                      (3, 0, 1, 1024)):      # one large map
            quads = pb.tessellate_uniform(depth, first=first, nquads=nq, params=p)
            h = pb.generate_height_maps(quads, dim, 18, p)
            cases.append((prec, p.noise_kind, dim, nq, float(h.double().sum())))
torch.cuda.synchronize()
for c in cases:
    print(c)
