"""The fused K2 + K4 kernel under ncu: ONE process, two GPUs of the box, GPU 0 computes a C2-sized batch and
pushes every finished tile to its own buffer and to a buffer on GPU 1 (peer access), which is what a rank
does for each of its peers.  ncu cannot wrap a multi-rank command; this is the same kernel and the same
stores.  usage: python tools/profile_gather_one_process.py [peers]   (peers <= visible GPUs - 1)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import planet_b200 as pb

npeers = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(0)
pb.init(0)
p = pb.fbm_params(8, 0.5, pb.FAST)
nq = 16384
quads = pb.tessellate_uniform(7, 0, nq, p)
out = torch.empty((nq, 32, 32), dtype=torch.float32, device="cuda:0")
peers = [torch.zeros((nq, 32, 32), dtype=torch.float32, device=f"cuda:{1 + k % (torch.cuda.device_count() - 1)}") for k in range(npeers)]
for t in peers:
    t.copy_(out)                                       # makes torch enable peer access 0 -> k
torch.cuda.synchronize()
for _ in range(2):
    pb.generate_height_maps_gathered(quads, 32, 18, out, peers, p)
torch.cuda.synchronize()
want = pb.generate_height_maps(quads, 32, 18, p)
ok = all(bool(torch.equal(t.to("cuda:0"), want)) for t in peers) and bool(torch.equal(out, want))
print("fused gather, one process:", npeers, "peer(s), bytes identical:", ok)
sys.exit(0 if ok else 1)
