"""BASELINE.json configs[3]: the full planet at ~1 B vertices (depth 8, patch 50 -> 52 x 52 maps,
fBm 12 octaves) sharded by patch range over the ranks of one box.
PLANET_DEPTH=7 PLANET_PATCH=30 PLANET_OCTAVES=8 gives configs[2] (100 M vertices) as a fixed total
split over the ranks, i.e. its strong-scaling form.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c4_run.py

Every rank tessellates, generates and shades its own contiguous range of the 393 216 leaf quads
(no data-path collective); rank 0 prints one JSON line with the max-over-ranks device time."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import planet_b200 as pb
from planet_b200.sharding import shard_range

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); pb.init(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
DEPTH, PATCH, OCT = (int(os.environ.get(k, d)) for k, d in (("PLANET_DEPTH", 8), ("PLANET_PATCH", 50), ("PLANET_OCTAVES", 12)))
DIM = PATCH + 2
total_quads = 6 * 4 ** DEPTH
lo, hi = shard_range(total_quads, rank, world)
nq = hi - lo
p = pb.fbm_params(OCT, 0.5, pb.FAST, patch_verts=PATCH)
nv, ni = pb.patch_vertex_count(PATCH), pb.patch_index_count(PATCH)
quads = torch.empty((nq, 13), dtype=torch.int64, device=dev)
idx = torch.empty(nq * ni, dtype=torch.int32, device=dev)
heights = torch.empty((nq, DIM, DIM), dtype=torch.float32, device=dev)
pos = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev); nrm = torch.empty_like(pos)
L, C = pb.lib(), pb.C
pp, sp = C.byref(p), C.c_void_p(torch.cuda.current_stream().cuda_stream)
cam = (C.c_double * 3)(0.0, 0.0, -6371010.0)

def step():
    pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, lo, nq, quads.data_ptr(), idx.data_ptr(), sp))
    pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), nq, DIM, 18, heights.data_ptr(), sp))
    pb._check(L.planet_gpu_shade(pp, quads.data_ptr(), nq, cam, heights.data_ptr(), -1.0, pos.data_ptr(), nrm.data_ptr(), sp))

for _ in range(3): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.mean(ts)), torch.cuda.max_memory_allocated() / 2 ** 30], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
finite = bool(torch.isfinite(heights).all()) and bool(torch.isfinite(nrm).all())
if rank == 0:
    verts = total_quads * DIM * DIM
    print(json.dumps({"config": f"full planet depth {DEPTH}, patch {PATCH} ({DIM}x{DIM} maps), fBm {OCT} octaves", "n_gpus": world,
                      "quads": total_quads, "vertices": verts, "ms_per_step_max_over_ranks": float(t[0]),
                      "gvert_s": verts / float(t[0]) / 1e6, "hbm_gib_per_gpu": float(t[1]),
                      "bytes_per_vertex_resident": 4 + 32 * nv / (DIM * DIM) + 4 * ni / (DIM * DIM), "finite": finite}))
if world > 1: dist.destroy_process_group()
