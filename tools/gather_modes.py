"""torchrun --nproc-per-node N tools/gather_modes.py : the bench's N>1 step (K1 quads, K2 + K4, K3, wait)
timed under each way of driving K4, in one process group:
  fused<k>        K2 pushes its own tiles as bulk copies, every k-th map left to K3 (0: none)
  concurrent_*    one unfused K2 launch publishing per-warp progress + the pusher kernel beside it
                  (tma: bulk copies through a staging slot; lsu: 16-byte loads/stores; every<n>: K2 publishes every n tiles)
Prints one JSON line per rank."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import planet_b200 as pb
from planet_b200.sharding import PatchGather, shard_range

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
pb.init(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
DEPTH, DIM, Q = 7, 32, 6 * 4 ** 7
lo, hi = shard_range(Q, rank, world)
nq = hi - lo
p = pb.fbm_params(8, 0.5, pb.FAST)
quads = torch.empty((nq, 13), dtype=torch.int64, device=dev)
g = PatchGather(Q, DIM, n_buffers=2, device=dev)
L, C = pb.lib(), pb.C
pp = C.byref(p)
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
side = torch.cuda.Stream(device=dev)
side_p = C.c_void_p(side.cuda_stream)
nv, ni = pb.patch_vertex_count(30), pb.patch_index_count(30)
pos = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
nrm = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
indices = torch.empty(nq * ni, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
camv = (C.c_double * 3)(0.0, 0.0, -6371010.0)
ev = lambda: torch.cuda.Event(enable_timing=True)


def step():
    fork = torch.cuda.Event(); fork.record(stream)
    pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, lo, nq, quads.data_ptr(), None, sp))
    pb._check(L.planet_gpu_gather_height_maps(g.handle, pp, quads.data_ptr(), nq, lo, DIM, 18, sp))
    side.wait_event(fork)
    pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, lo, nq, None, indices.data_ptr(), side_p))
    join = torch.cuda.Event(); join.record(side)
    pb._check(L.planet_gpu_gather_shade(g.handle, pp, quads.data_ptr(), nq, lo, camv, -1.0, pos.data_ptr(), nrm.data_ptr(), sp))
    stream.wait_event(join)
    pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))


def timed(n=12):
    ts = []
    for i in range(3 + n):
        dist.barrier(); torch.cuda.synchronize()
        flush.zero_()
        a, b = ev(), ev()
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


KNOBS = ("PLANET_PUSH_LSU", "PLANET_PUSH_EVERY", "PLANET_PUSH_CTAS")
MODES = [("fused0", 0, 0, {}), ("fused2", 0, 2, {}), ("fused_3of8_to_k3", 0, -3, {}), ("fused_5of8_to_k3", 0, -5, {}), ("fused_6of8_to_k3", 0, -6, {}),
         ("concurrent_tma", 2, 0, {}), ("concurrent_lsu", 2, 0, {"PLANET_PUSH_LSU": "1"})]
if os.environ.get("PLANET_MODES_ALL"):
    MODES += [("concurrent_tma_every2", 2, 0, {"PLANET_PUSH_EVERY": "2"}), ("concurrent_tma_every4", 2, 0, {"PLANET_PUSH_EVERY": "4"}),
              ("concurrent_lsu_every4", 2, 0, {"PLANET_PUSH_LSU": "1", "PLANET_PUSH_EVERY": "4"}),
              ("concurrent_tma_74ctas", 2, 0, {"PLANET_PUSH_CTAS": "74"})]
res = {"rank": rank, "world": world, "quads_per_gpu": nq}
for name, mode, share, env in MODES:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    pb._check(L.planet_gpu_gather_set_push_mode(g.handle, mode))
    pb._check(L.planet_gpu_gather_set_shade_share(g.handle, share))
    res[name + "_ms"] = timed()
    g.check()
# the bytes of the last mode against the plain collective
last = g.gathered().clone()
pb._check(L.planet_gpu_gather_set_push_mode(g.handle, 0))
pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), nq, DIM, 18, g.local(lo, nq, which=0).data_ptr(), sp))
g.nccl([(a, b - a) for a, b in (shard_range(Q, r, world) for r in range(world))], which=0)
torch.cuda.synchronize()
res["last_mode_identical_to_nccl"] = bool(torch.equal(last, g.gathered(which=0)))
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    for r in allr:
        print(json.dumps(r))
g.close()
dist.destroy_process_group()
