"""torchrun --nproc-per-node N tools/gather_probe.py : what bounds the fused gather?

  all     every rank computes its C3 shard and pushes it to all peers (the bench's step, K2 + K4 only)
  one     only rank 0 computes and pushes (its NVLink egress carries 7 x shard, nothing comes in)
  none    K2 without any peer (compute only)
Also prints whether the device reports multicast (NVLS) support, which decides whether a
multimem.st gather is possible on this box."""
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
quads = pb.tessellate_uniform(DEPTH, first=lo, nquads=nq, params=p)
g = PatchGather(Q, DIM, n_buffers=2, device=dev)
L, C = pb.lib(), pb.C
pp = C.byref(p)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
plain = torch.empty((nq, DIM, DIM), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, n=12):
    ts = []
    for i in range(3 + n):
        dist.barrier(); torch.cuda.synchronize()
        flush.zero_()
        a, b = ev(), ev()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def step_all():
    pb._check(L.planet_gpu_gather_height_maps(g.handle, pp, quads.data_ptr(), nq, lo, DIM, 18, sp))
    pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))


def step_one():
    pb._check(L.planet_gpu_gather_height_maps(g.handle, pp, quads.data_ptr(), nq if rank == 0 else 0, lo, DIM, 18, sp))
    pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))


CHUNKS = int(os.environ.get("PLANET_CE_CHUNKS", "6"))
PER = DIM * DIM * 4


def step_ce():
    """plain K2 in CHUNKS launches, every chunk handed to the copy engines, then wait"""
    pb._check(L.planet_gpu_gather_begin(g.handle, sp))
    base = L.planet_gpu_gather_buffer(g.handle, L.planet_gpu_gather_last_buffer(g.handle))
    for c in range(CHUNKS):
        a, b = nq * c // CHUNKS, nq * (c + 1) // CHUNKS
        pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr() + a * 104, b - a, DIM, 18, base + (lo + a) * PER, sp))
        pb._check(L.planet_gpu_gather_push(g.handle, (lo + a) * PER, (b - a) * PER, sp))
    pb._check(L.planet_gpu_gather_publish(g.handle))
    pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))


def step_ce_only():
    """the copies alone (data already in the buffer): what the copy engines reach all-to-all"""
    pb._check(L.planet_gpu_gather_begin(g.handle, sp))
    pb._check(L.planet_gpu_gather_push(g.handle, lo * PER, nq * PER, sp))
    pb._check(L.planet_gpu_gather_publish(g.handle))
    pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))


nv = pb.patch_vertex_count(30)
pos = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
nrm = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
camv = (C.c_double * 3)(0.0, 0.0, -6371010.0)


def make_step_split(every):
    """K2 + K3 + wait, every `every`-th map pushed by K3 (0: K2 pushes everything)"""
    def fn():
        pb._check(L.planet_gpu_gather_set_shade_share(g.handle, every))
        pb._check(L.planet_gpu_gather_height_maps(g.handle, pp, quads.data_ptr(), nq, lo, DIM, 18, sp))
        pb._check(L.planet_gpu_gather_shade(g.handle, pp, quads.data_ptr(), nq, lo, camv, -1.0, pos.data_ptr(), nrm.data_ptr(), sp))
        pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))
    return fn


def step_none():
    pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), nq, DIM, 18, plain.data_ptr(), sp))


res = {"rank": rank, "world": world, "all_ms": timed(step_all), "one_ms": timed(step_one), "none_ms": timed(step_none),
       "ce_chunks": CHUNKS, "ce_ms": timed(step_ce), "ce_only_ms": timed(step_ce_only)} if os.environ.get("PLANET_PROBE_ALL") else {"rank": rank, "world": world}
for every in (0, 4, 2):
    res[f"k2_k3_wait_share{every}_ms"] = timed(make_step_split(every))
# the link itself: K2 with ONE octave (almost no arithmetic) pushing everything = a pure all-to-all push
p1 = pb.fbm_params(1, 0.5, pb.FAST)
pp1 = C.byref(p1)


def step_push_only(only_rank0=False):
    def fn():
        n = nq if (not only_rank0 or rank == 0) else 0
        pb._check(L.planet_gpu_gather_height_maps(g.handle, pp1, quads.data_ptr(), n, lo, DIM, 18, sp))
        pb._check(L.planet_gpu_gather_wait(g.handle, 1, sp))
    return fn


pb._check(L.planet_gpu_gather_set_shade_share(g.handle, 0))
res["push_only_all_ms"] = timed(step_push_only(False))
res["push_only_rank0_ms"] = timed(step_push_only(True))
shard_mb0 = nq * DIM * DIM * 4 / 1e6
res["push_only_all_GBs"] = shard_mb0 * (world - 1) / res["push_only_all_ms"]
res["push_only_rank0_GBs"] = shard_mb0 * (world - 1) / res["push_only_rank0_ms"]
plain1 = timed(lambda: pb._check(L.planet_gpu_generate_height_maps(pp1, quads.data_ptr(), nq, DIM, 18, plain.data_ptr(), sp)))
res["k2_one_octave_no_push_ms"] = plain1
pb._check(L.planet_gpu_gather_set_shade_share(g.handle, 0))
g.check()
mc = None
try:
    from cuda.bindings import driver as cu
    err, val = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, local)
    mc = int(val)
    err, fab = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, local)
    err, fd = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, local)
    res["fabric_handles"] = int(fab); res["posix_fd_handles"] = int(fd)
except Exception as exc:                                            # noqa: BLE001
    mc = f"{type(exc).__name__}: {exc}"
res["multicast_supported"] = mc
shard_mb = nq * DIM * DIM * 4 / 1e6
if "all_ms" in res:
    res["egress_GBs_all"] = shard_mb * (world - 1) / res["all_ms"]
    res["egress_GBs_ce_only"] = shard_mb * (world - 1) / res["ce_only_ms"]
    res["egress_GBs_one"] = shard_mb * (world - 1) / res["one_ms"] if rank == 0 else None
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    for r in allr:
        print(json.dumps(r))
g.close()
dist.destroy_process_group()
