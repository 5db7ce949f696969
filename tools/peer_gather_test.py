"""torchrun --nproc-per-node N tools/peer_gather_test.py : PeerGather vs NCCL all-gather (bytes + time)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import planet_b200 as pb
from planet_b200.sharding import PeerGather

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); pb.init(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NQ, DIM, CH = 16384, 32, 8
p = pb.fbm_params(8, 0.5, pb.FAST)
quads = pb.tessellate_uniform(7, first=(rank % 6) * NQ, nquads=NQ, params=p)
pg = PeerGather((NQ, DIM, DIM), device=dev)
shard = pg.local_shard()
L, C = pb.lib(), pb.C; pp = C.byref(p)
main = torch.cuda.current_stream()
sp = C.c_void_p(main.cuda_stream)

def step_overlapped():
    n = NQ // CH
    for c in range(CH):
        pb._check(L.planet_gpu_generate_height_maps(pp, quads[c * n:(c + 1) * n].data_ptr(), n, DIM, 18, shard[c * n:(c + 1) * n].data_ptr(), sp))
        pg.push(c * n, (c + 1) * n, main)
    return pg.finish(main)

peer_shards = pg.peer_shards()
def step_fused():
    pb.generate_height_maps_gathered(quads, DIM, 18, shard, peer_shards, p)
    return pg.finish(main)

ref = torch.empty((world * NQ, DIM, DIM), dtype=torch.float32, device=dev)
def step_nccl():
    pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), NQ, DIM, 18, shard.data_ptr(), sp))
    dist.all_gather_into_tensor(ref, shard)
    torch.cuda.synchronize(); dist.barrier()

def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ts))], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

t_n = timeit(step_nccl)
t_o = timeit(step_overlapped)
t_f = timeit(step_fused)
got = step_overlapped().clone(); pg.gathered.zero_(); torch.cuda.synchronize(); dist.barrier()
got_f = step_fused().clone(); step_nccl()
same = torch.equal(got, ref) and torch.equal(got_f, ref)
flags = torch.tensor([int(same)], device=dev); dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "k2_plus_nccl_gather_ms": t_n, "k2_with_overlapped_peer_gather_ms": t_o, "k2_fused_peer_stores_ms": t_f, "bytes_identical_on_all_ranks": bool(flags.item())}))
dist.destroy_process_group()
