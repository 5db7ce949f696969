"""Does K3 (store-bound) hide behind K2 (issue-bound) when the batch is chunked over two streams?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import planet_b200 as pb
pb.init(0)
p = pb.fbm_params(8, 0.5, pb.FAST)
NQ = 16384
quads = torch.empty((NQ, 13), dtype=torch.int64, device="cuda")
idx = torch.empty(NQ * 2036, dtype=torch.int32, device="cuda")
h = torch.empty((NQ, 32, 32), dtype=torch.float32, device="cuda")
pos = torch.empty((NQ, 1020, 4), dtype=torch.float32, device="cuda"); nrm = torch.empty_like(pos)
L, C = pb.lib(), pb.C
pp = C.byref(p); cam = (C.c_double * 3)(0.0, 0.0, -6371010.0)
main = torch.cuda.current_stream(); side = torch.cuda.Stream()
def sp(s): return C.c_void_p(s.cuda_stream)

def serial():
    pb._check(L.planet_gpu_tessellate_uniform(pp, 7, 0, NQ, quads.data_ptr(), idx.data_ptr(), sp(main)))
    pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), NQ, 32, 18, h.data_ptr(), sp(main)))
    pb._check(L.planet_gpu_shade(pp, quads.data_ptr(), NQ, cam, h.data_ptr(), -1.0, pos.data_ptr(), nrm.data_ptr(), sp(main)))

def overlapped(chunks):
    # quads first (K2 needs them), the index stream goes to the side stream next to K2
    pb._check(L.planet_gpu_tessellate_uniform(pp, 7, 0, NQ, quads.data_ptr(), None, sp(main)))
    ev0 = torch.cuda.Event(); ev0.record(main); side.wait_event(ev0)
    pb._check(L.planet_gpu_tessellate_uniform(pp, 7, 0, NQ, None, idx.data_ptr(), sp(side)))
    n = NQ // chunks
    for c in range(chunks):
        lo = c * n
        q = quads[lo:lo + n]; hh = h[lo:lo + n]
        pb._check(L.planet_gpu_generate_height_maps(pp, q.data_ptr(), n, 32, 18, hh.data_ptr(), sp(main)))
        ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
        pb._check(L.planet_gpu_shade(pp, q.data_ptr(), n, cam, hh.data_ptr(), -1.0, pos[lo:lo + n].data_ptr(), nrm[lo:lo + n].data_ptr(), sp(side)))
    evj = torch.cuda.Event(); evj.record(side); main.wait_event(evj)

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps): fn()
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {"serial_ms": timeit(serial)}
ref_pos = pos.clone(); ref_idx = idx.clone()
for ch in (1, 2, 4, 8):
    res[f"overlap_{ch}_ms"] = timeit(lambda: overlapped(ch))
    assert torch.equal(pos, ref_pos) and torch.equal(idx, ref_idx)
print(json.dumps({"k3_warps": os.environ.get("PLANET_K3_WARPS", "8"), **res}))
