import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, planet_b200 as pb
pb.init(0)
p = pb.default_params(); L, C = pb.lib(), pb.C; pp = C.byref(p)
NQ = 16384
quads = torch.empty((NQ, 13), dtype=torch.int64, device="cuda"); idx = torch.empty(NQ * 2036, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn):
    ts = []
    for i in range(25):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3
print(json.dumps({
    "both_us": t(lambda: L.planet_gpu_tessellate_uniform(pp, 7, 0, NQ, quads.data_ptr(), idx.data_ptr(), sp)),
    "quads_only_us": t(lambda: L.planet_gpu_tessellate_uniform(pp, 7, 0, NQ, quads.data_ptr(), None, sp)),
    "indices_only_us": t(lambda: L.planet_gpu_tessellate_uniform(pp, 7, 0, NQ, None, idx.data_ptr(), sp)),
    "memset_133MB_us": t(lambda: idx.zero_()),
    "empty_launch_us": t(lambda: L.planet_gpu_tessellate_uniform(pp, 0, 0, 1, quads.data_ptr(), None, sp)),
}))
