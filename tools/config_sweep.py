"""Measure every BASELINE.json config on one B200 (per-GPU shard for the multi-GPU ones).
Writes one JSON object per line; copy the output into profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import planet_b200 as pb

pb.init(0)
CAM = (0.0, 0.0, -6371000.0 - 10.0)


def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def pipeline(name, quads_fn, params, dim, max_lod=18, shade=True, note=""):
    quads = quads_fn()
    nq = quads.shape[0]
    n = params.patch_verts
    heights = torch.empty((nq, dim, dim), dtype=torch.float32, device="cuda")
    t_k2 = timed(lambda: pb.generate_height_maps(quads, dim, max_lod, params, out=heights))
    out = {"config": name, "quads": nq, "dim": dim, "samples": nq * dim * dim, "k2_ms": t_k2,
           "k2_gvert_s": nq * dim * dim / t_k2 / 1e6, "note": note}
    if shade:
        nv = pb.patch_vertex_count(n)
        pos = torch.empty((nq, nv, 4), dtype=torch.float32, device="cuda"); nrm = torch.empty_like(pos)
        t_k3 = timed(lambda: pb.shade(quads, heights, CAM, params, pos=pos, nrm=nrm))
        out.update(k3_ms=t_k3, k3_gbs=(nq * dim * dim * 4 + nq * nv * 32) / t_k3 / 1e6)
    print(json.dumps(out), flush=True)


# C1: the reference's default frame (ridged, 6 + 12*depth/18 octaves), LOD selection on the GPU
p1 = pb.default_params()
t_lod = timed(lambda: pb.select_lod(CAM, 18, p1), reps=5, warm=2)
q1 = pb.select_lod(CAM, 18, p1)
print(json.dumps({"config": "C1 lod selection", "quads": int(q1.shape[0]), "ms": t_lod}), flush=True)
pipeline("C1 default planet frame, EXACT (bit-identical to the reference)", lambda: q1, p1, 32)
pipeline("C1 default planet frame, FAST", lambda: q1, pb.default_params(precision=pb.FAST), 32)

# C1 through the reference's own seam: one synchronous call per map / per point, host pointers
# (what GetHeightMapForQuad and ProcessQuad do, main.cpp:244, 552-555); wall clock per call
pb.set_params(p1)
hq1 = pb.quads_to_host(q1)
buf = np.empty((32, 32), np.float32)
for _ in range(20): pb.generate_height_map(hq1[0], 32, 18)
t0 = time.perf_counter()
for k in range(400): pb.generate_height_map(hq1[k % len(hq1)], 32, 18)
t_map = (time.perf_counter() - t0) / 400
pt = np.array([0.0, 0.0, -6371000.0])
for _ in range(20): pb.get_height_at(pt, 0, 1)
t0 = time.perf_counter()
for k in range(400): pb.get_height_at(pt, 0, 1)
t_pt = (time.perf_counter() - t0) / 400
print(json.dumps({"config": "C1 legacy seam (HeightMapGenerator's two pointers, synchronous, host buffers)",
                  "generate_height_map_32x32_us": t_map * 1e6, "get_height_at_us": t_pt * 1e6,
                  "note": "EXACT arithmetic; includes H2D of the quad, the launch, D2H of the map and the ctypes call"}), flush=True)

# C2: one face, depth 7, fBm 8
p2 = pb.fbm_params(8, 0.5, pb.FAST)
pipeline("C2 one face depth 7 fBm-8 FAST", lambda: pb.tessellate_uniform(7, 0, 16384, p2), p2, 32)
pipeline("C2 one face depth 7 fBm-8 EXACT", lambda: pb.tessellate_uniform(7, 0, 16384, p2), pb.fbm_params(8, 0.5, pb.EXACT), 32)

# C3: full planet 100M vertices on one GPU (98 304 quads) and the 1/8 shard
pipeline("C3 full planet depth 7 fBm-8 FAST (all 98304 quads on 1 GPU)", lambda: pb.tessellate_uniform(7, 0, 98304, p2), p2, 32)
pipeline("C3 1/8 shard (12288 quads)", lambda: pb.tessellate_uniform(7, 0, 12288, p2), p2, 32)

# C4: 1B vertices = depth 8, patch 50 (dim 52), fBm 12: the 1/8 shard of 49 152 quads
p4 = pb.fbm_params(12, 0.5, pb.FAST, patch_verts=50)
pipeline("C4 1/8 shard: 49152 quads depth 8, patch 50 (dim 52), fBm-12 FAST", lambda: pb.tessellate_uniform(8, 0, 49152, p4), p4, 52)

# C5: single-quad / small-batch latency, dim 64..4096, octaves 1..16.  Launches are captured
# into a CUDA graph (20 per graph) and the replay is timed, so the number is device time per
# launch, not the Python/ctypes launch path.
def graph_time(fn, per_graph=20, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / per_graph)
    return float(np.median(ts))


root = pb.tessellate_uniform(3, 0, 256, p2)
for dim in (64, 128, 256, 512, 1024, 2048, 4096):
    for octaves in (1, 8, 16):
        for batch in (1, 4, 16, 64, 256):
            if dim * dim * batch > 1 << 26:
                continue
            pp = pb.fbm_params(octaves, 0.5, pb.FAST)
            out = torch.empty((batch, dim, dim), dtype=torch.float32, device="cuda")
            ms = graph_time(lambda: pb.generate_height_maps(root[:batch], dim, 18, pp, out=out))
            print(json.dumps({"config": "C5", "dim": dim, "octaves": octaves, "batch": batch, "ms": ms,
                              "us_per_patch": ms * 1e3 / batch, "gvert_s": batch * dim * dim / ms / 1e6}), flush=True)

# C5, the shade kernel at the patch sizes it takes (a patch's map must fit the warp's shared-memory staging: dim <= 256)
for dim in (64, 128, 256):
    for batch in (1, 16, 256):
        pp = pb.fbm_params(8, 0.5, pb.FAST, patch_verts=dim - 2)
        maps = pb.generate_height_maps(root[:batch], dim, 18, pp)
        nv = pb.patch_vertex_count(dim - 2)
        pos = torch.empty((batch, nv, 4), dtype=torch.float32, device="cuda"); nrm = torch.empty_like(pos)
        ms = graph_time(lambda: pb.shade(root[:batch], maps, (0.0, 0.0, -6371010.0), pp, pos=pos, nrm=nrm))
        print(json.dumps({"config": "C5 K3", "dim": dim, "batch": batch, "ms": ms, "us_per_patch": ms * 1e3 / batch,
                          "gvert_s": batch * nv / ms / 1e6, "GBs_written": batch * nv * 32 / ms / 1e6}), flush=True)
