import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, numpy as np, planet_b200 as pb
    pb.init(0)
    p = pb.fbm_params(8, 0.5, pb.FAST)
    nq = int(os.environ.get("K3_NQ", "16384"))
    quads = pb.tessellate_uniform(7, 0, nq, p)
    h = pb.generate_height_maps(quads, 32, 18, p)
    pos = torch.empty((nq, 1020, 4), dtype=torch.float32, device="cuda"); nrm = torch.empty_like(pos)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(25):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pb.shade(quads, h, (0, 0, -6371010.0), p, pos=pos, nrm=nrm); e1.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts)); b = nq * (1024 * 4 + 1020 * 32 + 104)
    print(json.dumps({"warps": os.environ.get("PLANET_K3_WARPS"), "blocks_per_sm": os.environ.get("PLANET_K3_BLOCKS_PER_SM"), "nq": nq, "ms": ms, "gbs": b / ms / 1e6}))
else:
    for nq in ("16384", "98304"):
        for w, b in (("8", "3"), ("8", "2"), ("8", "1"), ("4", "4"), ("4", "6"), ("4", "8"), ("2", "8"), ("2", "16")):
            e = dict(os.environ); e.update(PLANET_K3_WARPS=w, PLANET_K3_BLOCKS_PER_SM=b, K3_NQ=nq)
            subprocess.run([sys.executable, __file__, "child"], env=e)
