"""Time K2 (C2 workload) for the tuning knobs exposed through the environment."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, planet_b200 as pb
    pb.init(0)
    kind = os.environ.get("K2_KIND", "fbm")
    p = pb.fbm_params(octaves=8, gain=0.5, precision=pb.FAST) if kind == "fbm" else pb.default_params(precision=pb.FAST, fixed_octaves=8)
    quads = pb.tessellate_uniform(7, first=0, nquads=16384, params=p)
    out = torch.empty((16384, 32, 32), dtype=torch.float32, device="cuda")
    for _ in range(5): pb.generate_height_maps(quads, 32, 18, p, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pb.generate_height_maps(quads, 32, 18, p, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PLANET_") or k.startswith("K2_")},
                      "ms": ms, "gvert_s": 16.777216e6 / ms / 1e6, "tflops": 16.777216e6 * 787 / ms / 1e9,
                      "checksum": float(out.double().sum())}))
else:
    for env in [dict(PLANET_K2_THREADS=t, K2_KIND=k) for k in ("fbm", "ridged") for t in ("512", "768", "1024")]:
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child"], env=e)
