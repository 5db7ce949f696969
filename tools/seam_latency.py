import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import planet_b200 as pb
pb.init(0)
p1 = pb.default_params(); pb.set_params(p1)
q1 = pb.quads_to_host(pb.select_lod((0.0, 0.0, -6371010.0), 18, p1))
for _ in range(50): pb.generate_height_map(q1[0], 32, 18)
t0 = time.perf_counter()
for k in range(1000): pb.generate_height_map(q1[k % len(q1)], 32, 18)
t_map = (time.perf_counter() - t0) / 1000
pt = np.array([0.0, 0.0, -6371000.0])
for _ in range(50): pb.get_height_at(pt, 0, 1)
t0 = time.perf_counter()
for k in range(1000): pb.get_height_at(pt, 0, 1)
t_pt = (time.perf_counter() - t0) / 1000
pf = pb.default_params(precision=pb.FAST); pb.set_params(pf)
for _ in range(50): pb.generate_height_map(q1[0], 32, 18)
t0 = time.perf_counter()
for k in range(1000): pb.generate_height_map(q1[k % len(q1)], 32, 18)
t_fast = (time.perf_counter() - t0) / 1000
print({"map_exact_us": t_map * 1e6, "point_exact_us": t_pt * 1e6, "map_fast_params_us": t_fast * 1e6})
