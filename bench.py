#!/usr/bin/env python
"""bench.py -- displaced+shaded vertices/s of the terrain hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic quads, per GPU:
    K1 tessellate  : 16 384 leaf quads of one cube-sphere face at depth 7 + merged strip indices
    K2 heights     : 16 384 x 32^2 = 16 777 216 height samples, fBm 8 octaves (FAST arithmetic)
    K3 shade       : 16 384 x 1 020 displaced positions + normals + Lambert term
(BASELINE.json configs[1]; with N GPUs rank r takes face r -- configs[2] sharded by face --
and faces 6, 7 reuse faces 0, 1 with a seed offset so every GPU always has one face of work:
weak scaling, no data-path collective.  The NCCL gather of finished patches is timed separately
and reported as `with_gather`.)

`value` counts height-map vertices (32^2 per quad, border included -- SURVEY.md 8d) per second
with inputs resident in HBM; `e2e` is the same metric through the reference-facing host-buffer
call (quads in pinned host memory, height maps back to pinned host memory).  `--impl reference`
times the reference's own CPU implementation (oracle/_ref, else the C port) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries write there too (with NCCL_DEBUG=VERSION in the
# environment NCCL prints its banner to stdout whatever NCCL_DEBUG_FILE says), so file descriptor 1
# is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    os.write(1 if _REAL_STDOUT is None else _REAL_STDOUT, (json.dumps(line) + "\n").encode())

DEPTH, DIM, PATCH, OCTAVES, GAIN, MAX_LOD = 7, 32, 30, 8, 0.5, 18
QUADS_PER_FACE = 4 ** DEPTH                      # 16 384
VERTS_PER_GPU = QUADS_PER_FACE * DIM * DIM       # 16 777 216
# algorithmic flop per height sample = octaves*95 + 27
FLOP_PER_VERTEX = OCTAVES * 95 + 27              # 787
METRIC = "displaced+shaded vertices/sec"
WORKLOAD = ("C2 per GPU: one cube-sphere face at depth 7 = 16384 quads x 32^2 = 16.78M vertices, "
            "fBm 8 octaves gain 0.5, patch 30 verts; N GPUs = N faces (C3 sharded by face)")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock/throttle sampling (NVML) while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:                                           # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        while not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                fn = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = fn(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:                                       # noqa: BLE001
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(nquads, threads, repeats=1):
    """The reference's CPU path (GenerateHeightMap over quads) on the host cores; returns
    (vertices/s, kind, seconds).  This is the one place bench.py executes oracle/."""
    from oracle.bindings import FBM, best_oracle, height_params, PortOracle
    orc = best_oracle()
    quads = PortOracle().uniform_quads(0, DEPTH)[:nquads]
    hp = height_params(kind=FBM, gain=GAIN, fixed_octaves=OCTAVES)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        maps = orc.generate_height_maps(quads, DIM, MAX_LOD, hp, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert np.isfinite(maps).all()
    cpu_reference_run.last_maps = maps
    return nquads * DIM * DIM / best, orc.kind, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    nquads = QUADS_PER_FACE                       # the whole C2 batch: ~6 core-seconds of CPU work
    times = []
    kind = None
    for i in range(args.warmup + args.steps):
        v, kind, dt = cpu_reference_run(nquads, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = nquads * DIM * DIM / (ms * 1e-3)
    sample = f"{nquads} quads x {DIM}^2 (the full C2 batch) per step, GenerateHeightMap only"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "vertices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path = height-map generation "
                   "(main.cpp:123-151); its displacement+normals exist only as a GLSL shader"},
        "cpu_baseline": {"value": value, "unit": "vertices/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import planet_b200 as pb

    torch.cuda.set_device(local_rank)
    pb.init(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    face = rank % 6
    seed = (0.0, 0.0, 0.0) if rank < 6 else (0.0, 64.5 * (rank // 6), 0.0)
    params = pb.fbm_params(octaves=OCTAVES, gain=GAIN, precision=pb.FAST, seed_offset=seed)
    nq, nv, ni = QUADS_PER_FACE, pb.patch_vertex_count(PATCH), pb.patch_index_count(PATCH)
    cam = (0.0, 0.0, -6371000.0 - 10.0)           # main.cpp:864

    quads = torch.empty((nq, 13), dtype=torch.int64, device=dev)
    indices = torch.empty(nq * ni, dtype=torch.int32, device=dev)
    heights = torch.empty((nq, DIM, DIM), dtype=torch.float32, device=dev)
    pos = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
    nrm = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # 2x the 126 MB L2
    L, C = pb.lib(), pb.C
    pp = C.byref(params)
    camv = (C.c_double * 3)(*cam)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def k1():
        pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, face * nq, nq, quads.data_ptr(), indices.data_ptr(), sp))

    def k2():
        pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), nq, DIM, MAX_LOD, heights.data_ptr(), sp))

    def k3():
        pb._check(L.planet_gpu_shade(pp, quads.data_ptr(), nq, camv, heights.data_ptr(), -1.0,
                                     pos.data_ptr(), nrm.data_ptr(), sp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)                      # noqa: E731
    for _ in range(max(args.warmup, 3)):
        k1(); k2(); k3()
    barrier()

    # ---- timed region: K steps, per-step CUDA events on the launching stream, L2 flushed between ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    marks = []
    launches0 = pb.launch_count()
    barrier()
    for _ in range(args.steps):
        flush.zero_()                                                      # evict L2 (not timed)
        e = [ev() for _ in range(4)]
        e[0].record(); k1(); e[1].record(); k2(); e[2].record(); k3(); e[3].record()
        marks.append(e)
    barrier()
    launches = pb.launch_count() - launches0
    t_k1 = np.array([m[0].elapsed_time(m[1]) for m in marks])
    t_k2 = np.array([m[1].elapsed_time(m[2]) for m in marks])
    t_k3 = np.array([m[2].elapsed_time(m[3]) for m in marks])
    t_step = np.array([m[0].elapsed_time(m[3]) for m in marks])

    # ---- the same K2 batch in EXACT arithmetic (bit-identical to the reference), for the record ----
    p_exact = pb.fbm_params(octaves=OCTAVES, gain=GAIN, precision=pb.EXACT, seed_offset=seed)
    ppe = C.byref(p_exact)
    ex_t = []
    for i in range(5):
        a, b_ = ev(), ev()
        a.record()
        pb._check(L.planet_gpu_generate_height_maps(ppe, quads.data_ptr(), nq, DIM, MAX_LOD, heights.data_ptr(), sp))
        b_.record(); torch.cuda.synchronize()
        if i >= 2:
            ex_t.append(a.elapsed_time(b_))
    ms_k2_exact = float(np.mean(ex_t))
    exact_maps_host = heights.cpu().numpy() if (rank == 0 and world == 1) else None
    k2()                                                               # restore the FAST heights for K3 / e2e
    # parity at full size, stated in the bench line: FAST against the bit-exact mode on all 16.8 M samples,
    # with the bound the tests use (1e-5 relative to the fractal's amplitude sum times the height scale)
    fast_err = None
    if exact_maps_host is not None:
        torch.cuda.synchronize()
        fast_err = float(np.abs(heights.cpu().numpy().astype(np.float64) - exact_maps_host).max())
    fast_tol = 1e-5 * 8848.0 * float(sum(np.float32(GAIN) ** k for k in range(OCTAVES)))

    # ---- e2e: the host-buffer call a reference-side caller makes: H2D quads, K2, D2H heights, and
    # K3 on the device-resident maps (the GL texture's role) while the last maps cross PCIe ----
    h_quads = torch.empty((nq, 13), dtype=torch.int64).pin_memory()
    h_quads.copy_(quads.cpu())
    h_out = torch.empty((nq, DIM, DIM), dtype=torch.float32).pin_memory()
    cam3 = (C.c_double * 3)(*cam)
    e2e_t = []
    for i in range(3 + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pb._check(L.planet_gpu_terrain_host(pp, h_quads.data_ptr(), nq, MAX_LOD, cam3, -1.0, h_out.data_ptr(),
                                            heights.data_ptr(), pos.data_ptr(), nrm.data_ptr()))
        torch.cuda.synchronize()
        if i >= 3:
            e2e_t.append(time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join()
    e2e_ms = 1e3 * float(np.mean(e2e_t))

    # ---- gather of finished patches (K4) ------------------------------------------------------
    # (a) NCCL all_gather_into_tensor after the step; (b) the gather fused into K2: every height is
    # also stored to the peers' IPC-mapped gathered buffers over NVLink while the kernel computes.
    gather_ms, fused_step_ms, gather_identical, fused_error = None, None, None, None
    if world > 1:
        from planet_b200.sharding import PeerGather
        allh = torch.empty((world * nq, DIM, DIM), dtype=torch.float32, device=dev)
        for i in range(3):
            dist.all_gather_into_tensor(allh, heights)
        barrier()
        g0, g1 = ev(), ev()
        g0.record()
        for _ in range(5):
            dist.all_gather_into_tensor(allh, heights)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1) / 5

        fused_error = None
        pg = None
        try:                                                        # mapping the peers' buffers (CUDA IPC)
            pg = PeerGather((nq, DIM, DIM), device=dev)
        except Exception as exc:                                    # noqa: BLE001
            fused_error = f"{type(exc).__name__}: {exc}"[:200]
        ok = torch.tensor([0 if pg is None else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                   # all ranks take the same branch
        try:
            if ok.item() == 0:
                raise RuntimeError(fused_error or "a peer rank could not map the gathered buffers")
            shard, peer_shards = pg.local_shard(), pg.peer_shards()
            peer_arr = (C.c_void_p * len(peer_shards))(*[t.data_ptr() for t in peer_shards])

            def fused_step():
                k1()
                pb._check(L.planet_gpu_generate_height_maps_gathered(pp, quads.data_ptr(), nq, DIM, MAX_LOD, shard.data_ptr(),
                                                                     peer_arr, len(peer_shards), sp))
                pb._check(L.planet_gpu_shade(pp, quads.data_ptr(), nq, camv, shard.data_ptr(), -1.0,
                                             pos.data_ptr(), nrm.data_ptr(), sp))
                pg.finish(stream)

            for _ in range(3):
                fused_step()
            ts = []
            for _ in range(args.steps):
                barrier()
                a, b_ = ev(), ev()
                a.record(); fused_step(); b_.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            fused_step_ms = float(np.mean(ts))
            k2(); dist.all_gather_into_tensor(allh, heights); torch.cuda.synchronize()
            gather_identical = bool(torch.equal(pg.gathered, allh))
            del shard, peer_shards
            pg.close()
        except Exception as exc:                                    # noqa: BLE001 -- e.g. CUDA IPC unavailable on this box
            fused_error = f"{type(exc).__name__}: {exc}"[:200]
            fused_step_ms, gather_identical = None, None

    # max over ranks (device-timed)
    stats = torch.tensor([t_step.mean(), t_k1.mean(), t_k2.mean(), t_k3.mean(), e2e_ms,
                          gather_ms or 0.0, fused_step_ms if fused_step_ms else 1e9, 0.0 if gather_identical else 1.0],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_step, ms_k1, ms_k2, ms_k3, e2e_ms, gather_ms, fused_step_ms, gather_bad = [float(x) for x in stats.tolist()]

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        fp32_tf, _ = pb.measure_fp32_peak(300.0)                           # live FFMA probe on this GPU
        info = pb.device_info()
        nominal_tf = info["sm_count"] * info["fp32_lanes_per_sm"] * 2 * info["clock_khz"] * 1e3 / 1e12
        k2_tf = VERTS_PER_GPU * FLOP_PER_VERTEX / (ms_k2 * 1e-3) / 1e12
        k1_bytes = nq * 104 + nq * ni * 4
        k3_bytes = nq * DIM * DIM * 4 + nq * nv * 32 + nq * 104
        total_verts = VERTS_PER_GPU * world
        cores = host_cores()
        cpu_v, cpu_kind, cpu_s = cpu_reference_run(QUADS_PER_FACE, cores)
        # both sides computed the same thing: the CPU maps are the bytes the EXACT-mode kernel produced
        same_bytes = None if exact_maps_host is None else (
            exact_maps_host.tobytes() == np.ascontiguousarray(cpu_reference_run.last_maps).tobytes())
        cpu1_v, _, cpu1_s = cpu_reference_run(1024, 1)                      # what the reference itself does: one thread
        line = {
            "metric": METRIC, "value": total_verts / (ms_step * 1e-3), "unit": "vertices/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "depth": DEPTH, "dim": DIM, "octaves": OCTAVES, "gain": GAIN,
                       "quads_per_gpu": nq, "vertices_per_gpu": VERTS_PER_GPU, "precision": "FAST",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); step working set 730 MB",
                       "step": "K1 tessellate + K2 heights + K3 shade"},
            "ms": {"k1_tessellate": ms_k1, "k2_heights": ms_k2, "k3_shade": ms_k3,
                   "k2_heights_exact_mode": ms_k2_exact},
            "parity": {"fast_vs_exact_max_abs_m": fast_err, "tolerance_m": fast_tol,
                       "exact_mode_bytes_equal_reference_cpu": same_bytes},
            "roofline": {"kernel": "k_height_maps_fast<768,32,gather=0,kind=fBm>", "bound": "fp32", "achieved": k2_tf, "peak": fp32_tf,
                         "unit": "TFLOP/s", "frac": k2_tf / fp32_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                         # (profiles/r01zf_ncu_summary_final.txt: 1.74 MB read + 12.81 MB written): the 67 MB of
                         # heights are still dirty in the 126 MB L2 when the kernel ends
                         "traffic": 14550016,
                         "peak_source": "FFMA probe measured in this run", "nominal_peak": nominal_tf,
                         "frac_of_nominal": k2_tf / nominal_tf, "flop_per_vertex": FLOP_PER_VERTEX,
                         "vertices_per_s": VERTS_PER_GPU / (ms_k2 * 1e-3)},
            "roofline_k1": {"kernel": "k_tessellate_fused", "bound": "hbm",
                            "achieved": k1_bytes / (ms_k1 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": k1_bytes / (ms_k1 * 1e-3) / 1e9 / hbm_peak, "bytes": k1_bytes,
                            "peak_source": peak_src},
            # algorithmic bytes; the 67 MB of height maps K2 just wrote are still in the 126 MB L2 when K3
            # reads them (L2 is flushed between steps, not between K2 and K3), so HBM sees ~536 MB of it
            "roofline_k3": {"kernel": "k_shade", "bound": "hbm", "achieved": k3_bytes / (ms_k3 * 1e-3) / 1e9,
                            "peak": hbm_peak, "unit": "GB/s", "frac": k3_bytes / (ms_k3 * 1e-3) / 1e9 / hbm_peak,
                            "bytes": k3_bytes, "peak_source": peak_src},
            "cpu_baseline": {"value": cpu_v, "unit": "vertices/s", "cores": cores, "kind": cpu_kind,
                             "sample": f"{QUADS_PER_FACE} quads x {DIM}^2 (full C2 batch), GenerateHeightMap only, "
                                       f"{cpu_s:.2f} s wall",
                             "single_thread_value": cpu1_v,
                             "single_thread_sample": f"1024 quads x {DIM}^2, {cpu1_s:.2f} s wall",
                             "same_bytes_as_gpu_exact_mode": same_bytes},
            "e2e": {"value": total_verts / (e2e_ms * 1e-3), "unit": "vertices/s",
                    "h2d_bytes_per_step": nq * 104, "d2h_bytes_per_step": nq * DIM * DIM * 4,
                    "ms_per_step": e2e_ms,
                    "path": "planet_gpu_terrain_host: pinned host quads -> K2 -> pinned host heights (8-chunk pipeline) + K3 on the resident maps"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        nccl_gather = {"value": total_verts / ((ms_step + gather_ms) * 1e-3) if world > 1 else None, "gather_ms": gather_ms,
                       "collective": "nccl all_gather_into_tensor of height maps after the step"}
        if world > 1 and fused_step_ms >= 1e8:                       # peer mapping failed on some rank: NCCL only
            line["with_gather"] = dict(nccl_gather, unit="vertices/s", note="fused peer-store gather unavailable: " + str(fused_error))
        elif world > 1:
            line["with_gather"] = {
                "value": total_verts / (fused_step_ms * 1e-3), "unit": "vertices/s", "ms_per_step": fused_step_ms,
                "method": "gather fused into K2: heights stored to every peer's CUDA-IPC-mapped buffer over NVLink "
                          "while the kernel computes (planet_gpu_generate_height_maps_gathered) + one barrier",
                "bytes_per_gpu": nq * DIM * DIM * 4, "identical_to_nccl_all_gather": gather_bad == 0.0,
                "nccl": nccl_gather}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
