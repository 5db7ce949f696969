#!/usr/bin/env python
"""bench.py -- displaced+shaded vertices/s of the terrain hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic quads:
    K1 tessellate  : leaf quads of the uniform depth-7 tree + merged strip indices
    K2 heights     : 32^2 height samples per quad, fBm 8 octaves (FAST arithmetic)
    K3 shade       : 1 020 displaced positions + normals + Lambert term per quad
    K4 gather      : (N > 1) every rank's finished height maps in one buffer on every rank

N = 1 runs BASELINE.json configs[1] (C2): one cube-sphere face, 16 384 quads, 16.8 M vertices.
N > 1 runs configs[2] (C3): the full planet, 6 faces = 98 304 quads = 100.7 M vertices as a FIXED
total, rank r taking the patch range shard_range(98 304, r, N) -- strong scaling --
and the timed step INCLUDES K4, fused into K2 (planet_gpu_gather_height_maps: finished tiles leave
the SM as bulk copies to every peer over NVLink while the kernel computes; arrival is signalled
GPU to GPU, no host barrier inside a step).  Compute without the gather and the plain NCCL
collective are extra keys.

Environment knobs (experiments; the defaults are what the numbers in DESIGN.md were measured with):
    PLANET_GATHER_MODE=fused|push|concurrent   how K4 is driven at N > 1
    PLANET_GATHER_SHADE_SHARE=k                K3's share of the pushes: 0, every k-th map (k >= 2) or -m = m maps in 8
    PLANET_GATHER_CHUNK_WAVES=a,b              chunk sizes of the `push` mode, in K2 waves
    PLANET_K1_BESIDE=1                         N = 1: K1's index stream by the slim kernel on a side stream beside K2

`value` counts height-map vertices (32^2 per quad, border included -- SURVEY.md 8d) per second
with inputs resident in HBM; `e2e` is the same metric through the reference-facing host-buffer
call (quads in pinned host memory, height maps back to pinned host memory).  `--impl reference`
times the reference's own CPU implementation (oracle/_ref) on the host cores.
"""
import argparse
import glob
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
# algorithmic flop per height sample = octaves*95 + 27
FLOP_PER_VERTEX = OCTAVES * 95 + 27              # 787
METRIC = "displaced+shaded vertices/sec"


def total_quads(world):
    """C2 (one face) on one GPU, C3 (the planet) as a fixed total on several."""
    return QUADS_PER_FACE if world == 1 else 6 * QUADS_PER_FACE


def shard_range(n_units, rank, world):                              # == planet_b200.sharding.shard_range
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def make_config(world):
    """The workload, identically worded by both arms (--impl ours / reference)."""
    nq = total_quads(world)
    if world == 1:
        workload = ("C2 (BASELINE.json configs[1]): one cube-sphere face at depth 7 = 16384 quads x 32^2 = 16.78M vertices, "
                    "fBm 8 octaves gain 0.5, patch 30 verts, 1 GPU")
    else:
        workload = (f"C3 (BASELINE.json configs[2]): full planet, 6 faces at depth 7 = 98304 quads x 32^2 = 100.66M vertices "
                    f"as a fixed total, fBm 8 octaves gain 0.5, patch 30 verts, split by patch range over {world} GPUs, "
                    f"finished height maps gathered on every GPU")
    per_gpu = [b - a for a, b in (shard_range(nq, r, world) for r in range(world))]
    # every key is emitted identically by both arms (the driver compares the two config objects), so the
    # entries that describe one arm only say which arm they describe
    return {"workload": workload, "depth": DEPTH, "dim": DIM, "octaves": OCTAVES, "gain": GAIN, "faces": nq // QUADS_PER_FACE,
            "quads": nq, "vertices": nq * DIM * DIM, "gpus": world,
            "quads_per_gpu": max(per_gpu), "vertices_per_gpu": max(per_gpu) * DIM * DIM,
            "precision": "GPU arm: FAST arithmetic (<= 1e-5 * height_scale * sum(gain^k) from the reference; the EXACT mode is "
                         "bit-identical and timed beside it); reference arm: the reference's own arithmetic",
            "l2": "GPU arm: 256 MiB buffer written between timed steps (L2 flush); reference arm: every step writes the whole "
                  "height-map batch (67 MB per face), larger than the host's last-level cache",
            "step": "GPU arm: K1 tessellate + K2 heights + K3 shade" + ("" if world == 1 else " + K4 gather (fused into K2/K3, wait for the peers' shards)") +
                    "; reference arm: GenerateHeightMap on every quad (the reference has no CPU displacement/normals)"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one K2 launch from the newest committed ncu
    capture (profiles/*k2_dram_traffic.json, written by tools/ncu_summary.py from an `ncu --set full`
    run of this same command); None when no capture is committed."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*k2_dram_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return int(d["dram_bytes_read"] + d["dram_bytes_write"]), os.path.relpath(files[-1], ROOT)


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


def reference_quads(world, nquads=None):
    """The workload's quads from the CPU checker (face-major emission order, main.cpp:604-624)."""
    from oracle.bindings import PortOracle
    port = PortOracle()
    faces = total_quads(world) // QUADS_PER_FACE
    quads = np.concatenate([port.uniform_quads(f, DEPTH) for f in range(faces)])
    return quads if nquads is None else quads[:nquads]


def cpu_reference_run(orc, quads, threads, repeats=1):
    """The reference's CPU path (GenerateHeightMap over quads) on the host cores; returns
    (vertices/s, seconds, maps).  This is the one place bench.py executes oracle/."""
    from oracle.bindings import FBM, height_params
    hp = height_params(kind=FBM, gain=GAIN, fixed_octaves=OCTAVES)
    best, maps = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        maps = orc.generate_height_maps(quads, DIM, MAX_LOD, hp, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert np.isfinite(maps).all()
    return len(quads) * DIM * DIM / best, best, maps


def run_reference(args, rank, world):
    """The reference arm: the reference's own CPU implementation of the path (oracle/_ref, built
    from /root/reference at -O3 where this host runs that build, else -O2, else the C port), all
    host cores, on the GPU arm's workload."""
    if rank != 0:
        return
    if "WORLD_SIZE" not in os.environ:                                  # launched without torchrun: --gpus names the workload
        world = max(args.gpus, 1)
    from oracle.bindings import fastest_oracle
    orc = fastest_oracle()
    cores = host_cores()
    quads = reference_quads(world)                # the whole workload: ~6 (C2) / ~40 (C3) core-seconds per step
    times = []
    for i in range(args.warmup + args.steps):
        _, dt, _ = cpu_reference_run(orc, quads, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = len(quads) * DIM * DIM / (ms * 1e-3)
    sample = (f"{len(quads)} quads x {DIM}^2 (the whole workload) per step, GenerateHeightMap only "
              f"(main.cpp:123-151; the reference's displacement + normals exist only as a GLSL shader), "
              f"built {getattr(orc, 'flags', 'gcc -O2 -ffp-contract=off (C port)')}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "vertices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(world),
        "cpu_baseline": {"value": value, "unit": "vertices/s", "cores": cores, "kind": orc.kind, "sample": sample},
        "e2e": {"value": value, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import planet_b200 as pb
    from planet_b200.sharding import PatchGather

    torch.cuda.set_device(local_rank)
    pb.init(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    Q = total_quads(world)
    lo, hi = shard_range(Q, rank, world)
    nq = hi - lo                                                            # this rank's quads
    params = pb.fbm_params(octaves=OCTAVES, gain=GAIN, precision=pb.FAST)
    nv, ni = pb.patch_vertex_count(PATCH), pb.patch_index_count(PATCH)
    cam = (0.0, 0.0, -6371000.0 - 10.0)           # main.cpp:864
    warmup = max(args.warmup, 3)

    quads = torch.empty((nq, 13), dtype=torch.int64, device=dev)
    indices = torch.empty(nq * ni, dtype=torch.int32, device=dev)
    pos = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
    nrm = torch.empty((nq, nv, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # 2x the 126 MB L2
    gather = PatchGather(Q, DIM, n_buffers=2, device=dev) if world > 1 else None
    # one rank: the height maps live in a plain buffer; several: in the rank's slice of the gathered buffers
    heights = torch.empty((nq, DIM, DIM), dtype=torch.float32, device=dev) if world == 1 else None
    L, C = pb.lib(), pb.C
    pp = C.byref(params)
    camv = (C.c_double * 3)(*cam)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def k1():
        pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, lo, nq, quads.data_ptr(), indices.data_ptr(), sp))

    # N > 1: K1 in its two independent halves -- the quads first (K2 needs them), the index stream
    # (pure HBM writes) on a side stream beside K3, whose share of the NVLink pushes leaves HBM idle
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    side_p = C.c_void_p(side.cuda_stream) if side is not None else None
    # K3 pushes every n-th map (0: K2 pushes all).  Worth it only where the step is NVLink-bound: every GPU
    # receives (N-1)/N of 403 MB at ~660 GB/s while K2 computes its 1/N of the planet at ~39 Gvert/s, so the
    # transfer outlasts K2 from N = 6 up (8 GPUs: 0.53 ms against 0.32 ms) and is hidden under it below
    nvlink_ms = (Q - nq) * DIM * DIM * 4 / 660e9 * 1e3
    k2_ms_est = nq * DIM * DIM / 39e9 * 1e3
    shade_share = int(os.environ.get("PLANET_GATHER_SHADE_SHARE", "2" if nvlink_ms > k2_ms_est else "0"))

    # How K4 is driven (PLANET_GATHER_MODE):
    #   fused  K2 pushes its finished tiles itself (planet_gpu_gather_height_maps), K3 its share
    #   push   K2 runs unfused in chunks into this rank's slice of the gathered buffer; every finished chunk is
    #          pushed to the peers by the library's small pusher kernel on a side stream, resident on the SMs beside
    #          the K2 CTAs of the next chunk and beside K3 (planet_gpu_gather_begin / _push / _publish)
    #   concurrent  ONE unfused K2 launch that publishes per-warp progress counters, and the pusher kernel following
    #          them on the side stream from the first finished tile on (planet_gpu_gather_height_maps in that push mode)
    gather_mode = os.environ.get("PLANET_GATHER_MODE", "fused") if world > 1 else "none"
    # chunk sizes in K2 "waves" (one 128-sample tile for each of the 148 x 24 resident warps = 444 maps of 32^2):
    # whole waves keep every warp equally loaded; a short first chunk starts the transfer early
    wave = 148 * 24 * 128 // (DIM * DIM)
    waves = [float(w) for w in os.environ.get("PLANET_GATHER_CHUNK_WAVES", "2,5").split(",")]
    chunks, at = [], 0
    while at < nq:
        w = waves[min(len(chunks), len(waves) - 1)]
        n = min(nq - at, max(1, int(round(w * wave))))
        if nq - at - n < wave // 2:                                        # no crumb at the end
            n = nq - at
        chunks.append((at, n))
        at += n

    # PLANET_K1_BESIDE=1: K1's index stream by the slim kernel on a side stream beside K2.  Measured neutral at C2
    # (step 0.5672 -> 0.5637 ms: the stream leaves the critical path, K2 beside it runs 1.6 % slower, and what stays
    # in front of K2 is the quads kernel's fp64 latency, 0.020 ms), so the default step keeps its three kernels in a row
    k1_beside = world == 1 and os.environ.get("PLANET_K1_BESIDE", "0") != "0"
    side1 = torch.cuda.Stream(device=dev) if k1_beside else None
    side1_p = C.c_void_p(side1.cuda_stream) if k1_beside else None

    def k1_quads():
        pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, lo, nq, quads.data_ptr(), None, sp))

    def k1_indices():
        pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, lo, nq, None, indices.data_ptr(), side_p))

    def k2(out):
        pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr(), nq, DIM, MAX_LOD, out.data_ptr(), sp))

    def k3(maps):
        ptr = maps if isinstance(maps, int) else maps.data_ptr()
        pb._check(L.planet_gpu_shade(pp, quads.data_ptr(), nq, camv, ptr, -1.0, pos.data_ptr(), nrm.data_ptr(), sp))

    # this rank's slice of each gathered buffer (device pointers; the memory belongs to the C++ gather object)
    shard_ptr = [gather.local(lo, nq, which=b).data_ptr() for b in range(2)] if gather is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)                      # noqa: E731

    def step(marks=None):
        """One step; with `marks`, CUDA events on the launching stream around every kernel."""
        e = [ev() for _ in range(5)] if marks is not None else None
        if e: e[0].record()
        if gather is None and k1_beside:
            # K1 in its two independent halves: the quads (K2 needs them) on the main stream, the merged index
            # stream -- nothing but HBM writes -- by the slim kernel on a side stream, resident on the SMs
            # BESIDE K2, which keeps their arithmetic busy and never touches HBM
            fork = torch.cuda.Event()
            fork.record(stream)
            side1.wait_event(fork)
            pb._check(L.planet_gpu_merged_indices_beside(pp, nq, indices.data_ptr(), side1_p))
            join = torch.cuda.Event()
            join.record(side1)
            k1_quads()
            if e: e[1].record()
            k2(heights)
            if e: e[2].record()
            k3(heights)
            stream.wait_event(join)                                         # the step is over when the index stream is too
            if e: e[3].record(); e[4].record()
        elif gather is None:
            k1()
            if e: e[1].record()
            k2(heights)
            if e: e[2].record()
            k3(heights)
            if e: e[3].record(); e[4].record()
        elif gather_mode == "push":
            fork = torch.cuda.Event()
            fork.record(stream)
            k1_quads()
            if e: e[1].record()
            pb._check(L.planet_gpu_gather_begin(gather.handle, sp))        # next buffer; the peers have released it
            b = L.planet_gpu_gather_last_buffer(gather.handle)
            per = DIM * DIM * 4
            for a, n in chunks:                                            # K2 chunk by chunk, each handed to the pusher kernel
                pb._check(L.planet_gpu_generate_height_maps(pp, quads.data_ptr() + a * 104, n, DIM, MAX_LOD, shard_ptr[b] + a * per, sp))
                pb._check(L.planet_gpu_gather_push(gather.handle, (lo + a) * per, n * per, sp))
            pb._check(L.planet_gpu_gather_publish(gather.handle))          # the peers are signalled when the last push is done
            if e: e[2].record()
            side.wait_event(fork)
            k1_indices()
            join = torch.cuda.Event()
            join.record(side)
            k3(shard_ptr[b])                                               # under the tail of the transfer
            stream.wait_event(join)
            if e: e[3].record()
            pb._check(L.planet_gpu_gather_wait(gather.handle, 1, sp))
            if e: e[4].record()
        else:
            fork = torch.cuda.Event()
            fork.record(stream)
            k1_quads()
            if e: e[1].record()
            # K2 + K4: compute, push finished tiles to all GPUs as bulk copies (the share left to K3 excepted)
            pb._check(L.planet_gpu_gather_height_maps(gather.handle, pp, quads.data_ptr(), nq, lo, DIM, MAX_LOD, sp))
            if e: e[2].record()
            side.wait_event(fork)
            k1_indices()                                                    # beside K3, on the side stream
            join = torch.cuda.Event()
            join.record(side)
            # K3 on this rank's patches (maps read from the gathered buffer) + its share of the pushes + signal
            pb._check(L.planet_gpu_gather_shade(gather.handle, pp, quads.data_ptr(), nq, lo, camv, -1.0,
                                                pos.data_ptr(), nrm.data_ptr(), sp))
            stream.wait_event(join)
            if e: e[3].record()
            pb._check(L.planet_gpu_gather_wait(gather.handle, 1, sp))       # every peer's shard is in this rank's buffer; release it
            if e: e[4].record()
        if e: marks.append(e)

    if gather is not None:
        pb._check(L.planet_gpu_gather_set_shade_share(gather.handle, shade_share if gather_mode == "fused" else 0))
        pb._check(L.planet_gpu_gather_set_push_mode(gather.handle, {"push": 1, "concurrent": 2}.get(gather_mode, 0)))
    for _ in range(warmup):
        step()
    barrier()

    # ---- timed region: K steps, per-step CUDA events on the launching stream, L2 flushed between ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    marks = []
    launches0 = pb.launch_count()
    barrier()
    for _ in range(args.steps):
        flush.zero_()                                                      # evict L2 (not timed)
        step(marks)
    barrier()
    launches = pb.launch_count() - launches0
    if gather is not None:
        gather.check()                                                     # no wait timed out
    seg = lambda a, b: np.array([m[a].elapsed_time(m[b]) for m in marks])  # noqa: E731
    t_k1, t_k2, t_k3, t_wait, t_step = seg(0, 1), seg(1, 2), seg(2, 3), seg(3, 4), seg(0, 4)

    extra = {}
    local_maps = heights if gather is None else gather.local(lo, nq)
    if gather is not None:
        fused_buffer = gather.gathered().clone()
        # ---- the same step without K4 (compute only), and with the plain NCCL collective after it ----
        plain = torch.empty((nq, DIM, DIM), dtype=torch.float32, device=dev)
        tc, tn = [], []
        spans = [(a, b - a) for a, b in (shard_range(Q, r, world) for r in range(world))]
        for i in range(3 + args.steps):
            flush.zero_()
            a, b_ = ev(), ev()
            a.record(); k1(); k2(plain); k3(plain); b_.record()
            torch.cuda.synchronize()
            if i >= 3: tc.append(a.elapsed_time(b_))
        for i in range(3 + min(args.steps, 10)):
            barrier()
            flush.zero_()
            a, b_ = ev(), ev()
            a.record(); k1(); k2(gather.local(lo, nq, which=0)); k3(gather.local(lo, nq, which=0)); gather.nccl(spans, which=0); b_.record()
            torch.cuda.synchronize()
            if i >= 3: tn.append(a.elapsed_time(b_))
        barrier()
        same_as_nccl = bool(torch.equal(fused_buffer, gather.gathered(which=0)))
        # the gathered buffer against the same 98 304 quads computed on THIS GPU alone
        allq = pb.tessellate_uniform(DEPTH, first=0, nquads=Q, params=params)
        single = pb.generate_height_maps(allq, DIM, MAX_LOD, params)
        same_as_single = bool(torch.equal(fused_buffer, single))
        del allq, single, fused_buffer
        extra = {"compute_ms": float(np.mean(tc)), "nccl_step_ms": float(np.mean(tn)),
                 "same_as_nccl": same_as_nccl, "same_as_single": same_as_single}
        local_maps = plain
        k2(plain)

    # ---- N = 1 extras: EXACT-mode K2, full-size parity figures, K1 at a size L2 cannot absorb ----
    ms_k2_exact, fast_err, exact_maps_host, k1_c3_ms, k1_c2_ms = None, None, None, None, None
    if world == 1:
        p_exact = pb.fbm_params(octaves=OCTAVES, gain=GAIN, precision=pb.EXACT)
        ppe = C.byref(p_exact)
        ex_t = []
        scratch = torch.empty_like(heights)
        for i in range(5):
            a, b_ = ev(), ev()
            a.record()
            pb._check(L.planet_gpu_generate_height_maps(ppe, quads.data_ptr(), nq, DIM, MAX_LOD, scratch.data_ptr(), sp))
            b_.record(); torch.cuda.synchronize()
            if i >= 2:
                ex_t.append(a.elapsed_time(b_))
        ms_k2_exact = float(np.mean(ex_t))
        exact_maps_host = scratch.cpu().numpy()
        # FAST against the bit-exact mode on all 16.8 M samples, with the bound the tests use
        fast_err = float(np.abs(heights.cpu().numpy().astype(np.float64) - exact_maps_host).max())
        del scratch
        # K1 (quads + merged index stream, one launch) alone at C2 size, L2 flushed before each
        t2 = []
        for i in range(3 + 10):
            flush.zero_()
            a, b_ = ev(), ev()
            a.record(); k1(); b_.record(); torch.cuda.synchronize()
            if i >= 3:
                t2.append(a.elapsed_time(b_))
        k1_c2_ms = float(np.mean(t2))
        # K1 at C3 size: 98 304 quads, 811 MB written -- six times the 126 MB L2
        q3 = 6 * QUADS_PER_FACE
        quads3 = torch.empty((q3, 13), dtype=torch.int64, device=dev)
        idx3 = torch.empty(q3 * ni, dtype=torch.int32, device=dev)
        t3 = []
        for i in range(3 + 10):
            flush.zero_()
            a, b_ = ev(), ev()
            a.record()
            pb._check(L.planet_gpu_tessellate_uniform(pp, DEPTH, 0, q3, quads3.data_ptr(), idx3.data_ptr(), sp))
            b_.record(); torch.cuda.synchronize()
            if i >= 3:
                t3.append(a.elapsed_time(b_))
        k1_c3_ms = float(np.mean(t3))
        del quads3, idx3
    fast_tol = 1e-5 * 8848.0 * float(sum(np.float32(GAIN) ** k for k in range(OCTAVES)))

    # ---- e2e: the host-buffer call a reference-side caller makes: H2D quads, K2, D2H heights, and
    # K3 on the device-resident maps (the GL texture's role) while the last maps cross PCIe ----
    h_quads = torch.empty((nq, 13), dtype=torch.int64).pin_memory()
    h_quads.copy_(quads.cpu())
    h_out = torch.empty((nq, DIM, DIM), dtype=torch.float32).pin_memory()
    cam3 = (C.c_double * 3)(*cam)
    e2e_t = []
    for i in range(3 + args.steps):
        barrier()
        t0 = time.perf_counter()
        pb._check(L.planet_gpu_terrain_host(pp, h_quads.data_ptr(), nq, MAX_LOD, cam3, -1.0, h_out.data_ptr(),
                                            local_maps.data_ptr(), pos.data_ptr(), nrm.data_ptr()))
        torch.cuda.synchronize()
        if i >= 3:
            e2e_t.append(time.perf_counter() - t0)
    # what the PCIe link alone does with the step's output: one D2H copy of the whole batch into the same pinned buffer
    d2h_t = []
    for i in range(2 + 5):
        barrier()                                                          # all ranks copy at once, as in the e2e step
        a, b_ = ev(), ev()
        a.record(); h_out.copy_(local_maps, non_blocking=True); b_.record()
        torch.cuda.synchronize()
        if i >= 2: d2h_t.append(a.elapsed_time(b_))
    d2h_ms = float(np.mean(d2h_t))
    sampler.stop_flag = True
    sampler.join()
    e2e_ms = 1e3 * float(np.mean(e2e_t))

    # max over ranks (device-timed)
    mine = [t_step.mean(), t_k1.mean(), t_k2.mean(), t_k3.mean(), t_wait.mean(), e2e_ms,
            extra.get("compute_ms", 0.0), extra.get("nccl_step_ms", 0.0),
            0.0 if extra.get("same_as_nccl", True) else 1.0, 0.0 if extra.get("same_as_single", True) else 1.0]
    stats = torch.tensor(mine, dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allstats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allstats, stats)
        per_rank = torch.stack(allstats).cpu().numpy()
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_step, ms_k1, ms_k2, ms_k3, ms_wait, e2e_ms, compute_ms, nccl_ms, nccl_bad, single_bad = [float(x) for x in stats.tolist()]

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        fp32_tf, _ = pb.measure_fp32_peak(300.0)                           # live FFMA probe on this GPU
        info = pb.device_info()
        nominal_tf = info["sm_count"] * info["fp32_lanes_per_sm"] * 2 * info["clock_khz"] * 1e3 / 1e12
        verts_rank = nq * DIM * DIM
        total_verts = Q * DIM * DIM
        k2_tf = verts_rank * FLOP_PER_VERTEX / (ms_k2 * 1e-3) / 1e12
        k1_bytes = nq * 104 + (nq * ni * 4 if world == 1 else 0)         # N > 1: the timed K1 segment is the quads half
        k3_write = nq * nv * 32
        k3_hbm = k3_write + nq * 104 + (0 if world == 1 else verts_rank * 4)   # cold reads; see roofline_k3.note
        k3_alg = k3_write + nq * 104 + verts_rank * 4
        traffic, traffic_src = profiled_traffic()
        config = make_config(world)
        line = {
            "metric": METRIC, "value": total_verts / (ms_step * 1e-3), "unit": "vertices/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "ms": (({"k1_quads": ms_k1, "k2_heights_with_k1_index_stream_beside_it": ms_k2, "k3_shade": ms_k3,
                     "k1_tessellate_alone": k1_c2_ms, "k2_heights_exact_mode": ms_k2_exact} if k1_beside else
                    {"k1_tessellate": ms_k1, "k2_heights": ms_k2, "k3_shade": ms_k3, "k2_heights_exact_mode": ms_k2_exact}) if world == 1 else
                   {"k1_quads": ms_k1, "k2_heights_with_gather": ms_k2, "k3_shade_with_gather_and_k1_indices_beside_it": ms_k3,
                    "wait_for_peers": ms_wait}),
            "parity": {"fast_vs_exact_max_abs_m": fast_err, "tolerance_m": fast_tol},
            # the dominant kernel; at N > 1 its duration includes the NVLink pushes it is fused with
            "roofline": {"kernel": "k_height_maps_fast<768,32,gather=%d,kind=fBm>" % (world > 1), "bound": "fp32",
                         "achieved": k2_tf, "peak": fp32_tf, "unit": "TFLOP/s", "frac": k2_tf / fp32_tf,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "FFMA probe measured in this run", "nominal_peak": nominal_tf,
                         "frac_of_nominal": k2_tf / nominal_tf, "flop_per_vertex": FLOP_PER_VERTEX,
                         "vertices_per_launch": verts_rank, "ms": ms_k2},
            "roofline_k1": {"kernel": "k_tessellate_bulk", "bound": "hbm",
                            "achieved": k1_bytes / ((k1_c2_ms or ms_k1) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": k1_bytes / ((k1_c2_ms or ms_k1) * 1e-3) / 1e9 / hbm_peak, "bytes": k1_bytes, "ms": k1_c2_ms or ms_k1,
                            "peak_source": peak_src,
                            "note": "the one-launch K1 (quads + index stream) timed alone, L2 flushed before; write-only, and at this size part of "
                                    "the stream is still in the 126 MB L2 when the kernel ends.  Inside the step the index stream runs as "
                                    "k_index_stream_slim beside K2 (ms.k2_heights_with_k1_index_stream_beside_it)" if k1_beside else
                                    "write-only; at this size part of the stream is still in the 126 MB L2 when the kernel ends"},
            # bytes that reach HBM: the two float4 streams written + the quads read.  At N = 1 the 67 MB of
            # height maps K2 wrote in the same step are L2 hits (L2 is flushed between steps, not between
            # K2 and K3); frac follows from bytes_hbm, bytes_algorithmic is printed beside it
            "roofline_k3": {"kernel": "k_shade", "bound": "hbm", "achieved": k3_hbm / (ms_k3 * 1e-3) / 1e9,
                            "peak": hbm_peak, "unit": "GB/s", "frac": k3_hbm / (ms_k3 * 1e-3) / 1e9 / hbm_peak,
                            "bytes_hbm": k3_hbm, "bytes_algorithmic": k3_alg, "ms": ms_k3, "peak_source": peak_src},
            "e2e": {"value": total_verts / (e2e_ms * 1e-3), "unit": "vertices/s",
                    "h2d_bytes_per_step": nq * 104, "d2h_bytes_per_step": nq * DIM * DIM * 4,
                    "ms_per_step": e2e_ms, "bytes_are": "per GPU",
                    "path": "planet_gpu_terrain_host: pinned host quads -> K2 -> pinned host heights (pipelined in chunks of 1, 2, 4, 8, 8 ... "
                            "kernel waves) + K3 on the resident maps",
                    "pcie_floor": {"d2h_copy_of_the_output_alone_ms": d2h_ms, "GBs": nq * DIM * DIM * 4 / (d2h_ms * 1e-3) / 1e9,
                                   "frac": d2h_ms / e2e_ms,
                                   "what": "one cudaMemcpyAsync of this rank's height maps to the same pinned buffer, all ranks copying at the same time, "
                                           "rank 0's figure: the part of e2e that is the PCIe link / the host's memory path; frac = that time / e2e time"}},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if world == 1:
            k1c3_bytes = 6 * QUADS_PER_FACE * (104 + ni * 4)
            k1c3_prof = sorted(glob.glob(os.path.join(ROOT, "profiles", "*k1_c3_dram_traffic.json")))
            k1c3_dram = json.load(open(k1c3_prof[-1]))["dram_bytes_write"] if k1c3_prof else None
            line["roofline_k1_c3"] = {"kernel": "k_tessellate_bulk", "bound": "hbm", "workload": "98304 quads (C3 size): 811 MB written, 6x the L2",
                                      "traffic": k1c3_dram, "traffic_source": os.path.relpath(k1c3_prof[-1], ROOT) if k1c3_prof else None,
                                      "traffic_note": "dram__bytes_write of one launch under ncu: what reaches HBM INSIDE the kernel; the rest of the 811 MB is "
                                                      "still dirty in the 126 MB L2 when it ends and is written back afterwards",
                                      "achieved": k1c3_bytes / (k1_c3_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": k1c3_bytes / (k1_c3_ms * 1e-3) / 1e9 / hbm_peak, "bytes": k1c3_bytes, "ms": k1_c3_ms,
                                      "peak_source": peak_src}
            # CPU baseline: the reference's own code on this box's host cores, on the whole C2 batch
            from oracle.bindings import RefOracle, best_oracle
            cores = host_cores()
            cq = reference_quads(1)
            o2 = best_oracle()                                             # oracle/_ref at -O2 (else the C port)
            o2_v, o2_s, o2_maps = cpu_reference_run(o2, cq, cores)
            fast, o3_v, o3_s, o3_maps = o2, o2_v, o2_s, o2_maps
            if RefOracle.o3_available():                                   # the same sources at -O3 / AVX2
                fast = RefOracle(o3=True)
                o3_v, o3_s, o3_maps = cpu_reference_run(fast, cq, cores)
            one_v, one_s, _ = cpu_reference_run(fast, cq[:1024], 1)         # what the reference itself does: one thread
            same_bytes = exact_maps_host.tobytes() == np.ascontiguousarray(o2_maps).tobytes()
            line["parity"]["exact_mode_bytes_equal_reference_cpu"] = same_bytes
            line["cpu_baseline"] = {
                "value": o3_v, "unit": "vertices/s", "cores": cores, "kind": fast.kind,
                "sample": f"{len(cq)} quads x {DIM}^2 (the whole C2 batch), GenerateHeightMap only, {o3_s:.2f} s wall, "
                          f"built {getattr(fast, 'flags', 'gcc -O2 -ffp-contract=off (C port)')}",
                "value_O2": o2_v, "flags_O2": getattr(o2, "flags", "gcc -O2 -ffp-contract=off (C port)"),
                "O3_bytes_equal_O2": bool(np.ascontiguousarray(o3_maps).tobytes() == np.ascontiguousarray(o2_maps).tobytes()),
                "single_thread_value": one_v, "single_thread_sample": f"1024 quads x {DIM}^2, {one_s:.2f} s wall",
                "same_bytes_as_gpu_exact_mode": same_bytes}
        else:
            line.pop("roofline_k1")                                          # the timed K1 segment at N > 1 is the quads half only
            line["e2e"]["per_rank_ms"] = [float(x) for x in per_rank[:, 5]]
            line["per_rank_step_ms"] = [float(x) for x in per_rank[:, 0]]
            gather_bytes_in = (Q - nq) * DIM * DIM * 4                       # what every GPU must receive per step
            line["gather"] = {
                "method": "fused into K2 and K3: K2 sends every finished 128-sample tile as one 512-byte bulk copy (cp.async.bulk) to "
                          "this GPU's buffer and to the same offset of every peer's CUDA-IPC-mapped buffer over NVLink; "
                          f"every {shade_share}-th map is left to K3, which sends it as one 4 KB bulk copy per peer from the "
                          "shared-memory copy it shades from; arrival and release are flags written GPU to GPU, two gathered "
                          "buffers, no host barrier in the step",
                "mode": gather_mode, "shade_share": shade_share if gather_mode == "fused" else 0,
                "k2_chunks": len(chunks) if gather_mode == "push" else 1,
                "bytes_received_per_gpu": gather_bytes_in,
                # roofline of the fused compute + transfer step as B200_PROFILING.md defines it: the slower of the compute
                # alone and the bytes that must cross NVLink at the measured per-direction rate (758 GB/s of payload: 900 raw
                # minus the 18.75 % protocol share ncu counts for these writes, profiles/r02zf_fused_gather_nvlink_bytes.txt;
                # the guide's measured peer copy is 770)
                "roofline_k4": {"bound": "nvlink" if gather_bytes_in / 758e9 * 1e3 > compute_ms else "compute",
                                "bytes_received_per_gpu": gather_bytes_in, "peak": 758.0, "unit": "GB/s",
                                "transfer_alone_ms": gather_bytes_in / 758e9 * 1e3, "compute_alone_ms": compute_ms,
                                "target_ms": max(gather_bytes_in / 758e9 * 1e3, compute_ms), "achieved_ms": ms_step,
                                "frac": max(gather_bytes_in / 758e9 * 1e3, compute_ms) / ms_step,
                                "peak_source": "ncu nvltx__bytes: payload / wire bytes of the kernel's own peer writes x 900 GB/s nominal"},
                "nvlink_floor_ms": gather_bytes_in / 660e9 * 1e3,
                "nvlink_floor_note": "bytes every GPU must receive / 660 GB/s: what this box's NVLink sustains per direction when all "
                                     "GPUs push to all peers with no arithmetic in the way (profiles/r02l_push_only_8gpu.txt; one "
                                     "GPU pushing alone: 700, copy engines all-to-all: 480)",
                "identical_to_nccl_all_gather": nccl_bad == 0.0,
                "identical_to_single_gpu_buffer": single_bad == 0.0,
                "compute_only": {"ms_per_step": compute_ms, "value": total_verts / (compute_ms * 1e-3),
                                 "what": "the same step without K4 (K1 + K2 + K3 on the rank's shard)"},
                "nccl": {"ms_per_step": nccl_ms, "value": total_verts / (nccl_ms * 1e-3),
                         "what": "K1 + K2 + K3, then ncclAllGather of the height maps (planet_gpu_gather_nccl)"}}
        emit(line)
    if gather is not None:
        del local_maps
        gather.close()
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
