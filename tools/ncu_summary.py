"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] [--json out.json]
--json writes the DRAM traffic of the first matching launch (what bench.py's roofline.traffic reads from
profiles/*k2_dram_traffic.json)."""
import csv, io, json, os, subprocess, sys, time
args = [a for a in sys.argv[1:] if a != "--json"]
json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
if json_out: args.remove(json_out)
rep = args[0]; filt = args[1] if len(args) > 1 else ""
wrote = False
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
stall = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if filt not in name: continue
    print("==", name[:90])
    if json_out and not wrote:
        get = lambda k: float(r[hdr.index(k)].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = get("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = get("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
        json.dump({"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "gpu_time_duration": r[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
                   "source": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on python tools/profile_step.py 1",
                   "written": time.strftime("%Y-%m-%d")}, open(json_out, "w"), indent=1)
        wrote = True
    for k in want:
        if k in hdr: print(f"  {k:78s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = sorted(((float(r[hdr.index(k)] or 0), k) for k in stall if r[hdr.index(k)] not in ("", "n/a")), reverse=True)[:9]
    for v, k in st: print(f"  stall {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):40s} {v:8.3f} warps/issue")
