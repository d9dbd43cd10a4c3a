"""One-page summary of an .ncu-rep (raw page): python tools/ncu_summary.py rep.ncu-rep
With `--json KEY FILE`: also merges {KEY: {dram_bytes_per_launch, duration_us, source}} into FILE
(profiles/r2_ncu_traffic.json, which bench.py reads for `roofline.traffic`)."""
import csv, json, os, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("%-82s %s %s" % (k, r[i], units[i]))
    print()

if "--json" in sys.argv:
    q = sys.argv.index("--json")
    key, path = sys.argv[q + 1], sys.argv[q + 2]
    r = rows[2]
    val = lambda k: float(r[hdr.index(k)].replace(",", ""))
    unit = lambda k: units[hdr.index(k)]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(val(k) * scale[unit(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    dur = val("gpu__time_duration.sum") * tscale[unit("gpu__time_duration.sum")]
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"dram_bytes_per_launch": dram, "duration_us": dur,
              "kernel": r[hdr.index("Kernel Name")][:120],
              "source": "%s, ncu --set full --clock-control none" % os.path.basename(sys.argv[1])}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
