"""Turns the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.
usage: [PROFILE_WORKLOAD=cfg2|cfg4|sinc] python tools/summarize_profiles.py <round tag> <launch csv> <rep:label> [<rep:label> ...]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
wl = os.environ.get("PROFILE_WORKLOAD", "cfg2")
reps = sys.argv[3:]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    v = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.csv"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --workload {wl} --steps 1 --warmup 1 --no-cpu-baseline\n")
    f.write(f"# cold-cache, serialised per-launch times: compare SHARES, not absolutes ({wl} workload, first 400 launches)\n")
    f.write("kernel,launches,total_us,share_pct\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{100 * v[1] / tot:.1f}\n")
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, one launch each (one time block; block size = PB200_TIME_BLOCK default)\n")
    f.write("report,kernel,metric,value,unit\n")
    for spec in reps:
        rep, label = spec.split(":")
        # a report, or the raw page already exported on the GPU box (tools/final_gpu_pass.sh)
        out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        h, r = rr[0], rr[2]
        kn = r[h.index("Kernel Name")]
        for w in want:
            if w in h:
                f.write(f"{label},\"{kn}\",{w},{r[h.index(w)]},{rr[1][h.index(w)]}\n")
print("ok")
