#!/usr/bin/env python
"""Turns gpurun_out/prof_<tag>.ncu-rep + launches_bench_<tag>.csv into the tracked summaries
under profiles/ (run in the build container; ncu reads reports without a GPU)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "c2"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"
# usage: summarize_profile.py <tag> <round> [raw.csv | report.ncu-rep] [launches.csv] [command text]
src = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
launches_csv = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", f"launches_bench_{tag}.csv")
cmd_text = sys.argv[5] if len(sys.argv) > 5 else f"bin/kbench {tag} --profile --reps 3 --warmup 1"
launch_cmd = sys.argv[6] if len(sys.argv) > 6 else f"python bench.py --workload {tag} --steps 20 --warmup 5 --no-cpu"
out_dir = os.path.join(ROOT, "profiles")

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "sm__cycles_elapsed.max"]

if src.endswith(".csv"):      # `ncu -i report --page raw --csv` already run on the GPU box (reports > 6 MB stay there)
    raw = open(src).read()
else:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(os.path.join(out_dir, f"{rnd}_{tag}_ncu_raw_selected.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])

# per-kernel means
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "")
    agg.setdefault(name, []).append(r)


def mean(rs, key):
    i = hdr.index(key)
    return sum(float(r[i].replace(",", "")) for r in rs) / len(rs)


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


traffic = {}
lines = [f"# ncu --set full summary, {tag}, round {rnd}",
         "",
         f"Command: `ncu --set full --clock-control none --import-source on -k regex:... {cmd_text}` "
         "(after the same command exited 0 without ncu).",
         "Times under ncu are cold-cache and serialised (compare shares / traffic, not absolutes).",
         "",
         "| kernel | launches | time us | dram read MB | dram write MB | traffic MB | dram % of ncu peak | L2 hit % | L1 hit % | warps active % | regs | grid x block |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, rs in agg.items():
    ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
    rd, wr = to_bytes(mean(rs, "dram__bytes_read.sum"), ur), to_bytes(mean(rs, "dram__bytes_write.sum"), uw)
    traffic[name] = rd + wr
    lines.append(f"| `{name}` | {len(rs)} | {mean(rs, 'gpu__time_duration.sum'):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                 f"{(rd + wr) / 1e6:.1f} | {mean(rs, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{mean(rs, 'lts__t_sector_hit_rate.pct'):.1f} | {mean(rs, 'l1tex__t_sector_hit_rate.pct'):.1f} | "
                 f"{mean(rs, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{int(mean(rs, 'launch__registers_per_thread'))} | {int(mean(rs, 'launch__grid_size'))} x {int(mean(rs, 'launch__block_size'))} |")

# launch list of bench.py
lpath = launches_csv
if os.path.exists(lpath):
    lrows = [r for r in csv.reader(open(lpath)) if len(r) > 5]
    lh = lrows[0]
    ki, vi = lh.index("Kernel Name"), lh.index("Metric Value")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in lrows[1:]:
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:80]
        tot[n] = tot.get(n, 0.0) + float(r[vi].replace(",", ""))
        cnt[n] += 1
    s = sum(tot.values())
    lines += ["", f"## launch list of `{launch_cmd}` "
              "(`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
              "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for n, v in tot.items():
        lines.append(f"| `{n}` | {cnt[n]} | {v / 1e3:.1f} | {v / cnt[n] / 1e3:.2f} | {100 * v / s:.1f}% |")
    import shutil
    shutil.copy(lpath, os.path.join(out_dir, f"{rnd}_{tag}_launches_bench.csv"))

with open(os.path.join(out_dir, f"{rnd}_{tag}_ncu_summary.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
tpath = os.path.join(out_dir, "traffic.json")
allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
allt[f"{rnd}_{tag}"] = traffic
json.dump(allt, open(tpath, "w"), indent=1)
print("\n".join(lines))
