"""Turns gpurun_out/<round>_launches.csv and <round>_window.ncu-rep into committed summaries under
profiles/ (run here, no GPU needed):
  profiles/<round>_launches.txt      per-kernel share of one bench run (ncu gpu__time_duration)
  profiles/<round>_window_kernels.txt key metrics of the spread / gather kernels (ncu --set full)
  profiles/<round>_roofline.json     dram traffic per launch of the dominant kernel (for bench.py)
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

# ---- launch list
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{R}_launches.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 2:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, f"{R}_launches.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-extras\n")
    f.write(f"# {len(data)} launches captured; per-launch times are cold-cache and serialised: compare SHARES.\n")
    f.write(f"{'kernel':80s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:80]:80s} {c:8d} {t/1e6:10.3f} {t/c/1e3:10.1f} {100*t/tot:6.1f}%\n")
print(open(os.path.join(OUT, f"{R}_launches.txt")).read())

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", f"{R}_window.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: j for j, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
roof = {}
with open(os.path.join(OUT, f"{R}_window_kernels.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:spread|gather -s 6 -c 2  python bench.py --steps 2 --warmup 3 --no-extras\n")
    f.write("# workload c4: 3D, N=128, m=4, n=2^24 uniform points, batch_size=4, 1 channel (B200)\n")
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        f.write(f"\n== {name}\n")
        for k in KEYS:
            if k in col:
                f.write(f"   {k:72s} {r[col[k]]:>20s} {units[col[k]]}\n")
        st = sorted(((float(r[col[h]] or 0), h) for h in STALLS), reverse=True)[:8]
        f.write("   warp stalls per issued instruction: " + ", ".join(
            f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st) + "\n")
        def num(k):
            v, u = float(r[col[k]].replace(",", "")), units[col[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        roof[name] = {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                      "duration_ms_under_ncu": float(r[col["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "s": 1e3}[units[col["gpu__time_duration.sum"]]]}
print(open(os.path.join(OUT, f"{R}_window_kernels.txt")).read())
json.dump({"source": f"profiles/{R}_window_kernels.txt (ncu --set full, one launch each)", "kernels": roof},
          open(os.path.join(OUT, f"{R}_roofline.json"), "w"), indent=1)
