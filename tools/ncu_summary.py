"""Summarise an .ncu-rep (raw page + SASS hot spots) -- used to produce profiles/*.txt."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 20e6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sectors.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:3]:
    for k in keys:
        if k in hdr:
            print("%-62s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
            except ValueError:
                pass
    print("stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) >= len(hdr) and r[0] != "Address":
        data.append(r)
tot_e = sum(int(r[iexec]) for r in data)
tot_s = sum(int(r[isamp]) for r in data)
print("SASS: %d instructions, %.1f warp-instr per interaction" % (len(data), tot_e / per))
c, s = Counter(), Counter()
for r in data:
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    c[op] += int(r[iexec])
    s[op] += int(r[isamp])
print("opcode mix per interaction:", ", ".join("%s %.1f" % (o, v / per) for o, v in c.most_common(14)))
print("top stall-sample instructions:")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    print("  %5.2f%%  x%.2f/int  %s" % (100 * int(r[isamp]) / tot_s, int(r[iexec]) / per, r[isrc].strip()[:80]))
