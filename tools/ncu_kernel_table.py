#!/usr/bin/env python3
"""Per-kernel table from an `ncu --set full` report: duration, DRAM bytes, throughput, occupancy, instructions.
usage: ncu_kernel_table.py report.ncu-rep [json_out]"""
import csv, io, json, re, subprocess, sys
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
     "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, units, rows = r[0], r[1], r[2:]
ix = {n: i for i, n in enumerate(h)}
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def val(row, m):
    if m not in ix: return None
    try: v = float(row[ix[m]].replace(",", ""))
    except ValueError: return None
    return v * scale.get(units[ix[m]], 1.0)
tab = []
for row in rows:
    name = re.sub(r"\(.*", "", row[ix["Kernel Name"]]).replace("colate::", "").replace("void ", "")
    tab.append({"kernel": name, **{m: val(row, m) for m in M}})
print("| kernel | grid x block | regs | us | DRAM read MB | DRAM write MB | DRAM % of peak | SM % | warps active % | issue active % | warp insts |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for t in tab:
    print("| %s | %d x %d | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.3g |" % (
        t["kernel"], t["launch__grid_size"], t["launch__block_size"], t["launch__registers_per_thread"], t["gpu__time_duration.sum"],
        t["dram__bytes_read.sum"] / 1e6, t["dram__bytes_write.sum"] / 1e6, t["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"] or 0,
        t["sm__throughput.avg.pct_of_peak_sustained_elapsed"] or 0, t["sm__warps_active.avg.pct_of_peak_sustained_active"] or 0,
        t["smsp__issue_active.avg.pct_of_peak_sustained_active"] or 0, t["smsp__inst_executed.sum"] or 0))
if len(sys.argv) > 2:
    json.dump(tab, open(sys.argv[2], "w"), indent=1)
