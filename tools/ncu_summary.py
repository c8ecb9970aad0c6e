"""Summarise gpurun_out/*.ncu-rep + launches.csv into small committed files under profiles/.
  python tools/ncu_summary.py <round-tag> [rep-name]"""
import csv
import io
import subprocess
import sys

tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else "prof_conv_tc"
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
raw = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(f"profiles/{tag}_{rep}_ncu_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
print("wrote", f"profiles/{tag}_{rep}_ncu_summary.csv", len(rows) - 2, "kernels")
# launch list: id, kernel, grid, duration
try:
    lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
    out = [("id", "kernel", "grid", "duration_ns")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            out.append((row["ID"], row["Kernel Name"].split("(")[0].replace("<unnamed>::", ""), row["Grid Size"],
                        row["Metric Value"].replace(",", "")))
    with open(f"profiles/{tag}_launches.csv", "w", newline="") as f:
        csv.writer(f).writerows(out)
    print("wrote", f"profiles/{tag}_launches.csv", len(out) - 1, "launches")
except FileNotFoundError:
    pass
