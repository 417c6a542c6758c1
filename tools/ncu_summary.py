"""Summarises an .ncu-rep (ncu --set full capture) into a small CSV + markdown table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([a for a, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
with open(out + ".md", "w") as f:
    f.write("| kernel | grid x block | regs | time | dram read | dram write | dram GB/s (r+w) | dram %% of ncu peak |\n|---|---|---|---|---|---|---|---|\n")
    g = {a: i for a, i in idx}
    for r in rows[2:]:
        def v(k):
            return r[g[k]]
        def tb(k):
            x = float(v(k)); u = units[g[k]]
            return x * (1000.0 if u.startswith("Tbyte") else 1.0 if u.startswith("Gbyte") else 1e-3)
        name = v("Kernel Name").split("(")[0].replace("void ", "")
        f.write("| %s | %s x %s | %s | %s %s | %s %s | %s %s | %.0f | %s |\n" % (
            name, v("launch__grid_size"), v("launch__block_size"), v("launch__registers_per_thread"),
            v("gpu__time_duration.sum"), units[g["gpu__time_duration.sum"]],
            v("dram__bytes_read.sum"), units[g["dram__bytes_read.sum"]], v("dram__bytes_write.sum"), units[g["dram__bytes_write.sum"]],
            tb("dram__bytes_read.sum.per_second") + tb("dram__bytes_write.sum.per_second"),
            v("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
print(open(out + ".md").read())
