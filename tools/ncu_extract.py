"""Reduce an `ncu --set full` report to the per-kernel table kept under profiles/: one row per distinct kernel (its last
captured launch) with time, DRAM bytes, L2 / tensor / XU / issue utilisation and registers.
Usage: python tools/ncu_extract.py gpurun_out/hot_kernels.ncu-rep > profiles/rNN_hot_kernels_ncu_full.csv"""
import csv
import io
import subprocess
import sys

COLS = ["launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
last = {}
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("ivv::", "")
    if "--by-grid" in sys.argv:  # one row per launch: the same kernel at several shapes
        name = f"{name} #{len(last)} grid={r[ix['launch__grid_size']]}"
    last[name] = r
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + [f"{c} [{units[ix[c]]}]" if units[ix[c]] else c for c in COLS if c in ix])
for name, r in last.items():
    w.writerow([name] + [r[ix[c]] for c in COLS if c in ix])
