"""Condense one `ncu --set full` report into the text summary committed under profiles/.
usage: ncu_summary.py <report.ncu-rep> "<command line that was profiled>" "<note>" > profiles/rNN_<name>_ncu_full.txt"""
import csv
import io
import subprocess
import sys

rep, cmd, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
WANT = ["dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {k: i for i, k in enumerate(hdr)}
print(f"ncu --set full --clock-control none --import-source on -k regex:... -c 1   {cmd}   {note}")
print("(per-launch numbers under ncu: cold cache, serialised; bench values are never taken under ncu)\n")
for k in WANT:
    if k in col:
        print(f"{k:78s} {vals[col[k]]} {units[col[k]]}")
print(f"{'kernel':78s} {vals[col['Kernel Name']]}")
try:
    def num(k):
        v = float(vals[col[k]].replace(",", ""))
        u = units[col[k]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}.get(u, 1.0)
    t = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    print(f"\ntraffic = dram__bytes_read.sum + dram__bytes_write.sum = {t / 1e9:.4f} GB per launch ({t:.0f} B)")
except Exception as e:      # noqa: BLE001
    print("traffic: n/a", e)
