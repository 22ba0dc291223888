"""Compact per-launch summary of an .ncu-rep (read on the build box, no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.csv"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'smsp__cycles_active.avg']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
idx = [h.index(w) for w in WANT if w in h]
w = csv.writer(sys.stdout)
w.writerow([h[i] for i in idx])
w.writerow([rows[1][i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
