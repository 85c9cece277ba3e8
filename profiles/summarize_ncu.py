"""Turns `ncu -i X.ncu-rep --page raw --csv` (stdin) into the compact per-kernel table committed here.
    ncu -i gpurun_out/r1_prof.ncu-rep --page raw --csv | python profiles/summarize_ncu.py > profiles/r1_ncu_summary.csv"""
import csv, sys
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"]
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] for i in idx])
