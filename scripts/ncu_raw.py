"""key raw metrics of an ncu report; usage: ncu_raw.py rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'sm__cycles_active.avg']
for r in rows[2:]:
    print({k: r[h.index(k)] for k in keys if k in h})
