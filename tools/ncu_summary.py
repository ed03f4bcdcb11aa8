"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/."""
import csv, subprocess, sys
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpc__cycles_elapsed.avg.per_second', 'sm__cycles_elapsed.avg',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum', 'launch__cluster_x', 'sm__inst_executed.sum',
        'smsp__inst_executed.avg.per_cycle_active']
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if h in KEEP]
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
for r in rows[2:]:
    d = {hdr[i]: r[i] for i in idx}
    print(d['Kernel Name'][:48], '| us', d.get('gpu__time_duration.sum'), '| tensor% elapsed', d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
          '| dram MB r/w', d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum'), '| dram%', d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
          '| GHz', d.get('gpc__cycles_elapsed.avg.per_second'), '| regs', d.get('launch__registers_per_thread'))
