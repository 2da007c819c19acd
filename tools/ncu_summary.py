"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the short text kept under profiles/."""
import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ('Kernel Name', 'dram__bytes', 'gpu__time_duration', 'launch__registers', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'pipe_tensor', 'pipe_fp64', 'sm__throughput', 'issue_active',
        'smsp__inst_executed.sum', 'bank_conflicts', 'warps_active', 'issue_stalled', 'l1tex__data_pipe_lsu_wavefronts',
        'inst_executed_pipe_xu', 'inst_executed_pipe_fp64', 'inst_executed_pipe_lsu', 'inst_executed_pipe_alu', 'inst_executed_pipe_fma')
with open(out, 'w') as f:
    f.write(title + '\n')
    for vals in rows[2:]:
        for h, u, v in zip(hdr, units, vals):
            if any(k in h for k in keep) and not h.startswith(('SM_', 'TPC')) and ('issue_stalled' not in h or 'per_issue_active' in h) \
                    and not any(x in h for x in ('.max.', '.min.', '.sum.p', 'per_second', '.peak_sustained', 'per_cycle_active', 'dram__bytes.')):
                f.write('%s | %s | %s\n' % (h, u, v))
