import csv, sys, subprocess
rep = sys.argv[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct','fp64','warps_active','registers_per_thread','bank_conflicts','lts__t_sector_hit_rate','issue_active','smsp__issue','stall','warp_issue_stalled', 'smsp__average_warp', 'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'occupancy']
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==== ', r[hdr.index('Kernel Name')][:60], 'id', r[0])
    for i,h in enumerate(hdr):
        if any(p in h for p in pats):
            print('  %-90s %s %s' % (h, r[i], units[i]))
