"""Summarise an ncu raw CSV export (ncu -i X.ncu-rep --page raw --csv) into a few key metrics."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
    r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__throughput.avg.pct|registers_per_thread|"
    r"sm__warps_active.avg.pct|occupancy_limit|pipe_fp64|pipe_fma.*pct|pipe_alu.*pct|pipe_lsu.*pct|bank_conflicts.*shared.sum|wavefronts_mem_shared.sum$|"
    r"issue_active.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|warp_issue_stalled.*_per_warp_active.pct|smsp__inst_executed.sum$|achieved_occupancy|"
    r"shared_mem_per_block|sm__cycles_elapsed.avg$|inst_executed_op_shared|smsp__cycles_active.avg$")
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, k in enumerate(hdr):
        if pat.search(k):
            print(f"  {k} = {r[i]} {units[i]}")
