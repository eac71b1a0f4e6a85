#!/bin/bash
# usage: tools/ncu_rows.sh N  -> key metrics of the flux + rows kernels
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:fv1_ -s 2 -c 2 --csv python tools/prof_one.py hex $1 lps fields gather 2>/dev/null | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[0].isdigit(): print(r[4][:24], r[12][:44], r[14])
"
