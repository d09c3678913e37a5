#!/bin/bash
# round-2 last ncu capture: the tcgen05 kernels of the N = 512 configuration at HEAD
mkdir -p gpurun_out
CMD="python tools/kbench.py largeN 12500 1"
timeout 120 $CMD > gpurun_out/f5_plain.log 2>&1 || { tail -5 gpurun_out/f5_plain.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_fb_tc|k_xi_tc" -s 3 -c 3 -o gpurun_out/prof_r02fin_tc -f $CMD > gpurun_out/f5_ncu.log 2>&1
tail -1 gpurun_out/f5_ncu.log
python tools/ncu_summary.py gpurun_out/prof_r02fin_tc.ncu-rep --ops > gpurun_out/r02fin_ncu_tc_summary.txt
ncu -i gpurun_out/prof_r02fin_tc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]
keys = [k for k in h if 'pipe_tensor' in k and 'pct' in k or k == 'lts__t_sector_hit_rate.pct' or k == 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:60], {k: r[h.index(k)] for k in keys})
" >> gpurun_out/r02fin_ncu_tc_summary.txt
head -12 gpurun_out/r02fin_ncu_tc_summary.txt | cut -c1-330; tail -4 gpurun_out/r02fin_ncu_tc_summary.txt | cut -c1-400
find gpurun_out -name "*.ncu-rep" -size +30M -delete
