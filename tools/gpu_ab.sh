#!/bin/bash
# round-2 GPU call AB: ncu of k_xi_tc on a partial-tile state count (N = 320) against a full-tile one (N = 384)
mkdir -p gpurun_out
for n in 320 384; do
  CMD="python tools/kbench.py custom:N=$n,T=200 6000 1"
  timeout 300 $CMD > gpurun_out/ab_plain_$n.log 2>&1 || { tail -5 gpurun_out/ab_plain_$n.log; exit 1; }
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_xi_tc" -s 1 -c 1 -o gpurun_out/prof_ab_$n -f $CMD > gpurun_out/ab_ncu_$n.log 2>&1
  tail -1 gpurun_out/ab_ncu_$n.log
  python tools/ncu_summary.py gpurun_out/prof_ab_$n.ncu-rep --ops > gpurun_out/ab_ncu_xi_$n.txt; head -40 gpurun_out/ab_ncu_xi_$n.txt
done
find gpurun_out -name "*.ncu-rep" -size +30M -delete
