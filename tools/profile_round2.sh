#!/bin/bash
# round-2 profiling recipe (B200_PROFILING.md): plain run, ncu launch list of the same command, one --set full capture
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --seqs 8000 --no-e2e --no-cpu --no-decode --no-extras"
$CMD > gpurun_out/s_plain.log 2>&1 || { tail -5 gpurun_out/s_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02s.csv $CMD > gpurun_out/s_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_stats2|k_emission_pk|k_forward3|k_backward3|k_xi_tc" -s 5 -c 5 -o gpurun_out/prof_r02s -f $CMD > gpurun_out/s_ncu_f.log 2>&1
tail -2 gpurun_out/s_ncu_f.log
python tools/dbench.py largeN 12000 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_viterbi_tile" -c 1 -o gpurun_out/prof_r02s_vt -f python tools/dbench.py largeN 12000 > gpurun_out/s_ncu_vt.log 2>&1
tail -1 gpurun_out/s_ncu_vt.log
