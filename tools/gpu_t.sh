#!/bin/bash
# round-2 GPU call T: ncu --set full of the emission / statistics sweeps at the config-5 (N = 512) and config-3 (N = 10)
# shapes; summaries are made on the box (the reports together exceed what gpurun copies back)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:k_emission|k_stats" -c 2 -o gpurun_out/prof_r02t_n512 -f python tools/kbench.py largeN 12500 1 > gpurun_out/t_ncu1.log 2>&1; tail -1 gpurun_out/t_ncu1.log
ncu --set full --clock-control none --import-source on -k "regex:k_emission|k_stats|k_forward2|k_backward2" -c 4 -o gpurun_out/prof_r02t_n10 -f python tools/kbench.py gmm100k 100000 1 2 > gpurun_out/t_ncu2.log 2>&1; tail -1 gpurun_out/t_ncu2.log
python tools/ncu_summary.py gpurun_out/prof_r02t_n512.ncu-rep --ops > gpurun_out/r02t_ncu_n512_summary.txt
python tools/ncu_summary.py gpurun_out/prof_r02t_n10.ncu-rep --ops > gpurun_out/r02t_ncu_n10_summary.txt
for r in n512 n10; do ncu -i gpurun_out/prof_r02t_$r.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/r02t_src_$r.csv.gz; done
ls -la gpurun_out/
find gpurun_out -name "*.ncu-rep" -size +20M -delete
