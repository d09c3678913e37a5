#!/bin/bash
# round-2 final 1-GPU call: GPU suite, bench (both arms), the profiling recipe at HEAD, all-configuration timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not two_gpus" > gpurun_out/r02fin_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02fin_pytest_gpu.log; tail -3 gpurun_out/r02fin_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/f_bench.log | tail -1 > gpurun_out/r02fin_bench_1gpu.json; cut -c1-600 gpurun_out/r02fin_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.log 2> gpurun_out/f_ref.err; grep '^{' gpurun_out/f_ref.log | tail -1 > gpurun_out/r02fin_bench_reference.json; cut -c1-400 gpurun_out/r02fin_bench_reference.json
# profiling recipe (B200_PROFILING.md): plain run, ncu launch list of the same command, one --set full capture
CMD="python bench.py --steps 1 --warmup 1 --seqs 8000 --no-e2e --no-cpu --no-decode --no-extras"
$CMD > gpurun_out/f_plain.log 2>&1 || { tail -5 gpurun_out/f_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02fin.csv $CMD > gpurun_out/f_ncu_l.log 2>&1
python tools/launches_summary.py gpurun_out/launches_r02fin.csv "$CMD" > gpurun_out/r02fin_launches_summary.txt; head -14 gpurun_out/r02fin_launches_summary.txt
ncu --set full --clock-control none --import-source on -k "regex:k_stats2|k_emission_px|k_forward3|k_backward3|k_xi_tc" -s 5 -c 5 -o gpurun_out/prof_r02fin -f $CMD > gpurun_out/f_ncu_f.log 2>&1
tail -1 gpurun_out/f_ncu_f.log
python tools/ncu_summary.py gpurun_out/prof_r02fin.ncu-rep --ops > gpurun_out/r02fin_ncu_summary.txt; head -16 gpurun_out/r02fin_ncu_summary.txt
# every BASELINE configuration: per-kernel ms of one EM iteration and decode passes
for w in "readme 1000" "robot 200" "gmm100k 100000" "scaleout 32000" "largeN 12500"; do python tools/kbench.py $w 3 >> gpurun_out/r02fin_all_configs.jsonl 2>> gpurun_out/f_err.log; done
for w in "readme 200" "gmm100k 100000" "scaleout 32000" "largeN 125000"; do python tools/dbench.py $w >> gpurun_out/r02fin_all_configs.jsonl 2>> gpurun_out/f_err.log; done
cat gpurun_out/r02fin_all_configs.jsonl | cut -c1-400
find gpurun_out -name "*.ncu-rep" -size +30M -delete
# in-between state counts (tcgen05 recursions / xi / statistics at any N from 65 to 512)
for w in "custom:N=100,T=200 50000" "custom:N=128,T=200 50000" "custom:N=200,T=200 25000" "custom:N=300,T=200 12500" "custom:N=399,T=200 12500" "custom:N=500,T=200 12500"; do python tools/kbench.py $w 2 >> gpurun_out/r02fin_any_n.jsonl 2>> gpurun_out/f_err.log; done
cut -c1-300 gpurun_out/r02fin_any_n.jsonl
