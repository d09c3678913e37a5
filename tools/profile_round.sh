set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r01f_pytest.log 2>&1
python bench.py > gpurun_out/r01f_bench_1gpu.json 2> gpurun_out/r01f_bench.err
python bench.py --impl reference > gpurun_out/r01f_bench_reference.json 2>> gpurun_out/r01f_bench.err
CMD="python bench.py --steps 1 --warmup 1 --seqs 8000 --no-e2e --no-cpu --no-decode"
$CMD > gpurun_out/r01f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv $CMD > gpurun_out/r01f_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_stats_pk|k_emission_pk|k_forward3|k_backward3|k_xi_tc" -s 5 -c 5 -o gpurun_out/prof_r01f -f $CMD > gpurun_out/r01f_ncu_f.log 2>&1
python tools/kbench.py largeN 18944 1 > gpurun_out/r01f_plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_fb_tc|k_xi_tc" -s 2 -c 3 -o gpurun_out/prof_r01f_tc -f python tools/kbench.py largeN 18944 1 > gpurun_out/r01f_ncu_tc.log 2>&1
for w in "readme 1000" "robot 200" "gmm100k 100000" "scaleout 125000" "largeN 12500" "largeN 100000"; do python tools/kbench.py $w 3; done > gpurun_out/r01f_all_configs.jsonl 2> gpurun_out/r01f_all_configs.err
for w in "readme 200" "gmm100k 100000" "scaleout 125000" "largeN 125000"; do python tools/dbench.py $w; done >> gpurun_out/r01f_all_configs.jsonl 2>> gpurun_out/r01f_all_configs.err
tail -3 gpurun_out/r01f_pytest.log
