#!/bin/bash
# round-2 GPU call D4: lane-per-sequence forward / backward, branch-free copies and steps
mkdir -p gpurun_out
rm -f gpurun_out/d2_kbench.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x -k "lane_per_sequence or randomised or fused_and_general" > gpurun_out/d2_pytest.log 2>&1; tail -4 gpurun_out/d2_pytest.log
for n in 100000 25000 12500; do
  for mode in 1 2; do
    DNBC_FB_SEQ_LANES=$mode python tools/kbench.py gmm100k $n 3 2 >> gpurun_out/d2_kbench.jsonl 2>> gpurun_out/d2_err.log
  done
  DNBC_B200_LIB=build/libdnbc_fbsareg.so DNBC_FB_SEQ_LANES=1 python tools/kbench.py gmm100k $n 3 2 >> gpurun_out/d2_kbench.jsonl 2>> gpurun_out/d2_err.log
done
for w in "readme 16000 3 2" "robot 16000 3 2"; do
  for mode in 1 2; do
    DNBC_FB_SEQ_LANES=$mode python tools/kbench.py $w >> gpurun_out/d2_kbench.jsonl 2>> gpurun_out/d2_err.log
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/d2_kbench.jsonl'):
    d = json.loads(l); m = d['ms_per_iter']; print(d['lib'], d['workload'], d['seqs'], 'fwd', m['forward'], 'bwd', m['backward'], 'total', m['total'], d['ll'])
PY
DNBC_FB_SEQ_LANES=1 ncu --set full --clock-control none --import-source on -k "regex:k_forward_s|k_backward_s" -c 2 -o gpurun_out/prof_r02d_fbs -f python tools/kbench.py gmm100k 100000 1 2 > gpurun_out/d2_ncu.log 2>&1; tail -1 gpurun_out/d2_ncu.log
python tools/ncu_summary.py gpurun_out/prof_r02d_fbs.ncu-rep --ops > gpurun_out/r02d_ncu_fbs_summary.txt; head -8 gpurun_out/r02d_ncu_fbs_summary.txt; grep -A13 "^void.*total executed" gpurun_out/r02d_ncu_fbs_summary.txt | head -60
ncu -i gpurun_out/prof_r02d_fbs.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/r02d_src_fbs.csv.gz
find gpurun_out -name "*.ncu-rep" -size +20M -delete
