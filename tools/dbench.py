"""Decode (Viterbi) timing: DNBC_B200_LIB=... python tools/dbench.py [workload] [seqs]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dnbc_scala_b200 import _native as nat, workloads
wl = workloads.by_name(sys.argv[1] if len(sys.argv) > 1 else "scaleout")
seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 32000
truth = workloads.true_params(wl)
ctx = nat.Context(wl.N, [wl.V] * wl.D, [wl.K] * wl.C, device=0)
ctx.set_params(truth["pi"], truth["A"], truth["B"], truth["w"], truth["mu"], truth["var"])
ctx.generate(seqs, wl.T, 4321)
path, score, tie = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32, want_tie=True)
ctx.profile_enable(True)
t0 = time.perf_counter()
for _ in range(3):
    ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32, want_score=True, want_tie=False)
dt = (time.perf_counter() - t0) / 3
ms, _ = ctx.profile_read()
_, _, _, labels = ctx.download(0, min(seqs, 256))
acc = float((path[:len(labels)] == labels).mean())
print(json.dumps({"workload": wl.name, "seqs": seqs, "wall_ms_per_pass": round(dt * 1e3, 2), "emission_ms": round(ms["emission"] / 3, 2),
                  "decode_ms": round(ms["decode"] / 3, 2), "seq_per_s": round(seqs / dt), "accuracy_vs_truth": round(acc, 4),
                  "ties": int((tie != 0).sum()), "score_sum": float(score.sum()), "path_crc": int(np.bitwise_xor.reduce(path.astype(np.int64) * (np.arange(path.size) % 65521)))}))
