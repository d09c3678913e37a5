"""Per-kernel-family timings of one EM iteration for A/B library builds.
usage: DNBC_B200_LIB=build/libdnbc_x.so python tools/kbench.py [workload] [seqs] [iters] [fused mode 0|1|2]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnbc_scala_b200 import _native as nat  # noqa: E402
from dnbc_scala_b200 import workloads  # noqa: E402

wl = workloads.by_name(sys.argv[1] if len(sys.argv) > 1 else "scaleout")
seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 32000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
truth = workloads.true_params(wl)
init = workloads.perturbed_params(wl, truth)
ctx = nat.Context(wl.N, [wl.V] * wl.D, [wl.K] * wl.C, device=0)
if len(sys.argv) > 4:
    ctx.set_option(nat.OPT_FUSED_ESTEP, int(sys.argv[4]))      # 1 = fused E-step kernel, 2 = general pipeline
if os.environ.get("DNBC_FB_SEQ_LANES"):
    ctx.set_option(nat.OPT_FB_SEQ_LANES, int(os.environ["DNBC_FB_SEQ_LANES"]))   # 1 = lane per sequence, 2 = lane per state
ctx.set_params(truth["pi"], truth["A"], truth["B"], truth["w"], truth["mu"], truth["var"])
ctx.generate(seqs, wl.T, 20260000 + wl.index * 1000)
ctx.set_params(init["pi"], init["A"], init["B"], init["w"], init["mu"], init["var"])
ll = ctx.em_iterate(2)
ctx.profile_enable(True)
ll = ctx.em_iterate(iters)
ms, _ = ctx.profile_read()
ms = {k: round(v / iters, 3) for k, v in ms.items() if v}
ms["total"] = round(sum(ms.values()), 3)
print(json.dumps({"lib": os.path.basename(nat.LIB_PATH), "workload": wl.name, "seqs": seqs, "ll": float(ll[-1]),
                  "ms_per_iter": ms}))
