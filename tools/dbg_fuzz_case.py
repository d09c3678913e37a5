"""Replays the draws of tests/fuzz_parity.py (FUZZ_LONG=1, seed 9001) up to case 190 -- the N = 1 shape whose emission
launch failed with `invalid argument` (48 KB of dynamic shared memory without the opt-in) -- and runs the stages of the
pipeline one by one with CUDA_LAUNCH_BLOCKING=1.  Kept as the template for the next failure the fuzz finds."""
import os, sys
os.environ["FUZZ_LONG"] = "1"
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np
from dnbc_scala_b200 import _native as nat
from oracle import oracle
from tests.util import ctx_for, perturb, random_model, sample, upload
rng = np.random.default_rng(9001)
LIST = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 17, 31, 32, 33, 48, 63, 64, 65, 70, 96, 100, 128, 130, 160, 200, 256, 300, 320]
for case in range(191):
    N = int(rng.choice(LIST)); D = int(rng.integers(0, 4)); C = int(rng.integers(0, 4)) if D else int(rng.integers(1, 4))
    V = [int(rng.integers(2, 12)) for _ in range(D)]; K = [int(rng.integers(1, 9)) for _ in range(C)]
    tps = int(rng.integers(1, min(N, 12) + 1)); mseed = int(rng.integers(1 << 30)); uv = bool(rng.integers(2))
    nseq = int(rng.integers(1, 50)); tmax = int(rng.choice([1, 3, 20, 90, 400]))
    lengths = [int(x) for x in rng.integers(0, tmax + 1, nseq)]
    if sum(lengths) == 0: lengths[0] = 2
    sseed = int(rng.integers(1 << 30)); pseed = int(rng.integers(1 << 30))
print("case", case, N, V, K, tps, lengths)
m = random_model(oracle, N, V, K, mseed, transitions_per_state=tps, unit_var=uv)
data = sample(oracle, m, lengths, sseed); q = perturb(oracle, m, pseed)
def stage(name, fn):
    try:
        r = fn(); print("ok  ", name, flush=True); return r
    except Exception as e:
        print("FAIL", name, e, flush=True)
for fused in (2, 0):
    with ctx_for(nat, q) as ctx:
        ctx.set_option(nat.OPT_FUSED_ESTEP, fused)
        upload(ctx, data, labels=False)
        print("fused option", fused)
        stage("emission_logprob f32", lambda: ctx.emission_logprob(nat.PREC_F32) if hasattr(ctx, "emission_logprob") else None)
        stage("em_estep", ctx.em_estep)
        stage("em_get_stats", ctx.em_get_stats)
        stage("em_posteriors", ctx.em_posteriors)
        stage("decode f32", lambda: ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32))
