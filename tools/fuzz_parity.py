"""Randomised parity sweep: random shapes / mixtures / ragged lengths, CUDA path vs the fp64 oracle.
usage: python tools/fuzz_parity.py [cases] [seed]   (needs a B200; prints the first failure and exits 1)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnbc_scala_b200 import _native as nat  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.util import ctx_for, perturb, random_model, sample, upload  # noqa: E402


def flagged_or_equal(path, opath, tie, otie, off, backtrace):
    """Paths identical except where a flagged tie explains the difference: at or before the point in
    REFERENCE mode (the recursion is greedy forward), anywhere in the sequence in BACKTRACE mode (a tie
    at step t changes the back-pointers the trace follows to every EARLIER step)."""
    diff = np.nonzero(path.astype(np.int64) != opath.astype(np.int64))[0]
    flagged = (tie | otie) != 0
    seq_of = np.searchsorted(off, diff, side="right") - 1
    return all(flagged[off[s]:(off[s + 1] if backtrace else n + 1)].any() for n, s in zip(diff, seq_of))


def one(case, rng):
    N = int(rng.choice([1, 2, 3, 7, 10, 16, 17, 31, 32, 33, 48, 63, 64, 65, 96, 128, 130, 200, 256, 300]))
    D = int(rng.integers(0, 4))
    C = int(rng.integers(0, 4)) if D else int(rng.integers(1, 4))
    V = [int(rng.integers(2, 12)) for _ in range(D)]
    K = [int(rng.integers(1, 9)) for _ in range(C)]
    tps = int(rng.integers(1, min(N, 12) + 1))
    m = random_model(oracle, N, V, K, int(rng.integers(1 << 30)), transitions_per_state=tps, unit_var=bool(rng.integers(2)))
    nseq = int(rng.integers(1, 50))
    tmax = int(rng.choice([1, 3, 20, 90, 400] if os.environ.get("FUZZ_LONG") else [1, 3, 20, 90]))
    lengths = [int(x) for x in rng.integers(0, tmax + 1, nseq)]
    if sum(lengths) == 0:
        lengths[0] = 2
    data = sample(oracle, m, lengths, int(rng.integers(1 << 30)))
    q = perturb(oracle, m, int(rng.integers(1 << 30)))
    tag = f"case {case}: N={N} V={V} K={K} tps={tps} lengths={lengths[:8]}... P={data.P}"
    want = oracle.estep(q, data)
    if not np.isfinite(want).all():
        return None          # a zero-probability sequence: outside the parity contract
    opath, oscore, otie = oracle.decode(q, data, oracle.DECODE_BACKTRACE)
    rpath, rscore, rtie = oracle.decode(q, data, oracle.DECODE_REFERENCE)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False)
        ctx.em_estep()
        got = ctx.em_get_stats()
        g = ctx.em_posteriors()
        p32, s32, t32 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32)
        p64, s64, t64 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F64)
        rp, rs, rt = ctx.decode(nat.DECODE_REFERENCE, nat.PREC_F64)
    sw, sg = oracle.split_stats(q, want), oracle.split_stats(q, got)
    assert abs(sg["ll"] - sw["ll"]) <= 1e-5 * abs(sw["ll"]) + 1e-6, (tag, "ll", sg["ll"], sw["ll"])
    for key in ("pi", "xi", "gamma", "cat"):
        np.testing.assert_allclose(sg[key], sw[key], rtol=2e-4, atol=2e-4, err_msg=tag + " " + key)
    if C:
        np.testing.assert_allclose(sg["gmm"][..., 0], sw["gmm"][..., 0], rtol=2e-4, atol=2e-3, err_msg=tag + " gmm s0")
    np.testing.assert_allclose(g, oracle.posteriors(q, data), atol=5e-5, err_msg=tag + " gamma")
    if not flagged_or_equal(p32, opath, t32, otie, data.off, True):
        diff = np.nonzero(p32.astype(np.int64) != opath.astype(np.int64))[0]
        n = int(diff[0]); sq = int(np.searchsorted(data.off, n, side="right") - 1)
        lo, hi = int(data.off[sq]), int(data.off[sq + 1])
        print(tag, "first diff at point", n, "sequence", sq, "range", lo, hi)
        print(" gpu f32 path", p32[lo:hi].tolist()); print(" gpu f64 path", p64[lo:hi].tolist()); print(" oracle path ", opath[lo:hi].tolist())
        print(" gpu tie", t32[lo:hi].tolist()); print(" oracle tie", otie[lo:hi].tolist())
        print(" scores", s32[sq], s64[sq], oscore[sq], "active", q.active.tolist())
        raise AssertionError((tag, "f32 backtrace path"))
    assert flagged_or_equal(p64, opath, t64, otie, data.off, True), (tag, "f64 backtrace path")
    assert flagged_or_equal(rp, rpath, rt, rtie, data.off, False), (tag, "reference-mode path")
    fin = np.isfinite(oscore)
    np.testing.assert_allclose(s32[fin], oscore[fin], rtol=1e-5, atol=1e-4, err_msg=tag + " f32 score")
    np.testing.assert_allclose(s64[fin], oscore[fin], rtol=1e-11, atol=1e-9, err_msg=tag + " f64 score")
    return tag


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
    done = skipped = 0
    for case in range(cases):
        if one(case, rng) is None:
            skipped += 1
        else:
            done += 1
    print(f"fuzz ok: {done} cases compared, {skipped} skipped (zero-probability data)")


if __name__ == "__main__":
    main()
