"""GPU parity at BASELINE.json's exact model shapes (`-m gpu`), context-reuse regressions and the
library-owned NCCL collective.

Round-1 review: config 4's 16 + 16 variables with 8-component mixtures at T = 1000 and config 5's
N = 512 with 5 + 5 variables at T = 200 were only ever checked through invariants at full size;
here the same shapes (fewer sequences, so the fp64 oracle finishes in seconds) are compared with
the oracle value for value.  Arithmetic being protected: the Gaussian-mixture E/M formulas of
Tools/ExpectationMaximization1D.java:92-132 and the naive-Bayes emission product of
DynamicNaiveBayesianClassifier.scala:52-56 inside the Baum-Welch recursion.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from dnbc_scala_b200 import _native as nat
from dnbc_scala_b200 import workloads
from tests.util import ctx_for, perturb, random_model, sample, upload

pytestmark = pytest.mark.gpu

TOL = 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_of(oracle, wl, p):
    m = oracle.Model(wl.N, [wl.V] * wl.D, [wl.K] * wl.C)
    for k in ("pi", "A", "B", "w", "mu", "var"):
        getattr(m, k)[...] = p[k].reshape(getattr(m, k).shape)
    return m


def _workload_case(oracle, name, S, seed):
    """Truth / perturbed start / S sequences of workload `name` at its exact variable counts."""
    wl = workloads.WORKLOADS[name]
    truth = workloads.true_params(wl)
    off, disc, cont, labels = workloads.sample_host(wl, truth, S, seed)
    data = oracle.Data(off, disc, cont.astype(np.float64), labels)
    q = _model_of(oracle, wl, workloads.perturbed_params(wl, truth))
    return wl, data, q


def _assert_params(ctx, m, rtol, atol):
    p = ctx.get_params()
    np.testing.assert_allclose(p["pi"], m.pi, rtol=rtol, atol=atol, err_msg="pi")
    np.testing.assert_allclose(p["A"], m.A, rtol=rtol, atol=atol, err_msg="A")
    np.testing.assert_allclose(p["B"], m.B, rtol=rtol, atol=atol, err_msg="B")
    for c in range(m.C):
        k = int(m.K[c])
        np.testing.assert_allclose(p["w"][c, :, :k], m.w[c, :, :k], rtol=rtol, atol=atol, err_msg=f"w[{c}]")
        np.testing.assert_allclose(p["mu"][c, :, :k], m.mu[c, :, :k], rtol=rtol, atol=2e-5, err_msg=f"mu[{c}]")
        np.testing.assert_allclose(p["var"][c, :, :k], m.var[c, :, :k], rtol=rtol, atol=2e-5, err_msg=f"var[{c}]")
    return p


def _paths_equal_except_ties(path, opath, tie, otie, off, backtrace):
    diff = np.nonzero(path.astype(np.int64) != opath.astype(np.int64))[0]
    flagged = (tie | otie) != 0
    seq_of = np.searchsorted(off, diff, side="right") - 1
    for n, s in zip(diff, seq_of):
        hi = off[s + 1] if backtrace else n + 1
        assert flagged[off[s]:hi].any(), f"unflagged path difference at point {n}"
    return len(diff)


# ---------------------------------------------------------------------------------
# BASELINE config 4: N=64, D=16 (V=10), C=16, K=8, T=1000
# ---------------------------------------------------------------------------------
def test_config4_exact_variables_three_em_iterations(oracle):
    wl, data, q = _workload_case(oracle, "scaleout", 12, 4104)
    assert (wl.N, wl.D, wl.C, wl.K, wl.T, wl.V) == (64, 16, 16, 8, 1000, 10)
    want = q.copy()
    st0 = oracle.estep(want, data)
    ll_want = oracle.em_iterate(want, data, 3)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False, dtype=np.float32)
        ctx.em_estep()
        got = ctx.em_get_stats()
        g = ctx.em_posteriors()
        ll = ctx.em_iterate(3)
        p = _assert_params(ctx, want, rtol=2e-4, atol=2e-5)
        floored = ctx.em_floored_components()
    assert floored == 0
    # per-iteration log-likelihood (north_star: ~1e-5 relative)
    np.testing.assert_allclose(ll, ll_want, rtol=TOL)
    assert (np.diff(ll) > -1e-6 * abs(ll[0])).all()
    # first E-step: statistics and posteriors against the oracle
    sw, sg = oracle.split_stats(q, st0), oracle.split_stats(q, got)
    scale = float(data.P)
    for key in ("pi", "xi", "gamma", "cat"):
        np.testing.assert_allclose(sg[key], sw[key], rtol=TOL, atol=TOL * scale / max(1, sw[key].size) + 1e-5, err_msg=key)
    np.testing.assert_allclose(sg["gmm"][..., 0], sw["gmm"][..., 0], rtol=TOL, atol=1e-3, err_msg="gmm s0")
    np.testing.assert_allclose(sg["gmm"][..., 1], sw["gmm"][..., 1], rtol=1e-4, atol=2e-3 * np.sqrt(scale), err_msg="gmm s1")
    np.testing.assert_allclose(sg["gmm"][..., 2], sw["gmm"][..., 2], rtol=1e-4, atol=1e-2 * np.sqrt(scale), err_msg="gmm s2")
    np.testing.assert_allclose(g, oracle.posteriors(q, data), atol=2e-5)
    # parameters that carry real mass agree to a few 1e-5 relative
    big = want.A > 1e-2
    np.testing.assert_allclose(p["A"][big], want.A[big], rtol=5 * TOL)
    heavy = want.w > 0.05
    np.testing.assert_allclose(p["mu"][heavy], want.mu[heavy], rtol=5 * TOL, atol=5e-5)


def test_config4_exact_variables_decode(oracle):
    wl, data, q = _workload_case(oracle, "scaleout", 6, 4105)
    opath, oscore, otie = oracle.decode(q, data, oracle.DECODE_BACKTRACE)
    rpath, rscore, rtie = oracle.decode(q, data, oracle.DECODE_REFERENCE)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False)
        p32, s32, t32 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32)
        p64, s64, t64 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F64)
        rp, rs, rt = ctx.decode(nat.DECODE_REFERENCE, nat.PREC_F64)
    _paths_equal_except_ties(p32, opath, t32, otie, data.off, True)
    _paths_equal_except_ties(p64, opath, t64, otie, data.off, True)
    _paths_equal_except_ties(rp, rpath, rt, rtie, data.off, False)
    np.testing.assert_allclose(s32, oscore, rtol=1e-5)
    np.testing.assert_allclose(s64, oscore, rtol=1e-11)
    np.testing.assert_allclose(rs, rscore, rtol=1e-11)
    assert (p32 == opath).mean() > 0.999


# ---------------------------------------------------------------------------------
# BASELINE config 5: N=512 dense transitions, D=5, C=5, K=3, T=200 (tcgen05 recursion)
# ---------------------------------------------------------------------------------
def test_config5_exact_shape_two_tensor_core_batches(oracle):
    """S = 168 sequences of T = 200: two 128-sequence tcgen05 batches (the second partial), the
    full 200-step dependency chain the bf16x3 split has to survive, the tcgen05 transition
    statistics, the N = 512 statistics sweeps, the max-plus Viterbi tiles and two EM iterations."""
    wl, data, q = _workload_case(oracle, "largeN", 168, 5105)
    assert (wl.N, wl.D, wl.C, wl.K, wl.T) == (512, 5, 5, 3, 200)
    st0 = oracle.estep(q, data)
    gamma0 = oracle.posteriors(q, data)
    opath, oscore, otie = oracle.decode(q, data, oracle.DECODE_BACKTRACE)
    want = q.copy()
    ll_want = oracle.em_iterate(want, data, 2)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False, dtype=np.float32)
        ctx.em_estep()
        got = ctx.em_get_stats()
        g = ctx.em_posteriors()
        p32, s32, t32 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32)
        ll = ctx.em_iterate(2)
        p = ctx.get_params()
    sw, sg = oracle.split_stats(q, st0), oracle.split_stats(q, got)
    assert abs(sg["ll"] - sw["ll"]) <= TOL * abs(sw["ll"])
    np.testing.assert_allclose(ll, ll_want, rtol=TOL)
    np.testing.assert_allclose(g, gamma0, atol=2e-5)
    np.testing.assert_allclose(sg["pi"], sw["pi"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sg["gamma"], sw["gamma"], rtol=1e-4, atol=1e-4)
    # xi entries are ~P/N^2 = 0.13 expected transitions each: absolute tolerance relative to the row mass
    np.testing.assert_allclose(sg["xi"], sw["xi"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sg["cat"], sw["cat"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sg["gmm"][..., 0], sw["gmm"][..., 0], rtol=1e-4, atol=1e-3)
    _paths_equal_except_ties(p32, opath, t32, otie, data.off, True)
    np.testing.assert_allclose(s32, oscore, rtol=1e-5)
    assert (p32 == opath).mean() > 0.995
    # after two EM iterations from the same start
    np.testing.assert_allclose(p["pi"], want.pi, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p["A"], want.A, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p["B"], want.B, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p["w"], want.w, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p["mu"], want.mu, rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(p["var"], want.var, rtol=5e-4, atol=1e-4)


def test_long_sequences_at_large_state_counts():
    """tests/fuzz_parity.py in its `large` mode: N in {256, 300, 512}, sequences up to T = 400
    (the regime where the split-bf16 recursion drifted before beta rescaling, DESIGN.md Numerics)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "6", "2028", "large"],
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---------------------------------------------------------------------------------
# context reuse (round-1 advisor findings)
# ---------------------------------------------------------------------------------
def test_decode_small_batch_then_larger_batch_on_one_context(oracle):
    """model.score(one sequence) followed by a batch decode on the same context: every decode
    output buffer has to grow with the data (the tie flags once kept the first call's size)."""
    m = random_model(oracle, 7, [5, 3], [2], 301)
    small = sample(oracle, m, [9], 302)
    big = sample(oracle, m, [60] * 40 + [3, 1], 303)
    with ctx_for(nat, m) as ctx:
        for data in (small, big, small):
            upload(ctx, data, labels=False)
            for mode, prec in ((nat.DECODE_REFERENCE, nat.PREC_F64), (nat.DECODE_BACKTRACE, nat.PREC_F32)):
                opath, oscore, otie = oracle.decode(m, data, mode)
                path, score, tie = ctx.decode(mode, prec)
                assert path.shape == (data.P,) and tie.shape == (data.P,)
                _paths_equal_except_ties(path, opath, tie, otie, data.off, mode == nat.DECODE_BACKTRACE)
                if prec == nat.PREC_F64:
                    np.testing.assert_array_equal(tie, otie)
                    np.testing.assert_allclose(score, oscore, rtol=1e-11)


def test_f64_planes_do_not_outlive_their_upload(oracle):
    """f64 upload, then an f32 upload / a generated data set on the same context: the F64 decode
    and emission paths must read the resident observations, not the earlier fp64 planes."""
    m = random_model(oracle, 6, [4], [3, 2], 311)
    first = sample(oracle, m, [50] * 20, 312)
    second = sample(oracle, m, [30] * 11 + [7], 313)
    with ctx_for(nat, m) as ctx:
        upload(ctx, first, labels=False, dtype=np.float64)
        ctx.decode(nat.DECODE_REFERENCE, nat.PREC_F64)
        upload(ctx, second, labels=False, dtype=np.float32)
        e = ctx.emission_logprob(nat.PREC_F64)
        np.testing.assert_allclose(e, oracle.emission_loglik(m, second), rtol=1e-12, atol=1e-12)
        path, score, tie = ctx.decode(nat.DECODE_REFERENCE, nat.PREC_F64)
        opath, oscore, otie = oracle.decode(m, second, oracle.DECODE_REFERENCE)
        _paths_equal_except_ties(path, opath, tie, otie, second.off, False)
        np.testing.assert_allclose(score, oscore, rtol=1e-11)
        # generated data after an f64 upload
        upload(ctx, first, labels=False, dtype=np.float64)
        ctx.generate(16, 25, 99)
        off, disc, cont, labels = ctx.download()
        gen = oracle.Data(off, disc, cont, labels)
        e = ctx.emission_logprob(nat.PREC_F64)
        np.testing.assert_allclose(e, oracle.emission_loglik(m, gen), rtol=1e-12, atol=1e-12)


def test_mstep_variance_floor(oracle):
    """A component that collapses onto one repeated value: the M-step raises its variance to the
    floor (relative to the variable's pooled variance), counts it, and agrees with the oracle's
    floored M-step; with the floor off the variance is ~0 as in the reference's EM1D."""
    m = random_model(oracle, 2, [], [2], 321)
    m.mu[0, :, :] = [[0.0, 50.0], [5.0, 50.0]]
    m.var[0, :, :] = [[1.0, 1e-3], [1.0, 1e-3]]
    data = sample(oracle, m, [40] * 6, 322)
    hit = np.arange(0, data.P, 7)
    data.cont[0, hit] = 50.0                      # identical outliers claimed by the second component
    data = oracle.Data(data.off, data.disc, data.cont, data.labels)
    for floor in (1e-3, 0.0):
        want = m.copy()
        st = oracle.estep(want, data)
        n_want = oracle.mstep(want, st, data.S, floor_rel=floor)
        with ctx_for(nat, m) as ctx:
            upload(ctx, data, labels=False, dtype=np.float32)
            ctx.em_set_variance_floor(floor)
            ctx.em_iterate(1)
            p = ctx.get_params()
            n_got = ctx.em_floored_components()
        assert n_got == n_want
        if floor > 0:
            assert n_got >= 2
            np.testing.assert_allclose(p["var"], want.var, rtol=1e-4, atol=1e-9)
            assert (p["var"] > 0).all()
        else:
            assert n_got == 0 or (p["var"] >= 0).all()
        np.testing.assert_allclose(p["mu"], want.mu, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------
# the collective behind the C ABI
# ---------------------------------------------------------------------------------
def test_single_rank_communicator_is_identity(oracle):
    """One rank: ncclCommInitRank / ncclAllReduce through the ABI leave the statistics unchanged and
    dnbc_em_iterate_sharded equals dnbc_em_iterate."""
    m = random_model(oracle, 10, [6], [2, 3], 331)
    data = sample(oracle, m, [40] * 12, 332)
    q = perturb(oracle, m, 333)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False, dtype=np.float32)
        ll_a = ctx.em_iterate(3)
        pa = ctx.get_params()
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False, dtype=np.float32)
        assert ctx.comm_info() == (1, 0)
        ctx.comm_init_rank(1, 0, nat.comm_unique_id())
        assert ctx.comm_info() == (1, 0)
        ctx.em_estep()
        before = ctx.em_get_stats()
        ctx.em_allreduce()
        np.testing.assert_array_equal(ctx.em_get_stats(), before)
        ll_b = ctx.em_iterate_sharded(3, data.S)
        pb = ctx.get_params()
        with pytest.raises(nat.DnbcError):
            ctx.comm_init_rank(1, 0, nat.comm_unique_id())      # already joined
        ctx.comm_destroy()
        assert ctx.comm_info() == (1, 0)
    np.testing.assert_allclose(ll_b, ll_a, rtol=1e-12)
    for k in ("pi", "A", "B", "w", "mu", "var"):
        np.testing.assert_allclose(pb[k], pa[k], rtol=1e-9, atol=1e-12)


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_two_gpus_one_process_group_allreduce(oracle):
    """dnbc_comm_init_all over two devices in one process (the reference's single driver thread
    holding both GPUs): shard sums equal the single-GPU statistics, the group iteration equals the
    single-GPU iteration and the fp64 oracle."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    m = random_model(oracle, 12, [5, 4], [3, 2], 341)
    data = sample(oracle, m, [50] * 30 + [13, 1], 342)
    q = perturb(oracle, m, 343)
    want = q.copy()
    st_want = oracle.estep(want, data)
    ll_want = oracle.em_iterate(want, data, 3)
    half = data.S // 2
    cut = int(data.off[half])
    shards = [oracle.Data(data.off[:half + 1], data.disc[:, :cut], data.cont[:, :cut]),
              oracle.Data(data.off[half:] - cut, data.disc[:, cut:], data.cont[:, cut:])]
    ctxs = [ctx_for(nat, q, device=i) for i in range(2)]
    try:
        for ctx, sh in zip(ctxs, shards):
            upload(ctx, sh, labels=False, dtype=np.float32)
        nat.comm_init_all(ctxs)
        assert [c.comm_info() for c in ctxs] == [(2, 0), (2, 1)]
        for ctx in ctxs:
            ctx.em_estep()
        nat.em_allreduce_group(ctxs)
        s0, s1 = ctxs[0].em_get_stats(), ctxs[1].em_get_stats()
        np.testing.assert_array_equal(s0, s1)                   # every rank holds the same sums
        np.testing.assert_allclose(s0[0], st_want[0], rtol=TOL)
        sw, sg = oracle.split_stats(q, st_want), oracle.split_stats(q, s0)
        for key in ("pi", "xi", "gamma", "cat"):
            np.testing.assert_allclose(sg[key], sw[key], rtol=TOL, atol=TOL * data.P / max(1, sw[key].size) + 1e-6, err_msg=key)
        ll = nat.em_iterate_group(ctxs, 3)
        np.testing.assert_allclose(ll, ll_want, rtol=TOL)
        p0, p1 = ctxs[0].get_params(), ctxs[1].get_params()
        for k in ("pi", "A", "B", "w", "mu", "var"):
            np.testing.assert_array_equal(p0[k], p1[k])         # replicated M-step, no broadcast needed
        np.testing.assert_allclose(p0["A"], want.A, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(p0["mu"], want.mu, rtol=2e-4, atol=2e-5)
    finally:
        for ctx in ctxs:
            ctx.close()


def test_two_gpus_two_processes_init_rank(oracle, tmp_path):
    """One process per GPU: rank 0 writes the unique id to a file (standing in for the Spark
    broadcast of the JVM shim), both ranks join with dnbc_comm_init_rank and iterate sharded."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.pop("OMP_NUM_THREADS", None)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "two_rank_nccl.py"), str(r), "2", str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "rank ok" in o, o[-3000:]


def test_c_abi_smoke_program():
    """tests/abi_smoke.c: a plain-C11 program that walks create -> upload -> mle -> decode ->
    em_iterate -> comm -> destroy in the order the JNI glue does (the closest stand-in for the JVM
    call path available without a JDK)."""
    exe = os.path.join(ROOT, "tests", "_build", "abi_smoke")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests"), "-s"], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "dnbc_scala_b200", "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "abi_smoke ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


# ---------------------------------------------------------------------------------
# N <= 16: the single-launch E-step (k_fused.cu) and the general pipeline, both forced
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("N,V,K,lengths", [
    (10, [10] * 5, [3, 1, 4, 2, 3], [200] * 9 + [37, 1, 0, 150]),      # config-1 variables, ragged, an empty sequence
    (12, [4], [], [200] * 7 + [3]),                                    # toy robot: discrete only
    (10, [], [3] * 5, [200] * 8),                                      # config-3 shape
    (1, [3], [2], [5, 1, 9]),                                          # a single state
    (3, [2, 2], [8], [40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50]),    # ten groups of three lanes, K = 8
    (7, [5], [5, 7], [64] * 6),                                        # odd N (zero pad entry), odd K
    (16, [6], [2], [90] * 5),                                          # the largest state count the kernel takes
])
@pytest.mark.parametrize("mode", [1, 2])
def test_fused_and_general_estep_match_oracle(oracle, N, V, K, lengths, mode):
    """dnbc_set_option(DNBC_OPT_FUSED_ESTEP): 1 forces the single-launch kernel, 2 the general pipeline;
    both against the fp64 oracle -- statistics, log-likelihood and three EM iterations."""
    m = random_model(oracle, N, V, K, 400 + N, transitions_per_state=min(N, 6))
    data = sample(oracle, m, lengths, 410 + N)
    q = perturb(oracle, m, 420 + N)
    want = oracle.estep(q, data)
    w3 = q.copy()
    ll_want = oracle.em_iterate(w3, data, 3)
    with ctx_for(nat, q) as ctx:
        ctx.set_option(nat.OPT_FUSED_ESTEP, mode)
        upload(ctx, data, labels=False, dtype=np.float32)
        l0 = ctx.launch_count()
        ctx.em_estep()
        launches = ctx.launch_count() - l0
        got = ctx.em_get_stats()
        ll = ctx.em_iterate(3)
        p = ctx.get_params()
    assert launches == 1 if mode == 1 else launches >= 4, launches     # one launch for the whole E-step vs the pipeline
    sw, sg = oracle.split_stats(q, want), oracle.split_stats(q, got)
    assert abs(sg["ll"] - sw["ll"]) <= TOL * abs(sw["ll"])
    for key in ("pi", "xi", "gamma", "cat"):
        np.testing.assert_allclose(sg[key], sw[key], rtol=1e-4, atol=1e-4, err_msg=key)
    if K:
        np.testing.assert_allclose(sg["gmm"][..., 0], sw["gmm"][..., 0], rtol=1e-4, atol=1e-3, err_msg="gmm s0")
        np.testing.assert_allclose(sg["gmm"][..., 1], sw["gmm"][..., 1], rtol=1e-4, atol=2e-3 * np.sqrt(data.P), err_msg="gmm s1")
        np.testing.assert_allclose(sg["gmm"][..., 2], sw["gmm"][..., 2], rtol=1e-4, atol=1e-2 * np.sqrt(data.P), err_msg="gmm s2")
    np.testing.assert_allclose(ll, ll_want, rtol=TOL)
    np.testing.assert_allclose(p["A"], w3.A, rtol=5e-4, atol=5e-5)
    np.testing.assert_allclose(p["B"], w3.B, rtol=5e-4, atol=5e-5)
    np.testing.assert_allclose(p["mu"], w3.mu, rtol=5e-4, atol=5e-4)


def test_fused_estep_outliers_and_long_sequences(oracle):
    """The fused kernel's exponent-shifted paths (points 15-25 sigma from every component) and its refusal of
    sequences whose rows do not fit shared memory (T = 4000 at N = 16 falls back to the general pipeline)."""
    m = random_model(oracle, 6, [5], [3, 2], 471)
    m.var[0, :, :3] = 4.0
    data = sample(oracle, m, [40] * 12 + [7, 1], 472)
    rng = np.random.default_rng(473)
    hit = rng.choice(data.P, size=37, replace=False)
    data.cont[0, hit] = rng.choice([-1.0, 1.0], 37) * rng.uniform(60, 90, 37)
    data = oracle.Data(data.off, data.disc, data.cont, data.labels)
    want = oracle.split_stats(m, oracle.estep(m, data))
    for mode in (1, 2):
        with ctx_for(nat, m) as ctx:
            ctx.set_option(nat.OPT_FUSED_ESTEP, mode)
            upload(ctx, data, labels=False, dtype=np.float32)
            ctx.em_estep()
            got = oracle.split_stats(m, ctx.em_get_stats())
        assert np.isfinite(got["ll"]) and abs(got["ll"] - want["ll"]) <= TOL * abs(want["ll"]), mode
        np.testing.assert_allclose(got["gamma"], want["gamma"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(got["gmm"][..., 0], want["gmm"][..., 0], rtol=1e-4, atol=1e-3)
    m2 = random_model(oracle, 16, [3], [1], 481, transitions_per_state=4)
    long = sample(oracle, m2, [4000, 10], 482)
    w2 = oracle.split_stats(m2, oracle.estep(m2, long))
    with ctx_for(nat, m2) as ctx:
        ctx.set_option(nat.OPT_FUSED_ESTEP, 1)              # asked for, but 2 x 4000 x 16 floats do not fit: general path
        upload(ctx, long, labels=False, dtype=np.float32)
        l0 = ctx.launch_count()
        ctx.em_estep()
        assert ctx.launch_count() - l0 >= 4
        g2 = oracle.split_stats(m2, ctx.em_get_stats())
    assert abs(g2["ll"] - w2["ll"]) <= TOL * abs(w2["ll"])


# ---------------------------------------------------------------------------------
# even N <= 12: forward / backward with one lane per sequence (k_fbs.cu), forced on small inputs
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("N,V,K,lengths", [
    (10, [10] * 5, [3, 1, 4, 2, 3], [200] * 9 + [37, 1, 0, 150] + [5, 6, 7, 8] * 9),   # config-1 variables, ragged, empty
    (12, [4], [], [200] * 7 + [3] + list(range(1, 60))),                               # toy robot; two warps of sequences
    (10, [], [3] * 5, [200] * 40),                                                     # config-3 shape
    (2, [3], [2], [5, 1, 9, 4, 4, 4, 13]),
    (4, [2, 2], [8], list(range(30, 70))),
    (6, [5], [5, 7], [64] * 6 + [1] * 30),
    (8, [6], [2], [90] * 5 + [3, 2, 1, 0, 0, 17]),
])
def test_lane_per_sequence_recursions_match_oracle(oracle, N, V, K, lengths):
    """dnbc_set_option(DNBC_OPT_FB_SEQ_LANES, 1): the lane-per-sequence forward / backward kernels on inputs far below
    their automatic threshold, against the fp64 oracle (statistics, posteriors, three EM iterations) and against
    the lane-per-state kernels (mode 2); several chunks as well."""
    m = random_model(oracle, N, V, K, 500 + N, transitions_per_state=min(N, 6))
    data = sample(oracle, m, lengths, 510 + N)
    q = perturb(oracle, m, 520 + N)
    want = oracle.estep(q, data)
    w3 = q.copy()
    ll_want = oracle.em_iterate(w3, data, 3)
    stats, post = {}, {}
    for mode, chunk in ((1, 0), (2, 0), (1, max(data.P // 3, 256))):
        with ctx_for(nat, q) as ctx:
            ctx.set_option(nat.OPT_FUSED_ESTEP, 2)
            ctx.set_option(nat.OPT_FB_SEQ_LANES, mode)
            if chunk:
                ctx.set_option(nat.OPT_MAX_CHUNK_POINTS, chunk)
            upload(ctx, data, labels=False, dtype=np.float32)
            ctx.em_estep()
            got = ctx.em_get_stats()
            if not chunk:
                post[mode] = ctx.em_posteriors()
            ll = ctx.em_iterate(3)
            p = ctx.get_params()
        stats[(mode, chunk)] = got
        sw, sg = oracle.split_stats(q, want), oracle.split_stats(q, got)
        assert abs(sg["ll"] - sw["ll"]) <= TOL * abs(sw["ll"]), (mode, chunk)
        for key in ("pi", "xi", "gamma", "cat"):
            np.testing.assert_allclose(sg[key], sw[key], rtol=1e-4, atol=1e-4, err_msg=f"{key} {mode} {chunk}")
        np.testing.assert_allclose(ll, ll_want, rtol=TOL)
        np.testing.assert_allclose(p["A"], w3.A, rtol=5e-4, atol=5e-5)
        np.testing.assert_allclose(p["pi"], w3.pi, rtol=5e-4, atol=5e-5)
    np.testing.assert_allclose(post[1], post[2], rtol=0, atol=5e-5)
    np.testing.assert_allclose(post[1], oracle.posteriors(q, data), rtol=0, atol=5e-5)


def test_state_counts_above_512(oracle):
    """N = 600 pads to 1024: beyond the tcgen05 recursion (256 / 512) and beyond the max-plus Viterbi tiles (two
    delta buffers of 32 sequences no longer fit one SM's shared memory), so the CTA-per-sequence kernels run."""
    m = random_model(oracle, 600, [4], [2], 491, transitions_per_state=12)
    data = sample(oracle, m, [25, 9, 1, 30], 492)
    q = perturb(oracle, m, 493)
    want = oracle.split_stats(q, oracle.estep(q, data))
    opath, oscore, otie = oracle.decode(q, data, oracle.DECODE_BACKTRACE)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False)
        ctx.em_estep()
        got = oracle.split_stats(q, ctx.em_get_stats())
        p32, s32, t32 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F32)
        p64, s64, t64 = ctx.decode(nat.DECODE_BACKTRACE, nat.PREC_F64)
    assert abs(got["ll"] - want["ll"]) <= TOL * abs(want["ll"])
    np.testing.assert_allclose(got["gamma"], want["gamma"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["xi"], want["xi"], rtol=1e-4, atol=1e-4)
    _paths_equal_except_ties(p32, opath, t32, otie, data.off, True)
    _paths_equal_except_ties(p64, opath, t64, otie, data.off, True)
    np.testing.assert_allclose(s32, oscore, rtol=1e-5)
    np.testing.assert_allclose(s64, oscore, rtol=1e-11)
