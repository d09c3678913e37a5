"""Shared helpers for the tests: seeded random models / data in the dense layout."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def random_model(orc, N, V, K, seed, transitions_per_state=None, unit_var=False):
    """Random parameters with the shapes of the reference's generator (TestUtils.scala:90-153)."""
    rng = np.random.default_rng(seed)
    m = orc.Model(N, V, K)
    m.pi[:] = rng.normal(100, 4, N).clip(1)
    m.pi /= m.pi.sum()
    for i in range(N):
        k = N if transitions_per_state is None else min(N, transitions_per_state)
        dst = rng.choice(N, size=k, replace=False)
        m.A[i, dst] = rng.random(k) + 0.05
        m.A[i] /= m.A[i].sum()
    for d in range(m.D):
        m.B[d, :, :V[d]] = rng.random((N, V[d])) + 0.02
        m.B[d] /= m.B[d].sum(-1, keepdims=True)
    for c in range(m.C):
        k = K[c]
        m.w[c, :, :k] = 1.0 / k
        m.mu[c, :, :k] = rng.integers(-10, 11, (N, k))
        m.var[c, :, :k] = 1.0 if unit_var else rng.integers(1, 5, (N, k))
    return m


def perturb(orc, m, seed):
    """Initial parameters for EM: the true ones, perturbed."""
    rng = np.random.default_rng(seed)
    q = m.copy()
    q.pi[:] = q.pi + 0.1 * rng.random(q.N) / q.N
    q.pi /= q.pi.sum()
    q.A[:] = q.A + 0.1 * rng.random(q.A.shape) / q.N
    q.A /= q.A.sum(-1, keepdims=True)
    for d in range(q.D):
        v = q.V[d]
        q.B[d, :, :v] += 0.1 * rng.random((q.N, v)) / v
        q.B[d, :, :v] /= q.B[d, :, :v].sum(-1, keepdims=True)
    for c in range(q.C):
        k = q.K[c]
        q.mu[c, :, :k] += rng.normal(0, 1, (q.N, k))
        q.var[c, :, :k] *= rng.uniform(0.5, 2.0, (q.N, k))
    return q


def sample(orc, m, lengths, seed, f32_values=True):
    """Draw sequences from model m (host-side sampler for tests)."""
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, np.int64)
    off = np.zeros(len(lengths) + 1, np.int64)
    off[1:] = np.cumsum(lengths)
    P = int(off[-1])
    labels = np.zeros(P, np.int32)
    disc = np.zeros((m.D, P), np.uint8)
    cont = np.zeros((m.C, P))
    for s, T in enumerate(lengths):
        h = 0
        for t in range(int(T)):
            n = off[s] + t
            p = m.pi if t == 0 else m.A[h]
            h = int(rng.choice(m.N, p=p / p.sum()))
            labels[n] = h
            for d in range(m.D):
                pv = m.B[d, h, :m.V[d]]
                disc[d, n] = rng.choice(m.V[d], p=pv / pv.sum())
            for c in range(m.C):
                k = int(rng.choice(m.K[c], p=m.w[c, h, :m.K[c]] / m.w[c, h, :m.K[c]].sum()))
                cont[c, n] = rng.normal(m.mu[c, h, k], np.sqrt(m.var[c, h, k]))
    if f32_values:  # values exactly representable in fp32, so f32 and f64 uploads agree
        cont = cont.astype(np.float32).astype(np.float64)
    return orc.Data(off, disc, cont, labels)


def ctx_for(nat, m, device=0):
    ctx = nat.Context(m.N, m.V, m.K, device)
    ctx.set_params(m.pi, m.A, m.B, m.w, m.mu, m.var, m.active)
    return ctx


def upload(ctx, data, labels=True, dtype=np.float64):
    ctx.upload(data.off, data.disc, data.cont, data.labels if labels else None, cont_dtype=dtype)
