"""The benchmark configurations of BASELINE.json and their synthetic-data distributions.

Parameter distributions restate the reference's performance data generator
(test-utils/src/main/scala/TestUtils.scala:75-153): initial probabilities ~ N(100, var 16)
normalised (:90-95); per state `transitionsPerNode` distinct destinations with weights
~ U(0,1) normalised (:97-111); per discrete variable and state 10 symbols with weights
~ U(0,1) normalised (:113-125); per continuous variable K equally weighted components with
integer means in [-10, 10] and integer VARIANCES in [1, 4] (:132-153 -- the value named
`sigma` is used as the 1x1 covariance).  The reference draws from the unseeded global
scala.util.Random, so there is no reference seed: a fixed numpy seed per configuration is
used instead (seed = 20260000 + config index).

Host-side numpy only: this module fills parameter arrays and (for the CPU baseline sample)
draws small data sets; it never computes likelihoods or statistics.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import numpy as np


@dataclass(frozen=True)
class Workload:
    name: str
    N: int
    D: int
    C: int
    K: int              # Gaussians per mixture (fixed-K configurations)
    T: int
    S: int              # sequences of the full configuration
    transitions_per_state: int
    V: int = 10         # TestUtils.scala:117
    index: int = 0

    @property
    def bytes_per_timestep(self) -> int:      # each observation read once
        return self.D + 4 * self.C

    @property
    def flops_per_timestep(self) -> int:      # SURVEY 8(d), one EM iteration, FMA = 2
        N, D, C, K = self.N, self.D, self.C, self.K
        return 6 * N * N + N * (2 * D + C * (15 * K + 1)) + 10 * N

    @property
    def mufu_per_timestep(self) -> int:
        return self.N * self.C * (self.K + 1) + self.N


# BASELINE.json configs, in order.  Config 1's README set draws K in 1..4 per variable; a
# fixed K=3 is used for the throughput shape (parity tests cover mixed K).
WORKLOADS: Dict[str, Workload] = {
    "readme": Workload("readme", N=10, D=5, C=5, K=3, T=200, S=1000, transitions_per_state=5, index=0),
    "robot": Workload("robot", N=12, D=1, C=0, K=1, T=200, S=200, transitions_per_state=4, V=4, index=1),
    "gmm100k": Workload("gmm100k", N=10, D=0, C=5, K=3, T=200, S=100_000, transitions_per_state=10, index=2),
    "scaleout": Workload("scaleout", N=64, D=16, C=16, K=8, T=1000, S=1_000_000, transitions_per_state=64, index=3),
    "largeN": Workload("largeN", N=512, D=5, C=5, K=3, T=200, S=100_000, transitions_per_state=512, index=4),
}


def by_name(name: str) -> Workload:
    """A BASELINE configuration by name, or an ad-hoc shape for kernel timings:
    "custom:N=128,D=5,C=5,K=3,T=200" (missing keys default to the largeN configuration's)."""
    if not name.startswith("custom:"):
        return WORKLOADS[name]
    kv = dict(item.split("=") for item in name[7:].split(",") if item)
    g = {k: int(v) for k, v in kv.items()}
    N = g.get("N", 512)
    return Workload(name, N=N, D=g.get("D", 5), C=g.get("C", 5), K=g.get("K", 3), T=g.get("T", 200), S=g.get("S", 100_000),
                    transitions_per_state=g.get("tps", N), index=9)


def true_params(wl: Workload) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(20260000 + wl.index)
    N, D, C, K, V = wl.N, wl.D, wl.C, wl.K, wl.V
    pi = rng.normal(100.0, 4.0, N)
    pi /= pi.sum()
    A = np.zeros((N, N))
    for i in range(N):
        dst = rng.choice(N, size=min(N, wl.transitions_per_state), replace=False)
        A[i, dst] = rng.random(len(dst))
        A[i] /= A[i].sum()
    B = rng.random((D, N, V))
    B /= np.maximum(B.sum(-1, keepdims=True), 1e-300)
    w = np.full((C, N, K), 1.0 / K)
    mu = rng.integers(-10, 11, (C, N, K)).astype(np.float64)
    var = rng.integers(1, 5, (C, N, K)).astype(np.float64)
    return dict(pi=pi, A=A, B=B, w=w, mu=mu, var=var)


def perturbed_params(wl: Workload, p: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """EM starting point: probability rows <- normalise(p + 0.1 U(0,1)/width),
    mu <- mu + N(0,1), var <- var * U(0.5, 2)."""
    rng = np.random.default_rng(20261000 + wl.index)
    q = {k: v.copy() for k, v in p.items()}
    q["pi"] = q["pi"] + 0.1 * rng.random(wl.N) / wl.N
    q["pi"] /= q["pi"].sum()
    q["A"] = q["A"] + 0.1 * rng.random(q["A"].shape) / wl.N
    q["A"] /= q["A"].sum(-1, keepdims=True)
    if wl.D:
        q["B"] = q["B"] + 0.1 * rng.random(q["B"].shape) / wl.V
        q["B"] /= q["B"].sum(-1, keepdims=True)
    q["mu"] = q["mu"] + rng.normal(0, 1, q["mu"].shape)
    q["var"] = q["var"] * rng.uniform(0.5, 2.0, q["var"].shape)
    return q


def _draw(cdf: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Row-wise inverse-CDF draw: first index with u <= cdf."""
    return np.minimum((u[:, None] > cdf).sum(1), cdf.shape[1] - 1)


def sample_host(wl: Workload, p: Dict[str, np.ndarray], S: int, seed: int):
    """Vectorised sampler following the reference's sampling loop (TestUtils.scala:171-185);
    returns (offsets i64 [S+1], disc u8 [D][P], cont f32 [C][P], labels i32 [P])."""
    rng = np.random.default_rng(seed)
    N, D, C, K, T = wl.N, wl.D, wl.C, wl.K, wl.T
    P = S * T
    labels = np.zeros((S, T), np.int32)
    disc = np.zeros((D, S, T), np.uint8)
    cont = np.zeros((C, S, T), np.float32)
    cpi, cA = np.cumsum(p["pi"]), np.cumsum(p["A"], -1)
    cB = np.cumsum(p["B"], -1) if D else None
    h = _draw(np.broadcast_to(cpi, (S, N)), rng.random(S))
    for t in range(T):
        if t:
            h = _draw(cA[h], rng.random(S))
        labels[:, t] = h
        for d in range(D):
            disc[d, :, t] = _draw(cB[d, h], rng.random(S))
        for c in range(C):
            k = rng.integers(0, K, S)
            cont[c, :, t] = rng.normal(p["mu"][c, h, k], np.sqrt(p["var"][c, h, k]))
    off = np.arange(S + 1, dtype=np.int64) * T
    return off, disc.reshape(D, P), cont.reshape(C, P), labels.reshape(P)
