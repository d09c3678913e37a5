"""ctypes binding of the fp64 CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/dnbc_oracle.h).  The product package
dnbc_scala_b200 never imports it.

Besides the raw bindings this module restates, in `mle()`, the orchestration of the
reference's supervised learner (DynamicNaiveBayesianClassifier.mle, reference
core/src/main/scala/DynamicNaiveBayesianClassifier.scala:124-161 and
ContinuousEdge.learnFinalize, core/src/main/scala/Edge.scala:87-104) on top of the
C primitives.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdnbc_oracle.so")

DECODE_REFERENCE = 0
DECODE_BACKTRACE = 1
QUIRK_TRANSITIONS = 1
DOUBLE_COUNT_FIRST = 2


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("dnbc_oracle.c", "dnbc_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Model(C.Structure):
    _fields_ = [("N", C.c_int32), ("D", C.c_int32), ("C", C.c_int32), ("Vmax", C.c_int32),
                ("Kmax", C.c_int32), ("V", C.c_void_p), ("K", C.c_void_p), ("pi", C.c_void_p),
                ("A", C.c_void_p), ("B", C.c_void_p), ("w", C.c_void_p), ("mu", C.c_void_p),
                ("var", C.c_void_p), ("active", C.c_void_p)]


class _Data(C.Structure):
    _fields_ = [("S", C.c_int64), ("off", C.c_void_p), ("disc", C.c_void_p),
                ("cont", C.c_void_p), ("labels", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_mle_counts.restype = C.c_int
        L.orc_edge_points.restype = C.c_int64
        L.orc_edge_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.orc_kmeans_1d.restype = C.c_int
        L.orc_kmeans_1d.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_em1d_initialize.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p] + [C.c_void_p] * 3
        L.orc_em1d_run.restype = C.c_int
        L.orc_em1d_run.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_double]
        for f in ("orc_jmef_density", "orc_mllib_pdf"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double] * 3
        L.orc_mixture_pdf.restype = C.c_double
        L.orc_mixture_pdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_stats_size.restype = C.c_int64
        L.orc_accuracy.restype = C.c_double
        L.orc_mstep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_mstep_floor.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_normalize_counts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Data:
    """Dense observations (same layout as the product C-ABI)."""

    def __init__(self, offsets, disc, cont, labels=None):
        self.off = np.ascontiguousarray(offsets, np.int64)
        P = int(self.off[-1])
        self.disc = np.ascontiguousarray(disc, np.uint8).reshape(-1, P) if np.size(disc) else np.zeros((0, P), np.uint8)
        self.cont = np.ascontiguousarray(cont, np.float64).reshape(-1, P) if np.size(cont) else np.zeros((0, P), np.float64)
        self.labels = None if labels is None else np.ascontiguousarray(labels, np.int32)
        self.S, self.P = len(self.off) - 1, P
        self.D, self.C = self.disc.shape[0], self.cont.shape[0]
        self._c = _Data(self.S, _p(self.off), _p(self.disc), _p(self.cont), _p(self.labels))

    @classmethod
    def of(cls, ds) -> "Data":
        """From a dnbc_scala_b200.dataset.DenseSequences (duck-typed)."""
        return cls(ds.offsets, ds.disc, ds.cont, ds.labels)

    @property
    def ref(self):
        return C.byref(self._c)


class Model:
    """Dense fp64 model (layout in dnbc_oracle.h)."""

    def __init__(self, N: int, V: Sequence[int], K: Sequence[int], Vmax: Optional[int] = None,
                 Kmax: Optional[int] = None):
        self.N, self.D, self.C = int(N), len(V), len(K)
        self.V = np.ascontiguousarray(V, np.int32)
        self.K = np.ascontiguousarray(K, np.int32)
        self.Vmax = int(Vmax if Vmax is not None else max([1] + list(V)))
        self.Kmax = int(Kmax if Kmax is not None else max([1] + list(K)))
        self.pi = np.zeros(self.N)
        self.A = np.zeros((self.N, self.N))
        self.B = np.zeros((self.D, self.N, self.Vmax))
        self.w = np.zeros((self.C, self.N, self.Kmax))
        self.mu = np.zeros((self.C, self.N, self.Kmax))
        self.var = np.ones((self.C, self.N, self.Kmax))
        self.active = np.ones(self.N, np.uint8)

    def copy(self) -> "Model":
        m = Model(self.N, self.V, self.K, self.Vmax, self.Kmax)
        for f in ("pi", "A", "B", "w", "mu", "var", "active"):
            getattr(m, f)[...] = getattr(self, f)
        return m

    @property
    def ref(self):
        for f in ("pi", "A", "B", "w", "mu", "var"):
            a = getattr(self, f)
            assert a.flags.c_contiguous and a.dtype == np.float64, f
        self._c = _Model(self.N, self.D, self.C, self.Vmax, self.Kmax, _p(self.V), _p(self.K),
                         _p(self.pi), _p(self.A), _p(self.B), _p(self.w), _p(self.mu), _p(self.var),
                         _p(self.active))
        return C.byref(self._c)


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------

def mle_counts(N, Vmax, data: Data, flags: int):
    cnt_pi = np.zeros(N, np.int64)
    cnt_A = np.zeros((N, N), np.int64)
    cnt_B = np.zeros((data.D, N, Vmax), np.int64)
    rc = lib().orc_mle_counts(C.c_int(N), C.c_int(data.D), C.c_int(Vmax), data.ref, C.c_int(flags),
                              C.c_void_p(_p(cnt_pi)), C.c_void_p(_p(cnt_A)), C.c_void_p(_p(cnt_B)))
    if rc != 0:
        raise ValueError("supervised learning needs labels")
    return cnt_pi, cnt_A, cnt_B


def normalize_counts(cnt: np.ndarray):
    cnt = np.ascontiguousarray(cnt, np.int64)
    rows, cols = int(np.prod(cnt.shape[:-1])), cnt.shape[-1]
    prob = np.zeros(cnt.shape)
    nonempty = np.zeros(rows, np.uint8)
    lib().orc_normalize_counts(_p(cnt), rows, cols, _p(prob), _p(nonempty))
    return prob, nonempty.reshape(cnt.shape[:-1])


def edge_points(data: Data, c: int, state: int, flags: int = 0) -> np.ndarray:
    n = lib().orc_edge_points(data.ref, c, state, flags, None, 0)
    out = np.zeros(n)
    lib().orc_edge_points(data.ref, c, state, flags, _p(out), n)
    return out


def kmeans_1d(x: np.ndarray, init_centroids: Sequence[float]):
    x = np.ascontiguousarray(x, np.float64)
    init = np.ascontiguousarray(init_centroids, np.float64)
    k = len(init)
    assign = np.zeros(len(x), np.int32)
    cen = np.zeros(k)
    its = lib().orc_kmeans_1d(_p(x), len(x), k, _p(init), _p(assign), _p(cen))
    return assign, cen, its


def em1d_initialize(x: np.ndarray, k: int, assign: np.ndarray):
    x = np.ascontiguousarray(x, np.float64)
    assign = np.ascontiguousarray(assign, np.int32)
    w, mu, var = np.zeros(k), np.zeros(k), np.zeros(k)
    lib().orc_em1d_initialize(_p(x), len(x), k, _p(assign), _p(w), _p(mu), _p(var))
    return w, mu, var


def em1d_run(x: np.ndarray, w, mu, var, max_iters: int = 30, rel_tol: float = 0.01):
    """Returns (w, mu, var, ll_series[0..iters]) -- reference EM1D.run."""
    x = np.ascontiguousarray(x, np.float64)
    w, mu, var = (np.array(a, np.float64, copy=True) for a in (w, mu, var))
    ll = np.zeros(max_iters + 1)
    its = lib().orc_em1d_run(_p(x), len(x), len(w), _p(w), _p(mu), _p(var), _p(ll), max_iters, rel_tol)
    return w, mu, var, ll[: its + 1].copy()


def emission_loglik(model: Model, data: Data) -> np.ndarray:
    out = np.zeros((data.P, model.N))
    lib().orc_emission_loglik(model.ref, data.ref, C.c_void_p(_p(out)))
    return out


def decode(model: Model, data: Data, mode: int = DECODE_REFERENCE):
    path = np.zeros(data.P, np.int32)
    score = np.zeros(data.S)
    tie = np.zeros(data.P, np.uint8)
    lib().orc_decode(model.ref, data.ref, C.c_int(mode), C.c_void_p(_p(path)), C.c_void_p(_p(score)),
                     C.c_void_p(_p(tie)))
    return path, score, tie


def accuracy(data: Data, path: np.ndarray) -> float:
    path = np.ascontiguousarray(path, np.int32)
    return float(lib().orc_accuracy(data.ref, C.c_void_p(_p(path))))


def stats_size(model: Model) -> int:
    return int(lib().orc_stats_size(model.ref))


def estep(model: Model, data: Data, clamp_labels: bool = False) -> np.ndarray:
    st = np.zeros(stats_size(model))
    lib().orc_estep(model.ref, data.ref, C.c_int(int(clamp_labels)), C.c_void_p(_p(st)))
    return st


def mstep(model: Model, stats: np.ndarray, S: int, floor_rel: Optional[float] = None) -> int:
    """M-step in place.  floor_rel: the variance floor relative to the variable's pooled variance
    (None = the product's default 1e-10, 0 = none, as the reference's EM1D); returns how many
    component variances were raised to the floor."""
    stats = np.ascontiguousarray(stats, np.float64)
    if floor_rel is None:
        lib().orc_mstep(model.ref, _p(stats), S)
        return 0
    n = C.c_int64(0)
    lib().orc_mstep_floor(model.ref, _p(stats), S, float(floor_rel), C.byref(n))
    return int(n.value)


def em_iterate(model: Model, data: Data, n_iters: int) -> np.ndarray:
    """n_iters Baum-Welch iterations in place; returns the log-likelihood BEFORE each M-step."""
    ll = np.zeros(n_iters)
    for i in range(n_iters):
        st = estep(model, data)
        ll[i] = st[0]
        mstep(model, st, data.S)
    return ll


def posteriors(model: Model, data: Data) -> np.ndarray:
    g = np.zeros((data.P, model.N))
    lib().orc_posteriors(model.ref, data.ref, C.c_void_p(_p(g)))
    return g


def split_stats(model: Model, st: np.ndarray):
    """Views into a packed statistics vector (layout in dnbc_oracle.h)."""
    N, D, Cn, V, K = model.N, model.D, model.C, model.Vmax, model.Kmax
    o = 1
    out = {"ll": st[0]}
    for name, shape in (("pi", (N,)), ("xi", (N, N)), ("gamma", (N,)), ("cat", (D, N, V)),
                        ("gmm", (Cn, N, K, 3))):
        n = int(np.prod(shape))
        out[name] = st[o:o + n].reshape(shape)
        o += n
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


# ---------------------------------------------------------------------------
# the reference's supervised learner, restated on the primitives
# ---------------------------------------------------------------------------

def pick_initial_centroids(x: np.ndarray, k: int, rng: np.random.Generator) -> np.ndarray:
    """KMeans.initialize (reference core/src/main/java/Tools/KMeans.java:46-72): k random
    points with pairwise distinct values.  The reference draws from an unseeded
    java.util.Random, so only the *distribution* can be restated; tests inject these."""
    cen = [x[rng.integers(len(x))]]
    guard = 0
    while len(cen) < k:
        cand = x[rng.integers(len(x))]
        guard += 1
        if guard > 100000:
            raise ValueError("fewer than k distinct points (the reference loops forever here)")
        if all(cand != c for c in cen):
            cen.append(cand)
    return np.array(cen)


@dataclass
class EdgeTrace:
    c: int
    state: int
    n: int
    init_centroids: np.ndarray
    init: tuple
    ll_series: np.ndarray


def mle(data: Data, N: int, V: Sequence[int], hints: Optional[Sequence[int]] = None, flags: int = 0,
        rng: Optional[np.random.Generator] = None, init_centroids=None, trace: Optional[List[EdgeTrace]] = None) -> Model:
    """Supervised maximum likelihood as the reference does it (DNBC.scala:124-259).

    hints: mixture size per continuous variable (default 1 each, DNBC.scala:279-285;
    validation messages DNBC.scala:270-277).  init_centroids[(c, state)] overrides the
    random k-means initialisation.
    """
    Cn = data.C
    if hints is None or len(hints) == 0:
        hints = [1] * Cn
    else:
        if len(hints) != Cn:
            raise ValueError("Number of continuous variable hints does not correspond to the number of continous variables")
        if any(h < 1 for h in hints):
            raise ValueError("Number of suggested normal distributions in a continuous variable must be at least one")
    rng = rng or np.random.default_rng(0)
    m = Model(N, V, hints)
    cnt_pi, cnt_A, cnt_B = mle_counts(N, m.Vmax, data, flags)
    m.pi[...] = normalize_counts(cnt_pi[None, :])[0][0]
    m.A[...], nonempty = normalize_counts(cnt_A)
    m.active[...] = nonempty            # state set = transitions.keys (DNBC.scala:51)
    m.B[...] = normalize_counts(cnt_B)[0]
    for c in range(Cn):
        k = int(hints[c])
        for j in range(N):
            x = edge_points(data, c, j, flags)
            if len(x) == 0:
                m.w[c, j, :] = 0.0
                continue
            cen0 = None if init_centroids is None else init_centroids.get((c, j))
            if cen0 is None:
                cen0 = pick_initial_centroids(x, k, rng)
            assign, _, _ = kmeans_1d(x, cen0)
            w0, mu0, var0 = em1d_initialize(x, k, assign)
            w, mu, var, ll = em1d_run(x, w0, mu0, var0)
            m.w[c, j, :k], m.mu[c, j, :k], m.var[c, j, :k] = w, mu, var
            if trace is not None:
                trace.append(EdgeTrace(c, j, len(x), np.array(cen0), (w0, mu0, var0), ll))
    return m
