"""Host-side mirror of the reference's model API over the C ABI.

Reference (core/src/main/scala/DynamicNaiveBayesianClassifier.scala):
  object DynamicNaiveBayesianClassifier.mle(sc, sequences, continuousVariableHints)   :124-161
  class  DynamicNaiveBayesianClassifier(initialEdge, transitions, discreteEmissions,
                                        continuousEmissions)                          :14-17
         .inferMostLikelyHiddenStates(observedStates): List[String]                   :24-28
         .score(observedStates): Double                                               :35-39
  ObservedState(discreteVariables, continuousVariables)                               :99-102
  State(hiddenState, observedState)                                                   :109-112

Same names, argument meaning and error behaviour (plain `Exception` with the reference's
messages); the SparkContext argument has no counterpart.  The String <-> index dictionaries
live here, exactly where the JVM shim keeps them; everything numeric runs
in libdnbc_b200.so on the GPU.  Batched calls and Baum-Welch are additive.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from .dataset import UNKNOWN_SYMBOL, DenseSequences
from .scala_compat import scala_map_order

MSG_INCONSISTENT = "Number of observed variables is inconsistent"
MSG_HINT_COUNT = "Number of continuous variable hints does not correspond to the number of continous variables"
MSG_HINT_VALUE = "Number of suggested normal distributions in a continuous variable must be at least one"


@dataclass(frozen=True)
class ObservedState:
    DiscreteVariables: Tuple[str, ...]
    ContinuousVariables: Tuple[float, ...]

    def __init__(self, discreteVariables: Sequence[str], continuousVariables: Sequence[float]):
        object.__setattr__(self, "DiscreteVariables", tuple(discreteVariables))
        object.__setattr__(self, "ContinuousVariables", tuple(float(x) for x in continuousVariables))


@dataclass(frozen=True)
class State:
    HiddenState: str
    ObservedState: ObservedState

    def __init__(self, hiddenState: str, observedState: ObservedState):
        object.__setattr__(self, "HiddenState", hiddenState)
        object.__setattr__(self, "ObservedState", observedState)


class DynamicNaiveBayesianClassifier:
    """Learned model.  Dense index order of the hidden states follows the iteration order of
    the reference's `transitions` map (scala_compat), so device tie-breaking (lowest index)
    equals the reference's `maxBy` (first maximum)."""

    def __init__(self, initialEdge: Dict[str, float], transitions: Dict[str, Dict[str, float]],
                 discreteEmissions: List[Dict[str, Dict[str, float]]],
                 continuousEmissions: List[Dict[str, Tuple[Sequence[float], Sequence[float], Sequence[float]]]],
                 device: int = 0, _dense=None):
        if _dense is not None:
            (self.states, self.symbols, self.ctx) = _dense
            return
        # the 4-map constructor (:14-17): state set and order come from `transitions`
        order = scala_map_order(list(transitions.keys()))
        extra = [s for s in initialEdge if s not in transitions]
        for row in transitions.values():
            extra += [s for s in row if s not in transitions and s not in extra]
        self.states = order + list(dict.fromkeys(extra))
        sid = {s: i for i, s in enumerate(self.states)}
        N = len(self.states)
        self.symbols = []
        for edge in discreteEmissions:
            syms: List[str] = []
            for tab in edge.values():
                syms += [v for v in tab if v not in syms]
            self.symbols.append(syms)
        V = [max(1, len(s)) for s in self.symbols]
        K = [max([1] + [len(g[0]) for g in edge.values()]) for edge in continuousEmissions]
        self.ctx = nat.Context(N, V, K, device)
        c = self.ctx
        pi = np.zeros(N)
        A = np.zeros((N, N))
        B = np.zeros((c.D, N, c.Vmax))
        w = np.zeros((c.C, N, c.Kmax))
        mu = np.zeros_like(w)
        var = np.ones_like(w)
        active = np.zeros(N, np.uint8)
        for s, p in initialEdge.items():
            pi[sid[s]] = p
        for s, row in transitions.items():
            active[sid[s]] = 1
            for t, p in row.items():
                A[sid[s], sid[t]] = p
        for d, edge in enumerate(discreteEmissions):
            for s, tab in edge.items():
                for v, p in tab.items():
                    B[d, sid[s], self.symbols[d].index(v)] = p
        for ci, edge in enumerate(continuousEmissions):
            for s, (ww, mm, vv) in edge.items():
                k = len(ww)
                w[ci, sid[s], :k], mu[ci, sid[s], :k], var[ci, sid[s], :k] = ww, mm, vv
        c.set_params(pi, A, B, w, mu, var, active)

    # ------------------------------------------------------------------ learning
    @classmethod
    def mle(cls, sequences: Iterable[Sequence[State]], continuousVariableHints: Optional[Sequence[int]] = None, *,
            transition_quirk: bool = True, double_count_first: Optional[bool] = None, init_centroids=None,
            device: int = 0) -> "DynamicNaiveBayesianClassifier":
        """Supervised maximum likelihood (reference `mle`, :124-161).

        transition_quirk: reproduce learnTransitions' membership test against the wrong map
          (:204-216).  double_count_first: the reference learns the first sequence
          twice when `sequences` is re-iterable (`firstSequence ++ sequences`, :129-133); by
          default that is inferred the same way (a list/tuple is re-iterable, an iterator is not).
        init_centroids: {(continuous variable, state name): [k initial centroids]} replacing the
          reference's unseeded random k-means initialisation (KMeans.java:46-72).
        """
        if double_count_first is None:
            double_count_first = isinstance(sequences, (list, tuple))
        seqs = [list(s) for s in sequences]
        if not seqs or not seqs[0]:
            raise Exception("empty learning set")
        first = seqs[0][0].ObservedState
        D, C = len(first.DiscreteVariables), len(first.ContinuousVariables)
        hints = list(continuousVariableHints or [])
        if hints:  # checkValidHints :270-277
            if len(hints) != C:
                raise Exception(MSG_HINT_COUNT)
            if any(h < 1 for h in hints):
                raise Exception(MSG_HINT_VALUE)
        else:
            hints = [1] * C
        from_order: List[str] = []
        all_states: List[str] = []
        symbols: List[List[str]] = [[] for _ in range(D)]
        for seq in seqs:
            for t, st in enumerate(seq):
                o = st.ObservedState
                if len(o.DiscreteVariables) != D or len(o.ContinuousVariables) != C:
                    raise Exception(MSG_INCONSISTENT)  # :146-148
                all_states.append(st.HiddenState)
                if t + 1 < len(seq):
                    from_order.append(st.HiddenState)
                for d, v in enumerate(o.DiscreteVariables):
                    if v not in symbols[d]:
                        symbols[d].append(v)
        dense = cls._densify(seqs, from_order, all_states, symbols, D, C)
        return cls._mle_dense(dense[0], dense[1], symbols, hints, transition_quirk, double_count_first,
                              init_centroids, device)

    @staticmethod
    def _densify(seqs, from_order, all_states, symbols, D, C):
        order = scala_map_order(from_order)
        seen = set(order)
        states = order + [s for s in dict.fromkeys(all_states) if s not in seen]
        sid = {s: i for i, s in enumerate(states)}
        symid = [{v: i for i, v in enumerate(sy)} for sy in symbols]
        P = sum(len(s) for s in seqs)
        off = np.zeros(len(seqs) + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        disc = np.zeros((D, P), np.uint8)
        cont = np.zeros((C, P))
        labels = np.zeros(P, np.int32)
        n = 0
        for seq in seqs:
            for st in seq:
                labels[n] = sid[st.HiddenState]
                for d, v in enumerate(st.ObservedState.DiscreteVariables):
                    disc[d, n] = symid[d][v]
                cont[:, n] = st.ObservedState.ContinuousVariables
                n += 1
        return states, DenseSequences(off, disc, cont, labels)

    @classmethod
    def _mle_dense(cls, states, data: DenseSequences, symbols, hints, transition_quirk, double_count_first,
                   init_centroids, device):
        for sy in symbols:
            if len(sy) > 254:
                raise Exception("more than 254 distinct symbols in one discrete variable")
        ctx = nat.Context(len(states), [max(1, len(s)) for s in symbols], hints, device)
        ctx.upload(data.offsets, data.disc, data.cont, data.labels)
        flags = (nat.QUIRK_TRANSITIONS if transition_quirk else 0) | (nat.DOUBLE_COUNT_FIRST if double_count_first else 0)
        cen = None
        if init_centroids is not None:
            cen = np.zeros((ctx.C, ctx.N, ctx.Kmax))
            if isinstance(init_centroids, dict):
                # evenly spread defaults for edges the caller did not name
                for c in range(ctx.C):
                    for j in range(ctx.N):
                        x = data.cont[c, data.labels == j]
                        k = int(hints[c])
                        if len(x):
                            cen[c, j, :k] = x.min() + (x.max() - x.min()) * ((np.arange(k) + 0.5) / k)
                for (c, name), vals in init_centroids.items():
                    cen[c, states.index(name), :len(vals)] = vals
            else:
                cen[...] = np.asarray(init_centroids).reshape(cen.shape)
        ll, its = ctx.mle(flags, cen)
        model = cls({}, {}, [], [], _dense=(list(states), [list(s) for s in symbols], ctx))
        model.gmm_loglik_series, model.gmm_iterations = ll, its
        return model

    @classmethod
    def mle_dense(cls, states: Sequence[str], symbols: Sequence[Sequence[str]], data: DenseSequences,
                  continuousVariableHints: Optional[Sequence[int]] = None, *, transition_quirk: bool = True,
                  double_count_first: bool = False, init_centroids=None, device: int = 0):
        """`mle` for data already in the dense layout (dataset.load); `states` must list the
        hidden-state names by dense index."""
        C = data.C
        hints = list(continuousVariableHints or [])
        if hints:
            if len(hints) != C:
                raise Exception(MSG_HINT_COUNT)
            if any(h < 1 for h in hints):
                raise Exception(MSG_HINT_VALUE)
        else:
            hints = [1] * C
        return cls._mle_dense(list(states), data, symbols, hints, transition_quirk, double_count_first,
                              init_centroids, device)

    # ------------------------------------------------------------------ the four maps
    def _params(self):
        return self.ctx.get_params()

    @property
    def initialEdge(self) -> Dict[str, float]:
        p = self._params()
        return {s: float(p["pi"][i]) for i, s in enumerate(self.states) if p["pi"][i] > 0}

    @property
    def transitions(self) -> Dict[str, Dict[str, float]]:
        p = self._params()
        return {s: {t: float(p["A"][i, j]) for j, t in enumerate(self.states) if p["A"][i, j] > 0}
                for i, s in enumerate(self.states) if p["active"][i]}

    @property
    def discreteEmissions(self) -> List[Dict[str, Dict[str, float]]]:
        p = self._params()
        return [{s: {v: float(p["B"][d, i, q]) for q, v in enumerate(self.symbols[d]) if p["B"][d, i, q] > 0}
                 for i, s in enumerate(self.states) if p["B"][d, i].sum() > 0} for d in range(self.ctx.D)]

    @property
    def continuousEmissions(self):
        """Per variable: {state: (weights, means, variances)} -- what LearnedContinuousEdge.getModel
        exposes as a GaussianMixtureModel (Edge.scala:120)."""
        p = self._params()
        out = []
        for c in range(self.ctx.C):
            k = int(self.ctx.K[c])
            out.append({s: (p["w"][c, i, :k].copy(), p["mu"][c, i, :k].copy(), p["var"][c, i, :k].copy())
                        for i, s in enumerate(self.states) if p["w"][c, i, :k].sum() > 0})
        return out

    # ------------------------------------------------------------------ inference
    def _encode(self, sequences: Sequence[Sequence[ObservedState]]) -> DenseSequences:
        D, C = self.ctx.D, self.ctx.C
        symid = [{v: i for i, v in enumerate(sy)} for sy in self.symbols]
        P = sum(len(s) for s in sequences)
        off = np.zeros(len(sequences) + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in sequences])
        disc = np.zeros((D, P), np.uint8)
        cont = np.zeros((C, P))
        n = 0
        for seq in sequences:
            for o in seq:
                if len(o.DiscreteVariables) != D or len(o.ContinuousVariables) != C:
                    raise Exception(MSG_INCONSISTENT)  # checkValidObservations :41-47
                for d, v in enumerate(o.DiscreteVariables):
                    disc[d, n] = symid[d].get(v, UNKNOWN_SYMBOL)
                cont[:, n] = o.ContinuousVariables
                n += 1
        return DenseSequences(off, disc, cont, None)

    def decode_dense(self, data: DenseSequences, mode: int = nat.DECODE_REFERENCE, precision: int = nat.PREC_F64):
        """Batched decode of dense sequences: (path u16 [P], score f64 [S], tie flags u8 [P])."""
        self.ctx.upload(data.offsets, data.disc, data.cont, data.labels)
        return self.ctx.decode(mode, precision)

    def inferMostLikelyHiddenStatesBatch(self, sequences, mode=nat.DECODE_REFERENCE, precision=nat.PREC_F64):
        data = self._encode(sequences)
        path, _, _ = self.decode_dense(data, mode, precision)
        return [[self.states[j] for j in path[data.offsets[s]:data.offsets[s + 1]]] for s in range(data.S)]

    def inferMostLikelyHiddenStates(self, observedStates: Sequence[ObservedState]) -> List[str]:
        """Reference :24-28 (per-step argmax of the reference's recursion)."""
        return self.inferMostLikelyHiddenStatesBatch([list(observedStates)])[0]

    def score(self, observedStates: Sequence[ObservedState]) -> float:
        """Reference :35-39: sum_j delta_{T-1}(j), 0.0 for a single observation."""
        data = self._encode([list(observedStates)])
        _, score, _ = self.decode_dense(data)
        return float(score[0])

    # ------------------------------------------------------------------ Baum-Welch (additive)
    def baum_welch(self, data: DenseSequences, n_iters: int) -> np.ndarray:
        """Unsupervised EM from the current parameters; returns the log-likelihood entering
        each iteration.  No reference counterpart."""
        self.ctx.upload(data.offsets, data.disc, data.cont, data.labels, cont_dtype=np.float32)
        return self.ctx.em_iterate(n_iters)

    infer_most_likely_hidden_states = inferMostLikelyHiddenStates
