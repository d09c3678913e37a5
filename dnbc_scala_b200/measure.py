"""Mirror of the reference's measurement harness.

Reference: `Performance.Measure(sparkContext, dataSetPath, hints, resourcesPath)`
(core/src/main/scala/Performance.scala:22-56): parse the learning part, `mle`, decode every
test sequence, report the mean per-sequence fraction of correctly decoded hidden states
(x100) and the learning / testing times.  Here the test set is decoded in ONE batched call.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native as nat
from . import dataset
from .model import DynamicNaiveBayesianClassifier
from .scala_compat import scala_map_order


@dataclass
class Performance:
    successRate: float        # average success rate on the test set, percent
    learningTime: int         # nanoseconds spent in mle
    testingTime: int          # nanoseconds spent decoding
    ties: int = 0             # decode steps the device flagged as tied


def accuracy(labels: np.ndarray, path: np.ndarray, offsets: np.ndarray) -> float:
    """Performance.scala:45-55: mean over sequences of the per-sequence fraction correct, x100."""
    ok = (labels == path.astype(np.int64)).astype(np.float64)
    csum = np.concatenate([[0.0], np.cumsum(ok)])
    lens = np.diff(offsets)
    keep = lens > 0
    frac = (csum[offsets[1:]] - csum[offsets[:-1]])[keep] / lens[keep]
    return float(frac.mean() * 100.0) if len(frac) else 0.0


def reorder_states_scala(ds: dataset.DataSet) -> dataset.DataSet:
    """Renumber hidden states in the iteration order of the reference's `transitions` map
    (first appearance as a transition source, then Scala's immutable-Map order), so device
    tie-breaking matches `maxBy` (scala_compat)."""
    L = ds.learn
    src_mask = np.ones(L.P, bool)
    src_mask[L.offsets[1:] - 1] = False            # the last step of a sequence is not a transition source
    src = L.labels[src_mask]
    _, first = np.unique(src, return_index=True)
    from_order = [ds.vocab.states[i] for i in src[np.sort(first)]]
    order = scala_map_order(from_order)
    order += [s for s in ds.vocab.states if s not in set(order)]
    remap = np.array([order.index(s) for s in ds.vocab.states], np.int32)
    for part in (ds.learn, ds.test):
        if part.labels is not None and part.P:
            lab = part.labels.copy()
            part.labels = np.where(lab >= 0, remap[np.maximum(lab, 0)], -1).astype(np.int32)
    ds.vocab.states = order
    ds.vocab.state_id = {s: i for i, s in enumerate(order)}
    return ds


def Measure(dataSetPath: str, hints: Optional[List[int]] = None, *, transition_quirk: bool = True,
            init_centroids=None, mode: int = nat.DECODE_REFERENCE, precision: int = nat.PREC_F64,
            device: int = 0) -> Performance:
    ds = reorder_states_scala(dataset.load(dataSetPath))
    t0 = time.perf_counter_ns()
    model = DynamicNaiveBayesianClassifier.mle_dense(ds.vocab.states, ds.vocab.symbols, ds.learn, hints,
                                                     transition_quirk=transition_quirk, double_count_first=False,
                                                     init_centroids=init_centroids, device=device)
    t1 = time.perf_counter_ns()
    path, _, tie = model.decode_dense(ds.test, mode, precision)
    t2 = time.perf_counter_ns()
    return Performance(accuracy(ds.test.labels, path, ds.test.offsets), t1 - t0, t2 - t1, int((tie != 0).sum()))
