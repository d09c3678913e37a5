"""Reader/writer for the reference's `.data` text format -> dense arrays.

Format (reference DataSet.scala:9-57, writer TestUtils.scala:155-188): the first line
lists column types, column 0 (always written as "discrete") being the hidden-state
label; one line per timestep `label v1 v2 ...`; a line "." ends a sequence; a line
".." separates the learning set from the testing set.

The dense form is the one the C-ABI takes (include/dnbc_b200.h): labels i32 [P],
discrete symbols u8 [D][P], continuous f64 [C][P], offsets i64 [S+1]; string <->
index dictionaries stay on the host, exactly where the JVM shim keeps them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

UNKNOWN_SYMBOL = 255  # never-seen symbol: probability 0 (reference Edge.scala:70-73)


@dataclass
class Vocabulary:
    """String <-> index dictionaries for hidden states and each discrete variable."""

    states: List[str] = field(default_factory=list)
    symbols: List[List[str]] = field(default_factory=list)
    state_id: Dict[str, int] = field(default_factory=dict)
    symbol_id: List[Dict[str, int]] = field(default_factory=list)

    def add_state(self, name: str) -> int:
        i = self.state_id.get(name)
        if i is None:
            i = len(self.states)
            self.states.append(name)
            self.state_id[name] = i
        return i

    def add_symbol(self, d: int, name: str) -> int:
        while len(self.symbols) <= d:
            self.symbols.append([])
            self.symbol_id.append({})
        i = self.symbol_id[d].get(name)
        if i is None:
            i = len(self.symbols[d])
            if i >= UNKNOWN_SYMBOL:
                raise ValueError("more than 254 distinct symbols in one discrete variable")
            self.symbols[d].append(name)
            self.symbol_id[d][name] = i
        return i


@dataclass
class DenseSequences:
    """A batch of sequences in the dense layout of the C-ABI."""

    offsets: np.ndarray            # int64 [S+1]
    disc: np.ndarray               # uint8 [D, P]
    cont: np.ndarray               # float64 [C, P]
    labels: Optional[np.ndarray]   # int32 [P] or None

    @property
    def S(self) -> int:
        return len(self.offsets) - 1

    @property
    def P(self) -> int:
        return int(self.offsets[-1])

    @property
    def D(self) -> int:
        return self.disc.shape[0]

    @property
    def C(self) -> int:
        return self.cont.shape[0]

    def select(self, seq_ids: Sequence[int]) -> "DenseSequences":
        """Sub-batch made of the given sequences (copy)."""
        idx = [np.arange(self.offsets[s], self.offsets[s + 1]) for s in seq_ids]
        lens = [len(i) for i in idx]
        cat = np.concatenate(idx) if idx else np.zeros(0, np.int64)
        off = np.zeros(len(lens) + 1, np.int64)
        off[1:] = np.cumsum(lens)
        return DenseSequences(off, np.ascontiguousarray(self.disc[:, cat]),
                              np.ascontiguousarray(self.cont[:, cat]),
                              None if self.labels is None else np.ascontiguousarray(self.labels[cat]))


@dataclass
class DataSet:
    types: List[str]               # per observed column: "discrete" | "continuous"
    vocab: Vocabulary
    learn: DenseSequences
    test: DenseSequences

    @property
    def N(self) -> int:
        return len(self.vocab.states)

    @property
    def V(self) -> List[int]:
        return [len(s) for s in self.vocab.symbols]


def _densify(rows_per_seq, types, vocab: Vocabulary, grow: bool) -> DenseSequences:
    D = sum(t == "discrete" for t in types)
    C = sum(t == "continuous" for t in types)
    lens = [len(r) for r in rows_per_seq]
    P = sum(lens)
    off = np.zeros(len(lens) + 1, np.int64)
    off[1:] = np.cumsum(lens)
    disc = np.zeros((D, P), np.uint8)
    cont = np.zeros((C, P), np.float64)
    labels = np.zeros(P, np.int32)
    n = 0
    for rows in rows_per_seq:
        for parts in rows:
            name = parts[0]
            if grow:
                labels[n] = vocab.add_state(name)
            else:
                # a label never seen while learning can never be predicted
                labels[n] = vocab.state_id.get(name, -1)
            d = c = 0
            for typ, tok in zip(types, parts[1:]):
                if typ == "discrete":
                    if grow:
                        disc[d, n] = vocab.add_symbol(d, tok)
                    else:
                        disc[d, n] = vocab.symbol_id[d].get(tok, UNKNOWN_SYMBOL) if d < len(vocab.symbol_id) else UNKNOWN_SYMBOL
                    d += 1
                else:
                    cont[c, n] = float(tok)
                    c += 1
            n += 1
    return DenseSequences(off, disc, cont, labels)


def parse_lines(lines: Iterable[str]) -> DataSet:
    """Parse a whole `.data` stream (reference DataSetIterable/DataSetIterator).

    Like the reference iterator (DataSet.scala:16-33) a sequence ends at "." or at end
    of input, the learning set ends at ".."; an empty sequence produced by a trailing
    "." at end of file is dropped (the reference's iterator reports `finished` first).
    """
    it = iter(lines)
    header = next(it).rstrip("\n").split(" ")
    types = header[1:]
    for t in types:
        if t not in ("discrete", "continuous"):
            raise ValueError(f"unknown column type {t!r}")
    parts_learn, parts_test = [], []
    cur, target = [], parts_learn
    for raw in it:
        line = raw.rstrip("\n").rstrip("\r")
        if line == ".":
            target.append(cur)
            cur = []
        elif line == "..":
            target.append(cur)
            cur = []
            target = parts_test
        elif line == "":
            continue
        else:
            parts = line.split(" ")
            if len(parts) != len(types) + 1:
                # reference: DNBC.scala:146-148
                raise ValueError("Number of observed variables is inconsistent")
            cur.append(parts)
    if cur:
        target.append(cur)
    parts_learn = [r for r in parts_learn if r]
    parts_test = [r for r in parts_test if r]
    vocab = Vocabulary()
    D = sum(t == "discrete" for t in types)
    for d in range(D):
        vocab.symbols.append([])
        vocab.symbol_id.append({})
    learn = _densify(parts_learn, types, vocab, grow=True)
    test = _densify(parts_test, types, vocab, grow=False)
    return DataSet(types, vocab, learn, test)


def reorder_states_scala(ds: DataSet) -> DataSet:
    """Renumber hidden states in the iteration order of the reference's `transitions` map
    (first appearance as a transition source, then Scala's immutable-Map order), so that the
    device's "lowest dense index wins" tie-breaking means what `maxBy` over that map means
    (DynamicNaiveBayesianClassifier.scala:75, :85, :89; scala_compat.py).  In place."""
    from .scala_compat import scala_map_order
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


def load(path: str, reference_state_order: bool = False) -> DataSet:
    """Parse a `.data` file.  Hidden states are numbered in first-seen order; with
    reference_state_order=True they are renumbered in the reference's Map iteration order
    (reorder_states_scala), which is what decides EXACT decode ties on the device -- pass it
    whenever paths are to be compared with the reference's at ties (Measure does)."""
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rt") as fh:
            ds = parse_lines(fh)
    else:
        with open(path, "r") as fh:
            ds = parse_lines(fh)
    return reorder_states_scala(ds) if reference_state_order else ds


def write(path: str, types: Sequence[str], vocab: Vocabulary, learn: DenseSequences,
          test: Optional[DenseSequences] = None) -> None:
    """Writer in the reference's format (TestUtils.scala:155-188)."""
    def emit(fh, ds: DenseSequences, first: bool):
        for s in range(ds.S):
            if not first:
                fh.write(".\n")
            first = False
            for n in range(int(ds.offsets[s]), int(ds.offsets[s + 1])):
                toks = [vocab.states[int(ds.labels[n])]]
                d = c = 0
                for t in types:
                    if t == "discrete":
                        toks.append(vocab.symbols[d][int(ds.disc[d, n])])
                        d += 1
                    else:
                        toks.append(repr(float(ds.cont[c, n])))
                        c += 1
                fh.write(" ".join(toks) + "\n")

    with open(path, "w") as fh:
        fh.write("discrete" + "".join(" " + t for t in types) + "\n")
        emit(fh, learn, True)
        if test is not None and test.S:
            fh.write("..\n")
            emit(fh, test, True)
