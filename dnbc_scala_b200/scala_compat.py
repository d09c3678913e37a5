"""Iteration order of Scala 2.11 immutable maps keyed by String.

The reference resolves decode ties by taking the FIRST maximum while iterating an
immutable `Map[String, _]` (`vprev.maxBy`, reference
core/src/main/scala/DynamicNaiveBayesianClassifier.scala:75, :85, :89).  To make "lowest
dense index wins" on the device mean the same thing, the host assigns dense state indices
in that iteration order:

* up to 4 keys: `Map1..Map4` iterate in insertion order;
* 5 or more keys: `HashMap.HashTrieMap`, whose sub-tries are stored by ascending 5-bit
  chunks of `improve(key.##)`, least-significant chunk first.

This is written from the published Scala 2.11.8 collection sources (SURVEY 8(c) "tie /
iteration-order recollection") and cannot be verified against a JVM in this environment;
it only matters at exact ties, which the device reports through the tie flags.
"""
from __future__ import annotations

from typing import List, Sequence

_M32 = 0xFFFFFFFF


def java_string_hash(s: str) -> int:
    """java.lang.String.hashCode as an unsigned 32-bit value (UTF-16 code units)."""
    h = 0
    data = s.encode("utf-16-be")
    for i in range(0, len(data), 2):
        h = (31 * h + ((data[i] << 8) | data[i + 1])) & _M32
    return h


def improve(hcode: int) -> int:
    """scala.collection.immutable.HashMap.improve (2.11)."""
    h = (hcode + (~(hcode << 9) & _M32)) & _M32
    h ^= h >> 14
    h = (h + (h << 4)) & _M32
    return h ^ (h >> 10)


def _trie_key(name: str):
    h = improve(java_string_hash(name))
    return tuple((h >> (5 * level)) & 31 for level in range(7))


def scala_map_order(keys_in_insertion_order: Sequence[str]) -> List[str]:
    """Keys in the order `Map.keys` / `Map.foreach` visits them."""
    keys = list(dict.fromkeys(keys_in_insertion_order))
    if len(keys) <= 4:
        return keys
    return sorted(keys, key=_trie_key)
