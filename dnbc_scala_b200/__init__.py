"""dnbc_scala_b200 -- B200-native learning / decoding hot path of lucivpav/dnbc-scala.

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/dnbc_b200.h), _native (ctypes
binding of that ABI), model (mirror of the reference's DynamicNaiveBayesianClassifier API),
dataset (`.data` text format <-> dense arrays), measure (Performance.Measure harness),
distributed (sequence sharding + the per-iteration statistics all-reduce).
There is no CPU compute path: without libdnbc_b200.so and a B200 every numeric call fails.
"""
from . import _native
from ._native import (DECODE_BACKTRACE, DECODE_REFERENCE, DOUBLE_COUNT_FIRST, PREC_F32, PREC_F64,
                      QUIRK_TRANSITIONS, Context, DnbcError)
from .dataset import DataSet, DenseSequences, Vocabulary, load, parse_lines, write
from .model import DynamicNaiveBayesianClassifier, ObservedState, State

__all__ = ["Context", "DnbcError", "DynamicNaiveBayesianClassifier", "ObservedState", "State", "DataSet",
           "DenseSequences", "Vocabulary", "load", "parse_lines", "write", "DECODE_REFERENCE", "DECODE_BACKTRACE",
           "PREC_F32", "PREC_F64", "QUIRK_TRANSITIONS", "DOUBLE_COUNT_FIRST"]
