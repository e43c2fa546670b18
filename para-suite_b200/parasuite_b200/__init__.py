"""parasuite_b200 -- B200-native hot path of PARA-suite (error profile + T>C pileup).

Host-side mirror of the reference tool classes over the C ABI of libparasuite_b200.so
(include/parasuite_b200.h).  The CUDA library is required; there is no CPU fallback.
"""
from . import abi
from .batch import PackedReference, ReadBatch, Record, parse_cigar

__all__ = ["abi", "PackedReference", "ReadBatch", "Record", "parse_cigar"]
