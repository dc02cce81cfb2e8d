"""goetia_b200 -- a B200-native (sm_100a CUDA) backend for goetia's k-mer ingest hot path.

Scope: 2-bit packing of read batches, Lemire cyclic rolling hash
(Fwd/CanLemireShifter), batched insert/query into BitStorage / ByteStorage / NibbleStorage,
scaled-MinHash sketching, behind the reference's dBG<StorageType, ShifterType> surface.
All compute goes through the C ABI of include/goetia_b200.h (libgoetia_b200.so); there is no
CPU fallback.
"""
from . import _capi
from ._capi import GoetiaB200Error, init, MODE_BLIND, MODE_FAST, MODE_EXACT  # noqa: F401
from .hashing import FwdLemireShifter, CanLemireShifter, Hash, Canonical  # noqa: F401
from .storage import BitStorage, ByteStorage, NibbleStorage, get_n_primes_near_x  # noqa: F401
from .dbg import dBG  # noqa: F401
from .sketch import SourmashSketch  # noqa: F401
from .parsing import FastxParser, Record, SplitPairedReader  # noqa: F401
from .filters import DiginormFilter, FilterProcessor, StreamingSolidFilter  # noqa: F401
from .processors import InserterProcessor  # noqa: F401

__version__ = "0.1.0"
