"""ShifterType mirror: FwdLemireShifter / CanLemireShifter (include/goetia/hashing/hashshifter.hh:74-205).

Batch hashing (``hash_sequences``, ``hashes``, ``hash``) runs the K1 kernel through the C ABI.
The cursor members (``hash_base`` / ``shift_right`` / ``shift_left`` / ``get``) are the
reference's single-k-mer latency path, which stays on the host by design (SURVEY.md section
3e); they are a few integer ops on Python ints with the same table constants.
"""
import numpy as np

from . import _capi

_M64 = (1 << 64) - 1
# include/goetia/hashing/rollinghash/characterhash.h:27-113 -- the entries a validated read can touch
_T = {"A": 16664410744025174816, "C": 15956807086001210932, "G": 9404339731978646439, "T": 836480985777824379}
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


class InvalidSequenceException(ValueError):
    """hashshifter.hh:148-150 -- sequence shorter than K."""


class UninitializedShifterException(RuntimeError):
    """hashshifter.hh:133-145 -- shift before hash_base."""


def _rotl(x, r):
    r &= 63
    return ((x << r) | (x >> (64 - r))) & _M64 if r else x


def _rotr(x, r):
    r &= 63
    return ((x >> r) | (x << (64 - r))) & _M64 if r else x


class Hash:
    """hashing/canonical.hh:31-60 ``Hash<uint64_t>``."""

    __slots__ = ("hash",)

    def __init__(self, h=0):
        self.hash = int(h)

    def value(self):
        return self.hash

    def __int__(self):
        return self.hash

    def __index__(self):
        return self.hash

    def __eq__(self, o):
        return int(self) == int(o)

    def __hash__(self):
        return hash(self.hash)

    def __lt__(self, o):
        return int(self) < int(o)

    def __repr__(self):
        return "<Hash h=%d>" % self.hash


class Canonical:
    """hashing/canonical.hh:70-147 ``Canonical<uint64_t>``: value() = min(fw, rc) (:124-126)."""

    __slots__ = ("fw_hash", "rc_hash")

    def __init__(self, fw=0, rc=0):
        self.fw_hash = int(fw)
        self.rc_hash = int(rc)

    def value(self):
        return self.fw_hash if self.fw_hash < self.rc_hash else self.rc_hash

    def sign(self):  # canonical.hh:120-122
        return self.fw_hash < self.rc_hash

    def __int__(self):
        return self.value()

    def __index__(self):
        return self.value()

    def __eq__(self, o):
        return int(self) == int(o)

    def __hash__(self):
        return hash(self.value())

    def __lt__(self, o):
        return int(self) < int(o)

    def __repr__(self):
        return "<Canonical fw=%d rc=%d>" % (self.fw_hash, self.rc_hash)


def _validated(seq):
    if isinstance(seq, bytes):
        seq = seq.decode("ascii")
    return seq


class _LemireShifter:
    """HashShifter<LemireShifterPolicy<...>> (hashshifter.hh:74-199)."""

    shifter_kind = None
    hash_type = None
    NAME = None

    def __init__(self, K, start=None):
        if start is not None and not isinstance(K, int):
            K, start = start, K  # (start, K) ctor order of the reference
        self.K = int(K)
        if not 1 <= self.K <= 65535:
            raise ValueError("K out of range")
        self._fw = 0
        self._rc = 0
        self._init = False
        if start is not None:
            self.hash_base(start)

    @classmethod
    def build(cls, K, *args):
        return cls(K, *args)

    # -- cursor path (host; reference keeps it scalar too) ---------------------------------
    def is_initialized(self):
        return self._init

    def hash_base(self, seq):
        """hash_base_impl: rollinghashshifter.hh:69-79 (fwd), :183-197 (canonical)."""
        seq = _validated(seq)
        if len(seq) < self.K:
            raise InvalidSequenceException("Sequence must at least length K")
        fw = rc = 0
        K = self.K
        for i in range(K):
            fw = _rotl(fw, 1) ^ _T[seq[i]]
            rc = _rotl(rc, 1) ^ _T[_COMP[seq[K - 1 - i]]]
        self._fw, self._rc, self._init = fw, rc, True
        return self.get()

    def get(self):
        if self.shifter_kind == _capi.SHIFTER_CAN:
            return Canonical(self._fw, self._rc)
        return Hash(self._fw)

    def shift_right(self, out, inc):
        """shift_right_impl: rollinghashshifter.hh:103-106, :203-208; cyclichash.h:85-101."""
        if not self._init:
            raise UninitializedShifterException()
        K = self.K
        self._fw = _rotl(self._fw, 1) ^ _rotl(_T[out], K) ^ _T[inc]
        self._rc = _rotr(self._rc ^ _rotl(_T[_COMP[inc]], K) ^ _T[_COMP[out]], 1)
        return self.get()

    def shift_left(self, inc, out):
        """shift_left_impl: rollinghashshifter.hh:97-100, :214-219."""
        if not self._init:
            raise UninitializedShifterException()
        K = self.K
        self._fw = _rotr(self._fw ^ _rotl(_T[inc], K) ^ _T[out], 1)
        self._rc = _rotl(self._rc, 1) ^ _rotl(_T[_COMP[out]], K) ^ _T[_COMP[inc]]
        return self.get()

    # -- batch path (GPU) -------------------------------------------------------------------
    def hash_sequences(self, bases, offsets):
        """KmerIterator over every read on the GPU.  Returns (fw, rc-or-None, status)."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        cap = int(np.maximum(lens - self.K + 1, 0).sum())
        fw = np.zeros(max(cap, 1), dtype=np.uint64)
        rc = np.zeros(max(cap, 1), dtype=np.uint64) if self.shifter_kind == _capi.SHIFTER_CAN else None
        status = np.zeros(max(n, 1), dtype=np.uint8)
        tot = _capi.check(L.gt_hash_sequences(self.shifter_kind, self.K, bases.ctypes.data, offsets.ctypes.data, n,
                                              fw.ctypes.data, rc.ctypes.data if rc is not None else None,
                                              status.ctypes.data), "gt_hash_sequences")
        return fw[:tot], (rc[:tot] if rc is not None else None), status[:n]

    def hashes(self, seq):
        """All k-mer hashes of one sequence, in order (pythonize_dbg.py:24-28)."""
        seq = _validated(seq)
        if len(seq) < self.K:
            raise InvalidSequenceException("Sequence must have length >= K")
        bases, offsets = _capi.reads_from_strings([seq])
        fw, rc, status = self.hash_sequences(bases, offsets)
        if status[0] & _capi.READ_INVALID:
            raise ValueError("sequence holds a non-ACGT character")
        if rc is None:
            return [Hash(int(h)) for h in fw]
        return [Canonical(int(a), int(b)) for a, b in zip(fw, rc)]

    def hash(self, seq, K=None):
        """hash of the first K characters (hashshifter.hh:172-194; longer input is truncated,
        tests/test_dbg.py:125-128)."""
        seq = _validated(seq)
        if len(seq) < self.K:
            raise InvalidSequenceException("Sequence must at least length K")
        return self.hashes(seq[:self.K])[0]


class FwdLemireShifter(_LemireShifter):
    shifter_kind = _capi.SHIFTER_FWD
    hash_type = Hash
    NAME = "FwdLemireShifter"


class CanLemireShifter(_LemireShifter):
    shifter_kind = _capi.SHIFTER_CAN
    hash_type = Canonical
    NAME = "CanLemireShifter"


types = [FwdLemireShifter, CanLemireShifter]
typenames = [(t, t.NAME) for t in types]
