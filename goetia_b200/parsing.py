"""FASTX front end: mirror of goetia's parsing surface over the C ABI (gt_fastx_*).

Reference: ``FastxParser<DNA_SIMPLE>`` (include/goetia/parsing/readers.hh:86-219) over kseq
(parsing/kseq.h:174-214), ``Record`` (parsing/parsing.hh:26-58) and the exceptions the parser
raises (parsing/parsing.hh, goetia.hh:140-178).  The parser itself is C++ inside
libgoetia_b200.so (csrc/fastx_host.inc): a reader thread inflates ahead, records are located with
memchr, and ``next_batch`` hands whole batches to the device pipeline.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import GoetiaB200Error


class NoMoreReadsAvailable(GoetiaB200Error):
    pass


class InvalidRead(GoetiaB200Error):
    pass


class GoetiaFileException(GoetiaB200Error):
    pass


class InvalidCharacterException(GoetiaB200Error):
    pass


_ERRORS = {-2: InvalidRead, -3: GoetiaFileException, -4: InvalidCharacterException, -5: NoMoreReadsAvailable}


class Record:
    """parsing.hh:26-58"""
    __slots__ = ("name", "sequence", "quality")

    def __init__(self, name="", sequence="", quality=""):
        self.name, self.sequence, self.quality = name, sequence, quality

    def write_fastx(self, out):
        if self.quality:
            out.write("@%s\n%s\n+\n%s\n" % (self.name, self.sequence, self.quality))
        else:
            out.write(">%s\n%s\n" % (self.name, self.sequence))

    def __repr__(self):
        return "<Sequence name=%s seq=%s>" % (self.name, self.sequence)


class FastxParser:
    """FastxParser<DNA_SIMPLE>(infile, strict=False, min_length=0)."""

    def __init__(self, infile, strict=False, min_length=0):
        L = _capi.load()  # host-only: no GPU needed to parse
        self._h = L.gt_fastx_open(str(infile).encode(), int(bool(strict)), int(min_length))
        if not self._h:
            raise GoetiaFileException(_capi.last_error())
        self.filename = str(infile)

    @classmethod
    def build(cls, filename, strict=False, min_length=0):
        return cls(filename, strict, min_length)

    @property
    def handle(self):
        return self._h

    def _raise(self, rc):
        raise _ERRORS.get(int(rc), GoetiaB200Error)(_capi.last_error())

    def next(self):
        """std::optional<Record>: a Record, or None for a skipped record / the end of the file."""
        L = _capi.load()
        p = [C.c_char_p() for _ in range(3)]
        n = [C.c_uint64() for _ in range(3)]
        rc = L.gt_fastx_next_record(self._h, C.byref(p[0]), C.byref(n[0]), C.byref(p[1]), C.byref(n[1]),
                                    C.byref(p[2]), C.byref(n[2]))
        if rc < 0:
            self._raise(rc)
        if rc == 0:
            return None
        s = [C.string_at(p[i], n[i].value).decode("latin-1") if n[i].value else "" for i in range(3)]
        return Record(s[0], s[1], s[2])

    def __iter__(self):
        while not self.is_complete():
            r = self.next()
            if r is not None:
                yield r

    def next_batch(self, max_bases=64 << 20, max_reads=None):
        """(bases uint8, offsets uint64) of the next kept records; offsets.size == 1 at the end of the file."""
        L = _capi.load()
        max_reads = int(max_reads) if max_reads else max(1024, max_bases // 32)
        bases = np.empty(max_bases, dtype=np.uint8)
        offsets = np.empty(max_reads + 1, dtype=np.uint64)
        n = L.gt_fastx_next_batch(self._h, bases.ctypes.data, max_bases, offsets.ctypes.data, max_reads)
        if n < 0:
            self._raise(n)
        return bases[:int(offsets[n])], offsets[:n + 1]

    def next_packed_batch(self, max_bases=64 << 20, max_reads=None):
        """The next kept records 2-bit packed, as gt_insert_sequences_packed / dBG.insert_sequences_packed take them:
        (words uint64, offsets uint64, flags uint8, n_records).  Entries flagged READ_INVALID are alignment gaps between the
        pieces the parser's workers packed, not records; offsets.size == 1 at the end of the file."""
        L = _capi.load()
        max_reads = int(max_reads) if max_reads else max(1024, max_bases // 32)
        words = np.zeros(max_bases // 32 + 2, dtype=np.uint64)
        offsets = np.zeros(max_reads + 1, dtype=np.uint64)
        flags = np.zeros(max_reads, dtype=np.uint8)
        n_real = C.c_uint64(0)
        n = L.gt_fastx_next_packed_batch(self._h, words.ctypes.data, words.size, offsets.ctypes.data, flags.ctypes.data, max_reads,
                                         C.byref(n_real))
        if n < 0:
            self._raise(n)
        return words[:(int(offsets[n]) + 31) // 32], offsets[:n + 1], flags[:n], int(n_real.value)

    def _stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_int()
        _capi.check(_capi.load().gt_fastx_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), "gt_fastx_stats")
        return a.value, b.value, bool(c.value)

    def n_parsed(self):
        return self._stats()[0]

    def n_skipped(self):
        return self._stats()[1]

    def is_complete(self):
        return self._stats()[2]

    def close(self):
        if getattr(self, "_h", None):
            _capi.load().gt_fastx_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _split_on_first(name, delims=" \t"):
    """split_on_first, parsing.cc:31-54, statement for statement (including its quirk: a delimiter at index 0
    leaves `left` empty and the scan goes on to the next one)."""
    left = right = ""
    for i, c in enumerate(name):
        if left == "":
            if c in delims:
                left = name[:i]
        elif c not in delims:
            right = name[i:]
            break
    if left == "":
        left = name
    return left, right


def check_is_pair(left, right):
    """parsing.cc:56-86: do two record names belong to one read pair (/1 /2, or Illumina '1:' '2:' comments)?"""
    lf, ls = _split_on_first(left)
    rf, rs = _split_on_first(right)
    if lf.endswith("/1") and rf.endswith("/2"):
        return _split_on_first(lf, "/")[0] != "" and _split_on_first(lf, "/")[0] == _split_on_first(rf, "/")[0]
    if lf == rf and ls.endswith("1:") and rs.endswith("2:"):
        return True
    if lf == rf and ls.endswith("/1") and rs.endswith("/2"):
        return _split_on_first(ls, "/")[0] != "" and _split_on_first(ls, "/")[0] == _split_on_first(rs, "/")[0]
    return False


class SplitPairedReader:
    """SplitPairedReader<FastxParser<DNA_SIMPLE>> (parsing/readers.hh:234-340): two files read in lock step."""

    def __init__(self, left, right, strict=False, min_length=0, force_name_match=False):
        self.left_parser = FastxParser(left, strict, min_length)
        self.right_parser = FastxParser(right, strict, min_length)
        self._strict, self._force_name_match, self._n_skipped = bool(strict), bool(force_name_match), 0

    @classmethod
    def build(cls, left, right, strict=False, min_length=0, force_name_match=False):
        return cls(left, right, strict, min_length, force_name_match)

    def is_complete(self):
        if self.left_parser.is_complete() != self.right_parser.is_complete():
            raise GoetiaB200Error("Mismatched split paired files.")
        return self.left_parser.is_complete()

    def next(self):
        """RecordPair: (left or None, right or None)."""
        if self.is_complete():
            raise NoMoreReadsAvailable("NoMoreReadsAvailable")
        pair, errs = [], []
        for p in (self.left_parser, self.right_parser):
            try:
                pair.append(p.next())
                errs.append(None)
            except (InvalidCharacterException, InvalidRead) as e:
                pair.append(None)
                errs.append(e)
        if self._strict:
            for e in errs:
                if e is not None:
                    raise e
        left, right = pair
        if self._force_name_match and left is not None and right is not None and not check_is_pair(left.name, right.name):
            if self._strict:
                raise GoetiaB200Error("Unpaired reads")
            self._n_skipped += 2
            return None, None
        return left, right

    def __iter__(self):
        while not self.is_complete():
            yield self.next()

    def n_skipped(self):
        return self._n_skipped + self.left_parser.n_skipped() + self.right_parser.n_skipped()

    def close(self):
        self.left_parser.close()
        self.right_parser.close()
