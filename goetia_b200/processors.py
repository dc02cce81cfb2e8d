"""FileProcessor / InserterProcessor: mirror of goetia's file-level drivers over the C ABI.

Reference: ``FileProcessor<Derived, ParserType>`` (include/goetia/processors.hh:46-270) and
``InserterProcessor<InserterType, ParserType>`` (processors.hh:280-343) with the ``IntervalCounter`` they
keep time with (metrics.hh:104-145).  "Time" is the k-mers consumed: ``advance`` returns as soon as the
k-mers processed since the last interval boundary reach ``interval`` -- the record that crosses the
boundary is the last one consumed -- or the file ends (processors.hh:208-229).  The counter that did not reach the
interval when a file ended carries over to the next file, exactly like the reference's member ``timer``.

The records never pass through Python: ``gt_insert_fastx_advance`` / ``gt_insert_fastx`` parse into pinned
buffers and feed the device pipeline.
"""
import ctypes as C

from . import _capi
from ._capi import MODE_BLIND
from .parsing import FastxParser, _ERRORS


class InserterProcessor:
    """InserterProcessor<dBG<...>, FastxParser<>>.build(inserter, interval=500000, verbose=False)."""

    DEFAULT_INTERVAL = 500000  # IntervalCounter::DEFAULT_INTERVAL, metrics.hh:112

    def __init__(self, inserter, interval=DEFAULT_INTERVAL, verbose=False, mode=MODE_BLIND):
        if int(interval) <= 0:
            raise ValueError("interval must be positive")
        self.inserter = inserter
        self._interval = int(interval)
        self._verbose = bool(verbose)
        self._mode = int(mode)
        self._counter = 0      # IntervalCounter::_counter
        self._total = 0        # IntervalCounter::_total
        self._n_sequences = 0

    @classmethod
    def build(cls, inserter, interval=DEFAULT_INTERVAL, verbose=False, mode=MODE_BLIND):
        return cls(inserter, interval, verbose, mode)

    def n_sequences(self):
        return self._n_sequences

    def time_elapsed(self):
        return self._total

    def interval(self):
        return self._interval

    def advance(self, parser):
        """-> (total sequences processed, total time passed, whether sequences remain)."""
        g = self.inserter
        n_seqs, remaining = C.c_uint64(0), C.c_int(0)
        nk = _capi.lib().gt_insert_fastx_advance(g.S.handle, g.hasher.shifter_kind, g.K, parser.handle, self._mode,
                                                 self._interval - self._counter, C.byref(n_seqs), C.byref(remaining))
        if nk < 0:
            raise _ERRORS.get(int(nk), _capi.GoetiaB200Error)("gt_insert_fastx_advance: " + _capi.last_error())
        self._n_sequences += int(n_seqs.value)
        self._total += int(nk)
        self._counter = 0 if remaining.value else self._counter + int(nk)
        if self._mode == MODE_BLIND:
            g.S.flush()  # the caller may look at the graph between two intervals
        return self._n_sequences, self._total, bool(remaining.value)

    def process(self, filename_or_parser, strict=False, min_length=0, by_interval=False):
        """Consume a whole file -> (sequences processed, time passed), both cumulative over this processor's life.
        The reference loops over ``advance``; nothing observes the boundaries in between, so by default the file goes
        through the whole-file pipeline ``gt_insert_fastx`` and a fresh interval starts afterwards.
        ``by_interval=True`` runs the reference's loop literally (the interval counter then carries over exactly)."""
        own = not isinstance(filename_or_parser, FastxParser)
        parser = FastxParser(filename_or_parser, strict, min_length) if own else filename_or_parser
        try:
            if by_interval:
                remaining = True
                while remaining:
                    _, _, remaining = self.advance(parser)
                return self._n_sequences, self._total
            g = self.inserter
            n_seqs = C.c_uint64(0)
            nk = _capi.lib().gt_insert_fastx(g.S.handle, g.hasher.shifter_kind, g.K, parser.handle, self._mode, 0, C.byref(n_seqs))
            if nk < 0:
                raise _ERRORS.get(int(nk), _capi.GoetiaB200Error)("gt_insert_fastx: " + _capi.last_error())
            if self._mode == MODE_BLIND:
                g.S.flush()
            self._n_sequences += int(n_seqs.value)
            self._total += int(nk)
            self._counter = 0
            return self._n_sequences, self._total
        finally:
            if own:
                parser.close()
