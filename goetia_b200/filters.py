"""Read filters on the GPU dBG: DiginormFilter (include/goetia/diginorm.hh:25-125), StreamingSolidFilter
(include/goetia/solidifier.hh:33-77) and their FilterProcessor (include/goetia/processors.hh:347-430).

The reference judges and inserts one read at a time.  By DEFAULT the filters here give exactly that result for a whole
batch: DiginormFilter through gt_diginorm_sequences_serial (rounds of judge / claim / verify / insert that decide what
the serial loop decides), StreamingSolidFilter through gt_insert_and_query_sequences (the serial post-insert counts of
every k-mer of the batch).  ``DiginormFilter(..., batch_synchronous=True)`` opts into the one-round approximation
(every read of a batch judged against the table state at the start of the batch -- SURVEY.md section 8a): faster, but
redundancy INSIDE a batch is not removed, so its output differs from the reference's.
"""
import ctypes as C

import numpy as np

from . import _capi
from .parsing import FastxParser


class DiginormFilter:
    """``DiginormFilter[dBG].Filter``: ``build(graph, cutoff)``, ``filter_sequence``."""

    def __init__(self, graph, cutoff, batch_synchronous=False):
        self.graph, self.K, self.cutoff = graph, graph.K, int(cutoff)
        self.batch_synchronous = bool(batch_synchronous)

    @classmethod
    def build(cls, graph, cutoff, batch_synchronous=False):
        return cls(graph, cutoff, batch_synchronous)

    @staticmethod
    def median_count_at_least(sequence, cutoff, graph):
        """diginorm.hh:35-68"""
        bases, offsets = _capi.reads_from_strings([sequence])
        return bool(graph.median_count_at_least(bases, offsets, cutoff)[0])

    def filter_sequences(self, bases, offsets):
        """One batch: (keep uint8[n_reads], k-mers judged).  Kept reads have been inserted.  Serial semantics (the
        reference's) unless the filter was built with batch_synchronous=True."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        keep = np.zeros(max(n, 1), dtype=np.uint8)
        n_kept = C.c_uint64(0)
        fn, name = ((L.gt_diginorm_sequences, "gt_diginorm_sequences") if self.batch_synchronous
                    else (L.gt_diginorm_sequences_serial, "gt_diginorm_sequences_serial"))
        nk = _capi.check(fn(self.graph.S.handle, self.graph.hasher.shifter_kind, self.K, bases.ctypes.data, offsets.ctypes.data, n,
                            self.cutoff, keep.ctypes.data, C.byref(n_kept)), name)
        return keep[:n], int(nk)

    def filter_sequence(self, sequence):
        """std::tuple<bool, uint64_t> filter_sequence(const std::string&), diginorm.hh:111-119."""
        from .dbg import SequenceLengthException
        if len(sequence) < self.K:
            raise SequenceLengthException("Sequence must have length >= K")
        bases, offsets = _capi.reads_from_strings([sequence])
        keep, nk = self.filter_sequences(bases, offsets)
        return bool(keep[0]), len(sequence) - self.K + 1


class StreamingSolidFilter:
    """``StreamingSolidFilter[dBG].Filter`` (include/goetia/solidifier.hh:33-77): every read is inserted; it passes when
    fewer than ``1 - min_prop_solid`` of its k-mers have an insert_and_query count below ``solid_threshold``.  The counts are
    the reference's serial ones (each k-mer sees every earlier k-mer: dbg.hh:327-340), produced for a whole batch by
    gt_insert_and_query_sequences."""

    def __init__(self, graph, min_prop_solid=0.75, solid_threshold=1):
        self.graph, self.K = graph, graph.K
        self.min_prop_solid, self.solid_threshold = float(min_prop_solid), int(solid_threshold)

    @classmethod
    def build(cls, graph, min_prop_solid=0.75, solid_threshold=1):
        return cls(graph, min_prop_solid, solid_threshold)

    def _decide(self, n_not_solid, n_kmers):
        # arithmetic as the reference writes it (solidifier.hh:68): a float quotient against a double difference
        # whose subtrahend is the float member min_prop_solid
        return not (float(np.float32(n_not_solid) / np.float32(n_kmers)) >= 1.0 - float(np.float32(self.min_prop_solid)))

    def filter_sequence(self, sequence):
        counts = self.graph.insert_and_query_sequence(sequence)
        n_kmers = len(sequence) - self.K + 1
        n_not_solid = sum(1 for c in counts if c < self.solid_threshold)
        return self._decide(n_not_solid, n_kmers), n_kmers

    def filter_sequences(self, bases, offsets):
        """One batch, one library call: -> (keep uint8[n_reads], k-mers judged).  Reads shorter than K or holding a
        foreign symbol are skipped, as the processor swallows their exceptions (processors.hh:395-403)."""
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        counts, status = self.graph.insert_and_query_sequences(bases, offsets, want_status=True)
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        nk = np.where(status == 0, np.maximum(lens - self.K + 1, 0), 0)
        ends = np.cumsum(nk)
        not_solid = np.concatenate([[0], np.cumsum(counts < self.solid_threshold)])
        keep = np.zeros(n, dtype=np.uint8)
        for r in np.nonzero(nk > 0)[0]:
            e = int(ends[r])
            keep[r] = 1 if self._decide(int(not_solid[e] - not_solid[e - int(nk[r])]), int(nk[r])) else 0
        return keep, int(nk.sum())


class FilterProcessor:
    """FilterProcessor<Filter>: stream a FASTX file through the filter, write the passing records."""

    def __init__(self, filt, output_filename, batch_reads=100000):
        self.filter, self.output_filename, self.batch_reads = filt, output_filename, int(batch_reads)
        self.n_sequences = self.n_passed = self.time = 0

    @classmethod
    def build(cls, filt, output_filename, batch_reads=100000):
        return cls(filt, output_filename, batch_reads)

    def process(self, filename, strict=False, min_length=0):
        """-> (sequences processed, k-mers judged), as FileProcessor::process returns."""
        parser = FastxParser(filename, strict, min_length)
        with open(self.output_filename, "w") as out:
            batch = []
            while True:
                done = parser.is_complete()
                if not done:
                    rec = parser.next()
                    if rec is not None:
                        batch.append(rec)
                if batch and (done or len(batch) >= self.batch_reads):
                    bases, offsets = _capi.reads_from_strings([r.sequence for r in batch])
                    keep, nk = self.filter.filter_sequences(bases, offsets)
                    for r, k in zip(batch, keep):
                        if k:
                            r.write_fastx(out)
                    self.n_sequences += len(batch)
                    self.n_passed += int(keep.sum())
                    self.time += nk
                    batch = []
                if done:
                    break
        parser.close()
        return self.n_sequences, self.time
