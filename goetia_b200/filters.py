"""Read filters on the GPU dBG: DiginormFilter (include/goetia/diginorm.hh:25-125) and its
FilterProcessor (include/goetia/processors.hh:347-430).

The reference judges and inserts one read at a time.  Here a whole batch is judged against the table
state at the start of the batch and the passing reads are then inserted (the *batch-synchronous* rule,
SURVEY.md section 8a); ``batch_reads=1`` reproduces the reference exactly; the parity tests run the CPU
checker with the same batch size.
"""
import ctypes as C

import numpy as np

from . import _capi
from .parsing import FastxParser


class DiginormFilter:
    """``DiginormFilter[dBG].Filter``: ``build(graph, cutoff)``, ``filter_sequence``."""

    def __init__(self, graph, cutoff):
        self.graph, self.K, self.cutoff = graph, graph.K, int(cutoff)

    @classmethod
    def build(cls, graph, cutoff):
        return cls(graph, cutoff)

    @staticmethod
    def median_count_at_least(sequence, cutoff, graph):
        """diginorm.hh:35-68"""
        bases, offsets = _capi.reads_from_strings([sequence])
        return bool(graph.median_count_at_least(bases, offsets, cutoff)[0])

    def filter_sequences(self, bases, offsets):
        """One batch: (keep uint8[n_reads], k-mers judged).  Kept reads have been inserted."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        keep = np.zeros(max(n, 1), dtype=np.uint8)
        n_kept = C.c_uint64(0)
        nk = _capi.check(L.gt_diginorm_sequences(self.graph.S.handle, self.graph.hasher.shifter_kind, self.K,
                                                 bases.ctypes.data, offsets.ctypes.data, n, self.cutoff,
                                                 keep.ctypes.data, C.byref(n_kept)), "gt_diginorm_sequences")
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
    """``StreamingSolidFilter[dBG].Filter`` (include/goetia/solidifier.hh:33-77): a read passes when fewer than
    ``1 - min_prop_solid`` of its k-mers have an insert_and_query count below ``solid_threshold``.  The counts are
    the reference's serial ones (each k-mer sees the earlier k-mers of the read: dbg.hh:327-340), so this filter
    runs on the per-k-mer latency path; it is here for API completeness, not throughput."""

    def __init__(self, graph, min_prop_solid=0.75, solid_threshold=1):
        self.graph, self.K = graph, graph.K
        self.min_prop_solid, self.solid_threshold = float(min_prop_solid), int(solid_threshold)

    @classmethod
    def build(cls, graph, min_prop_solid=0.75, solid_threshold=1):
        return cls(graph, min_prop_solid, solid_threshold)

    def filter_sequence(self, sequence):
        counts = self.graph.insert_and_query_sequence(sequence)
        n_kmers = len(sequence) - self.K + 1
        n_not_solid = sum(1 for c in counts if c < self.solid_threshold)
        # arithmetic as the reference writes it (solidifier.hh:68): a float quotient against a double difference
        # whose subtrahend is the float member min_prop_solid
        if float(np.float32(n_not_solid) / np.float32(n_kmers)) >= 1.0 - float(np.float32(self.min_prop_solid)):
            return False, n_kmers
        return True, n_kmers

    def filter_sequences(self, bases, offsets):
        """Reads one after the other (the serial semantics do not batch); -> (keep, k-mers judged)."""
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        keep = np.zeros(n, dtype=np.uint8)
        judged = 0
        for r in range(n):
            seq = bases[int(offsets[r]):int(offsets[r + 1])].tobytes().decode("ascii")
            if len(seq) < self.K:
                continue  # SequenceLengthException, swallowed by the processor (processors.hh:395-403)
            try:
                ok, nk = self.filter_sequence(seq)
            except ValueError:
                continue  # InvalidCharacterException, swallowed likewise
            keep[r] = 1 if ok else 0
            judged += nk
        return keep, judged


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
