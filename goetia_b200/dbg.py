"""dBG<StorageType, ShifterType> mirror (include/goetia/dbg.hh:39-438) over the GPU backend.

``dBG[BitStorage, CanLemireShifter].build(storage, hasher)`` reads like the cppyy surface
(goetia/dbg.py:43-44, tests/utils.py:88-93).  Sequence members run the fused hash+insert /
hash+query kernels; the batch members (``insert_sequences`` / ``query_sequences`` /
``median_count_at_least``) take whole read batches and are the path the processors use.
"""
import numpy as np

from . import _capi
from .hashing import Canonical, Hash, InvalidSequenceException
from .storage import _hash_value

MODE_BLIND, MODE_FAST, MODE_EXACT = _capi.MODE_BLIND, _capi.MODE_FAST, _capi.MODE_EXACT


class SequenceLengthException(InvalidSequenceException):
    """kmeriterator.hh:57-59"""


class InvalidCharacterException(ValueError):
    """sequences/exceptions.hh:16-36"""


class dBG:
    storage_type = None
    shifter_type = None
    _specialisations = {}

    def __class_getitem__(cls, params):
        storage_t, shifter_t = params
        key = (storage_t, shifter_t)
        if key not in cls._specialisations:
            name = "dBG<%s,%s>" % (storage_t.NAME, shifter_t.NAME)
            cls._specialisations[key] = type(name, (dBG,), {"storage_type": storage_t, "shifter_type": shifter_t})
        return cls._specialisations[key]

    def __init__(self, storage, hasher_or_K, mode=MODE_EXACT):
        self.S = storage
        if isinstance(hasher_or_K, int):
            if self.shifter_type is None:
                raise TypeError("use dBG[Storage, Shifter].build(storage, K)")
            self.hasher = self.shifter_type(hasher_or_K)
        else:
            self.hasher = type(hasher_or_K)(hasher_or_K.K)
        if self.storage_type is None:
            self.storage_type = type(storage)
        if self.shifter_type is None:
            self.shifter_type = type(self.hasher)
        self.K = self.hasher.K
        self.mode = mode
        self.hash_type = self.shifter_type.hash_type

    @classmethod
    def build(cls, storage, hasher_or_K, *args):
        return cls(storage, hasher_or_K)

    # -- clones (dbg.hh:97-101, :128-131) ------------------------------------------------------
    def clone(self):
        return type(self)(self.S.clone(), self.hasher, self.mode)

    shallow_clone = clone  # pythonize_dbg.py:21-22

    def reference_copy(self):
        return type(self)(self.S, self.hasher, self.mode)

    # -- shifter side ------------------------------------------------------------------------
    def hash(self, kmer):
        return self.hasher.hash(kmer)

    def hashes(self, sequence):
        return iter(self.hasher.hashes(sequence))

    def get_hash_iter(self, sequence):
        return iter(self.hasher.hashes(sequence))

    def get_hasher(self):
        return type(self.hasher)(self.K)

    # the graph IS-A shifter in the reference (dbg.hh:39-41 derives from ShifterType): the cursor members of
    # hashshifter.hh:74-199 are reachable on it, and pythonize_dbg.py:50 exposes the cursor's hash as get_hash()
    def hash_base(self, kmer):
        return self.hasher.hash_base(kmer)

    set_cursor = hash_base

    def shift_right(self, out, inc):
        return self.hasher.shift_right(out, inc)

    def shift_left(self, inc, out):
        return self.hasher.shift_left(inc, out)

    def get_hash(self):
        return self.hasher.get()

    def is_initialized(self):
        return self.hasher.is_initialized()

    def _value(self, item):
        if isinstance(item, (str, bytes)):
            return self.hash(item).value()
        return _hash_value(item)

    # -- single k-mer members (dbg.hh:140-177) -------------------------------------------------
    def insert(self, kmer):
        return self.S.insert(self._value(kmer))

    def insert_and_query(self, kmer):
        return self.S.insert_and_query(self._value(kmer))

    def query(self, kmer):
        return self.S.query(self._value(kmer))

    get = query  # pythonize_dbg.py:53

    def add(self, item):
        """pythonize_dbg.py:8-14"""
        if not isinstance(item, int) and not isinstance(item, (Hash, Canonical)) and len(item) < self.K:
            raise ValueError()
        if isinstance(item, (int, Hash, Canonical)) or len(item) == self.K:
            return self.insert(item)
        return self.insert_sequence(item)

    # -- stats ---------------------------------------------------------------------------------
    def n_unique(self):
        return self.S.n_unique_kmers()

    def n_occupied(self):
        return self.S.n_occupied()

    def estimated_fp(self):
        return self.S.estimated_fp()

    def get_raw(self):
        return self.S.get_raw_tables()

    def suffix(self, kmer):
        return kmer[len(kmer) - self.K + 1:]

    def prefix(self, kmer):
        return kmer[:self.K - 1]

    def save(self, filename):
        self.S.save(filename, self.K)

    def load(self, filename):
        self.S.load(filename)

    def reset(self):
        self.S.reset()

    # -- batch members (the GPU hot path) --------------------------------------------------------
    def insert_sequences(self, bases, offsets, mode=None, want_n_new=False, want_status=False):
        """dBG::insert_sequence over a read batch.  Returns k-mers consumed (+ n_new, status)."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        mode = self.mode if mode is None else mode
        n_new = np.zeros(max(n, 1), dtype=np.uint64) if want_n_new else None
        status = np.zeros(max(n, 1), dtype=np.uint8) if want_status else None
        tot = _capi.check(L.gt_insert_sequences(self.S.handle, self.hasher.shifter_kind, self.K, bases.ctypes.data,
                                                offsets.ctypes.data, n, mode,
                                                n_new.ctypes.data if want_n_new else None,
                                                status.ctypes.data if want_status else None), "gt_insert_sequences")
        out = [int(tot)]
        if want_n_new:
            out.append(n_new[:n])
        if want_status:
            out.append(status[:n])
        return out[0] if len(out) == 1 else tuple(out)

    def insert_sequences_dev(self, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, mode=None):
        """Same as insert_sequences for ASCII bases + uint64 offsets already resident in HBM
        (raw device pointers, e.g. torch tensors' data_ptr())."""
        mode = self.mode if mode is None else mode
        return int(_capi.check(_capi.lib().gt_insert_sequences_dev(self.S.handle, self.hasher.shifter_kind, self.K,
                                                                   d_bases_ptr, d_offsets_ptr, n_reads, n_bases,
                                                                   mode), "gt_insert_sequences_dev"))

    def insert_sequences_dev_async(self, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, mode=None, d_kmer_total_ptr=None,
                                   n_kmers_upper=None):
        """insert_sequences_dev without the host wait: queued on the library's compute stream; the k-mers
        consumed are added to the device uint64 at d_kmer_total_ptr (optional).  n_kmers_upper: the caller's bound on the
        batch's k-mers (equal-length reads know it exactly); without it the library assumes one k-mer per base when it
        decides how many batches fit a pending store before the store has to be applied."""
        mode = self.mode if mode is None else mode
        if n_kmers_upper is not None:
            _capi.check(_capi.lib().gt_storage_hint_kmers(self.S.handle, int(n_kmers_upper)), "gt_storage_hint_kmers")
        _capi.check(_capi.lib().gt_insert_sequences_dev_async(self.S.handle, self.hasher.shifter_kind, self.K, d_bases_ptr,
                                                              d_offsets_ptr, n_reads, n_bases, mode, d_kmer_total_ptr),
                    "gt_insert_sequences_dev_async")

    def insert_sequences_packed(self, words, offsets, flags, mode=MODE_BLIND):
        """dBG::insert_sequence over a batch the host has already validated and 2-bit packed (goetia_b200.batch.pack_reads_host
        or the FASTX front end): only 0.25 B/base cross PCIe.  Returns the k-mers consumed."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        n = offsets.size - 1
        return int(_capi.check(_capi.lib().gt_insert_sequences_packed(self.S.handle, self.hasher.shifter_kind, self.K, words.ctypes.data,
                                                                      offsets.ctypes.data, flags.ctypes.data, n, mode),
                               "gt_insert_sequences_packed"))

    def insert_packed_dev_async(self, d_words_ptr, n_words_alloc, d_offsets_ptr, d_flags_ptr, n_reads, n_bases, mode=MODE_BLIND,
                                d_kmer_total_ptr=None):
        """The same for a packed batch already resident in HBM; queued on the compute stream, no host wait."""
        _capi.check(_capi.lib().gt_insert_packed_dev_async(self.S.handle, self.hasher.shifter_kind, self.K, d_words_ptr, n_words_alloc,
                                                           d_offsets_ptr, d_flags_ptr, n_reads, n_bases, mode, d_kmer_total_ptr),
                    "gt_insert_packed_dev_async")

    def flush(self):
        self.S.flush()

    def process_fastx(self, parser_or_filename, mode=MODE_BLIND, max_reads=0, strict=False, min_length=0):
        """FileProcessor<InserterProcessor<dBG>>::process / advance (processors.hh:112-127, 208-229,
        304-331): stream a FASTX file (or an open FastxParser; ``max_reads`` > 0 stops after that many
        records, like one ``advance`` interval) into the graph.  Returns (sequences processed, k-mers
        consumed); parsing of batch n+1 overlaps the GPU work of batch n (gt_insert_fastx)."""
        import ctypes as C
        from .parsing import FastxParser, _ERRORS
        own = not isinstance(parser_or_filename, FastxParser)
        parser = FastxParser(parser_or_filename, strict, min_length) if own else parser_or_filename
        try:
            n_seqs = C.c_uint64(0)
            nk = _capi.lib().gt_insert_fastx(self.S.handle, self.hasher.shifter_kind, self.K, parser.handle, int(mode),
                                             int(max_reads), C.byref(n_seqs))
            if nk < 0:
                raise _ERRORS.get(int(nk), _capi.GoetiaB200Error)("gt_insert_fastx: " + _capi.last_error())
            if mode == MODE_BLIND:
                self.S.flush()
            return int(n_seqs.value), int(nk)
        finally:
            if own:
                parser.close()

    def query_sequences(self, bases, offsets, want_status=False):
        """dBG::query_sequence over a read batch: counts of all k-mers, reads back to back."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        cap = int(np.maximum(lens - self.K + 1, 0).sum())
        counts = np.zeros(max(cap, 1), dtype=np.int16)
        status = np.zeros(max(n, 1), dtype=np.uint8)
        tot = _capi.check(L.gt_query_sequences(self.S.handle, self.hasher.shifter_kind, self.K, bases.ctypes.data,
                                               offsets.ctypes.data, n, counts.ctypes.data, status.ctypes.data),
                          "gt_query_sequences")
        return (counts[:tot], status[:n]) if want_status else counts[:tot]

    def median_count_at_least(self, bases, offsets, cutoff):
        """DiginormFilter::median_count_at_least per read (diginorm.hh:35-68) -> uint8[n_reads]."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        out = np.zeros(max(n, 1), dtype=np.uint8)
        _capi.check(L.gt_median_count_at_least(self.S.handle, self.hasher.shifter_kind, self.K, bases.ctypes.data,
                                               offsets.ctypes.data, n, int(cutoff), out.ctypes.data, None),
                    "gt_median_count_at_least")
        return out[:n]

    # -- sequence members (dbg.hh:249-394) ---------------------------------------------------------
    def _one(self, sequence):
        if isinstance(sequence, bytes):
            sequence = sequence.decode("ascii")
        if len(sequence) < self.K:
            raise SequenceLengthException("Sequence must have length >= K")
        return _capi.reads_from_strings([sequence])

    def insert_sequence(self, sequence, want_n_new=False):
        """Returns len-K+1 (dbg.hh:296-305); with want_n_new also the new-k-mer count (:307-318)."""
        bases, offsets = self._one(sequence)
        if want_n_new:
            tot, n_new, status = self.insert_sequences(bases, offsets, want_n_new=True, want_status=True)
        else:
            tot, status = self.insert_sequences(bases, offsets, want_status=True)
        if status[0] & _capi.READ_INVALID:
            raise InvalidCharacterException("sequence holds a non-ACGT character")
        return (tot, int(n_new[0])) if want_n_new else tot

    def _hashes_and_values(self, sequence):
        hs = self.hasher.hashes(sequence if isinstance(sequence, str) else sequence.decode("ascii"))
        return hs, np.array([h.value() for h in hs], dtype=np.uint64)

    def insert_sequence_hashes(self, sequence):
        """insert_sequence(sequence, std::vector<hash_type>& hashes), dbg.hh:282-294 -> (len-K+1, hashes)."""
        hs, vals = self._hashes_and_values(sequence)
        self.S.insert_many(vals, mode=MODE_EXACT, want_new=False)
        return len(hs), hs

    def insert_sequence_new_kmers(self, sequence):
        """insert_sequence(sequence, std::set<hash_type>& new_kmers), dbg.hh:267-280 -> (len-K+1, set of the
        hashes whose insert() returned true).  GT_MODE_EXACT gives the reference's serial first-toucher rule."""
        hs, vals = self._hashes_and_values(sequence)
        is_new = self.S.insert_many(vals, mode=MODE_EXACT)
        return len(hs), {h for h, nw in zip(hs, is_new) if nw}

    def insert_sequence_counts(self, sequence):
        """insert_sequence(sequence, hashes, counts), dbg.hh:249-265: counts[j] = insert_and_query(k-mer j),
        each k-mer seeing the earlier ones of the same sequence -> (len-K+1, hashes, counts)."""
        hs, _ = self._hashes_and_values(sequence)
        return len(hs), hs, [self.S.insert_and_query(h.value()) for h in hs]

    def query_sequence_hashes(self, sequence, want_new=False):
        """query_sequence(sequence, counts, hashes[, new_hashes]), dbg.hh:364-394: new_hashes = the hashes whose
        count is 0."""
        hs, vals = self._hashes_and_values(sequence)
        counts = [int(c) for c in self.S.query_many(vals)]
        if want_new:
            return counts, hs, {h for h, c in zip(hs, counts) if c == 0}
        return counts, hs

    def query_sequence(self, sequence):
        bases, offsets = self._one(sequence)
        counts, status = self.query_sequences(bases, offsets, want_status=True)
        if status[0] & _capi.READ_INVALID:
            raise InvalidCharacterException("sequence holds a non-ACGT character")
        return [int(c) for c in counts]

    def insert_and_query_sequences(self, bases, offsets, want_status=False):
        """dBG::insert_and_query_sequence over a read batch with the reference's serial semantics (each k-mer's count is
        the one AFTER its own insert and sees every earlier k-mer of the batch): one call, a few kernel rounds
        (gt_insert_and_query_sequences).  counts are laid out like query_sequences."""
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        n = offsets.size - 1
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        cap = int(np.maximum(lens - self.K + 1, 0).sum())
        counts = np.zeros(max(cap, 1), dtype=np.int16)
        status = np.zeros(max(n, 1), dtype=np.uint8)
        tot = _capi.check(L.gt_insert_and_query_sequences(self.S.handle, self.hasher.shifter_kind, self.K, bases.ctypes.data,
                                                          offsets.ctypes.data, n, counts.ctypes.data, status.ctypes.data),
                          "gt_insert_and_query_sequences")
        return (counts[:tot], status[:n]) if want_status else counts[:tot]

    def insert_and_query_sequence(self, sequence):
        """dbg.hh:327-340: the counts after insertion, each k-mer seeing the earlier k-mers of the same sequence."""
        bases, offsets = self._one(sequence)
        counts, status = self.insert_and_query_sequences(bases, offsets, want_status=True)
        if status[0] & _capi.READ_INVALID:
            raise InvalidCharacterException("sequence holds a non-ACGT character")
        return [int(c) for c in counts]

    def get_kmer_counts(self, sequence):
        return self.query_sequence(sequence)
