"""SourmashSketch mirror (include/goetia/sketches/sourmash_sketch.hh:22-85) over the GPU sketch kernel.

``SourmashSketch.Sketch(n, K, is_protein, dayhoff, hp, seed, scaled)`` has the reference's constructor
and the ``sourmash::MinHash`` members goetia uses (sketches/sourmash/sourmash.hpp:67-166): ``add_hash``,
``add_sequence``, ``merge``, ``count_common``, ``size``, ``mins``, ``num``, ``seed``, ``ksize``,
``max_hash``.  Only the DNA path exists (is_protein / dayhoff / hp must be False -- goetia's CLI never
sets them: goetia/cli/signature_runner.py).  All arithmetic runs in k_sketch through the C ABI; there
is no CPU fallback.
"""
import numpy as np

from . import _capi


class Sketch:
    """SourmashSketch::Sketch (sourmash_sketch.hh:24-82)."""

    def __init__(self, n, K, is_protein=False, dayhoff=False, hp=False, seed=42, scaled=0):
        if is_protein or dayhoff or hp:
            raise NotImplementedError("only the DNA sketch is on the GPU path")
        L = _capi.lib()
        self.K = int(K)
        self._num, self._seed = int(n), int(seed)
        self._max_hash = self.max_hash_from_scaled(int(scaled))
        self._h = L.gt_sketch_create(self._num, self.K, self._seed, self._max_hash)
        if not self._h:
            raise _capi.GoetiaB200Error("gt_sketch_create: " + _capi.last_error())

    @classmethod
    def build(cls, n, K, is_protein=False, dayhoff=False, hp=False, seed=42, scaled=0):
        return cls(n, K, is_protein, dayhoff, hp, seed, scaled)

    # -- sourmash_sketch.hh:52-69 ----------------------------------------------------------------
    @staticmethod
    def max_hash_from_scaled(scaled):
        return int(_capi.load().gt_max_hash_from_scaled(int(scaled)))

    @staticmethod
    def scaled_from_max_hash(max_hash):
        return 0 if max_hash == 0 else (2**64 - 1) // int(max_hash)

    # -- MinHash accessors (sourmash.hpp:116-157) ---------------------------------------------------
    @property
    def handle(self):
        return self._h

    def num(self):
        return self._num

    def seed(self):
        return self._seed

    def ksize(self):
        return self.K

    def max_hash(self):
        return self._max_hash

    def is_protein(self):
        return False

    def dayhoff(self):
        return False

    def hp(self):
        return False

    def track_abundance(self):
        return False

    def size(self):
        return int(_capi.check(_capi.lib().gt_sketch_size(self._h), "gt_sketch_size"))

    __len__ = size

    def mins(self):
        """Ascending, duplicate-free uint64 array (sourmash.hpp:151-157)."""
        L = _capi.lib()
        n = self.size()
        out = np.zeros(max(n, 1), dtype=np.uint64)
        got = _capi.check(L.gt_sketch_mins(self._h, out.ctypes.data, out.size), "gt_sketch_mins")
        return out[:got]

    get_mins = mins

    # -- updates ----------------------------------------------------------------------------------
    def add_hash(self, h):
        self.add_hashes([int(h)])

    def add_hashes(self, hashes):
        hs = np.ascontiguousarray(hashes, dtype=np.uint64)
        _capi.check(_capi.lib().gt_sketch_add_hashes(self._h, hs.ctypes.data, hs.size), "gt_sketch_add_hashes")

    def add_sequence(self, sequence, force=True):
        """kmerminhash_add_sequence (sourmash.hpp:92).  force=False is the strict form: a sequence
        holding a non-ACGT byte raises instead of having those windows skipped."""
        if isinstance(sequence, str):
            sequence = sequence.encode("ascii")
        if not force and any(c not in b"ACGTacgt" for c in sequence):
            raise ValueError("sourmash.rs: invalid DNA given to kmerminhash_add_sequence")
        bases, offsets = _capi.reads_from_strings([sequence])
        self.insert_sequences(bases, offsets)

    def insert_sequence(self, sequence):
        """sourmash_sketch.hh:71-81: add_sequence(seq, force=true); returns len - K + 1."""
        self.add_sequence(sequence, True)
        return len(sequence) - self.K + 1

    def insert_sequences(self, bases, offsets):
        """A whole read batch (concatenated bytes + offsets) in one call; returns the sum of len-K+1
        over reads with len >= K (what InserterProcessor<Sketch> accumulates, processors.hh:304-331)."""
        bases, offsets = _capi.as_reads(bases, offsets)
        return int(_capi.check(_capi.lib().gt_sketch_add_sequences(self._h, bases.ctypes.data, offsets.ctypes.data,
                                                                   offsets.size - 1), "gt_sketch_add_sequences"))

    def insert_sequences_dev(self, d_bases_ptr, d_offsets_ptr, n_reads, n_bases):
        """Same for ASCII bases + uint64 offsets already resident in HBM (raw device pointers)."""
        return int(_capi.check(_capi.lib().gt_sketch_add_sequences_dev(self._h, d_bases_ptr, d_offsets_ptr, n_reads,
                                                                       n_bases), "gt_sketch_add_sequences_dev"))

    def merge(self, other):
        _capi.check(_capi.lib().gt_sketch_merge(self._h, other._h), "gt_sketch_merge")

    def count_common(self, other, downsample=False):
        return int(_capi.check(_capi.lib().gt_sketch_count_common(self._h, other._h), "gt_sketch_count_common"))

    def jaccard(self, other):
        """|A n B| / |A u B| of the two hash sets (what signature_runner.py:131-157 tracks between
        snapshots via sourmash's MinHash.similarity for scaled sketches)."""
        common = self.count_common(other)
        union = self.size() + other.size() - common
        return common / union if union else 0.0

    similarity = jaccard

    def reset(self):
        _capi.check(_capi.lib().gt_sketch_reset(self._h), "gt_sketch_reset")

    def copy(self):
        c = Sketch.__new__(Sketch)
        c.K, c._num, c._seed, c._max_hash = self.K, self._num, self._seed, self._max_hash
        c._h = _capi.lib().gt_sketch_create(self._num, self.K, self._seed, self._max_hash)
        if not c._h:
            raise _capi.GoetiaB200Error("gt_sketch_create: " + _capi.last_error())
        c.merge(self)
        return c

    # -- streaming: what `goetia sketch sourmash` does around the sketch (goetia/cli/signature_runner.py:131-157:
    # a snapshot every `interval` k-mers and the distance between consecutive snapshots) -----------------------
    def stream_fastx(self, parser_or_filename, interval=1_000_000, strict=False, min_length=0, batch_bases=64 << 20):
        """Generator over the intervals of a FASTX stream: yields dicts ``{t, sequences, size, similarity,
        distance}`` -- t = k-mers consumed so far (the processors' clock), size = hashes in the sketch,
        similarity = Jaccard of this snapshot with the previous one (None for the first), distance = 1 - it.
        Whole intervals are sketched in one device call each; the last, partial interval is reported too."""
        from .parsing import FastxParser
        own = not isinstance(parser_or_filename, FastxParser)
        parser = FastxParser(parser_or_filename, strict, min_length) if own else parser_or_filename
        prev, t, n_seqs, since = None, 0, 0, 0
        try:
            while True:
                # a batch holds about one interval's worth of k-mers (reads carry len - K + 1 each)
                bases, offsets = parser.next_batch(batch_bases, max_reads=max(1, interval // 64))
                done = offsets.size == 1
                if not done:
                    # cut the batch where the interval fills up
                    kmers = np.maximum((offsets[1:] - offsets[:-1]).astype(np.int64) - self.K + 1, 0)
                    cum = np.cumsum(kmers)
                    r0 = 0
                    while r0 < kmers.size:
                        room = interval - since
                        base = cum[r0 - 1] if r0 else 0
                        r1 = int(np.searchsorted(cum, base + room, side="left")) + 1
                        r1 = min(max(r1, r0 + 1), kmers.size)
                        b = bases[int(offsets[r0]):int(offsets[r1])]
                        o = offsets[r0:r1 + 1] - offsets[r0]
                        nk = self.insert_sequences(b, o)
                        t += nk
                        since += nk
                        n_seqs += r1 - r0
                        r0 = r1
                        if since >= interval:
                            snap = self.copy()
                            sim = snap.jaccard(prev) if prev is not None else None
                            yield {"t": t, "sequences": n_seqs, "size": snap.size(), "similarity": sim,
                                   "distance": None if sim is None else 1.0 - sim}
                            if prev is not None:
                                prev.close()
                            prev, since = snap, 0
                if done:
                    if since:
                        sim = self.jaccard(prev) if prev is not None else None
                        yield {"t": t, "sequences": n_seqs, "size": self.size(), "similarity": sim,
                               "distance": None if sim is None else 1.0 - sim}
                    break
        finally:
            if prev is not None:
                prev.close()
            if own:
                parser.close()

    # -- multi-GPU: every rank sketches its shard of the reads; the sketch of the whole set is the
    # union ----------------------------------------------------------------
    def allgather_merge(self, group=None):
        """Union this rank's hash set with every other rank's (torch.distributed; NCCL or gloo).
        Exact because a sketch is a set: afterwards every rank holds the sketch of all the reads."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group)
        if world == 1:
            return
        backend = dist.get_backend(group)
        me = dist.get_rank(group)
        if backend == "nccl":
            # the sets never leave the devices: count, export, all-gather, add
            L = _capi.lib()
            dev = torch.device("cuda", torch.cuda.current_device())
            n_mine = _capi.check(L.gt_sketch_live_count(self._h), "gt_sketch_live_count")
            ns = torch.zeros(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(ns, torch.tensor([n_mine], dtype=torch.int64, device=dev), group=group)
            ns = [int(x) for x in ns.cpu()]
            cap = max(1, max(ns))
            buf = torch.zeros(cap, dtype=torch.int64, device=dev)
            _capi.check(L.gt_sketch_export_dev(self._h, buf.data_ptr(), cap), "gt_sketch_export_dev")
            bufs = torch.empty(world * cap, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(bufs, buf, group=group)
            torch.cuda.synchronize()
            for q in range(world):
                if q != me and ns[q]:
                    _capi.check(L.gt_sketch_add_hashes_dev(self._h, bufs.data_ptr() + 8 * q * cap, ns[q]), "gt_sketch_add_hashes_dev")
            return
        mine = self.mins()
        dev = torch.device("cpu")
        n = torch.tensor([mine.size], dtype=torch.int64, device=dev)
        ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        cap = max(1, max(int(x.item()) for x in ns))
        buf = torch.zeros(cap, dtype=torch.int64, device=dev)
        buf[:mine.size] = torch.from_numpy(mine.view(np.int64)).to(dev)
        bufs = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(bufs, buf, group=group)
        for q in range(world):
            if q != me and int(ns[q].item()):
                self.add_hashes(bufs[q][:int(ns[q].item())].numpy().view(np.uint64))

    def close(self):
        if getattr(self, "_h", None):
            _capi.load().gt_sketch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SourmashSketch:
    """Namespace struct of the reference (sourmash_sketch.hh:22): ``SourmashSketch.Sketch``."""
    Sketch = Sketch
