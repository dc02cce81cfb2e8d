"""Device-resident 2-bit packed read batches (the product of the parsing-to-device pipeline)."""
from . import _capi


class PackedBatch:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_host(cls, bases, offsets):
        L = _capi.lib()
        bases, offsets = _capi.as_reads(bases, offsets)
        h = L.gt_batch_pack(bases.ctypes.data, offsets.ctypes.data, offsets.size - 1)
        if not h:
            raise _capi.GoetiaB200Error("gt_batch_pack: " + _capi.last_error())
        return cls(h)

    @classmethod
    def from_device(cls, d_bases_ptr, d_offsets_ptr, n_reads, n_bases):
        """ASCII + offsets already in HBM (e.g. torch tensors' data_ptr())."""
        h = _capi.lib().gt_batch_pack_dev(d_bases_ptr, d_offsets_ptr, n_reads, n_bases)
        if not h:
            raise _capi.GoetiaB200Error("gt_batch_pack_dev: " + _capi.last_error())
        return cls(h)

    @property
    def handle(self):
        return self._h

    def n_reads(self):
        return int(_capi.lib().gt_batch_n_reads(self._h))

    def n_bases(self):
        return int(_capi.lib().gt_batch_n_bases(self._h))

    def n_kmers(self, K):
        return int(_capi.check(_capi.lib().gt_batch_n_kmers(self._h, K), "gt_batch_n_kmers"))

    def insert_into(self, graph, mode=None, stream=None):
        """Queue the fused hash+insert kernel for this batch; asynchronous."""
        mode = graph.mode if mode is None else mode
        return int(_capi.check(_capi.lib().gt_insert_batch(graph.S.handle, graph.hasher.shifter_kind, graph.K,
                                                           self._h, mode, stream), "gt_insert_batch"))

    def close(self):
        if self._h:
            _capi.load().gt_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_reads_host(bases, offsets, n_threads=0, words_out=None, flags_out=None):
    """Validate + fold case + 2-bit pack a read batch on the host (gt_pack_reads_host): -> (words uint64, flags uint8).
    The layout is the device's: base p of the batch at bits 2*(p%32) of word p/32, A=0 C=1 G=2 T=3; flags[r] = READ_INVALID
    for a read holding a byte outside ACGTacgt.  No GPU is needed."""
    import numpy as np
    bases, offsets = _capi.as_reads(bases, offsets)
    n = offsets.size - 1
    n_bases = int(offsets[-1] - offsets[0]) if n else 0
    words = np.zeros((n_bases + 31) // 32 + 1, dtype=np.uint64) if words_out is None else words_out
    flags = np.zeros(max(n, 1), dtype=np.uint8) if flags_out is None else flags_out
    _capi.check(_capi.load().gt_pack_reads_host(bases.ctypes.data, offsets.ctypes.data, n, words.ctypes.data, flags.ctypes.data,
                                                int(n_threads)), "gt_pack_reads_host")
    return words, flags
