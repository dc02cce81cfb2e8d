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
