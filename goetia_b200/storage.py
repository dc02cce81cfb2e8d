"""StorageType mirror: BitStorage / ByteStorage / NibbleStorage with tables resident in HBM.

Same member names and meaning as the reference classes (include/goetia/storage/bitstorage.hh:93-243,
bytestorage.hh:96-290, nibblestorage.hh:89-232) and their pythonizations
(goetia/pythonizors/pythonize_storage.py:12-46); every compute member goes through the C ABI.
"""
import gzip
import struct

import numpy as np

from . import _capi

count_t = np.int16  # storage/storage.hh:80


class GoetiaException(_capi.GoetiaB200Error):
    pass


class GoetiaFileException(GoetiaException):
    pass


def get_n_primes_near_x(n, x):
    """storage/storage.hh:166-190 (host arithmetic inside the library; no GPU needed)."""
    out = np.zeros(max(int(n), 1), dtype=np.uint64)
    k = _capi.check(_capi.load().gt_primes_near(int(n), int(x), out.ctypes.data_as(_capi.u64p)), "gt_primes_near")
    return [int(v) for v in out[:k]]


def _hash_value(h):
    """Accept raw ints and hash_type objects (Hash / Canonical -> value())."""
    v = getattr(h, "value", None)
    return int(v()) if callable(v) else int(h)


class _Storage:
    kind = None
    NAME = None
    is_probabilistic = True
    is_counting = False
    bits_per_slot = 1
    params_type = tuple
    default_params = (1000000, 4)  # StorageTraits<...>::default_params, bitstorage.hh:75
    _saved_type = None
    _max_count = 1

    def __init__(self, max_table_or_sizes, N=None):
        if N is None:
            sizes = [int(s) for s in max_table_or_sizes]
        else:
            sizes = get_n_primes_near_x(int(N), int(max_table_or_sizes))
        self._sizes = np.asarray(sizes, dtype=np.uint64)
        L = _capi.lib()
        self._h = L.gt_storage_create(self.kind, self._sizes.ctypes.data_as(_capi.u64p), len(sizes))
        if not self._h:
            raise GoetiaException("gt_storage_create: " + _capi.last_error())
        self.build_params = (max_table_or_sizes, N)

    # -- construction ----------------------------------------------------------------------
    @staticmethod
    def make_params(*args):
        return tuple(args)

    @classmethod
    def build(cls, *args):
        if len(args) == 0:
            args = cls.default_params
        elif len(args) == 1 and isinstance(args[0], tuple):
            args = args[0]
        st = cls(*args)
        st.build_params = args
        return st

    def clone(self):
        """Same table sizes, empty (bitstorage.hh:136-138)."""
        return type(self)([int(s) for s in self._sizes])

    def close(self):
        if getattr(self, "_h", None):
            _capi.load().gt_storage_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection -----------------------------------------------------------------------
    @property
    def handle(self):
        return self._h

    def get_tablesizes(self):
        return [int(s) for s in self._sizes]

    def n_tables(self):
        return len(self._sizes)

    def _stats(self):
        a = np.zeros(2, dtype=np.uint64)
        _capi.check(_capi.lib().gt_storage_stats(self._h, a[0:].ctypes.data_as(_capi.u64p),
                                                 a[1:].ctypes.data_as(_capi.u64p)), "gt_storage_stats")
        return int(a[0]), int(a[1])

    def n_unique_kmers(self):
        return self._stats()[0]

    def n_occupied(self):
        return self._stats()[1]

    def estimated_fp(self):
        """bitstorage.hh:183-188: (occupied / size0) ** n_tables."""
        return (float(self.n_occupied()) / float(self._sizes[0])) ** self.n_tables()

    def table_bytes(self, i):
        return int(_capi.lib().gt_storage_table_bytes(self._h, i))

    def get_raw_tables(self):
        """Host copies of the tables, byte-identical to the reference's (storage.hh:126)."""
        L = _capi.lib()
        out = []
        for i in range(self.n_tables()):
            buf = np.empty(self.table_bytes(i), dtype=np.uint8)
            _capi.check(L.gt_storage_download_table(self._h, i, buf.ctypes.data), "gt_storage_download_table")
            out.append(buf)
        return out

    def checksum(self, i):
        """Position-weighted checksum of table i computed on the device (gt_storage_checksum)."""
        import ctypes as C
        out = C.c_uint64(0)
        _capi.check(_capi.lib().gt_storage_checksum(self._h, int(i), C.byref(out)), "gt_storage_checksum")
        return int(out.value)

    def reset(self):
        _capi.check(_capi.lib().gt_storage_reset(self._h), "gt_storage_reset")

    def flush(self):
        """Apply every pending write-combined blind insert (include/goetia_b200.h, gt_storage_flush)."""
        _capi.check(_capi.lib().gt_storage_flush(self._h), "gt_storage_flush")

    def pending_info(self):
        a = np.zeros(8, dtype=np.uint64)
        _capi.check(_capi.lib().gt_storage_pending_info(self._h, a.ctypes.data), "gt_storage_pending_info")
        keys = ("built", "n_buckets", "slice_shift", "budget_kmers", "entries", "pending_kmers", "n_direct",
                "apply_grid")
        return dict(zip(keys, (int(v) for v in a)))

    # -- single-hash members (one launch each; the batch members are the fast path) ---------
    def insert(self, khash, mode=_capi.MODE_EXACT):
        return bool(self.insert_many([_hash_value(khash)], mode=mode)[0])

    def query(self, khash):
        return int(self.query_many([_hash_value(khash)])[0])

    def insert_and_query(self, khash):
        """Bit: always 1 (bitstorage.cc:78-84); counting: 1 if new else query (bytestorage.cc:142-150)."""
        new = self.insert(khash)
        if self.kind == _capi.STORAGE_BIT or new:
            return 1
        return self.query(khash)

    # -- batch members -----------------------------------------------------------------------
    def insert_many(self, hashes, mode=_capi.MODE_EXACT, want_new=True):
        hs = np.ascontiguousarray(hashes, dtype=np.uint64)
        want_new = want_new and mode != _capi.MODE_BLIND
        is_new = np.zeros(max(hs.size, 1), dtype=np.uint8) if want_new else None
        _capi.check(_capi.lib().gt_insert_hashes(self._h, hs.ctypes.data, hs.size, mode,
                                                 is_new.ctypes.data if want_new else None), "gt_insert_hashes")
        return is_new[:hs.size] if want_new else None

    def query_many(self, hashes):
        hs = np.ascontiguousarray(hashes, dtype=np.uint64)
        counts = np.zeros(max(hs.size, 1), dtype=np.int16)
        _capi.check(_capi.lib().gt_query_hashes(self._h, hs.ctypes.data, hs.size, counts.ctypes.data),
                    "gt_query_hashes")
        return counts[:hs.size]

    # -- OXLI v4 files (bitstorage.cc:151-307, bytestorage.cc:153-667, nibblestorage.cc:132-278) ----
    def save(self, filename, ksize):
        tables = self.get_raw_tables()
        opener = gzip.open if (self.kind == _capi.STORAGE_BYTE and str(filename).endswith(".gz")) else open
        with opener(filename, "wb") as f:
            f.write(b"OXLI")
            f.write(struct.pack("<BB", 4, self._saved_type))
            if self.kind == _capi.STORAGE_BYTE:
                f.write(struct.pack("<B", 0))  # use_bigcount (off: storage.hh:110)
            f.write(struct.pack("<IBQ", int(ksize), self.n_tables(), self.n_occupied()))
            for size, t in zip(self._sizes, tables):
                f.write(struct.pack("<Q", int(size)))
                f.write(t.tobytes())
            if self.kind == _capi.STORAGE_BYTE:
                f.write(struct.pack("<Q", 0))  # n_bigcounts

    def load(self, filename):
        """Replace this storage's contents from an OXLI file; returns the saved ksize."""
        opener = gzip.open if (self.kind == _capi.STORAGE_BYTE and str(filename).endswith(".gz")) else open
        try:
            with opener(filename, "rb") as f:
                data = f.read()
        except OSError as e:
            raise GoetiaFileException(str(e))
        if len(data) < 6 or data[:4] != b"OXLI":
            raise GoetiaFileException("Does not start with signature for a oxli binary file")
        version, ht_type = data[4], data[5]
        if version != 4:
            raise GoetiaFileException("Incorrect file format version %d" % version)
        if ht_type != self._saved_type:
            raise GoetiaFileException("Incorrect file format type %d" % ht_type)
        pos = 6
        if self.kind == _capi.STORAGE_BYTE:
            pos += 1
        ksize, n_tables, occupied = struct.unpack_from("<IBQ", data, pos)
        pos += 13
        sizes, tables = [], []
        for _ in range(n_tables):
            (size,) = struct.unpack_from("<Q", data, pos)
            pos += 8
            nb = size // 8 + 1 if self.kind == 0 else size if self.kind == 1 else size // 2 + 1
            if pos + nb > len(data):
                raise GoetiaFileException("truncated table")
            sizes.append(size)
            tables.append(np.frombuffer(data, dtype=np.uint8, count=nb, offset=pos))
            pos += nb
        L = _capi.lib()
        if sizes != self.get_tablesizes():
            # build the replacement first and swap only on success: a failed create must not leave a freed handle behind
            new_sizes = np.asarray(sizes, dtype=np.uint64)
            h = L.gt_storage_create(self.kind, new_sizes.ctypes.data_as(_capi.u64p), len(sizes))
            if not h:
                raise GoetiaException("gt_storage_create: " + _capi.last_error())
            old, self._h, self._sizes = self._h, h, new_sizes
            L.gt_storage_destroy(old)
        for i, t in enumerate(tables):
            t = np.ascontiguousarray(t)
            _capi.check(L.gt_storage_upload_table(self._h, i, t.ctypes.data), "gt_storage_upload_table")
        # n_unique_kmers is not part of the file; the reference leaves it at 0
        _capi.check(L.gt_storage_set_n_unique(self._h, 0), "gt_storage_set_n_unique")
        return ksize

    def describe(self):
        return "%s\n- Params: %s\n- Probabilistic: %s\n- Counting:      %s" % (
            type(self).__name__, self.build_params, self.is_probabilistic, self.is_counting)


class BitStorage(_Storage):
    """Bloom-style nodegraph (bitstorage.hh:93-243)."""
    kind = _capi.STORAGE_BIT
    NAME = "BitStorage"
    is_counting = False
    bits_per_slot = 1
    _saved_type = 2  # SAVED_HASHBITS, storage.hh:66
    _max_count = 1

    def update_from(self, other):
        """bitstorage.cc:103-137"""
        if self.get_tablesizes() != other.get_tablesizes():
            raise GoetiaException("both nodegraphs must have same table sizes")
        _capi.check(_capi.lib().gt_storage_update_from(self._h, other._h), "gt_storage_update_from")


class ByteStorage(_Storage):
    """8-bit count-min (bytestorage.hh:96-290)."""
    kind = _capi.STORAGE_BYTE
    NAME = "ByteStorage"
    is_counting = True
    bits_per_slot = 8
    _saved_type = 1  # SAVED_COUNTING_HT
    _max_count = 255


class NibbleStorage(_Storage):
    """4-bit count-min (nibblestorage.hh:89-232)."""
    kind = _capi.STORAGE_NIBBLE
    NAME = "NibbleStorage"
    is_counting = True
    bits_per_slot = 4
    _saved_type = 7  # SAVED_SMALLCOUNT
    _max_count = 15


types = [BitStorage, ByteStorage, NibbleStorage]
typenames = [(t, t.NAME) for t in types]
