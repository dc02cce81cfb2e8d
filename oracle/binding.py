"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Both classes expose the same small surface so that tests can run the same assertions against
the plain-C restatement (``Port``) and, where it was built, the unmodified reference (``Ref``).

storage kinds: 0 = BitStorage, 1 = ByteStorage, 2 = NibbleStorage.  ``can``: 0 = FwdLemireShifter,
1 = CanLemireShifter.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libgoetia_ref.so")
REFERENCE_ROOT = "/root/reference"

u64p = C.POINTER(C.c_uint64)
i16p = C.POINTER(C.c_int16)
u8p = C.POINTER(C.c_uint8)


def build_oracle(force=False):
    """gcc the plain-C restatement (needs nothing but gcc)."""
    src = os.path.join(_HERE, "goetia_oracle.c")
    if force or not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _PORT_SO


def build_ref(force=False):
    """Compile the unmodified reference where it lies (only possible where /root/reference exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return _REF_SO if os.path.exists(_REF_SO) else None
    harness = os.path.join(_HERE, "ref_harness.cc")
    if force or not os.path.exists(_REF_SO) or os.path.getmtime(_REF_SO) < os.path.getmtime(harness):
        subprocess.check_call(["make", "-C", _HERE, "-j8", "ref"], stdout=subprocess.DEVNULL)
    return _REF_SO


def have_ref():
    return os.path.exists(_REF_SO)


def _as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _bytes_of(seq):
    if isinstance(seq, str):
        return seq.encode("ascii")
    if isinstance(seq, np.ndarray):
        return seq.tobytes()
    return bytes(seq)


def synth_reads(n_reads, length, seed):
    """SURVEY.md section 8d synthetic reads: default_rng(seed).integers(0,4,(n,L)) -> "ACGT"[code].

    Returns (bases: uint8[n*L] ASCII, offsets: uint64[n+1]).
    """
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, (n_reads, length), dtype=np.uint8)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(length)
    return np.ascontiguousarray(bases), offsets


class _GraphBase:
    """dBG<Storage, Shifter> as one object; subclasses bind the C entry points."""

    kind = None
    can = None
    K = None
    sizes = None

    def table_nbytes(self, i):
        s = int(self.sizes[i])
        return s // 8 + 1 if self.kind == 0 else s if self.kind == 1 else s // 2 + 1

    def n_kmers(self, seq):
        return max(0, len(seq) - self.K + 1)


# ----------------------------------------------------------------------------- Port (plain C)
class _PortLib:
    _lib = None

    @classmethod
    def get(cls):
        if cls._lib is None:
            build_oracle()
            L = C.CDLL(_PORT_SO)
            L.orc_hash_sequence.restype = C.c_int64
            L.orc_hash_sequence.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint64, u64p, u64p]
            L.orc_primes_near.restype = C.c_int
            L.orc_primes_near.argtypes = [C.c_uint32, C.c_uint64, u64p]
            L.orc_storage_create.restype = C.c_void_p
            L.orc_storage_create.argtypes = [C.c_int, u64p, C.c_int]
            L.orc_storage_destroy.argtypes = [C.c_void_p]
            L.orc_storage_reset.argtypes = [C.c_void_p]
            L.orc_storage_table_bytes.restype = C.c_uint64
            L.orc_storage_table_bytes.argtypes = [C.c_void_p, C.c_int]
            L.orc_storage_table.restype = C.c_void_p
            L.orc_storage_table.argtypes = [C.c_void_p, C.c_int]
            L.orc_storage_stats.argtypes = [C.c_void_p, u64p, u64p]
            L.orc_insert.restype = C.c_int
            L.orc_insert.argtypes = [C.c_void_p, C.c_uint64]
            L.orc_query.restype = C.c_int16
            L.orc_query.argtypes = [C.c_void_p, C.c_uint64]
            L.orc_insert_and_query.restype = C.c_int16
            L.orc_insert_and_query.argtypes = [C.c_void_p, C.c_uint64]
            L.orc_insert_hashes.argtypes = [C.c_void_p, u64p, C.c_uint64, u8p]
            L.orc_query_hashes.argtypes = [C.c_void_p, u64p, C.c_uint64, i16p]
            for fn in ("orc_insert_sequence",):
                getattr(L, fn).restype = C.c_int64
                getattr(L, fn).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_uint64, u64p]
            for fn in ("orc_query_sequence", "orc_insert_and_query_sequence"):
                getattr(L, fn).restype = C.c_int64
                getattr(L, fn).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_uint64, i16p]
            L.orc_insert_reads.restype = C.c_int64
            L.orc_insert_reads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, u64p, C.c_uint64,
                                           u64p, C.POINTER(C.c_double)]
            L.orc_query_reads.restype = C.c_int64
            L.orc_query_reads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, u64p, C.c_uint64, i16p]
            L.orc_median_count_at_least.restype = C.c_int
            L.orc_median_count_at_least.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_uint64,
                                                    C.c_uint]
            L.orc_diginorm_reads.restype = C.c_int64
            L.orc_diginorm_reads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, u64p, C.c_uint64,
                                             C.c_uint, C.c_uint64, u8p]
            L.orc_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, u64p]
            L.orc_max_hash_from_scaled.restype = C.c_uint64
            L.orc_max_hash_from_scaled.argtypes = [C.c_uint64]
            L.orc_sketch_create.restype = C.c_void_p
            L.orc_sketch_create.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64]
            L.orc_sketch_destroy.argtypes = [C.c_void_p]
            L.orc_sketch_size.restype = C.c_uint64
            L.orc_sketch_size.argtypes = [C.c_void_p]
            L.orc_sketch_mins.restype = C.c_void_p
            L.orc_sketch_mins.argtypes = [C.c_void_p]
            L.orc_sketch_add_hash.argtypes = [C.c_void_p, C.c_uint64]
            L.orc_sketch_add_sequence.restype = C.c_int64
            L.orc_sketch_add_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
            L.orc_sketch_add_reads.restype = C.c_int64
            L.orc_sketch_add_reads.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.POINTER(C.c_double)]
            L.orc_fnv1a.restype = C.c_uint64
            L.orc_fnv1a.argtypes = [C.c_void_p, C.c_uint64]
            L.orc_storage_save.restype = C.c_int
            L.orc_storage_save.argtypes = [C.c_void_p, C.c_char_p, C.c_uint16]
            cls._lib = L
        return cls._lib


class Port(_GraphBase):
    """dBG over the plain-C restatement (oracle/goetia_oracle.c)."""

    name = "port"

    def __init__(self, kind, can, K, sizes):
        self.L = _PortLib.get()
        self.kind, self.can, self.K = int(kind), int(can), int(K)
        self.sizes = _as_u64(sizes)
        self.h = self.L.orc_storage_create(self.kind, self.sizes.ctypes.data_as(u64p), len(self.sizes))
        if not self.h:
            raise MemoryError("orc_storage_create failed")

    def close(self):
        if self.h:
            self.L.orc_storage_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- static helpers
    @staticmethod
    def hash_sequence(can, K, seq):
        L = _PortLib.get()
        b = _bytes_of(seq)
        n = max(1, len(b))
        fw = np.zeros(n, dtype=np.uint64)
        rc = np.zeros(n, dtype=np.uint64)
        r = L.orc_hash_sequence(int(can), int(K), b, len(b), fw.ctypes.data_as(u64p), rc.ctypes.data_as(u64p))
        if r < 0:
            raise ValueError("orc_hash_sequence: %d" % r)
        return fw[:r].copy(), rc[:r].copy()

    @staticmethod
    def primes_near(n, x):
        L = _PortLib.get()
        out = np.zeros(max(1, n), dtype=np.uint64)
        k = L.orc_primes_near(int(n), int(x), out.ctypes.data_as(u64p))
        return [int(v) for v in out[:k]]

    @staticmethod
    def murmur3_x64_128(key, seed):
        L = _PortLib.get()
        b = _bytes_of(key)
        out = np.zeros(2, dtype=np.uint64)
        L.orc_murmur3_x64_128(b, len(b), int(seed), out.ctypes.data_as(u64p))
        return int(out[0]), int(out[1])

    @staticmethod
    def max_hash_from_scaled(scaled):
        return int(_PortLib.get().orc_max_hash_from_scaled(int(scaled)))

    @staticmethod
    def fnv1a(arr):
        a = np.ascontiguousarray(arr, dtype=np.uint8)
        return int(_PortLib.get().orc_fnv1a(a.ctypes.data, a.size))

    # -- graph members
    def insert_sequence(self, seq):
        b = _bytes_of(seq)
        nn = C.c_uint64(0)
        n = self.L.orc_insert_sequence(self.h, self.can, self.K, b, len(b), C.byref(nn))
        if n < 0:
            raise ValueError("insert_sequence: %d" % n)
        return int(n), int(nn.value)

    def query_sequence(self, seq):
        b = _bytes_of(seq)
        out = np.zeros(max(1, len(b)), dtype=np.int16)
        n = self.L.orc_query_sequence(self.h, self.can, self.K, b, len(b), out.ctypes.data_as(i16p))
        if n < 0:
            raise ValueError("query_sequence: %d" % n)
        return out[:n].copy()

    def insert_and_query_sequence(self, seq):
        b = _bytes_of(seq)
        out = np.zeros(max(1, len(b)), dtype=np.int16)
        n = self.L.orc_insert_and_query_sequence(self.h, self.can, self.K, b, len(b), out.ctypes.data_as(i16p))
        if n < 0:
            raise ValueError("insert_and_query_sequence: %d" % n)
        return out[:n].copy()

    def insert(self, h):
        return bool(self.L.orc_insert(self.h, int(h)))

    def query(self, h):
        return int(self.L.orc_query(self.h, int(h)))

    def insert_and_query(self, h):
        return int(self.L.orc_insert_and_query(self.h, int(h)))

    def insert_hashes(self, hashes):
        hs = _as_u64(hashes)
        out = np.zeros(max(1, hs.size), dtype=np.uint8)
        self.L.orc_insert_hashes(self.h, hs.ctypes.data_as(u64p), hs.size, out.ctypes.data_as(u8p))
        return out[:hs.size]

    def query_hashes(self, hashes):
        hs = _as_u64(hashes)
        out = np.zeros(max(1, hs.size), dtype=np.int16)
        self.L.orc_query_hashes(self.h, hs.ctypes.data_as(u64p), hs.size, out.ctypes.data_as(i16p))
        return out[:hs.size]

    def insert_reads(self, bases, offsets, n_threads=1, want_n_new=False):
        """Streams reads in order (the reference's processor loop).  Returns (k-mers, seconds[, n_new])."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = _as_u64(offsets)
        n = offsets.size - 1
        secs = C.c_double(0)
        nn = np.zeros(max(1, n), dtype=np.uint64) if want_n_new else None
        tot = self.L.orc_insert_reads(self.h, self.can, self.K, bases.ctypes.data, offsets.ctypes.data_as(u64p),
                                      n, nn.ctypes.data_as(u64p) if want_n_new else None, C.byref(secs))
        if want_n_new:
            return int(tot), secs.value, nn[:n]
        return int(tot), secs.value

    def query_reads(self, bases, offsets):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = _as_u64(offsets)
        n = offsets.size - 1
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        nk = np.maximum(lens - self.K + 1, 0)
        out = np.zeros(max(1, int(nk.sum())), dtype=np.int16)
        tot = self.L.orc_query_reads(self.h, self.can, self.K, bases.ctypes.data, offsets.ctypes.data_as(u64p), n,
                                     out.ctypes.data_as(i16p))
        return out[:tot]

    def median_count_at_least(self, seq, cutoff):
        b = _bytes_of(seq)
        r = self.L.orc_median_count_at_least(self.h, self.can, self.K, b, len(b), int(cutoff))
        if r < 0:
            raise ValueError("median_count_at_least: %d" % r)
        return bool(r)

    def diginorm_reads(self, bases, offsets, cutoff, batch=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = _as_u64(offsets)
        n = offsets.size - 1
        keep = np.zeros(max(1, n), dtype=np.uint8)
        self.L.orc_diginorm_reads(self.h, self.can, self.K, bases.ctypes.data, offsets.ctypes.data_as(u64p), n,
                                  int(cutoff), int(batch), keep.ctypes.data_as(u8p))
        return keep[:n]

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.L.orc_storage_stats(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def table(self, i):
        n = int(self.L.orc_storage_table_bytes(self.h, i))
        p = self.L.orc_storage_table(self.h, i)
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(n,)).copy()

    def tables(self):
        return [self.table(i) for i in range(len(self.sizes))]

    def reset(self):
        self.L.orc_storage_reset(self.h)

    def save(self, fn):
        if self.L.orc_storage_save(self.h, fn.encode(), self.K) != 0:
            raise IOError(fn)


class PortSketch:
    """SourmashSketch restatement (PARITY UNPINNED -- see goetia_oracle.c header)."""

    def __init__(self, num, K, seed=42, scaled=0, max_hash=None):
        self.L = _PortLib.get()
        self.K = int(K)
        self.max_hash = Port.max_hash_from_scaled(scaled) if max_hash is None else int(max_hash)
        self.h = self.L.orc_sketch_create(int(num), int(K), int(seed), self.max_hash)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_sketch_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_hash(self, h):
        self.L.orc_sketch_add_hash(self.h, int(h))

    def insert_sequence(self, seq):
        b = _bytes_of(seq)
        return int(self.L.orc_sketch_add_sequence(self.h, b, len(b)))

    def add_reads(self, bases, offsets):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = _as_u64(offsets)
        secs = C.c_double(0)
        tot = self.L.orc_sketch_add_reads(self.h, bases.ctypes.data, offsets.ctypes.data_as(u64p),
                                          offsets.size - 1, C.byref(secs))
        return int(tot), secs.value

    def mins(self):
        n = int(self.L.orc_sketch_size(self.h))
        if n == 0:
            return np.zeros(0, dtype=np.uint64)
        p = self.L.orc_sketch_mins(self.h)
        return np.ctypeslib.as_array(C.cast(p, u64p), shape=(n,)).copy()


# ----------------------------------------------------------------------------- Ref (the real thing)
class _RefLib:
    _lib = None

    @classmethod
    def get(cls):
        if cls._lib is None:
            if not os.path.exists(_REF_SO):
                raise FileNotFoundError(_REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
            L = C.CDLL(_REF_SO)
            L.ref_dbg_create.restype = C.c_void_p
            L.ref_dbg_create.argtypes = [C.c_int, C.c_int, C.c_int, u64p, C.c_int]
            L.ref_dbg_destroy.argtypes = [C.c_void_p]
            L.ref_dbg_insert_sequence.restype = C.c_int64
            L.ref_dbg_insert_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u64p]
            for fn in ("ref_dbg_query_sequence", "ref_dbg_insert_and_query_sequence"):
                getattr(L, fn).restype = C.c_int64
                getattr(L, fn).argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, i16p]
            L.ref_dbg_insert_hash.restype = C.c_int
            L.ref_dbg_insert_hash.argtypes = [C.c_void_p, C.c_uint64]
            L.ref_dbg_query_hash.restype = C.c_int16
            L.ref_dbg_query_hash.argtypes = [C.c_void_p, C.c_uint64]
            L.ref_dbg_insert_and_query_hash.restype = C.c_int16
            L.ref_dbg_insert_and_query_hash.argtypes = [C.c_void_p, C.c_uint64]
            L.ref_dbg_stats.argtypes = [C.c_void_p, u64p, u64p]
            L.ref_dbg_table_bytes.restype = C.c_uint64
            L.ref_dbg_table_bytes.argtypes = [C.c_void_p, C.c_int]
            L.ref_dbg_table.restype = C.c_void_p
            L.ref_dbg_table.argtypes = [C.c_void_p, C.c_int]
            L.ref_dbg_reset.argtypes = [C.c_void_p]
            L.ref_dbg_save.restype = C.c_int
            L.ref_dbg_save.argtypes = [C.c_void_p, C.c_char_p]
            L.ref_dbg_load.restype = C.c_int
            L.ref_dbg_load.argtypes = [C.c_void_p, C.c_char_p]
            L.ref_dbg_process_file.restype = C.c_int64
            L.ref_dbg_process_file.argtypes = [C.c_void_p, C.c_char_p, u64p, u64p]
            if hasattr(L, "ref_dbg_advance_trace"):
                L.ref_dbg_advance_trace.restype = C.c_int64
                L.ref_dbg_advance_trace.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64]
            L.ref_dbg_median_count_at_least.restype = C.c_int
            L.ref_dbg_median_count_at_least.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint]
            L.ref_dbg_insert_reads.restype = C.c_int64
            L.ref_dbg_insert_reads.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_int,
                                               C.POINTER(C.c_double)]
            L.ref_hash_sequence.restype = C.c_int64
            L.ref_hash_sequence.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint64, u64p, u64p]
            L.ref_hash_kmer.restype = C.c_int
            L.ref_hash_kmer.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint64, u64p, u64p]
            L.ref_primes_near.restype = C.c_int
            L.ref_primes_near.argtypes = [C.c_uint32, C.c_uint64, u64p]
            L.ref_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, u64p]
            L.ref_char_table.argtypes = [u64p]
            L.ref_parse_file.restype = C.c_int64
            L.ref_parse_file.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, u64p,
                                         C.c_uint64, u64p]
            cls._lib = L
        return cls._lib


class Ref(_GraphBase):
    """dBG<Storage, Shifter> of the unmodified reference (oracle/_ref/libgoetia_ref.so)."""

    name = "reference"

    def __init__(self, kind, can, K, sizes):
        self.L = _RefLib.get()
        self.kind, self.can, self.K = int(kind), int(can), int(K)
        self.sizes = _as_u64(sizes)
        self.h = self.L.ref_dbg_create(self.kind, self.can, self.K, self.sizes.ctypes.data_as(u64p),
                                       len(self.sizes))
        if not self.h:
            raise MemoryError("ref_dbg_create failed")

    def close(self):
        if self.h:
            self.L.ref_dbg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def hash_sequence(can, K, seq):
        L = _RefLib.get()
        b = _bytes_of(seq)
        n = max(1, len(b))
        fw = np.zeros(n, dtype=np.uint64)
        rc = np.zeros(n, dtype=np.uint64)
        r = L.ref_hash_sequence(int(can), int(K), b, len(b), fw.ctypes.data_as(u64p), rc.ctypes.data_as(u64p))
        if r < 0:
            raise ValueError("ref_hash_sequence: %d" % r)
        return fw[:r].copy(), rc[:r].copy()

    @staticmethod
    def hash_kmer(can, K, seq):
        L = _RefLib.get()
        b = _bytes_of(seq)
        fw, rc = C.c_uint64(0), C.c_uint64(0)
        if L.ref_hash_kmer(int(can), int(K), b, len(b), C.byref(fw), C.byref(rc)) != 0:
            raise ValueError("ref_hash_kmer")
        return int(fw.value), int(rc.value)

    @staticmethod
    def primes_near(n, x):
        L = _RefLib.get()
        out = np.zeros(max(1, n), dtype=np.uint64)
        k = L.ref_primes_near(int(n), int(x), out.ctypes.data_as(u64p))
        return [int(v) for v in out[:k]]

    @staticmethod
    def murmur3_x64_128(key, seed):
        L = _RefLib.get()
        b = _bytes_of(key)
        out = np.zeros(2, dtype=np.uint64)
        L.ref_murmur3_x64_128(b, len(b), int(seed), out.ctypes.data_as(u64p))
        return int(out[0]), int(out[1])

    @staticmethod
    def char_table():
        L = _RefLib.get()
        out = np.zeros(256, dtype=np.uint64)
        L.ref_char_table(out.ctypes.data_as(u64p))
        return out

    @staticmethod
    def parse_file(fn, strict=False, min_length=0):
        """FastxParser<DNA_SIMPLE> over a file -> (bases, offsets, n_skipped)."""
        L = _RefLib.get()
        cap = max(1 << 20, 4 * os.path.getsize(fn) * (40 if fn.endswith(".gz") else 1))
        buf = np.zeros(cap, dtype=np.uint8)
        maxrec = cap // 2 + 2
        offsets = np.zeros(maxrec + 1, dtype=np.uint64)
        sk = C.c_uint64(0)
        n = L.ref_parse_file(fn.encode(), int(strict), int(min_length), buf.ctypes.data, cap,
                             offsets.ctypes.data_as(u64p), maxrec, C.byref(sk))
        if n < 0:
            raise ValueError("ref_parse_file: %d" % n)
        offsets = offsets[:n + 1].copy()
        return buf[:int(offsets[-1])].copy(), offsets, int(sk.value)

    def insert_sequence(self, seq):
        b = _bytes_of(seq)
        nn = C.c_uint64(0)
        n = self.L.ref_dbg_insert_sequence(self.h, b, len(b), C.byref(nn))
        if n < 0:
            raise ValueError("insert_sequence: %d" % n)
        return int(n), int(nn.value)

    def query_sequence(self, seq):
        b = _bytes_of(seq)
        out = np.zeros(max(1, len(b)), dtype=np.int16)
        n = self.L.ref_dbg_query_sequence(self.h, b, len(b), out.ctypes.data_as(i16p))
        if n < 0:
            raise ValueError("query_sequence: %d" % n)
        return out[:n].copy()

    def insert_and_query_sequence(self, seq):
        b = _bytes_of(seq)
        out = np.zeros(max(1, len(b)), dtype=np.int16)
        n = self.L.ref_dbg_insert_and_query_sequence(self.h, b, len(b), out.ctypes.data_as(i16p))
        if n < 0:
            raise ValueError("insert_and_query_sequence: %d" % n)
        return out[:n].copy()

    def insert(self, h):
        return bool(self.L.ref_dbg_insert_hash(self.h, int(h)))

    def query(self, h):
        return int(self.L.ref_dbg_query_hash(self.h, int(h)))

    def insert_and_query(self, h):
        return int(self.L.ref_dbg_insert_and_query_hash(self.h, int(h)))

    def insert_hashes(self, hashes):
        return np.array([self.insert(int(h)) for h in hashes], dtype=np.uint8)

    def query_hashes(self, hashes):
        return np.array([self.query(int(h)) for h in hashes], dtype=np.int16)

    def insert_reads(self, bases, offsets, n_threads=1, want_n_new=False):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = _as_u64(offsets)
        n = offsets.size - 1
        if want_n_new:
            nn = np.zeros(n, dtype=np.uint64)
            tot = 0
            for r in range(n):
                s = bases[int(offsets[r]):int(offsets[r + 1])].tobytes()
                if len(s) >= self.K:
                    k, nn[r] = self.insert_sequence(s)
                    tot += k
            return tot, 0.0, nn
        secs = C.c_double(0)
        tot = self.L.ref_dbg_insert_reads(self.h, bases.ctypes.data, offsets.ctypes.data_as(u64p), n,
                                          int(n_threads), C.byref(secs))
        return int(tot), secs.value

    def query_reads(self, bases, offsets):
        offsets = _as_u64(offsets)
        out = []
        for r in range(offsets.size - 1):
            s = bases[int(offsets[r]):int(offsets[r + 1])].tobytes()
            if len(s) >= self.K:
                out.append(self.query_sequence(s))
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int16)

    def median_count_at_least(self, seq, cutoff):
        b = _bytes_of(seq)
        r = self.L.ref_dbg_median_count_at_least(self.h, b, len(b), int(cutoff))
        if r < 0:
            raise ValueError("median_count_at_least: %d" % r)
        return bool(r)

    def process_file(self, fn):
        """InserterProcessor<dBG>::process(filename) -> (k-mers consumed, n_seqs, n_skipped)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        t = self.L.ref_dbg_process_file(self.h, fn.encode(), C.byref(a), C.byref(b))
        return int(t), int(a.value), int(b.value)

    def advance_trace(self, fn, interval, strict=False, min_length=0, cap=1 << 16):
        """FileProcessor::advance until nothing remains -> [(n_sequences, time_total, remaining), ...]."""
        out = np.zeros((cap, 3), dtype=np.uint64)
        n = self.L.ref_dbg_advance_trace(self.h, fn.encode(), int(interval), int(bool(strict)), int(min_length), out.ctypes.data, cap)
        if n < 0:
            raise ValueError("advance_trace failed")
        return [(int(a), int(b), bool(c)) for a, b, c in out[:min(n, cap)]]

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.L.ref_dbg_stats(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def table(self, i):
        n = int(self.L.ref_dbg_table_bytes(self.h, i))
        p = self.L.ref_dbg_table(self.h, i)
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(n,)).copy()

    def tables(self):
        return [self.table(i) for i in range(len(self.sizes))]

    def reset(self):
        self.L.ref_dbg_reset(self.h)

    def save(self, fn):
        if self.L.ref_dbg_save(self.h, fn.encode()) != 0:
            raise IOError(fn)

    def load(self, fn):
        if self.L.ref_dbg_load(self.h, fn.encode()) != 0:
            raise IOError(fn)
