"""The drop-in boundary, compiled: the UNMODIFIED reference templates goetia::dBG<StorageType, ShifterType>,
KmerIterator, UnitigWalker and FileProcessor / InserterProcessor instantiated over the GPU-backed StorageType of
adapter/goetia_gpu_storage.hh (adapter/adapter_harness.cc, built by adapter/Makefile against /root/reference/include).

The reference's own fixture -- tests/test-data/random-20-a.fa, whose reads are committed in tests/golden/golden.json --
is streamed through the reference's own InserterProcessor::process into dBG<GpuXStorage, XLemireShifter>, and
get_raw_tables() must match the golden FNVs the CPU storages produced (15765873264133613085 ...).
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests.test_oracle import fnv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adapter", "_build", "libgoetia_adapter.so")
u64p = C.POINTER(C.c_uint64)

SYMBOLS = ["gad_last_error", "gad_create", "gad_destroy", "gad_defer", "gad_process_file", "gad_insert_sequence",
           "gad_query_sequence", "gad_insert_and_query_sequence", "gad_stats", "gad_table_bytes", "gad_table", "gad_save",
           "gad_load", "gad_reset"]


def load_adapter():
    if not os.path.exists(LIB):
        pytest.skip("adapter/_build/libgoetia_adapter.so not built (make -C adapter; needs /root/reference)")
    L = C.CDLL(LIB)
    L.gad_last_error.restype = C.c_char_p
    L.gad_create.restype = C.c_void_p
    L.gad_create.argtypes = [C.c_int, C.c_int, C.c_int, u64p, C.c_int]
    L.gad_destroy.argtypes = [C.c_void_p]
    L.gad_defer.argtypes = [C.c_void_p, C.c_uint64]
    L.gad_process_file.restype = C.c_int64
    L.gad_process_file.argtypes = [C.c_void_p, C.c_char_p, u64p]
    for name in ("gad_insert_sequence", "gad_query_sequence", "gad_insert_and_query_sequence"):
        getattr(L, name).restype = C.c_int64
        getattr(L, name).argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p]
    L.gad_stats.argtypes = [C.c_void_p, u64p, u64p]
    L.gad_table_bytes.restype = C.c_uint64
    L.gad_table_bytes.argtypes = [C.c_void_p, C.c_int]
    L.gad_table.restype = C.POINTER(C.c_uint8)
    L.gad_table.argtypes = [C.c_void_p, C.c_int]
    L.gad_save.argtypes = [C.c_void_p, C.c_char_p]
    L.gad_load.argtypes = [C.c_void_p, C.c_char_p]
    L.gad_reset.argtypes = [C.c_void_p]
    return L


def test_adapter_library_exports():
    """CPU: the compiled boundary exists and exports its driver (no compute call)."""
    L = load_adapter()
    for s in SYMBOLS:
        assert hasattr(L, s), s


class Graph:
    def __init__(self, L, kind, can, K, sizes):
        self.L, self.n = L, len(sizes)
        arr = (C.c_uint64 * len(sizes))(*[int(x) for x in sizes])
        self.h = L.gad_create(kind, can, K, arr, len(sizes))
        assert self.h, L.gad_last_error()

    def close(self):
        if self.h:
            self.L.gad_destroy(self.h)
            self.h = None

    def tables(self):
        out = []
        for i in range(self.n):
            nb = int(self.L.gad_table_bytes(self.h, i))
            p = self.L.gad_table(self.h, i)
            assert p, self.L.gad_last_error()
            out.append(np.ctypeslib.as_array(p, shape=(nb,)).copy())
        return out

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        assert self.L.gad_stats(self.h, C.byref(a), C.byref(b)) == 0, self.L.gad_last_error()
        return int(a.value), int(b.value)


@pytest.fixture(scope="module")
def fasta(tmp_path_factory, golden):
    p = tmp_path_factory.mktemp("adapter") / "random-20-a.fa"
    with open(p, "w") as f:
        for i, s in enumerate(golden["random20a"]["reads"]):
            f.write(">read%d\n%s\n" % (i, s))
    return str(p)


@pytest.mark.gpu
@pytest.mark.parametrize("deferred", [0, 1000])
def test_reference_processor_over_gpu_storage(gb, golden, fasta, deferred):
    """InserterProcessor<dBG<GpuStorage, Shifter>>::process(file) == the CPU storages' golden tables."""
    L = load_adapter()
    for c in golden["random20a"]["cases"]:
        g = Graph(L, c["kind"], c["can"], c["K"], c["sizes"])
        if deferred:
            assert L.gad_defer(g.h, deferred) == 0
        n_seqs = C.c_uint64(0)
        nk = L.gad_process_file(g.h, fasta.encode(), C.byref(n_seqs))
        assert nk == c["n_kmers"], L.gad_last_error()
        assert int(n_seqs.value) == c["n_seqs"]
        tabs = g.tables()
        assert [t.size for t in tabs] == c["tables"]["nbytes"]
        assert [str(fnv(t)) for t in tabs] == c["tables"]["fnv"]
        assert g.stats() == (c["n_unique"], c["n_occupied"])
        g.close()


@pytest.mark.gpu
def test_dbg_members_over_gpu_storage(gb, golden, tmp_path):
    """dBG::insert_sequence(seq, n_new) / query_sequence / insert_and_query_sequence / save / load / reset through the
    unmodified template, against the oracle (one k-mer at a time, the reference's own loop)."""
    from tests.util import Port
    L = load_adapter()
    reads = golden["random20a"]["reads"]
    for kind, K in [(0, 31), (1, 21), (2, 25)]:
        sizes = gb.get_n_primes_near_x(4, 100_003)
        g = Graph(L, kind, 1, K, sizes)
        ref = Port(kind, 1, K, sizes)
        for s in reads[:12] + reads[:4]:
            b = s.encode()
            n_new = C.c_uint64(0)
            nk = L.gad_insert_sequence(g.h, b, len(b), C.byref(n_new))
            ek, en = ref.insert_sequence(s)
            assert (nk, int(n_new.value)) == (ek, en)
        s = reads[20]
        counts = np.zeros(len(s), dtype=np.int16)
        n = L.gad_insert_and_query_sequence(g.h, s.encode(), len(s), counts.ctypes.data)
        assert np.array_equal(counts[:n], ref.insert_and_query_sequence(s))
        for s in (reads[0], reads[20], reads[50]):
            n = L.gad_query_sequence(g.h, s.encode(), len(s), counts.ctypes.data)
            assert np.array_equal(counts[:n], ref.query_sequence(s))
        for a, b in zip(g.tables(), ref.tables()):
            assert np.array_equal(a, b)
        assert g.stats() == ref.stats()
        # OXLI round trip through the adapter; the file equals what the Python mirror writes (itself pinned to the
        # reference's files by test_oxli_files_match_reference)
        fn = str(tmp_path / ("adapter_%d.oxli" % kind))
        assert L.gad_save(g.h, fn.encode()) == 0, L.gad_last_error()
        st = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](sizes)
        for i, t in enumerate(ref.tables()):
            gb._capi.check(gb._capi.lib().gt_storage_upload_table(st.handle, i, np.ascontiguousarray(t).ctypes.data), "upload")
        fn2 = str(tmp_path / ("mirror_%d.oxli" % kind))
        st.save(fn2, K)
        assert open(fn, "rb").read() == open(fn2, "rb").read()
        assert L.gad_reset(g.h) == 0
        assert g.stats()[1] == 0
        assert L.gad_load(g.h, fn.encode()) == K, L.gad_last_error()
        for a, b in zip(g.tables(), ref.tables()):
            assert np.array_equal(a, b)
        g.close()
        ref.close()
