"""Shared helpers for the parity tests."""
import numpy as np

from oracle.binding import Port, synth_reads  # noqa: F401  (the checker)

STORAGES = [(0, "BitStorage"), (1, "ByteStorage"), (2, "NibbleStorage")]
SHIFTERS = [(0, "FwdLemireShifter"), (1, "CanLemireShifter")]


def storage_cls(gb, kind):
    return [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind]


def shifter_cls(gb, can):
    return [gb.FwdLemireShifter, gb.CanLemireShifter][can]


def make_graph(gb, kind, can, K, sizes):
    return gb.dBG[storage_cls(gb, kind), shifter_cls(gb, can)].build(storage_cls(gb, kind)(sizes), K)


def ragged_reads(n_reads, min_len, max_len, seed, alphabet=b"ACGT"):
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, n_reads)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    codes = rng.integers(0, len(alphabet), int(offsets[-1]), dtype=np.uint8)
    bases = np.frombuffer(alphabet, dtype=np.uint8)[codes]
    return np.ascontiguousarray(bases), offsets


def genome_reads(n_reads, read_len, genome_len, seed, sub_rate=0.01):
    """Reads sampled from a random genome (both strands, ~1 % substitutions) so that counts > 1."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    starts = rng.integers(0, genome_len - read_len + 1, n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    codes = genome[idx]
    flip = rng.random(n_reads) < 0.5
    codes[flip] = (3 - codes[flip])[:, ::-1]
    subs = rng.random(codes.shape) < sub_rate
    codes[subs] = (codes[subs] + rng.integers(1, 4, int(subs.sum()), dtype=np.uint8)) % 4
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return np.ascontiguousarray(bases), offsets


def read_str(bases, offsets, r):
    return bases[int(offsets[r]):int(offsets[r + 1])].tobytes().decode()


def assert_tables_equal(gpu_tables, cpu_tables):
    assert len(gpu_tables) == len(cpu_tables)
    for i, (a, b) in enumerate(zip(gpu_tables, cpu_tables)):
        assert a.shape == b.shape, "table %d size" % i
        if not np.array_equal(a, b):
            bad = np.nonzero(a != b)[0]
            raise AssertionError("table %d differs at %d bytes, first at %d: gpu=%d cpu=%d"
                                 % (i, bad.size, bad[0], a[bad[0]], b[bad[0]]))


def table_checksum(table_bytes):
    """numpy rendering of gt_storage_checksum (include/goetia_b200.h): sum of word * weight(index) mod 2^64."""
    b = np.ascontiguousarray(table_bytes, dtype=np.uint8)
    if b.size % 4:
        b = np.concatenate([b, np.zeros(4 - b.size % 4, dtype=np.uint8)])
    words = b.view("<u4").astype(np.uint64)
    with np.errstate(over="ignore"):
        k = np.arange(words.size, dtype=np.uint64) + np.uint64(0x9e3779b97f4a7c15)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xff51afd7ed558ccd)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xc4ceb9fe1a85ec53)
        k ^= k >> np.uint64(33)
        k |= np.uint64(1)
        return int((words * k).sum(dtype=np.uint64))
