"""Counter-based synthetic reads for harnesses (bench.py, full-size parity tests, golden generation).

The base at GLOBAL index i of stream `seed` is a pure function of (seed, i) -- so the same reads exist on the
CPU (this module, numpy), on one GPU and on every rank of a sharded run (gt_synth_bases_dev, kernels.cuh
k_synth_bases; same arithmetic):

    z = splitmix64(seed + (i // 32 + 1) * 0x9E3779B97F4A7C15);  base(i) = "ACGT"[(z >> 2 * (i % 32)) & 3]

Equal-length reads are cut from the flat stream: read r = bases [r * L, (r + 1) * L).
"""
import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_LUT = np.frombuffer(b"ACGT", dtype=np.uint8)
_SHIFTS = (np.arange(32, dtype=np.uint64) * np.uint64(2))


def synth_words(seed, w0, n_words):
    """splitmix64 words w0 .. w0+n_words-1 of stream `seed` (uint64 array)."""
    with np.errstate(over="ignore"):
        w = np.arange(w0 + 1, w0 + 1 + n_words, dtype=np.uint64)
        z = np.uint64(seed) + w * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def synth_bases(seed, first_base, n_bases):
    """ASCII bytes of the global bases [first_base, first_base + n_bases) of stream `seed` (uint8 array)."""
    if n_bases <= 0:
        return np.zeros(0, dtype=np.uint8)
    w0 = first_base // 32
    w1 = (first_base + n_bases + 31) // 32
    out = np.empty((w1 - w0) * 32, dtype=np.uint8)
    step = 1 << 20  # words per block: bounds the temporaries
    for a in range(w0, w1, step):
        b = min(w1, a + step)
        z = synth_words(seed, a, b - a)
        codes = ((z[:, None] >> _SHIFTS[None, :]) & np.uint64(3)).astype(np.uint8)
        out[(a - w0) * 32:(b - w0) * 32] = _LUT[codes].reshape(-1)
    lo = first_base - w0 * 32
    return out[lo:lo + n_bases]


def synth_reads(seed, first_read, n_reads, read_len):
    """(bases, offsets) of reads [first_read, first_read + n_reads) of `read_len` bases each."""
    bases = synth_bases(seed, first_read * read_len, n_reads * read_len)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return np.ascontiguousarray(bases), offsets
