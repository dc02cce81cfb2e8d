#!/usr/bin/env python
"""Generate tests/golden/fullsize_golden.json from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile): position-weighted table checksums (gt_storage_checksum's definition,
tests/util.table_checksum) and n_occupied of the BASELINE.json workloads at their FULL sizes, on the
counter-based synthetic reads of goetia_b200/synth.py -- the same reads bench.py and tests/test_gpu_fullsize.py
feed the GPU path, on one GPU or sharded over any number of ranks.

    python tests/golden/make_fullsize_golden.py [case ...]      (runs in this container; minutes to hours)

BitStorage (atomic OR) and NibbleStorage (per-table mutex, nibblestorage.cc:60-100) are exact under threads, so
those run one dBG copy per core over one shared storage (dbg.hh:97-101); ByteStorage's increment is a plain
read-modify-write (bytestorage.cc:81-86), so it runs on one thread.  Existing entries of the JSON are kept.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from goetia_b200.synth import synth_reads  # noqa: E402
from oracle import binding  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize_golden.json")

# name: (storage kind, K, x, n_tables, first read, n reads, read length, seed)
CASES = {
    "c1": (0, 31, int(1e9), 4, 0, 100_000, 150, 42),
    "c3_first_1m": (0, 31, int(8e9), 4, 0, 1_000_000, 150, 44),
    "c3": (0, 31, int(8e9), 4, 0, 50_000_000, 150, 44),
    "c2_first_1m": (1, 21, int(4e9), 4, 0, 1_000_000, 150, 43),
    "c2": (1, 21, int(4e9), 4, 0, 20_000_000, 150, 43),
    "c5_first_125k": (2, 25, int(8e9), 4, 0, 125_000, 10_000, 46),
    "c5": (2, 25, int(8e9), 4, 0, 1_000_000, 10_000, 46),
}


def checksum_chunked(table_bytes, chunk_words=1 << 26):
    b = table_bytes
    n_words = (b.size + 3) // 4
    total = np.uint64(0)
    with np.errstate(over="ignore"):
        for w0 in range(0, n_words, chunk_words):
            w1 = min(n_words, w0 + chunk_words)
            part = b[w0 * 4:min(b.size, w1 * 4)]
            if part.size % 4:
                part = np.concatenate([part, np.zeros(4 - part.size % 4, dtype=np.uint8)])
            words = part.view("<u4").astype(np.uint64)
            k = np.arange(w0, w1, dtype=np.uint64) + np.uint64(0x9e3779b97f4a7c15)
            k ^= k >> np.uint64(33)
            k *= np.uint64(0xff51afd7ed558ccd)
            k ^= k >> np.uint64(33)
            k *= np.uint64(0xc4ceb9fe1a85ec53)
            k ^= k >> np.uint64(33)
            k |= np.uint64(1)
            total += (words * k).sum(dtype=np.uint64)
    return int(total)


def run_case(name):
    kind, K, x, n_tables, r0, n_reads, L, seed = CASES[name]
    sizes = binding.Ref.primes_near(n_tables, x)
    threads = 1 if kind == 1 else (os.cpu_count() or 1)
    ref = binding.Ref(kind, 1, K, sizes)
    step = max(1, (1 << 28) // L)  # ~256 MB of ASCII per chunk
    t0 = time.time()
    nk = 0
    for a in range(r0, r0 + n_reads, step):
        n = min(step, r0 + n_reads - a)
        b, o = synth_reads(seed, a, n, L)
        k, _ = ref.insert_reads(b, o, n_threads=threads)
        nk += k
        print("  %s: %d / %d reads, %.0f s" % (name, a - r0 + n, n_reads, time.time() - t0), flush=True)
    assert nk == n_reads * (L - K + 1)
    n_unique, n_occ = ref.stats()
    sums = [checksum_chunked(ref.table(i)) for i in range(n_tables)]
    ref.close()
    return {"kind": kind, "K": K, "x": x, "tablesizes": [int(s) for s in sizes], "first_read": r0, "reads": n_reads,
            "read_len": L, "seed": seed, "kmers": nk, "n_occupied": int(n_occ), "checksums": sums,
            "generator": "goetia_b200/synth.py (counter-based splitmix64)", "threads": threads,
            "made_by": "oracle/_ref (unmodified reference), dBG<S, CanLemireShifter>::insert_sequence per read",
            "seconds": round(time.time() - t0, 1)}


def main():
    assert binding.have_ref() or binding.build_ref() or binding.have_ref(), "oracle/_ref is needed (make -C oracle ref)"
    want = sys.argv[1:] or list(CASES)
    for name in want:
        print("case", name, flush=True)
        res = run_case(name)
        gold = {}
        if os.path.exists(OUT):
            with open(OUT) as f:
                gold = json.load(f)
        gold[name] = res
        with open(OUT, "w") as f:
            json.dump(gold, f, indent=1, sort_keys=True)
        print("  ->", res["checksums"], res["n_occupied"], res["seconds"], "s", flush=True)


if __name__ == "__main__":
    main()
