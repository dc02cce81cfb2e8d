#!/usr/bin/env python
"""FASTX front end measured: parser throughput (serial / chunk-parallel engine) and file -> tables end to end
(gt_insert_fastx: parse of batch n+1 overlaps the GPU work of batch n), with the reference's own
FileProcessor<InserterProcessor<dBG>>::process over the same file as the CPU baseline (bounded sample).

    python scripts/bench_fastx.py [--reads N] [--out profiles/r1_fastx_bench.json]

The oracle (oracle/_ref) appears only as that baseline, never on the product path.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def write_fastq(fn, n_reads, length, seed):
    rng = np.random.default_rng(seed)
    with open(fn, "wb") as f:
        for r0 in range(0, n_reads, 1_000_000):
            n = min(1_000_000, n_reads - r0)
            w = 15 + length + 3 + length + 1
            rec = np.empty((n, w), dtype=np.uint8)
            rec[:, :15] = np.frombuffer(b"@read0000000/1\n", dtype=np.uint8)
            rec[:, 15:15 + length] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, length))]
            rec[:, 15 + length:18 + length] = np.frombuffer(b"\n+\n", dtype=np.uint8)
            rec[:, 18 + length:18 + 2 * length] = ord("I")
            rec[:, -1] = ord("\n")
            f.write(rec.tobytes())
    return os.path.getsize(fn)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--length", type=int, default=150)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import goetia_b200 as gb
    from goetia_b200 import _capi
    L = _capi.load()
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    fn = os.path.join(d, "gt_bench_%d.fq" % os.getpid())
    size = write_fastq(fn, args.reads, args.length, 7)
    out = {"file_bytes": size, "reads": args.reads, "read_len": args.length, "host_cores": os.cpu_count(), "parser": []}
    try:
        bases = np.empty(512 << 20, dtype=np.uint8)
        bases[:] = 0
        offs = np.empty((8 << 20) + 1, dtype=np.uint64)
        for thr in sorted({1, 2, 4, 8, min(16, os.cpu_count() or 1)}):
            os.environ["GT_FASTX_THREADS"] = str(thr)
            p = gb.FastxParser(fn)
            t = time.perf_counter()
            nb = nr = 0
            while True:
                n = L.gt_fastx_next_batch(p.handle, bases.ctypes.data, bases.size, offs.ctypes.data, 8 << 20)
                if n <= 0:
                    break
                nb += int(offs[n])
                nr += n
            dt = time.perf_counter() - t
            p.close()
            assert nr == args.reads
            out["parser"].append({"threads": thr, "engine": "serial" if thr == 1 else "chunk-parallel",
                                  "file_GB_per_s": size / dt / 1e9, "Gbases_per_s": nb / dt / 1e9})
        os.environ.pop("GT_FASTX_THREADS", None)
        # file -> tables on the GPU
        K = 31
        sizes = gb.get_n_primes_near_x(4, int(8e9))
        g = gb.dBG[gb.BitStorage, gb.CanLemireShifter].build(gb.BitStorage(sizes), K)
        g.process_fastx(fn, max_reads=200_000)  # warm-up: store allocation, first launches
        t = time.perf_counter()
        n_seqs, n_kmers = g.process_fastx(fn)
        dt = time.perf_counter() - t
        out["file_to_tables"] = {"workload": "dBG<BitStorage,CanLemireShifter> K=31, 4 x 8e9 bits", "sequences": n_seqs,
                                 "kmers": n_kmers, "seconds": dt, "kmers_per_s": n_kmers / dt,
                                 "file_GB_per_s": size / dt / 1e9,
                                 "api": "gt_fastx_open + gt_insert_fastx(GT_MODE_BLIND) + gt_storage_flush"}
        del g
        # the reference's own streaming driver on a bounded prefix of the same file
        from oracle import binding
        if binding.have_ref():
            n_ref = min(args.reads, 400_000)
            fn2 = fn + ".head"
            with open(fn, "rb") as f, open(fn2, "wb") as o:
                o.write(f.read(n_ref * (15 + 2 * args.length + 4)))
            ref = binding.Ref(0, 1, K, sizes)
            t = time.perf_counter()
            res = ref.process_file(fn2)
            dt = time.perf_counter() - t
            os.remove(fn2)
            out["cpu_reference"] = {"kind": "reference", "sample_reads": n_ref, "result": [int(x) for x in res],
                                    "seconds": dt, "kmers_per_s": n_ref * (args.length - K + 1) / dt,
                                    "api": "FileProcessor<InserterProcessor<dBG>>::process (processors.hh:112-127)"}
            ref.close()
    finally:
        os.remove(fn)
    line = json.dumps(out)
    print(line)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            f.write(line + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
