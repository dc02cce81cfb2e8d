"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libgoetia_ref.so).

Run only where /root/reference exists:   python tests/golden/make_golden.py
The JSON is committed; tests read it on boxes where the reference is absent.  Every value in
it was produced by the compiled reference code (dBG / storages / shifters / FastxParser /
vendored MurmurHash3) -- nothing comes from our own restatement or CUDA path.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.binding import Port, Ref, build_ref, synth_reads  # noqa: E402
from tests.util import genome_reads, ragged_reads, read_str  # noqa: E402

REF_DATA = "/root/reference/tests/test-data"
fnv = Port.fnv1a  # plain FNV-1a-64 over bytes (checksum only)

CONFIGS = [  # (kind, can, K)
    (0, 1, 31), (0, 0, 31), (1, 1, 21), (1, 0, 21), (2, 1, 25), (2, 0, 25),
]


def table_summary(g):
    tabs = g.tables()
    return {
        "fnv": [str(fnv(t)) for t in tabs],
        "bytesum": [int(t.astype(np.uint64).sum()) for t in tabs],
        "popcount": [int(np.unpackbits(t).sum()) for t in tabs],
        "nbytes": [int(t.size) for t in tabs],
    }


def graph_case(kind, can, K, sizes, bases, offsets, passes=1):
    g = Ref(kind, can, K, sizes)
    tot = 0
    n_new_all = None
    for _ in range(passes):
        t, _, n_new = g.insert_reads(bases, offsets, want_n_new=True)
        tot += t
        if n_new_all is None:
            n_new_all = n_new
    n_unique, n_occ = g.stats()
    first = read_str(bases, offsets, 0)
    out = {
        "kind": kind, "can": can, "K": K, "sizes": [int(s) for s in sizes], "passes": passes,
        "n_kmers": tot, "n_unique": n_unique, "n_occupied": n_occ,
        "n_new_first_pass_fnv": str(fnv(n_new_all.astype(np.uint64).view(np.uint8))),
        "n_new_first_pass_head": [int(v) for v in n_new_all[:16]],
        "tables": table_summary(g),
        "query_first_read": [int(c) for c in g.query_sequence(first)] if len(first) >= K else [],
    }
    return out


def main():
    build_ref()
    G = {"_generator": "tests/golden/make_golden.py against oracle/_ref (unmodified reference)"}

    # -- hashing known answers --------------------------------------------------------------
    kat = "TCACCTGTGTTGTGCTACTTGCGGCGC"
    fw, rc = Ref.hash_sequence(1, 27, kat)
    G["kat"] = {"seq": kat, "K": 27, "fw": str(int(fw[0])), "rc": str(int(rc[0])),
                "fwd_only": str(int(Ref.hash_sequence(0, 27, kat)[0][0])),
                "static_fw_rc": [str(v) for v in Ref.hash_kmer(1, 27, kat)]}
    tab = Ref.char_table()
    G["char_table"] = {c: str(int(tab[ord(c)])) for c in "ACGTN"}
    G["char_table"]["NUL"] = str(int(tab[0]))
    G["char_table"]["fnv_all_256"] = str(fnv(tab.view(np.uint8)))

    hv = {}
    seq150 = read_str(*synth_reads(1, 260, seed=1234), 0)
    for K in [1, 2, 21, 27, 31, 32, 33, 63, 64, 65, 101, 200]:
        for can in (0, 1):
            fw, rc = Ref.hash_sequence(can, K, seq150)
            hv["K%d_can%d" % (K, can)] = {
                "n": int(fw.size), "fw_fnv": str(fnv(fw.view(np.uint8))),
                "rc_fnv": str(fnv(rc.view(np.uint8))) if can else None,
                "fw_head": [str(int(v)) for v in fw[:3]], "rc_head": [str(int(v)) for v in rc[:3]] if can else None,
            }
    G["hash_vectors"] = {"seq": seq150, "cases": hv}

    # -- table sizing --------------------------------------------------------------------------
    G["primes"] = {"%d,%d" % (n, x): Ref.primes_near(n, x) for n, x in
                   [(4, 10**6), (4, 10**8), (4, 10**9), (4, 8 * 10**9), (4, 2 * 10**6), (4, 3 * 10**6), (2, 100), (5, 20),
                    (3, 8), (1, 3), (2, 2), (3, 1), (4, 100000)]}

    # -- MurmurHash3_x64_128 (vendored src/goetia/hashing/smhasher/MurmurHash3.cc) ---------------
    keys = ["ACG", "A" * 21, "ACGT" * 8, "T" * 31, "GATTACA", "", "C" * 15, "G" * 16, "ACGTTGCA" * 6 + "AC", "N" * 17]
    G["murmur"] = [{"key": k, "seed": s, "h": [str(v) for v in Ref.murmur3_x64_128(k, s)]} for k in keys for s in (42, 0)]

    # -- the reference's own fixture: tests/test-data/random-20-a.fa through InserterProcessor -----
    fa = os.path.join(REF_DATA, "random-20-a.fa")
    bases, offsets, n_skipped = Ref.parse_file(fa)
    reads = [read_str(bases, offsets, r) for r in range(offsets.size - 1)]
    sizes6 = Ref.primes_near(4, 10**6)
    fx = {"reads": reads, "n_skipped": n_skipped, "cases": []}
    for kind, can, K in CONFIGS:
        g = Ref(kind, can, K, sizes6)
        n_kmers, n_seqs, _ = g.process_file(fa)  # the reference's own streaming driver
        n_unique, n_occ = g.stats()
        case = {"kind": kind, "can": can, "K": K, "sizes": sizes6, "n_kmers": n_kmers, "n_seqs": n_seqs,
                "n_unique": n_unique, "n_occupied": n_occ, "tables": table_summary(g)}
        fwh, rch = Ref.hash_sequence(can, K, reads[0])
        case["first_read_hashes"] = [[str(int(fwh[i])), str(int(rch[i])) if can else None] for i in range(3)]
        if kind != 0:
            g.process_file(fa)
            g.process_file(fa)
            case["query_first_read_after_3_passes"] = [int(c) for c in g.query_sequence(reads[0])]
            case["n_unique_after_3_passes"] = g.stats()[0]
        fx["cases"].append(case)
    G["random20a"] = fx

    # -- synthetic inputs (regenerated by the tests from the same seeded numpy code) ---------------
    sizes2m = Ref.primes_near(4, 2 * 10**6)
    syn = []
    b, o = synth_reads(2000, 150, seed=42)
    for kind, can, K in CONFIGS:
        c = graph_case(kind, can, K, sizes2m, b, o)
        c["input"] = {"gen": "synth_reads", "n_reads": 2000, "length": 150, "seed": 42}
        syn.append(c)
    b, o = genome_reads(3000, 100, 4000, seed=77)
    for kind, can, K in CONFIGS:
        c = graph_case(kind, can, K, sizes2m, b, o, passes=2 if kind else 1)
        c["input"] = {"gen": "genome_reads", "n_reads": 3000, "read_len": 100, "genome_len": 4000, "seed": 77}
        syn.append(c)
    b, o = ragged_reads(1500, 0, 300, seed=5)
    for kind, can, K in [(0, 1, 31), (1, 1, 21), (2, 0, 25)]:
        c = graph_case(kind, can, K, sizes2m, b, o)
        c["input"] = {"gen": "ragged_reads", "n_reads": 1500, "min_len": 0, "max_len": 300, "seed": 5}
        syn.append(c)
    # saturation: one read repeated past the counter maximum
    one, _ = synth_reads(1, 60, seed=1)
    for kind, reps in [(1, 300), (2, 60)]:
        bb = np.tile(one, reps)
        oo = np.arange(reps + 1, dtype=np.uint64) * np.uint64(60)
        c = graph_case(kind, 1, 21, Ref.primes_near(4, 100000), bb, oo)
        c["input"] = {"gen": "tiled", "seed": 1, "length": 60, "reps": reps}
        syn.append(c)
    G["synthetic"] = syn

    # -- insert_and_query_sequence (dbg.hh:327-340) on a sequence with repeated k-mers ---------------
    rep = read_str(*synth_reads(1, 80, seed=3), 0)
    rep = rep + rep[:50] + rep[10:70]
    iq = []
    for kind, can, K in [(0, 1, 21), (1, 1, 21), (2, 0, 21)]:
        g = Ref(kind, can, K, sizes6)
        first = [int(c) for c in g.insert_and_query_sequence(rep)]
        second = [int(c) for c in g.insert_and_query_sequence(rep)]
        iq.append({"kind": kind, "can": can, "K": K, "sizes": sizes6, "seq": rep, "first": first, "second": second,
                   "n_unique": g.stats()[0]})
    G["insert_and_query_sequence"] = iq

    # -- DiginormFilter::median_count_at_least (diginorm.hh:35-68) ------------------------------------
    b, o = genome_reads(2000, 100, 3000, seed=21)
    qb, qo = genome_reads(200, 100, 3000, seed=22, sub_rate=0.08)
    med = []
    for kind in (1, 2):
        g = Ref(kind, 1, 21, sizes2m)
        g.insert_reads(b, o)
        res = {}
        for cutoff in (1, 5, 14, 40):
            res[str(cutoff)] = [int(g.median_count_at_least(read_str(qb, qo, r), cutoff)) for r in range(200)]
        med.append({"kind": kind, "can": 1, "K": 21, "sizes": sizes2m, "pass": res})
    G["median_count_at_least"] = {"insert": {"gen": "genome_reads", "n_reads": 2000, "read_len": 100, "genome_len": 3000, "seed": 21},
                                  "query": {"gen": "genome_reads", "n_reads": 200, "read_len": 100, "genome_len": 3000, "seed": 22,
                                            "sub_rate": 0.08},
                                  "cases": med}

    # -- OXLI v4 files written by the reference's save() ------------------------------------------------
    b, o = synth_reads(200, 100, seed=9)
    ox = []
    with tempfile.TemporaryDirectory() as td:
        for kind in (0, 1, 2):
            g = Ref(kind, 1, 21, Ref.primes_near(3, 20000))
            g.insert_reads(b, o)
            fn = os.path.join(td, "t%d.oxli" % kind)
            g.save(fn)
            data = np.fromfile(fn, dtype=np.uint8)
            ox.append({"kind": kind, "can": 1, "K": 21, "sizes": Ref.primes_near(3, 20000), "file_fnv": str(fnv(data)),
                       "file_bytes": int(data.size), "head_hex": data[:32].tobytes().hex()})
    G["oxli"] = {"input": {"gen": "synth_reads", "n_reads": 200, "length": 100, "seed": 9}, "files": ox}

    # -- FastxParser<DNA_SIMPLE> over the reference's fixtures (parsing/readers.hh:150-219) -------------
    pr = {}
    for name in ("random-20-a.fa", "test-fastq-reads.fq", "left.fq", "right.fq"):
        bb, oo, sk = Ref.parse_file(os.path.join(REF_DATA, name))
        pr[name] = {"n_reads": int(oo.size - 1), "n_skipped": sk, "n_bases": int(bb.size), "seq_fnv": str(fnv(bb)),
                    "len_fnv": str(fnv((oo[1:] - oo[:-1]).astype(np.uint64).view(np.uint8)))}
    # our own small FASTA with lower case, an invalid read and a short read (tests/test_parsing.py:34-51 shapes)
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "mixed.fa")
        txt = ">r1\nACGTacgtACGTTTGA\nCCGTA\n>r2 bad\nACGTNACGT\n>r3\nacgtacgt\n>r4\nAC\n>r5\nGGGGCCCCAAAATTTT\n"
        with open(fn, "w") as f:
            f.write(txt)
        bb, oo, sk = Ref.parse_file(fn)
        pr["_mixed.fa"] = {"text": txt, "reads": [read_str(bb, oo, r) for r in range(oo.size - 1)], "n_skipped": sk}
        fq = os.path.join(td, "mixed.fq")
        txtq = "@q1\nACGTAC\n+\nIIIIII\n@q2\nacgtnn\n+\nIIIIII\n@q3\nTTTTGGGG\n+q3\nIIIIIIII\n"
        with open(fq, "w") as f:
            f.write(txtq)
        bb, oo, sk = Ref.parse_file(fq)
        pr["_mixed.fq"] = {"text": txtq, "reads": [read_str(bb, oo, r) for r in range(oo.size - 1)], "n_skipped": sk}
    G["parser"] = pr

    out = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
