"""FASTX front end (gt_fastx_*, goetia_b200/parsing.py) against the reference parser.

CPU tests: the parser is host code inside libgoetia_b200.so and needs no GPU.  Expected values come from
tests/golden/fastx_golden.json and tests/golden/golden.json, both produced by the UNMODIFIED reference
(FastxParser<DNA_SIMPLE>, parsing/readers.hh:150-219) -- see tests/golden/make_fastx_golden.py.
The GPU test streams a file through gt_insert_fastx (FileProcessor + InserterProcessor) and compares the
tables with the oracle's after the same records.
"""
import json
import os

import numpy as np
import pytest

from oracle.binding import Port
from tests.fastx_cases import cases, write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "fastx_golden.json")))


def parse_all(fn, strict, min_length, max_bases=1 << 16, max_reads=None, by_record=False):
    from goetia_b200.parsing import FastxParser
    p = FastxParser(fn, strict, min_length)
    seqs = []
    if by_record:
        while not p.is_complete():
            r = p.next()
            if r is not None:
                seqs.append(r.sequence.encode())
    else:
        while True:
            b, o = p.next_batch(max_bases, max_reads)
            if o.size == 1:
                break
            seqs.extend(b[int(o[i]):int(o[i + 1])].tobytes() for i in range(o.size - 1))
    stats = (p.n_parsed(), p.n_skipped(), p.is_complete())
    p.close()
    return seqs, stats


def check(seqs, stats, want):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    assert len(seqs) == want["n_reads"]
    assert stats[1] == want["n_skipped"]
    assert stats[2] is True
    assert int(lens.sum()) == want["n_bases"]
    assert str(Port.fnv1a(np.frombuffer(b"".join(seqs), dtype=np.uint8))) == want["seq_fnv"]
    assert str(Port.fnv1a(lens.view(np.uint8))) == want["len_fnv"]


@pytest.fixture(params=["serial", "parallel"])
def engine(request, monkeypatch):
    """Both parser engines on every case.  The parallel one (uncompressed files read batch-wise) is forced onto
    tiny files with 997-byte chunks, so that most speculative chunk starts are wrong and get re-parsed."""
    if request.param == "serial":
        monkeypatch.setenv("GT_FASTX_THREADS", "1")
    else:
        monkeypatch.setenv("GT_FASTX_THREADS", "4")
        monkeypatch.setenv("GT_FASTX_CHUNK_BYTES", "997")
        monkeypatch.setenv("GT_FASTX_PARALLEL_MIN_BYTES", "0")
    return request.param


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("case", [c for c in cases()], ids=lambda c: c[0])
def test_parser_matches_reference(tmp_path, case, gz, engine):
    from goetia_b200 import parsing
    name, data, min_length, strict = case
    fn = write_case(str(tmp_path), name, data, gz)
    want = GOLD[os.path.basename(fn)]
    if "error" in want:
        exc = parsing.InvalidCharacterException if strict else parsing.InvalidRead
        with pytest.raises(exc):
            parse_all(fn, strict, min_length, max_bases=8 << 20)
        return
    big = want["n_bases"] > (1 << 16)
    seqs, stats = parse_all(fn, strict, min_length, max_bases=(8 << 20) if big else (1 << 16))
    check(seqs, stats, want)
    if not big:
        # tiny batches (records carried over between calls) and the record-at-a-time API give the same stream
        longest = max([len(s) for s in seqs] + [1])
        seqs2, stats2 = parse_all(fn, strict, min_length, max_bases=longest, max_reads=3)
        assert seqs2 == seqs and stats2 == stats
        seqs3, stats3 = parse_all(fn, strict, min_length, by_record=True)
        assert seqs3 == seqs and stats3 == stats


@pytest.mark.parametrize("block", [0xff00, 700])
@pytest.mark.parametrize("case", [c for c in cases()], ids=lambda c: c[0])
def test_bgzf_input_matches_reference(tmp_path, case, block, monkeypatch):
    """bgzip-compressed input: the blocks are inflated in parallel (fx::Bgzf) and the parser must see the bytes zlib's
    gzread hands the reference, so every case gives the reference's golden records (full-size and 700-byte blocks:
    thousands of blocks per ring block, records and lines cut by block boundaries everywhere)."""
    from goetia_b200 import parsing
    from tests.fastx_cases import write_bgzf_case
    monkeypatch.setenv("GT_FASTX_THREADS", "4")
    name, data, min_length, strict = case
    fn = write_bgzf_case(str(tmp_path), name, data, block)
    want = GOLD[os.path.basename(fn)]
    if "error" in want:
        exc = parsing.InvalidCharacterException if strict else parsing.InvalidRead
        with pytest.raises(exc):
            parse_all(fn, strict, min_length, max_bases=8 << 20)
        return
    seqs, stats = parse_all(fn, strict, min_length, max_bases=8 << 20)
    check(seqs, stats, want)
    seqs2, stats2 = parse_all(fn, strict, min_length, max_bases=8 << 20, by_record=want["n_reads"] < 5000)
    assert seqs2 == seqs and stats2 == stats


def test_bgzf_mixed_truncated_and_corrupt_streams(tmp_path, monkeypatch):
    """Where a file stops being well-formed BGZF the reader carries on as zlib's gzread would: a plain gzip member
    appended to BGZF blocks is read, trailing garbage is ignored, a truncated block ends the input after what could be
    inflated, a block with a wrong CRC is a read error.  Compared with the compiled reference parser where it is built."""
    import gzip
    import zlib
    from goetia_b200 import parsing
    from tests.fastx_cases import bgzf_bytes
    monkeypatch.setenv("GT_FASTX_THREADS", "4")
    rng = np.random.default_rng(5)

    def fasta(n, tag):
        return b"".join(b">%s%d\n%s\n" % (tag, i, bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(40, 200))).astype(np.uint8)))
                        for i in range(n))

    a, b = fasta(3000, b"a"), fasta(500, b"b")
    files = {
        "mixed.gz": bgzf_bytes(a, eof_marker=False) + gzip.compress(b, 1) + bgzf_bytes(a[:5000]),
        "garbage_tail.gz": bgzf_bytes(a) + b"this is not gzip",
        "truncated.gz": bgzf_bytes(a)[:-3000],
        "plain_member_first.gz": gzip.compress(b, 1) + bgzf_bytes(a),
    }
    want_text = {"mixed.gz": a + b + a[:5000], "garbage_tail.gz": a, "plain_member_first.gz": b + a}
    try:
        from oracle import binding
        Ref = binding.Ref if binding.have_ref() else None
    except Exception:
        Ref = None
    for name, blob in files.items():
        fn = os.path.join(str(tmp_path), name)
        with open(fn, "wb") as f:
            f.write(blob)
        seqs, stats = parse_all(fn, False, 0, max_bases=8 << 20)
        if name in want_text:
            plain = os.path.join(str(tmp_path), name + ".txt")
            with open(plain, "wb") as f:
                f.write(want_text[name])
            seqs_p, stats_p = parse_all(plain, False, 0, max_bases=8 << 20)
            assert seqs == seqs_p and stats == stats_p, name
        else:  # truncated: a prefix of the records, the last one possibly cut short
            full = [ln for ln in a.split(b"\n") if ln and not ln.startswith(b">")]
            assert 0 < len(seqs) < len(full), name
            assert seqs[:-1] == full[:len(seqs) - 1] and full[len(seqs) - 1].startswith(seqs[-1]), name
        if Ref is not None:
            bb, oo, sk = Ref.parse_file(fn, strict=False, min_length=0)
            ref_seqs = [bb[int(oo[i]):int(oo[i + 1])].tobytes() for i in range(oo.size - 1)]
            assert seqs == ref_seqs and stats[1] == sk, name
    # a corrupted block: CRC mismatch -> the reference's "Error reading stream"
    blob = bytearray(bgzf_bytes(a))
    blob[len(blob) // 2] ^= 0x55
    fn = os.path.join(str(tmp_path), "corrupt.gz")
    with open(fn, "wb") as f:
        f.write(bytes(blob))
    with pytest.raises(parsing.GoetiaFileException):
        parse_all(fn, False, 0, max_bases=8 << 20)


def unpack_batch(words, offsets, flags):
    """The records of a packed batch as text (entries flagged READ_INVALID are alignment gaps, not records)."""
    n = int(offsets[-1])
    codes = ((np.repeat(words, 32)[:n] >> (2 * (np.arange(n, dtype=np.uint64) % np.uint64(32)))) & np.uint64(3)).astype(np.uint8)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    return [text[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(offsets.size - 1) if not flags[i]]


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("case", [c for c in cases()], ids=lambda c: c[0])
def test_packed_batches_match_reference(tmp_path, case, gz, engine):
    """gt_fastx_next_packed_batch (what gt_insert_fastx feeds the device with): unpacked again, the records are the
    reference parser's.  Kept records hold ACGT only (after case folding), so the 2-bit form loses nothing."""
    from goetia_b200 import parsing
    name, data, min_length, strict = case
    fn = write_case(str(tmp_path), name, data, gz)
    want = GOLD[os.path.basename(fn)]
    p = parsing.FastxParser(fn, strict, min_length)
    seqs, n_real = [], 0
    try:
        while True:
            w, o, f, nr = p.next_packed_batch(8 << 20)
            if o.size == 1:
                break
            got = unpack_batch(w, o, f)
            assert len(got) == nr
            seqs.extend(got)
            n_real += nr
        assert "error" not in want
    except (parsing.InvalidRead, parsing.InvalidCharacterException):
        assert "error" in want
        return
    finally:
        stats = (p.n_parsed(), p.n_skipped(), p.is_complete())
        p.close()
    check(seqs, stats, want)


def test_host_pack_scalar_equals_avx2(tmp_path):
    """The portable packer and the AVX2 one give the same words (the choice is made once per process, so the scalar
    run is a second process)."""
    import subprocess
    import sys
    name, data, min_length, strict = [c for c in cases() if c[0] == "reads.fq"][0]
    fn = write_case(str(tmp_path), name, data, False)
    prog = ("import sys; sys.path.insert(0, %r); import numpy as np; from goetia_b200.parsing import FastxParser; "
            "from oracle.binding import Port; p = FastxParser(%r); w, o, f, n = p.next_packed_batch(1 << 20); "
            "print(n, Port.fnv1a(w.view(np.uint8)), Port.fnv1a(o.view(np.uint8)))" % (ROOT, fn))
    outs = []
    for scalar in ("0", "1"):
        env = dict(os.environ, GT_HOST_PACK_SCALAR=scalar, GT_FASTX_THREADS="1")
        outs.append(subprocess.check_output([sys.executable, "-c", prog], env=env).decode().split())
    assert outs[0] == outs[1] and int(outs[0][0]) == GOLD["reads.fq"]["n_reads"]


def test_records_and_errors(tmp_path, golden):
    from goetia_b200 import parsing
    for key in ("_mixed.fa", "_mixed.fq"):
        g = golden["parser"][key]
        fn = os.path.join(str(tmp_path), key)
        with open(fn, "w") as f:
            f.write(g["text"])
        p = parsing.FastxParser.build(fn)
        recs = list(p)
        assert [r.sequence for r in recs] == g["reads"]
        assert p.n_skipped() == g["n_skipped"]
        if key.endswith(".fq"):
            assert recs[0].name == "q1" and recs[0].quality == "IIIIII"
        else:
            assert recs[0].name == "r1" and recs[0].quality == ""
        with pytest.raises(parsing.NoMoreReadsAvailable):  # readers.hh:151-153
            p.next()
    with pytest.raises(parsing.GoetiaFileException):
        parsing.FastxParser(os.path.join(str(tmp_path), "does-not-exist.fa"))
    # a truncated quality string raises InvalidRead and counts as skipped; the parser stays usable (readers.hh:198-201)
    fn = os.path.join(str(tmp_path), "t.fq")
    with open(fn, "w") as f:
        f.write("@a\nACGT\n+\nIIIIII\n@b\nGGCC\n+\nIIII\n")
    p = parsing.FastxParser(fn)
    with pytest.raises(parsing.InvalidRead):
        p.next()
    assert p.n_skipped() == 1
    assert p.next().sequence == "GGCC"


@pytest.mark.gpu
def test_insert_fastx_matches_oracle(tmp_path, gb):
    """FileProcessor<InserterProcessor<dBG>>::process over a gz FASTQ == the oracle fed the parser's records."""
    from goetia_b200.parsing import FastxParser
    from tests.util import make_graph
    name, data, min_length, strict = [c for c in cases() if c[0] == "big.fq"][0]
    fn = write_case(str(tmp_path), name, data, gz=True)
    K, sizes = 31, gb.get_n_primes_near_x(4, 5_000_000)
    for kind in (0, 1, 2):
        g = make_graph(gb, kind, 1, K, sizes)
        os.environ["GT_FASTX_BATCH_BASES"] = str(1 << 20)  # several batches: the parse-ahead thread is exercised
        try:
            n_seqs, n_kmers = g.process_fastx(fn)
        finally:
            os.environ.pop("GT_FASTX_BATCH_BASES", None)
        seqs, _ = parse_all(fn, strict, min_length, max_bases=8 << 20)
        assert n_seqs == len(seqs) == GOLD["big.fq.gz"]["n_reads"]
        ref = Port(kind, 1, K, sizes)
        bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
        offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([len(s) for s in seqs])
        nk_ref, _ = ref.insert_reads(bases, offsets)
        assert n_kmers == nk_ref
        for a, b in zip(g.get_raw(), ref.tables()):
            assert np.array_equal(a, b)
        ref.close()


def test_split_paired_reader(tmp_path):
    """SplitPairedReader (readers.hh:234-340; reference tests/test_parsing.py:68-80): lock-step pairs, skipped mates
    come back as None, unequal files raise, force_name_match drops pairs whose names do not match."""
    from goetia_b200 import parsing
    d = str(tmp_path)
    L = ["@p%d/1\n%s\n+\n%s\n" % (i, "ACGT" * 5 if i != 2 else "ACGN" * 5, "I" * 20) for i in range(5)]
    R = ["@p%d/2\n%s\n+\n%s\n" % (i if i != 3 else 9, "TTGCA" * 4, "I" * 20) for i in range(5)]
    open(d + "/l.fq", "w").write("".join(L))
    open(d + "/r.fq", "w").write("".join(R))
    pairs = list(parsing.SplitPairedReader.build(d + "/l.fq", d + "/r.fq"))
    got = [(a.name if a else None, b.name if b else None) for a, b in pairs if a or b]
    assert got == [("p0/1", "p0/2"), ("p1/1", "p1/2"), (None, "p2/2"), ("p3/1", "p9/2"), ("p4/1", "p4/2")]
    rd = parsing.SplitPairedReader(d + "/l.fq", d + "/r.fq", force_name_match=True)
    kept = [(a.name, b.name) for a, b in rd if a and b]
    assert kept == [("p0/1", "p0/2"), ("p1/1", "p1/2"), ("p4/1", "p4/2")]
    assert rd.n_skipped() == 2 + 1  # the unmatched pair, and the mate with the N
    assert parsing.check_is_pair("r1 1:N:0:1", "r1 2:N:0:1") is False  # kseq names stop at the first blank
    assert parsing.check_is_pair("r/1", "r/2") and not parsing.check_is_pair("/1", "/2")
    open(d + "/short.fq", "w").write("".join(R[:3]))
    bad = parsing.SplitPairedReader(d + "/l.fq", d + "/short.fq")
    with pytest.raises(parsing.GoetiaB200Error):
        list(bad)


def _ref_or_skip():
    from oracle import binding
    if not binding.have_ref():
        pytest.skip("oracle/_ref (the compiled reference) is not built here")
    return binding.Ref


def test_parser_fuzz_against_live_reference(tmp_path, engine, monkeypatch):
    """Random FASTA / FASTQ-like text -- stray '@' '+' '>' at line starts, blank lines, CR LF, tabs, lower case,
    foreign symbols, truncated tails -- through the product parser and through the compiled reference parser:
    same kept sequences, same n_skipped, same error-or-not.  (Runs where oracle/_ref exists.)"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from goetia_b200 import parsing
    Ref = _ref_or_skip()
    fn = os.path.join(str(tmp_path), "fuzz.fx")
    if engine == "parallel":
        monkeypatch.setenv("GT_FASTX_CHUNK_BYTES", "64")  # several speculative chunks even in these tiny files

    line = st.text(alphabet="ACGTacgtN@+> \t", min_size=0, max_size=12)
    seqline = st.text(alphabet="ACGTacgtn", min_size=0, max_size=30)
    rec_fa = st.tuples(st.just(">"), line, st.lists(seqline, min_size=0, max_size=3))
    rec_fq = st.tuples(st.just("@"), line, st.lists(seqline, min_size=1, max_size=2), st.text(alphabet="I#@+>5", min_size=0, max_size=40))

    def render(recs, eol, tail):
        out = []
        for r in recs:
            if r[0] == ">":
                out.append(">" + r[1] + eol + "".join(s + eol for s in r[2]))
            else:
                seq = "".join(r[2])
                qual = (r[3] * 40)[:len(seq)] if len(r[3]) % 3 else r[3]  # mostly matching lengths, sometimes not
                out.append("@" + r[1] + eol + "".join(s + eol for s in r[2]) + "+" + eol + qual + eol)
        text = "".join(out)
        return text[:len(text) - tail] if tail else text

    @settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(st.one_of(rec_fa, rec_fq), min_size=0, max_size=8), st.sampled_from(["\n", "\r\n"]),
           st.integers(0, 5), st.integers(0, 12))
    def run(recs, eol, tail, min_length):
        with open(fn, "w", newline="") as f:
            f.write(render(recs, eol, tail))
        try:
            bb, oo, sk = Ref.parse_file(fn, strict=False, min_length=min_length)
            want = ([bb[int(oo[i]):int(oo[i + 1])].tobytes() for i in range(oo.size - 1)], sk)
        except ValueError:
            want = None
        try:
            seqs, stats = parse_all(fn, False, min_length, max_bases=1 << 12)
            got = (seqs, stats[1])
        except parsing.InvalidRead:
            got = None
        assert got == want

    run()
