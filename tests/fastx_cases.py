"""Deterministic FASTA / FASTQ files exercising the kseq grammar and FastxParser's skip rules.

Used twice: tests/golden/make_fastx_golden.py feeds them to the UNMODIFIED reference parser
(oracle/_ref) and records what it returns; tests/test_fastx.py feeds the same bytes to the product
parser (libgoetia_b200.so) and compares.
"""
import gzip
import os

import numpy as np


def _seq(rng, n, alphabet=b"ACGT"):
    return np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].tobytes()


def _wrap(s, width, eol=b"\n"):
    return eol.join(s[i:i + width] for i in range(0, len(s), width)) + eol if s else eol


def cases():
    """-> list of (file name, bytes, min_length, strict)"""
    out = []
    rng = np.random.default_rng(20191021)
    # 1. plain single-line FASTA with comments, lower case, Ns, an empty record, no trailing newline
    t = b">r1 first read\nACGTACGTAC\n>r2\nacgtacgtacgt\n>r3 has N\nACGTNNACGT\n>empty\n>r5\nGGGG\n>r6\tTAB comment\nTTTTACGT"
    out.append(("simple.fa", t, 0, False))
    out.append(("simple_min5.fa", t, 5, False))
    # 2. multi-line FASTA, blank lines inside, CRLF line ends
    recs = []
    for i in range(40):
        s = _seq(rng, int(rng.integers(1, 400)))
        recs.append(b">m%d len=%d\r\n" % (i, len(s)) + _wrap(s, 60, b"\r\n") + (b"\r\n" if i % 7 == 0 else b""))
    out.append(("multiline_crlf.fa", b"".join(recs), 0, False))
    # 3. FASTQ, 4-line, with '@' and '+' and '>' at the start of quality lines, mixed case, Ns
    recs = []
    for i in range(60):
        n = int(rng.integers(1, 200))
        s = _seq(rng, n, b"ACGTacgt" if i % 5 else b"ACGTN")
        q = bytearray(_seq(rng, n, b"IJ#5@+>"))
        if i % 3 == 0:
            q[0] = ord("@")
        if i % 4 == 0:
            q[0] = ord("+")
        recs.append(b"@q%d/1 comment\n" % i + s + b"\n+" + (b"q%d" % i if i % 2 else b"") + b"\n" + bytes(q) + b"\n")
    out.append(("reads.fq", b"".join(recs), 0, False))
    out.append(("reads_min50.fq", b"".join(recs), 50, False))
    # 4. multi-line FASTQ (sequence and quality wrapped)
    recs = []
    for i in range(25):
        n = int(rng.integers(50, 500))
        s, q = _seq(rng, n), _seq(rng, n, b"FGHI")
        recs.append(b"@w%d\n" % i + _wrap(s, 70) + b"+\n" + _wrap(q, 70))
    out.append(("wrapped.fq", b"".join(recs), 0, False))
    # 5. junk before the first header, header-only file end, white space in sequences
    out.append(("junk.fa", b"garbage line\n\n>a\nACGT\nAC GT\n>b\nACGT\n>c", 0, False))
    # 6. truncated quality (kseq returns -2 -> InvalidRead)
    out.append(("truncated.fq", b"@a\nACGTACGT\n+\nIIIIIIII\n@b\nACGTACGT\n+\nIII\n", 0, False))
    # 7. strict parser hits a foreign symbol
    out.append(("strict.fa", b">a\nACGT\n>b\nACNT\n>c\nACGT\n", 0, True))
    # 8. records and lines crossing the 4 MiB read blocks; one 5 Mbp single-line record
    recs = []
    for i in range(30000):
        s = _seq(rng, 150, b"ACGT" if i % 97 else b"ACGTN")
        recs.append(b"@big%d\n" % i + s + b"\n+\n" + b"I" * 150 + b"\n")
    out.append(("big.fq", b"".join(recs), 0, False))
    long = _seq(rng, 5_000_000)
    out.append(("long.fa", b">chr1\n" + long + b"\n>chr2\n" + _wrap(long[:1_000_000], 80) + b">chr3\nACGT\n", 0, False))
    # 9. empty file, header-less file
    out.append(("empty.fa", b"", 0, False))
    out.append(("noheader.fa", b"ACGT\nACGT\n", 0, False))
    return out


def write_case(dirname, name, data, gz=False):
    fn = os.path.join(dirname, name + (".gz" if gz else ""))
    if gz:
        with gzip.open(fn, "wb", compresslevel=1) as f:
            f.write(data)
    else:
        with open(fn, "wb") as f:
            f.write(data)
    return fn


def bgzf_bytes(data, block=0xff00, level=1, eof_marker=True):
    """`data` as a BGZF stream (the bgzip / htslib container): gzip members of <= 64 KiB whose extra field 'BC' holds
    the member's size - 1, closed by the empty end-of-file block.  zlib's gzread -- and so the reference parser --
    reads it as one ordinary multi-member gzip stream."""
    import struct
    import zlib
    out = []
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_marker else [])
    for c in chunks:
        z = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = z.compress(c) + z.flush()
        bsize = 12 + 6 + len(comp) + 8 - 1
        assert bsize < 65536
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) +
                   comp + struct.pack("<II", zlib.crc32(c) & 0xffffffff, len(c) & 0xffffffff))
    return b"".join(out)


def write_bgzf_case(dirname, name, data, block=0xff00):
    fn = os.path.join(dirname, name + ".gz")
    with open(fn, "wb") as f:
        f.write(bgzf_bytes(data, block))
    return fn

