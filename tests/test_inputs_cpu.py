"""CPU: the DNA input kinds of the `fasim` command line (plain FASTA as the reference's readDna parses it, gzip FASTA, UCSC
.2bit with regions) through `fasim --list-records`, which needs no GPU."""
import gzip
import os
import struct
import zlib

import pytest

import fasim_b200 as fb
from _harness import splitmix_bases


def write_twobit(path, seqs, byteswap=False):
    """seqs: list of (name, sequence with ACGTN, upper or lower case).  UCSC .2bit version 0."""
    e = ">" if byteswap else "<"
    code = {"T": 0, "C": 1, "A": 2, "G": 3, "N": 0}
    recs = []
    for name, s in seqs:
        up = s.upper()
        nblocks, mblocks = [], []
        for blocks, pred in ((nblocks, lambda c: c in "Nn"), (mblocks, lambda c: c.islower())):
            i = 0
            while i < len(s):
                if pred(s[i]):
                    j = i
                    while j < len(s) and pred(s[j]):
                        j += 1
                    blocks.append((i, j - i))
                    i = j
                else:
                    i += 1
        packed = bytearray((len(up) + 3) // 4)
        for i, c in enumerate(up):
            packed[i // 4] |= code[c] << (6 - 2 * (i % 4))
        body = struct.pack(e + "II", len(up), len(nblocks))
        body += b"".join(struct.pack(e + "I", a) for a, _ in nblocks) + b"".join(struct.pack(e + "I", n) for _, n in nblocks)
        body += struct.pack(e + "I", len(mblocks))
        body += b"".join(struct.pack(e + "I", a) for a, _ in mblocks) + b"".join(struct.pack(e + "I", n) for _, n in mblocks)
        body += struct.pack(e + "I", 0) + bytes(packed)
        recs.append(body)
    header = struct.pack(e + "IIII", 0x1A412743, 0, len(seqs), 0)
    index_size = sum(1 + len(n) + 4 for n, _ in seqs)
    off = len(header) + index_size
    index = b""
    for (name, _), body in zip(seqs, recs):
        index += struct.pack("B", len(name)) + name.encode() + struct.pack(e + "I", off)
        off += len(body)
    open(path, "wb").write(header + index + b"".join(recs))


def list_records(args, cwd, rna=None):
    if rna is None:
        rna = "_q.fa"
        open(os.path.join(cwd, rna), "w").write(">q\nACGT\n")
    r = fb.run_cli(list(args) + ["-f2", rna, "--list-records"], cwd=cwd)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [l.split("\t")[1:] for l in r.stdout.splitlines() if l.startswith("record\t")]
    return [(sp, ch, int(st), int(n), crc) for sp, ch, st, n, crc in rows]


def list_queries(args, cwd):
    r = fb.run_cli(list(args) + ["--list-records"], cwd=cwd)
    assert r.returncode == 0, r.stdout + r.stderr
    return [tuple(l.split("\t")[1:]) for l in r.stdout.splitlines() if l.startswith("query\t")]


def crc(s):
    return "%08x" % (zlib.crc32(s.encode()) & 0xFFFFFFFF)


@pytest.fixture(scope="module")
def seqs():
    a = list(splitmix_bases(7001, 10007))
    for lo, hi in ((0, 13), (4999, 5003), (9000, 9400)):
        a[lo:hi] = "N" * (hi - lo)
    a = "".join(a)
    a = a[:2000] + a[2000:2600].lower() + a[2600:]            # a soft-masked stretch (lower case in the .2bit's FASTA view)
    return [("chr1", a), ("chrUn_x", splitmix_bases(7002, 37)), ("chr2", splitmix_bases(7003, 6001))]


def test_fasta_plain_and_gzip_give_the_same_records(tmp_path, seqs):
    fb.build()
    d = str(tmp_path)
    text = "".join(">hg|%s|%d-%d\n%s\n" % (n, 101 + k, 100 + k + len(s), "\n".join(s.upper()[i:i + 60] for i in range(0, len(s), 60)))
                   for k, (n, s) in enumerate(seqs))
    open(os.path.join(d, "x.fa"), "w").write(text)
    with gzip.open(os.path.join(d, "x.fa.gz"), "wt") as f:
        f.write(text.replace("\n", "\r\n"))                     # CR/LF line ends are stripped like the reference does
    expect = [("hg", n, 101 + k, len(s), crc(s.upper())) for k, (n, s) in enumerate(seqs)]
    assert list_records(["-f1", "x.fa"], d) == expect
    assert list_records(["-f1", "x.fa.gz"], d) == expect


@pytest.mark.parametrize("byteswap", [False, True])
def test_twobit_whole_sequences_and_regions(tmp_path, seqs, byteswap):
    fb.build()
    d = str(tmp_path)
    write_twobit(os.path.join(d, "mini.2bit"), seqs, byteswap)
    whole = list_records(["-f1", "mini.2bit"], d)
    assert whole == [("mini", n, 1, len(s), crc(s.upper())) for n, s in seqs]
    a = seqs[0][1].upper()
    got = list_records(["-f1", "mini.2bit", "--seq", "chr1:5-5010,chr2", "--seq", "chr1:9399-20000", "--species", "hg38"], d)
    assert got == [("hg38", "chr1", 5, 5006, crc(a[4:5010])), ("hg38", "chr2", 1, 6001, crc(seqs[2][1])),
                   ("hg38", "chr1", 9399, len(a) - 9398, crc(a[9398:]))]
    r = fb.run_cli(["-f1", "mini.2bit", "--seq", "chrZ", "--list-records"], cwd=d)
    assert r.returncode != 0 and "chrZ" in r.stderr


def test_rna_readers(tmp_path):
    """-f2 as the reference reads it (first line = name with every '>' removed, every other line appended: readRna,
    Fasim-LongTarget.cpp:174-200) and with --queries (one lncRNA per '>' record, empty records dropped)."""
    fb.build()
    d = str(tmp_path)
    open(os.path.join(d, "dna.fa"), "w").write(">sp|chr1|1-8\nACGTACGT\n")
    a, b = splitmix_bases(4001, 150), splitmix_bases(4002, 77)
    open(os.path.join(d, "two.fa"), "w").write(">lncA extra>words\r\n%s\r\n%s\n>lncB\n%s\n>empty\n\n" % (a[:60], a[60:], b))
    multi = list_queries(["-f1", "dna.fa", "-f2", "two.fa", "--queries"], d)
    assert multi == [("lncA extrawords", "150", crc(a)), ("lncB", "77", crc(b))]
    # without --queries the file is ONE lncRNA, later header lines and all (what the reference does with such a file)
    single = list_queries(["-f1", "dna.fa", "-f2", "two.fa"], d)
    assert single == [("lncA extrawords", str(150 + len(">lncB") + 77 + len(">empty")), crc(a + ">lncB" + b + ">empty"))]
    r = fb.run_cli(["-f1", "dna.fa", "-f2", "missing.fa", "--list-records"], cwd=d)
    assert r.returncode != 0 and "missing.fa" in r.stderr


def test_softmask_policies_and_corrupt_twobit(tmp_path, seqs):
    """--softmask: lower-case (soft-masked) bases are upper-cased (default, with a notice), scored as N, or refused — in FASTA and
    in .2bit (mask blocks) alike.  A .2bit file whose counts do not fit the file is reported as corrupt, not allocated for."""
    fb.build()
    d = str(tmp_path)
    a = seqs[0][1]                                             # holds a lower-case stretch and N runs
    open(os.path.join(d, "m.fa"), "w").write(">hg|chr1|1-%d\n%s\n" % (len(a), a))
    write_twobit(os.path.join(d, "m.2bit"), [("chr1", a)])
    as_n = "".join("N" if c.islower() else c for c in a)
    for f1 in ("m.fa", "m.2bit"):
        sp = "hg" if f1 == "m.fa" else "m"
        assert list_records(["-f1", f1], d) == [(sp, "chr1", 1, len(a), crc(a.upper()))]
        assert list_records(["-f1", f1, "--softmask", "n"], d) == [(sp, "chr1", 1, len(a), crc(as_n))]
        r = fb.run_cli(["-f1", f1, "-f2", "_q.fa", "--list-records", "--softmask", "error"], cwd=d)
        assert r.returncode != 0 and "soft-masked" in r.stderr
        r = fb.run_cli(["-f1", f1, "-f2", "_q.fa", "--list-records"], cwd=d)
        assert "upper-cased" in r.stderr
    # a region that does not start on a byte boundary of the packed data
    assert list_records(["-f1", "m.2bit", "--seq", "chr1:7-3001", "--softmask", "n"], d) == [("m", "chr1", 7, 2995, crc(as_n[6:3001]))]
    raw = bytearray(open(os.path.join(d, "m.2bit"), "rb").read())
    off = struct.unpack("<I", raw[16 + 1 + 4:16 + 1 + 4 + 4])[0]                     # header 16, index entry: size byte, "chr1", offset
    bad = bytearray(raw)
    bad[off + 4:off + 8] = struct.pack("<I", 0x7FFFFFF0)                            # nBlockCount
    open(os.path.join(d, "bad.2bit"), "wb").write(bytes(bad))
    r = fb.run_cli(["-f1", "bad.2bit", "-f2", "_q.fa", "--list-records"], cwd=d)
    assert r.returncode != 0 and "corrupt" in r.stderr
    bad = bytearray(raw)
    bad[8:12] = struct.pack("<I", 0x7FFFFFFF)                                        # sequenceCount
    open(os.path.join(d, "bad2.2bit"), "wb").write(bytes(bad))
    r = fb.run_cli(["-f1", "bad2.2bit", "-f2", "_q.fa", "--list-records"], cwd=d)
    assert r.returncode != 0 and "corrupt" in r.stderr
