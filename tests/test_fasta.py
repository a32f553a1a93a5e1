"""CPU: parse_fasta semantics (reference src/parse.rs:54-99), pinned by the reference's own four
tests (parse.rs:166-251), for both the oracle restatement and the product's host parser."""
import os

import pytest


def _write(tmp_path, name, data: bytes):
    p = tmp_path / name
    p.write_bytes(data)
    return str(p)


def _parsers(oracle):
    from sequencealigning_b200 import parse_fasta

    def product(path):
        try:
            recs, err = parse_fasta(path)
        except ValueError:
            return None
        return [r.name for r in recs], [r.seq for r in recs], err

    def orc(path):
        r = oracle.parse_fasta_path(path)
        return None if r is None else (r.names, r.seqs, r.err_chars)

    return [product, orc]


def test_parse_good_fasta(tmp_path, oracle):  # parse.rs:167-186
    path = _write(tmp_path, "a.fa", b">Record1\nATGCATGCATGCATGCATGCATGCATGC\n>Record2\nATGCATGCGTGCAGTGACCACA")
    for parse in _parsers(oracle):
        names, seqs, err = parse(path)
        assert len(names) == 2 and len(names[0]) == 8 and len(seqs[0]) == 28 and err == b""
        assert names[1] == b">Record2" and seqs[1] == b"ATGCATGCGTGCAGTGACCACA"


def test_parse_bad_header(tmp_path, oracle):  # parse.rs:189-215
    path = _write(tmp_path, "b.fa", b">Record1\nATGCATGCATGCATGCATGCATGCATGC\nRecord2\nATGCATGCGTGCAGTGACCACA")
    for parse in _parsers(oracle):
        names, seqs, err = parse(path)
        assert err == b"Record2"
        assert names == [b">Record1"]
        assert seqs == [b"ATGCATGCATGCATGCATGCATGCATGCATGCATGCGTGCAGTGACCACA"]


def test_parse_bad_nt(tmp_path, oracle):  # parse.rs:218-238
    path = _write(tmp_path, "c.fa", b">Record1\nATGCATGCAKGCATGCATGCANNNGCATGC")
    for parse in _parsers(oracle):
        names, seqs, err = parse(path)
        assert err == b"K" and names == [b">Record1"] and seqs == [b"ATGCATGCAGCATGCATGCANNNGCATGC"]


def test_parse_false_file(tmp_path, oracle):  # parse.rs:241-251
    path = _write(tmp_path, "d.txt", b">x\nACGT\n")
    for parse in _parsers(oracle):
        assert parse(path) is None
    for parse in _parsers(oracle):
        assert parse(str(tmp_path / "missing.fa")) is None


@pytest.mark.parametrize("ext", ["fa", "fasta", "fna"])
def test_quirks(tmp_path, oracle, ext):
    # text before the first '>' lands in the default record that parse.rs:91 removes; '\r' and
    # lowercase are rejected characters; a header at EOF yields an empty sequence
    path = _write(tmp_path, f"e.{ext}", b"ACGT\n>r1 desc\r\nAC\r\nacGT\n>r2")
    for parse in _parsers(oracle):
        names, seqs, err = parse(path)
        assert names == [b">r1 desc\r", b">r2"]
        assert seqs == [b"ACGT", b""]
        assert err == b"\rac"
    empty = _write(tmp_path, f"f.{ext}", b"")
    for parse in _parsers(oracle):
        assert parse(empty) == ([], [], b"")


def test_large_file_takes_the_parallel_path_and_agrees_with_the_oracle(tmp_path, oracle):
    """Files above 1 MB are cut at '>' bytes and parsed by several threads; the quirks must
    survive the cuts: '>' inside header lines and inside sequence lines, bytes outside ACGTN
    (reported in order), CRLF, blank lines, residues before the first header, no final newline."""
    import random
    rng = random.Random(5)
    parts = [b"ACGTNN\nGG\n"]  # default record: dropped (parse.rs:91)
    for i in range(30000):
        name = b">r%d" % i
        k = rng.random()
        if k < 0.02:
            name += b" with > inside"        # starts another record (a header of its own)
        elif k < 0.04:
            name += b"\r"
        parts.append(name + b"\n")
        for _ in range(rng.randint(0, 3)):
            line = bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(0, 70)))
            if rng.random() < 0.03:
                pos = rng.randint(0, len(line))
                line = line[:pos] + rng.choice([b"x", b"K", b"\r", b" ", b">mid"]) + line[pos:]
            parts.append(line + (b"\n" if rng.random() < 0.98 else b"\n\n"))
    parts.append(b">last\nACGT")
    data = b"".join(parts)
    assert len(data) > (1 << 20)
    path = _write(tmp_path, "big.fasta", data)
    product, orc = _parsers(oracle)
    a, b = product(path), orc(path)
    assert a[0] == b[0]
    assert a[1] == b[1]
    assert a[2] == b[2] and len(a[2]) > 100
    assert len(a[0]) > 30000


def test_fused_parse_and_pack(tmp_path, oracle):
    """sa_parse_fasta_packed: the same records as parse_fasta, plus the 2-bit image of the output
    buffer; the record offsets address both formats; 'N' clears all_acgt (then the byte format is
    what a batch must use)."""
    import random
    import numpy as np
    from sequencealigning_b200.engine import PairBatch, parse_fasta, parse_fasta_packed
    rng = random.Random(11)
    parts = []
    for i in range(40000):   # > 1 MB: the chunk-parallel path, pack slices on 4-residue boundaries
        parts.append(b">read%d\n" % i)
        parts.append(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 90))) + b"\n")
    data = b"".join(parts)
    assert len(data) > (1 << 20)
    path = _write(tmp_path, "reads.fa", data)
    recs, err, out, packed, index, all_acgt = parse_fasta_packed(path)
    plain, err2 = parse_fasta(path)
    assert [r.seq for r in recs] == [r.seq for r in plain] and [r.name for r in recs] == [r.name for r in plain]
    assert err == err2 == b"" and all_acgt
    # a batch on the packed image with the index's offsets decodes to the same sequences
    so, sl = index[:, 2].copy(), index[:, 3].astype(np.uint32)
    pb = PairBatch(packed, so, sl, so, sl, packing=1)
    for r in list(range(0, len(recs), 997)) + [len(recs) - 1]:
        assert pb.query(r) == recs[r].seq
    # the image is the packer's output on the whole buffer (names code as 0)
    codes = np.zeros(256, np.uint8)
    codes[ord("C")], codes[ord("G")], codes[ord("T")] = 1, 2, 3
    c = codes[out]
    pad = np.concatenate([c, np.zeros((-len(c)) % 4, np.uint8)]).reshape(-1, 4)
    expect = (pad[:, 0] | (pad[:, 1] << 2) | (pad[:, 2] << 4) | (pad[:, 3] << 6)).astype(np.uint8)
    assert np.array_equal(packed, expect)
    # an 'N' in a sequence: still parsed, but not packable
    path_n = _write(tmp_path, "n.fa", b">a\nACGT\n>b\nACNT\n")
    recs, err, out, packed, index, all_acgt = parse_fasta_packed(path_n)
    assert [r.seq for r in recs] == [b"ACGT", b"ACNT"] and not all_acgt
    # an 'N' in a header does not matter
    assert parse_fasta_packed(_write(tmp_path, "h.fa", b">Nome\nACGT\n"))[5]


def test_multithreaded_packer_matches_the_scalar_one():
    import numpy as np
    from sequencealigning_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(3)
    for n in (0, 1, 3, 4, 5, 17, 4099, (1 << 21) + 3):
        src = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
        a = np.zeros((n + 3) // 4 + 1, np.uint8)
        b = np.zeros_like(a)
        assert lib.sa_pack_2bit(src.ctypes.data, n, a.ctypes.data, 0) == 0
        for threads in (1, 3, 0):
            b[:] = 0
            assert lib.sa_pack_2bit_mt(src.ctypes.data, n, b.ctypes.data, threads) == 0
            assert np.array_equal(a, b), (n, threads)
        if n > 4:
            bad = src.copy()
            bad[n // 2] = ord("N")
            assert lib.sa_pack_2bit_mt(bad.ctypes.data, n, b.ctypes.data, 0) == -2
