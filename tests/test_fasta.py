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
