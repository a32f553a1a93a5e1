"""ctypes binding of oracle/build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")

OK, REF_PANIC, REF_NO_CONVERGENCE, NOT_IMPLEMENTED, REF_PANIC_EARLY, REF_NO_OUTPUT = range(6)
OP_M, OP_I, OP_D = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc via oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if force else []), check=True)
    return _LIB_PATH


class Scheme(C.Structure):
    _fields_ = [("match_", C.c_int32), ("mismatch", C.c_int32), ("gap_opening", C.c_int32), ("gap_extension", C.c_int32)]


class _AffineResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("score", C.c_int32),
        ("end_m", C.c_int32),
        ("end_i", C.c_int32),
        ("end_d", C.c_int32),
        ("any_panic", C.c_int32),
        ("n_cooptimal", C.c_int64),
        ("cigar_len", C.c_uint32),
        ("n_columns", C.c_uint32),
    ]


class _LinearResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("score", C.c_int32),
        ("n_hits", C.c_int64),
        ("cigar_len", C.c_uint32),
        ("n_columns", C.c_uint32),
        ("start1", C.c_uint32),
        ("start2", C.c_uint32),
        ("end1", C.c_uint32),
        ("end2", C.c_uint32),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        u8p, u32p, u64p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32))
        _lib.sao_affine_align.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(Scheme), C.POINTER(_AffineResult), u32p]
        _lib.sao_affine_align.restype = C.c_int
        _lib.sao_affine_score.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(Scheme)]
        _lib.sao_affine_score.restype = C.c_int32
        _lib.sao_affine_print_all.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(Scheme), C.c_uint64, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
        _lib.sao_affine_print_all.restype = C.c_int64
        _lib.sao_affine_matrices.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(Scheme), i32p, i32p, i32p, u8p]
        _lib.sao_affine_matrices.restype = C.c_int
        _lib.sao_affine_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Scheme), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        _lib.sao_affine_batch.restype = C.c_int
        _lib.sao_linear_align.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.POINTER(_LinearResult), u32p]
        _lib.sao_linear_align.restype = C.c_int
        _lib.sao_linear_matrices.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, i32p, u8p, u8p]
        _lib.sao_linear_matrices.restype = C.c_int
        _lib.sao_linear_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        _lib.sao_linear_batch.restype = C.c_int
        _lib.sao_linear_batch_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        _lib.sao_linear_batch_ex.restype = C.c_int
        _lib.sao_linear_print_hits.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.c_uint64, C.c_char_p, C.c_size_t,
                                               C.POINTER(C.c_uint64)]
        _lib.sao_linear_print_hits.restype = C.c_int64
        _lib.sao_parse_fasta_mem.argtypes = [C.c_char_p, C.c_size_t, u8p, C.c_size_t, u64p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib.sao_parse_fasta_mem.restype = C.c_int64
        _lib.sao_parse_fasta_path.argtypes = [C.c_char_p, u8p, C.c_size_t, u64p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib.sao_parse_fasta_path.restype = C.c_int64
        if hasattr(_lib, "sao_wfa_literal"):
            _bind_wfa(_lib)
    return _lib


def _bind_wfa(l):  # filled in by oracle/wfa.c's section below
    u8p, u32p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32))
    l.sao_wfa_literal.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(_WfaResult)]
    l.sao_wfa_literal.restype = C.c_int
    l.sao_wfa_gotoh_cost.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32]
    l.sao_wfa_gotoh_cost.restype = C.c_int64
    l.sao_wfa_standard.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32]
    l.sao_wfa_standard.restype = C.c_int64
    l.sao_wfa_literal_ex.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(_WfaResult), i32p, C.c_uint32, u32p, i32p]
    l.sao_wfa_literal_ex.restype = C.c_int
    l.sao_wfa_print.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_int32)]
    l.sao_wfa_print.restype = C.c_int64
    l.sao_wfa_literal_cap.argtypes = [C.c_uint32, C.c_uint32]
    l.sao_wfa_literal_cap.restype = C.c_uint32
    l.sao_wfa_literal_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    l.sao_wfa_literal_batch.restype = C.c_int


class _WfaResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("printed_score", C.c_int32),
        ("panic_line", C.c_int32),
        ("n_wavefronts", C.c_int32),
    ]


@dataclass
class AffineResult:
    status: int
    score: int
    end_m: int
    end_i: int
    end_d: int
    any_panic: bool
    n_cooptimal: int
    cigar: List[int]
    n_columns: int


def _scheme(s) -> Optional[Scheme]:
    if s is None:
        return None
    if isinstance(s, Scheme):
        return s
    return Scheme(*s)


def affine_align(seq1: bytes, seq2: bytes, scheme=None) -> AffineResult:
    r = _AffineResult()
    cig = (C.c_uint32 * (len(seq1) + len(seq2) + 1))()
    sc = _scheme(scheme)
    rc = lib().sao_affine_align(seq1, len(seq1), seq2, len(seq2), C.byref(sc) if sc else None, C.byref(r), cig)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return AffineResult(r.status, r.score, r.end_m, r.end_i, r.end_d, bool(r.any_panic), r.n_cooptimal, list(cig[: r.cigar_len]), r.n_columns)


def affine_score(seq1: bytes, seq2: bytes, scheme=None) -> int:
    sc = _scheme(scheme)
    return lib().sao_affine_score(seq1, len(seq1), seq2, len(seq2), C.byref(sc) if sc else None)


def affine_print_all(seq1: bytes, seq2: bytes, scheme=None, max_alignments: int = 1 << 20) -> Tuple[str, int, bool]:
    sc = _scheme(scheme)
    n = C.c_uint64()
    pan = C.c_int32()
    need = lib().sao_affine_print_all(seq1, len(seq1), seq2, len(seq2), C.byref(sc) if sc else None, max_alignments, None, 0, C.byref(n), C.byref(pan))
    buf = C.create_string_buffer(need + 1)
    lib().sao_affine_print_all(seq1, len(seq1), seq2, len(seq2), C.byref(sc) if sc else None, max_alignments, buf, need + 1, C.byref(n), C.byref(pan))
    return buf.value.decode("latin1"), n.value, bool(pan.value)


def affine_matrices(seq1: bytes, seq2: bytes, scheme=None):
    n1, n2 = len(seq1), len(seq2)
    shape = (n2 + 1, n1 + 1)
    m = np.zeros(shape, np.int32)
    i = np.zeros(shape, np.int32)
    d = np.zeros(shape, np.int32)
    par = np.zeros(shape, np.uint8)
    sc = _scheme(scheme)
    i32p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib().sao_affine_matrices(seq1, n1, seq2, n2, C.byref(sc) if sc else None, m.ctypes.data_as(i32p), i.ctypes.data_as(i32p), d.ctypes.data_as(i32p), par.ctypes.data_as(u8p))
    return m, i, d, par


@dataclass
class BatchResult:
    score: np.ndarray
    status: np.ndarray
    cigar_len: np.ndarray
    cigar_pool: Optional[np.ndarray]  # [n_pairs, stride]

    def cigar(self, p: int) -> List[int]:
        return [int(v) for v in self.cigar_pool[p, : self.cigar_len[p]]]


def _batch_args(residues, q_off, q_len, d_off, d_len):
    residues = np.ascontiguousarray(residues, np.uint8)
    q_off = np.ascontiguousarray(q_off, np.uint64)
    d_off = np.ascontiguousarray(d_off, np.uint64)
    q_len = np.ascontiguousarray(q_len, np.uint32)
    d_len = np.ascontiguousarray(d_len, np.uint32)
    return residues, q_off, q_len, d_off, d_len


def affine_batch(residues, q_off, q_len, d_off, d_len, scheme=None, cigar_stride: int = 0, n_threads: int = 1) -> BatchResult:
    residues, q_off, q_len, d_off, d_len = _batch_args(residues, q_off, q_len, d_off, d_len)
    n = len(q_len)
    score = np.zeros(n, np.int32)
    status = np.zeros(n, np.uint8)
    clen = np.zeros(n, np.uint32)
    pool = np.zeros((n, cigar_stride), np.uint32) if cigar_stride else None
    sc = _scheme(scheme)
    rc = lib().sao_affine_batch(
        residues.ctypes.data, q_off.ctypes.data, q_len.ctypes.data, d_off.ctypes.data, d_len.ctypes.data, n,
        C.byref(sc) if sc else None, score.ctypes.data, status.ctypes.data, clen.ctypes.data,
        pool.ctypes.data if pool is not None else None, cigar_stride, n_threads)
    if rc != 0:
        raise MemoryError("oracle batch failed")
    return BatchResult(score, status, clen, pool)


@dataclass
class LinearResult:
    status: int
    score: int
    n_hits: int
    cigar: List[int]
    n_columns: int
    start1: int
    start2: int
    end1: int = 0
    end2: int = 0


def linear_align(seq1: bytes, seq2: bytes, local: bool = False) -> LinearResult:
    r = _LinearResult()
    cig = (C.c_uint32 * (len(seq1) + len(seq2) + 1))()
    rc = lib().sao_linear_align(seq1, len(seq1), seq2, len(seq2), int(local), C.byref(r), cig)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return LinearResult(r.status, r.score, r.n_hits, list(cig[: r.cigar_len]), r.n_columns, r.start1, r.start2, r.end1, r.end2)


def linear_matrices(seq1: bytes, seq2: bytes, local: bool = False):
    n1, n2 = len(seq1), len(seq2)
    shape = (n1 + 1, n2 + 1)
    s = np.zeros(shape, np.int32)
    mv = np.zeros(shape, np.uint8)
    g = np.zeros(shape, np.uint8)
    lib().sao_linear_matrices(seq1, n1, seq2, n2, int(local), s.ctypes.data_as(C.POINTER(C.c_int32)), mv.ctypes.data_as(C.POINTER(C.c_uint8)), g.ctypes.data_as(C.POINTER(C.c_uint8)))
    return s, mv, g


def linear_batch(residues, q_off, q_len, d_off, d_len, cigar_stride: int = 0, n_threads: int = 1, local: bool = False) -> BatchResult:
    residues, q_off, q_len, d_off, d_len = _batch_args(residues, q_off, q_len, d_off, d_len)
    n = len(q_len)
    score = np.zeros(n, np.int32)
    status = np.zeros(n, np.uint8)
    clen = np.zeros(n, np.uint32)
    end1 = np.zeros(n, np.uint32)
    end2 = np.zeros(n, np.uint32)
    pool = np.zeros((n, cigar_stride), np.uint32) if cigar_stride else None
    rc = lib().sao_linear_batch_ex(
        residues.ctypes.data, q_off.ctypes.data, q_len.ctypes.data, d_off.ctypes.data, d_len.ctypes.data, n, int(local),
        score.ctypes.data, status.ctypes.data, clen.ctypes.data, pool.ctypes.data if pool is not None else None, cigar_stride,
        end1.ctypes.data, end2.ctypes.data, n_threads)
    if rc != 0:
        raise MemoryError("oracle batch failed")
    out = BatchResult(score, status, clen, pool)
    out.end1, out.end2 = end1, end2
    return out


def linear_print_hits(seq1: bytes, seq2: bytes, local: bool = False, max_hits: int = 1 << 20):
    """The reference's stdout for the hits of one pair (needleman_wunsch.rs:106-116, :155-178, :205-254).
    Returns (text, n_printed)."""
    n = C.c_uint64()
    need = lib().sao_linear_print_hits(seq1, len(seq1), seq2, len(seq2), int(local), max_hits, None, 0, C.byref(n))
    if need < 0:
        raise MemoryError("oracle allocation failed")
    buf = C.create_string_buffer(need + 1)
    lib().sao_linear_print_hits(seq1, len(seq1), seq2, len(seq2), int(local), max_hits, buf, need + 1, C.byref(n))
    return buf.raw[:need].decode("latin1"), n.value


@dataclass
class FastaRecords:
    names: List[bytes]
    seqs: List[bytes]
    err_chars: bytes  # non-empty <=> the reference returns CharError carrying these records


def parse_fasta_bytes(contents: bytes) -> FastaRecords:
    n = len(contents)
    out = (C.c_uint8 * max(n, 1))()
    idx_cap = contents.count(b">") + 1
    idx = (C.c_uint64 * (4 * idx_cap))()
    err = (C.c_uint8 * max(n, 1))()
    nerr = C.c_size_t()
    nrec = lib().sao_parse_fasta_mem(contents, n, out, max(n, 1), idx, idx_cap, err, max(n, 1), C.byref(nerr))
    raw = bytes(out)
    names, seqs = [], []
    for r in range(nrec):
        no, nl, so, sl = idx[4 * r : 4 * r + 4]
        names.append(raw[no : no + nl])
        seqs.append(raw[so : so + sl])
    return FastaRecords(names, seqs, bytes(err[: nerr.value]))


def parse_fasta_path(path: str) -> Optional[FastaRecords]:
    """None <=> AlignerError::FastaError (bad extension or unreadable file)."""
    lib()
    ok_ext = os.path.splitext(path)[1] in (".fa", ".fasta", ".fna")
    size = os.path.getsize(path) if os.path.exists(path) else 0
    out = (C.c_uint8 * max(size, 1))()
    idx_cap = 1
    if ok_ext and os.path.exists(path):
        with open(path, "rb") as f:
            idx_cap = f.read().count(b">") + 1
    idx = (C.c_uint64 * (4 * idx_cap))()
    err = (C.c_uint8 * max(size, 1))()
    nerr = C.c_size_t()
    nrec = lib().sao_parse_fasta_path(path.encode(), out, max(size, 1), idx, idx_cap, err, max(size, 1), C.byref(nerr))
    if nrec < 0:
        return None
    raw = bytes(out)
    names, seqs = [], []
    for r in range(nrec):
        no, nl, so, sl = idx[4 * r : 4 * r + 4]
        names.append(raw[no : no + nl])
        seqs.append(raw[so : so + sl])
    return FastaRecords(names, seqs, bytes(err[: nerr.value]))


@dataclass
class WfaResult:
    status: int
    printed_score: int
    panic_line: int
    n_wavefronts: int


def wfa_literal_cap(n1: int, n2: int) -> int:
    return lib().sao_wfa_literal_cap(n1, n2)


def wfa_literal(seq1: bytes, seq2: bytes, max_score: Optional[int] = None) -> WfaResult:
    r = _WfaResult()
    if max_score is None:
        max_score = wfa_literal_cap(len(seq1), len(seq2))
    lib().sao_wfa_literal(seq1, len(seq1), seq2, len(seq2), max_score, C.byref(r))
    return WfaResult(r.status, r.printed_score, r.panic_line, r.n_wavefronts)


def wfa_literal_ex(seq1: bytes, seq2: bytes, max_score: Optional[int] = None):
    """(result, [(lo, hi) lines], converged element (offset, state, parents) or None)"""
    r = _WfaResult()
    if max_score is None:
        max_score = wfa_literal_cap(len(seq1), len(seq2))
    cap = max_score + 8
    lohi = (C.c_int32 * (2 * cap))()
    n = C.c_uint32()
    conv = (C.c_int32 * 6)()
    lib().sao_wfa_literal_ex(seq1, len(seq1), seq2, len(seq2), max_score, C.byref(r), lohi, cap, C.byref(n), conv)
    lines = [(lohi[2 * k], lohi[2 * k + 1]) for k in range(min(n.value, cap))]
    ce = None
    if r.status == OK:
        ce = (conv[0], "MDI"[conv[1]], tuple("MDI"[conv[3 + k]] for k in range(conv[2])))
    return WfaResult(r.status, r.printed_score, r.panic_line, r.n_wavefronts), lines, ce


def wfa_print(seq1: bytes, seq2: bytes, max_score: Optional[int] = None):
    """The reference's stdout for one pair under `-a wfa` (wfa.rs:23-42): (text, status)."""
    if max_score is None:
        max_score = wfa_literal_cap(len(seq1), len(seq2))
    st = C.c_int32()
    need = lib().sao_wfa_print(seq1, len(seq1), seq2, len(seq2), max_score, None, 0, C.byref(st))
    if need < 0:
        raise MemoryError("oracle allocation failed")
    buf = C.create_string_buffer(need + 1)
    lib().sao_wfa_print(seq1, len(seq1), seq2, len(seq2), max_score, buf, need + 1, C.byref(st))
    return buf.raw[:need].decode("latin1"), st.value


def wfa_literal_batch(residues, q_off, q_len, d_off, d_len):
    residues, q_off, q_len, d_off, d_len = _batch_args(residues, q_off, q_len, d_off, d_len)
    n = len(q_len)
    score = np.zeros(n, np.int32)
    status = np.zeros(n, np.uint8)
    rc = lib().sao_wfa_literal_batch(residues.ctypes.data, q_off.ctypes.data, q_len.ctypes.data, d_off.ctypes.data,
                                     d_len.ctypes.data, n, score.ctypes.data, status.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle wfa batch failed")
    return score, status


def wfa_gotoh_cost(seq1: bytes, seq2: bytes, x: int = 4, o: int = 2, e: int = 6) -> int:
    return lib().sao_wfa_gotoh_cost(seq1, len(seq1), seq2, len(seq2), x, o, e)


def wfa_standard(seq1: bytes, seq2: bytes, x: int = 4, o: int = 2, e: int = 6) -> int:
    return lib().sao_wfa_standard(seq1, len(seq1), seq2, len(seq2), x, o, e)
