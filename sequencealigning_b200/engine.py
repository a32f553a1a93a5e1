"""Host-side mirror of the reference's aligner interface over the C ABI.

The reference's only API is three functions called once per (query, db) pair from the loop in
/root/reference/src/main.rs:61-79:
    n_w_align(seq1: &Record, seq2: &Record, verbose, mode)   needleman_wunsch_affine.rs:424
    wfa_align(seq1, seq2, mode)                              wfa.rs:23
Here the same call is made once per BATCH of pairs.  `Record` keeps the reference's field
names (parse.rs:135-139).  All computation happens in libsa_engine.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import (ALIGNMENT_OMITTED, ALGO_NW_AFFINE, ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD, MODE_GLOBAL, MODE_LOCAL, MODE_SEMIGLOBAL,
                    NOT_IMPLEMENTED, OK, REF_NO_CONVERGENCE, REF_NO_OUTPUT, REF_PANIC, REF_PANIC_EARLY)

STATUS_NAMES = {
    OK: "OK", REF_PANIC: "REF_PANIC", REF_NO_CONVERGENCE: "REF_NO_CONVERGENCE",
    NOT_IMPLEMENTED: "NOT_IMPLEMENTED", REF_PANIC_EARLY: "REF_PANIC_EARLY", REF_NO_OUTPUT: "REF_NO_OUTPUT",
}
ALGOS = {"needleman-wunsch": ALGO_NW_AFFINE, "needleman-wunsch-linear": ALGO_NW_LINEAR, "wfa": ALGO_WFA,
         "wfa-standard": ALGO_WFA_STANDARD}
MODES = {"global": MODE_GLOBAL, "local": MODE_LOCAL, "semi-global": MODE_SEMIGLOBAL}


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sa_engine error {code}: {msg}")
        self.code = code


@dataclass
class Record:
    """parse.rs:135-139.  `name` includes the leading '>'."""
    seq: bytes
    name: bytes = b""


@dataclass
class PairBatch:
    """Pairs as (offset, length) views into one residue buffer -- the layout of sa_batch_t."""
    residues: np.ndarray  # uint8
    q_off: np.ndarray     # uint64
    q_len: np.ndarray     # uint32
    d_off: np.ndarray     # uint64
    d_len: np.ndarray     # uint32
    packing: int = 0      # 0 = bytes, 1 = 2-bit codes (offsets then count residues)

    def __post_init__(self):
        self.residues = np.ascontiguousarray(self.residues, np.uint8)
        self.q_off = np.ascontiguousarray(self.q_off, np.uint64)
        self.d_off = np.ascontiguousarray(self.d_off, np.uint64)
        self.q_len = np.ascontiguousarray(self.q_len, np.uint32)
        self.d_len = np.ascontiguousarray(self.d_len, np.uint32)
        n = len(self.q_len)
        if not (len(self.q_off) == len(self.d_off) == len(self.d_len) == n):
            raise ValueError("offset/length arrays differ in length")

    @property
    def n_pairs(self) -> int:
        return len(self.q_len)

    @property
    def cells(self) -> int:
        return int((self.q_len.astype(np.uint64) * self.d_len.astype(np.uint64)).sum())

    def _seq(self, o: int, l: int) -> bytes:
        if getattr(self, "packing", 0) == 0:
            return self.residues[o:o + l].tobytes()
        idx = np.arange(o, o + l)
        codes = (self.residues[idx >> 2] >> (2 * (idx & 3)).astype(np.uint8)) & 3
        return np.frombuffer(b"ACGT", np.uint8)[codes].tobytes()

    def query(self, p: int) -> bytes:
        return self._seq(int(self.q_off[p]), int(self.q_len[p]))

    def db(self, p: int) -> bytes:
        return self._seq(int(self.d_off[p]), int(self.d_len[p]))

    def select(self, idx: np.ndarray) -> "PairBatch":
        return PairBatch(self.residues, self.q_off[idx], self.q_len[idx], self.d_off[idx], self.d_len[idx],
                         getattr(self, "packing", 0))

    def packed(self) -> "PairBatch":
        """The same pairs in the 2-bit format (sa_pack_2bit); raises ValueError on non-ACGT bytes."""
        if getattr(self, "packing", 0) == 1:
            return self
        n = int(self.residues.size)
        out = np.zeros((n + 3) // 4 + 1, np.uint8)
        rc = _capi.lib().sa_pack_2bit_mt(self.residues.ctypes.data, n, out.ctypes.data, 0)
        if rc != 0:
            raise ValueError("2-bit packing needs A/C/G/T only")
        return PairBatch(out, self.q_off, self.q_len, self.d_off, self.d_len, 1)

    @staticmethod
    def from_pairs(pairs: Iterable[Tuple[bytes, bytes]]) -> "PairBatch":
        chunks, q_off, q_len, d_off, d_len, cur = [], [], [], [], [], 0
        for q, d in pairs:
            q_off.append(cur); q_len.append(len(q)); chunks.append(q); cur += len(q)
            d_off.append(cur); d_len.append(len(d)); chunks.append(d); cur += len(d)
        res = np.frombuffer(b"".join(chunks), np.uint8) if cur else np.zeros(0, np.uint8)
        return PairBatch(res, np.array(q_off, np.uint64), np.array(q_len, np.uint32),
                         np.array(d_off, np.uint64), np.array(d_len, np.uint32))

    @staticmethod
    def from_records(query: Sequence[Record], db: Sequence[Record]) -> "PairBatch":
        """The db-major cross product of main.rs:61-62: for d in db { for q in query { .. } }."""
        seqs = [r.seq for r in query] + [r.seq for r in db]
        offs = np.zeros(len(seqs) + 1, np.uint64)
        offs[1:] = np.cumsum([len(s) for s in seqs])
        res = np.frombuffer(b"".join(seqs), np.uint8) if offs[-1] else np.zeros(0, np.uint8)
        nq, nd = len(query), len(db)
        qi = np.tile(np.arange(nq), nd)
        di = np.repeat(np.arange(nd), nq) + nq
        lens = np.array([len(s) for s in seqs], np.uint32)
        return PairBatch(res, offs[qi], lens[qi], offs[di], lens[di])


@dataclass
class AlignResult:
    score: np.ndarray      # int32
    status: np.ndarray     # uint8 (sa_status_t)
    cigar_off: np.ndarray  # uint64
    cigar_len: np.ndarray  # uint32
    cigar: np.ndarray      # uint32 pool
    end1: Optional[np.ndarray] = None  # uint32: cell the traceback starts from (global: n1, n2;
    end2: Optional[np.ndarray] = None  # local: first row-major maximum, needleman_wunsch.rs:256-272)

    def cigar_of(self, p: int) -> List[int]:
        o, l = int(self.cigar_off[p]), int(self.cigar_len[p])
        return [int(v) for v in self.cigar[o:o + l]]

    def cigar_string(self, p: int) -> str:
        return "".join(f"{w >> 2}{'MID'[w & 3]}" for w in self.cigar_of(p))


class PinnedBuffer:
    """Page-locked host memory from sa_alloc_pinned, viewed as a numpy array."""

    def __init__(self, n: int, dtype):
        self._lib = _capi.lib()
        self.nbytes = max(int(n) * np.dtype(dtype).itemsize, 1)
        self.ptr = self._lib.sa_alloc_pinned(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"sa_alloc_pinned({self.nbytes}) failed")
        raw = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=dtype, count=int(n))

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._lib.sa_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedResult:
    """Caller-owned pinned result buffers for Engine.align(..., out=...)."""

    def __init__(self, n_pairs: int, cigar_capacity: int):
        self.n_pairs, self.cigar_capacity = n_pairs, cigar_capacity
        self._bufs = [PinnedBuffer(n_pairs, np.int32), PinnedBuffer(n_pairs, np.uint8), PinnedBuffer(n_pairs, np.uint64),
                      PinnedBuffer(n_pairs, np.uint32), PinnedBuffer(max(cigar_capacity, 1), np.uint32),
                      PinnedBuffer(n_pairs, np.uint32), PinnedBuffer(n_pairs, np.uint32)]
        self.score, self.status, self.cigar_off, self.cigar_len, self.cigar, self.end1, self.end2 = (b.array for b in self._bufs)

    def free(self):
        for b in self._bufs:
            b.free()


def pin_batch(batch: "PairBatch"):
    """Copies a batch into pinned memory (what a packer would write into directly)."""
    bufs = []
    arrs = []
    for a in (batch.residues, batch.q_off, batch.q_len, batch.d_off, batch.d_len):
        pb = PinnedBuffer(a.size, a.dtype)
        pb.array[:] = a
        bufs.append(pb)
        arrs.append(pb.array)
    out = PairBatch.__new__(PairBatch)
    out.residues, out.q_off, out.q_len, out.d_off, out.d_len = arrs
    out.packing = getattr(batch, "packing", 0)
    out._pinned = bufs
    return out


def status_name(status: int) -> str:
    """Name of a per-pair status byte; the ALIGNMENT_OMITTED flag (0x80) is shown as a suffix."""
    base = STATUS_NAMES.get(int(status) & 0x7F, f"status {int(status) & 0x7F}")
    return base + ("|ALIGNMENT_OMITTED" if int(status) & ALIGNMENT_OMITTED else "")


def _scheme(s) -> Optional[_capi.Scheme]:
    if s is None:
        return None
    if isinstance(s, _capi.Scheme):
        return s
    return _capi.Scheme(*s)


class Engine:
    """One engine per process.  `device` = one GPU; `devices=[..]` = ONE engine over several GPUs
    (sa_engine_create_multi): align() then shards the pair list over them inside the call and
    returns everything in input order (see shards())."""

    def __init__(self, device: int = 0, devices: Optional[Sequence[int]] = None):
        self._lib = _capi.lib()
        self._h = C.c_void_p()
        if devices is not None:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._lib.sa_engine_create_multi(ids, len(devices), C.byref(self._h))
            device = int(devices[0]) if len(devices) else 0
        else:
            rc = self._lib.sa_engine_create(device, C.byref(self._h))
        if rc != 0:
            msg = self._lib.sa_last_error(self._h).decode() if self._h else "allocation failed"
            if self._h:
                self._lib.sa_engine_destroy(self._h)
                self._h = C.c_void_p()
            raise EngineError(rc, msg)
        self.device = device
        self.devices = list(devices) if devices is not None else [device]

    @property
    def device_count(self) -> int:
        return int(self._lib.sa_engine_device_count(self._h))

    def shards(self) -> List[dict]:
        """How the last align() was split over the devices (sa_last_shards)."""
        n = C.c_int()
        arr = (_capi.ShardInfo * 64)()
        self._check(self._lib.sa_last_shards(self._h, arr, 64, C.byref(n)))
        return [{k: getattr(arr[i], k) for k, _ in _capi.ShardInfo._fields_} for i in range(n.value)]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sa_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc < 0:
            raise EngineError(rc, self._lib.sa_last_error(self._h).decode())

    @staticmethod
    def _c_batch(b: PairBatch) -> _capi.Batch:
        return _capi.Batch(b.residues.ctypes.data, b.residues.size, b.q_off.ctypes.data, b.q_len.ctypes.data,
                           b.d_off.ctypes.data, b.d_len.ctypes.data, b.n_pairs, getattr(b, "packing", 0))

    @staticmethod
    def _alloc_result(n: int, cigar_capacity: int):
        score = np.zeros(n, np.int32)
        status = np.zeros(n, np.uint8)
        off = np.zeros(n, np.uint64)
        ln = np.zeros(n, np.uint32)
        pool = np.zeros(max(cigar_capacity, 0), np.uint32)
        end1 = np.zeros(n, np.uint32)
        end2 = np.zeros(n, np.uint32)
        res = _capi.Result(score.ctypes.data, status.ctypes.data, off.ctypes.data, ln.ctypes.data,
                           pool.ctypes.data if cigar_capacity > 0 else None, max(cigar_capacity, 0), 0,
                           end1.ctypes.data, end2.ctypes.data)
        return res, (score, status, off, ln, pool, end1, end2)

    def align(self, batch: PairBatch, algo: int = ALGO_NW_AFFINE, mode: int = MODE_GLOBAL, scheme=None,
              cigar: bool = True, cigar_capacity: Optional[int] = None, out: Optional["PinnedResult"] = None) -> AlignResult:
        """sa_align_batch: host buffers in, host buffers out (the reference-facing call)."""
        n = batch.n_pairs
        if out is not None:
            sc = _scheme(scheme)
            cb = self._c_batch(batch)
            # end cells: only local mode computes them; a global alignment ends at (n1, n2), which the caller
            # already holds (asking the engine for them costs two n-element copies per call)
            local = mode == MODE_LOCAL
            res = _capi.Result(out.score.ctypes.data, out.status.ctypes.data, out.cigar_off.ctypes.data,
                               out.cigar_len.ctypes.data, out.cigar.ctypes.data if cigar else None,
                               out.cigar_capacity if cigar else 0, 0,
                               out.end1.ctypes.data if local else None, out.end2.ctypes.data if local else None)
            self._check(self._lib.sa_align_batch(self._h, algo, mode, C.byref(sc) if sc else None, C.byref(cb), C.byref(res)))
            return AlignResult(out.score, out.status, out.cigar_off, out.cigar_len, out.cigar[: int(res.cigar_used)],
                               out.end1 if local else batch.q_len, out.end2 if local else batch.d_len)
        cap = 0
        if cigar:
            # 32 runs per pair covers read pairs; long pairs get a share of their length
            cap = cigar_capacity if cigar_capacity is not None else max(
                64, 32 * n, int(np.minimum(batch.q_len, batch.d_len).sum(dtype=np.int64)) // 8)
        sc = _scheme(scheme)
        cb = self._c_batch(batch)
        while True:
            res, arrs = self._alloc_result(n, cap)
            rc = self._lib.sa_align_batch(self._h, algo, mode, C.byref(sc) if sc else None, C.byref(cb), C.byref(res))
            if rc == _capi.E_CIGAR_CAPACITY and cigar_capacity is None:
                cap = int(res.cigar_used) + 16
                continue
            self._check(rc)
            break
        score, status, off, ln, pool, end1, end2 = arrs
        return AlignResult(score, status, off, ln, pool[: int(res.cigar_used)] if cap else pool, end1, end2)

    def all_alignments(self, seq1: bytes, seq2: bytes, scheme=None, max_alignments: int = 1 << 20):
        """Every co-optimal alignment of one pair as the reference prints them.
        Returns (text, n_printed, panicked)."""
        sc = _scheme(scheme)
        n = C.c_uint64()
        pan = C.c_int32()
        args = (self._h, seq1, len(seq1), seq2, len(seq2), C.byref(sc) if sc else None, max_alignments)
        need = self._lib.sa_affine_all_alignments(*args, None, 0, C.byref(n), C.byref(pan))
        self._check(need)
        buf = C.create_string_buffer(need + 1)
        self._check(self._lib.sa_affine_all_alignments(*args, buf, need + 1, C.byref(n), C.byref(pan)))
        return buf.value.decode("latin1"), n.value, bool(pan.value)

    def linear_all_hits(self, seq1: bytes, seq2: bytes, local: bool = False, scheme=None, max_hits: int = 1 << 20):
        """Every hit the reference's linear aligner prints for one pair, as its text (sa_linear_all_hits).
        Returns (text, n_printed)."""
        sc = _scheme(scheme)
        n = C.c_uint64()
        args = (self._h, seq1, len(seq1), seq2, len(seq2), int(local), C.byref(sc) if sc else None, max_hits)
        need = self._lib.sa_linear_all_hits(*args, None, 0, C.byref(n))
        self._check(need)
        buf = C.create_string_buffer(need + 1)
        self._check(self._lib.sa_linear_all_hits(*args, buf, need + 1, C.byref(n)))
        return buf.raw[:need].decode("latin1"), n.value

    def wfa_reference_stdout(self, seq1: bytes, seq2: bytes):
        """The reference's complete stdout for one pair under `-a wfa` (sa_wfa_reference_stdout).
        Returns (text, status)."""
        st = C.c_int32()
        need = self._lib.sa_wfa_reference_stdout(self._h, seq1, len(seq1), seq2, len(seq2), None, 0, C.byref(st))
        self._check(need)
        buf = C.create_string_buffer(need + 1)
        self._check(self._lib.sa_wfa_reference_stdout(self._h, seq1, len(seq1), seq2, len(seq2), buf, need + 1, C.byref(st)))
        return buf.raw[:need].decode("latin1"), st.value

    # ---- device-resident path (benchmarks: inputs already in HBM) -------------------------
    def count_cooptimal(self, batch: PairBatch, scheme=None) -> np.ndarray:
        """Per pair, how many alignments the reference's traceback prints when nothing panics
        (sa_affine_count_cooptimal; saturates at INT64_MAX // 4)."""
        counts = np.zeros(batch.n_pairs, np.int64)
        cb = self._c_batch(batch)
        sc = _scheme(scheme)
        self._check(self._lib.sa_affine_count_cooptimal(self._h, C.byref(sc) if sc else None, C.byref(cb), counts.ctypes.data))
        return counts

    def upload(self, batch: PairBatch) -> "ResidentBatch":
        h = C.c_void_p()
        cb = self._c_batch(batch)
        self._check(self._lib.sa_batch_upload(self._h, C.byref(cb), C.byref(h)))
        self._check(self._lib.sa_engine_synchronize(self._h))
        return ResidentBatch(self, h, batch.n_pairs)

    def synchronize(self):
        self._check(self._lib.sa_engine_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.sa_engine_stream(self._h) or 0)

    def timing(self) -> dict:
        t = _capi.Timing()
        self._lib.sa_last_timing(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in t._fields_}


class ResidentBatch:
    def __init__(self, eng: Engine, handle, n_pairs: int):
        self._eng, self._h, self.n_pairs = eng, handle, n_pairs

    def align(self, algo: int = ALGO_NW_AFFINE, mode: int = MODE_GLOBAL, scheme=None, cigar: bool = True):
        sc = _scheme(scheme)
        e = self._eng
        e._check(e._lib.sa_align_resident(e._h, algo, mode, C.byref(sc) if sc else None, self._h, int(cigar)))

    def download(self, cigar_capacity: Optional[int] = None) -> AlignResult:
        e = self._eng
        n = self.n_pairs
        cap = cigar_capacity if cigar_capacity is not None else max(64, 32 * n)
        while True:
            res, arrs = e._alloc_result(n, cap)
            rc = e._lib.sa_resident_download(e._h, self._h, C.byref(res))
            if rc == _capi.E_CIGAR_CAPACITY and cigar_capacity is None:
                cap = int(res.cigar_used) + 16
                continue
            e._check(rc)
            break
        score, status, off, ln, pool, end1, end2 = arrs
        return AlignResult(score, status, off, ln, pool[: int(res.cigar_used)], end1, end2)

    def free(self):
        if self._h:
            self._eng._lib.sa_batch_free(self._eng._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def render_affine(seq1: bytes, seq2: bytes, cigar: Sequence[int]) -> str:
    """Reference stdout for one alignment (needleman_wunsch_affine.rs:283-286, :390-411)."""
    l = _capi.lib()
    arr = (C.c_uint32 * max(len(cigar), 1))(*cigar)
    need = l.sa_render_affine(seq1, len(seq1), seq2, len(seq2), arr, len(cigar), None, 0)
    if need < 0:
        raise ValueError("CIGAR does not fit the sequences")
    buf = C.create_string_buffer(need + 1)
    l.sa_render_affine(seq1, len(seq1), seq2, len(seq2), arr, len(cigar), buf, need + 1)
    return buf.value.decode("latin1")


def render_linear_hit(seq1: bytes, seq2: bytes, cigar: Sequence[int], end1: Optional[int] = None, end2: Optional[int] = None) -> str:
    """Reference stdout for one hit of the linear aligner (needleman_wunsch.rs:155-178, :207-216)."""
    l = _capi.lib()
    end1 = len(seq1) if end1 is None else int(end1)
    end2 = len(seq2) if end2 is None else int(end2)
    arr = (C.c_uint32 * max(len(cigar), 1))(*cigar)
    need = l.sa_render_linear_hit(seq1, len(seq1), seq2, len(seq2), arr, len(cigar), end1, end2, None, 0)
    if need < 0:
        raise ValueError("CIGAR / end cell do not fit the sequences")
    buf = C.create_string_buffer(need + 1)
    l.sa_render_linear_hit(seq1, len(seq1), seq2, len(seq2), arr, len(cigar), end1, end2, buf, need + 1)
    return buf.raw[:need].decode("latin1")


def parse_fasta_packed(path: str):
    """sa_parse_fasta_packed: (records, err_chars, out, packed, index, all_acgt).  `out` is the parser's
    output buffer (names and sequences), `packed` its 2-bit image, `index` the (name_off, name_len,
    seq_off, seq_len) rows: a PairBatch built on `packed` with packing = 1 uses seq_off / seq_len as is."""
    import os
    l = _capi.lib()
    if not os.path.isfile(path):
        raise ValueError(f"FastaError: {path}")
    size = os.path.getsize(path)
    out = np.zeros(max(size, 1), np.uint8)
    with open(path, "rb") as f:
        idx_cap = f.read().count(b">") + 1 if size else 1
    idx = np.zeros(4 * idx_cap, np.uint64)
    err = np.zeros(max(size, 1), np.uint8)
    packed = np.zeros((max(size, 1) + 3) // 4 + 1, np.uint8)
    nerr = C.c_size_t()
    acgt = C.c_int()
    out_len = C.c_uint64()
    n = l.sa_parse_fasta_packed(path.encode(), out.ctypes.data, out.size, idx.ctypes.data, idx_cap, err.ctypes.data, err.size,
                                C.byref(nerr), packed.ctypes.data, packed.size, C.byref(acgt), C.byref(out_len))
    if n < 0:
        raise ValueError(f"FastaError: {path}")
    raw = out.tobytes()
    index = idx[: 4 * n].reshape(n, 4)
    recs = [Record(seq=raw[int(so):int(so + sl)], name=raw[int(no):int(no + nl)]) for no, nl, so, sl in index]
    return recs, err[: nerr.value].tobytes(), out[: out_len.value], packed[: (out_len.value + 3) // 4], index, bool(acgt.value)


def parse_fasta(path: str):
    """parse_fasta (parse.rs:54-99).  Returns (records, err_chars); raises ValueError for the
    reference's FastaError.  Non-empty err_chars <=> the reference returns CharError carrying
    these records (main.rs:29-35 then continues with them)."""
    import os
    l = _capi.lib()
    if not os.path.isfile(path):
        raise ValueError(f"FastaError: {path}")  # parse.rs:62 `read(path)?`
    size = os.path.getsize(path)
    out = np.zeros(max(size, 1), np.uint8)
    with open(path, "rb") as f:
        idx_cap = f.read().count(b">") + 1 if size else 1
    idx = np.zeros(4 * idx_cap, np.uint64)
    err = np.zeros(max(size, 1), np.uint8)
    nerr = C.c_size_t()
    n = l.sa_parse_fasta(path.encode(), out.ctypes.data, out.size, idx.ctypes.data, idx_cap, err.ctypes.data, err.size, C.byref(nerr))
    if n < 0:
        raise ValueError(f"FastaError: {path}")
    raw = out.tobytes()
    recs = []
    for r in range(n):
        no, nl, so, sl = (int(v) for v in idx[4 * r:4 * r + 4])
        recs.append(Record(seq=raw[so:so + sl], name=raw[no:no + nl]))
    return recs, err[: nerr.value].tobytes()
