"""ctypes binding of the C ABI in include/sa_engine.h (libsa_engine.so, built in-tree).

The library is the product: if it is missing or cannot be loaded this module raises -- there
is no Python or CPU fallback for the alignment path.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lib", "libsa_engine.so")

# sa_status_t
OK, REF_PANIC, REF_NO_CONVERGENCE, NOT_IMPLEMENTED, REF_PANIC_EARLY, REF_NO_OUTPUT = range(6)
ALIGNMENT_OMITTED = 0x80  # flag ORed into a per-pair status (score and status exact, no CIGAR)
E_CUDA, E_ARG, E_NOMEM, E_CIGAR_CAPACITY, E_UNSUPPORTED = -1, -2, -3, -4, -5
ALGO_NW_AFFINE, ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD = 0, 1, 2, 3
MODE_GLOBAL, MODE_LOCAL, MODE_SEMIGLOBAL = 0, 1, 2
OP_M, OP_I, OP_D = 0, 1, 2

# every symbol include/sa_engine.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "sa_abi_version", "sa_engine_create", "sa_engine_destroy", "sa_last_error", "sa_align_batch",
    "sa_batch_upload", "sa_batch_free", "sa_align_resident", "sa_resident_download",
    "sa_engine_synchronize", "sa_engine_stream", "sa_last_timing", "sa_alloc_pinned", "sa_free_pinned",
    "sa_partition_lpt", "sa_parse_fasta", "sa_render_affine", "sa_pack_2bit", "sa_affine_all_alignments",
    "sa_affine_count_cooptimal", "sa_engine_create_multi", "sa_engine_device_count", "sa_last_shards", "sa_plan_shards",
    "sa_pack_2bit_mt", "sa_parse_fasta_packed", "sa_render_linear_hit", "sa_wfa_reference_stdout", "sa_host_register", "sa_host_unregister", "sa_linear_all_hits",
]


class Scheme(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_ext", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [
        ("residues", C.c_void_p), ("residues_len", C.c_uint64),
        ("q_off", C.c_void_p), ("q_len", C.c_void_p), ("d_off", C.c_void_p), ("d_len", C.c_void_p),
        ("n_pairs", C.c_uint64), ("packing", C.c_uint32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("score", C.c_void_p), ("status", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar_len", C.c_void_p),
        ("cigar", C.c_void_p), ("cigar_capacity", C.c_uint64), ("cigar_used", C.c_uint64),
        ("end1", C.c_void_p), ("end2", C.c_void_p),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("kernels_ms", C.c_double), ("fill_ms", C.c_double), ("long_fwd_ms", C.c_double), ("long_back_ms", C.c_double),
        ("wall_ms", C.c_double), ("cells", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("pairs_rerun", C.c_uint64), ("pairs_fallback", C.c_uint64),
        ("wfa_cells", C.c_uint64), ("wfa_extended", C.c_uint64),
    ]


class ShardInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("contiguous", C.c_uint32), ("first_pair", C.c_uint64), ("pairs", C.c_uint64),
        ("cells", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("device_ms", C.c_double), ("host_ms", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sequencealigning_b200.build` "
            "(nvcc, sm_100a). The alignment path has no CPU fallback."
        )
    l = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    l.sa_abi_version.restype = C.c_int
    l.sa_engine_create.argtypes = [C.c_int, C.POINTER(vp)]
    l.sa_engine_create.restype = C.c_int
    l.sa_engine_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    l.sa_engine_create_multi.restype = C.c_int
    l.sa_engine_device_count.argtypes = [vp]
    l.sa_engine_device_count.restype = C.c_int
    l.sa_last_shards.argtypes = [vp, C.POINTER(ShardInfo), C.c_int, C.POINTER(C.c_int)]
    l.sa_last_shards.restype = C.c_int
    l.sa_plan_shards.argtypes = [vp, vp, C.c_uint64, C.c_int, vp, vp, C.POINTER(C.c_int)]
    l.sa_plan_shards.restype = C.c_int
    l.sa_engine_destroy.argtypes = [vp]
    l.sa_engine_destroy.restype = C.c_int
    l.sa_last_error.argtypes = [vp]
    l.sa_last_error.restype = C.c_char_p
    l.sa_align_batch.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Scheme), C.POINTER(Batch), C.POINTER(Result)]
    l.sa_align_batch.restype = C.c_int
    l.sa_batch_upload.argtypes = [vp, C.POINTER(Batch), C.POINTER(vp)]
    l.sa_batch_upload.restype = C.c_int
    l.sa_batch_free.argtypes = [vp, vp]
    l.sa_batch_free.restype = C.c_int
    l.sa_align_resident.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Scheme), vp, C.c_int]
    l.sa_align_resident.restype = C.c_int
    l.sa_resident_download.argtypes = [vp, vp, C.POINTER(Result)]
    l.sa_resident_download.restype = C.c_int
    l.sa_engine_synchronize.argtypes = [vp]
    l.sa_engine_synchronize.restype = C.c_int
    l.sa_engine_stream.argtypes = [vp]
    l.sa_engine_stream.restype = vp
    l.sa_last_timing.argtypes = [vp, C.POINTER(Timing)]
    l.sa_last_timing.restype = C.c_int
    l.sa_alloc_pinned.argtypes = [C.c_size_t]
    l.sa_alloc_pinned.restype = vp
    l.sa_free_pinned.argtypes = [vp]
    l.sa_free_pinned.restype = None
    l.sa_host_register.argtypes = [vp, C.c_size_t]
    l.sa_host_register.restype = C.c_int
    l.sa_host_unregister.argtypes = [vp]
    l.sa_host_unregister.restype = C.c_int
    l.sa_partition_lpt.argtypes = [vp, vp, C.c_uint64, C.c_int, vp]
    l.sa_partition_lpt.restype = C.c_int
    l.sa_parse_fasta.argtypes = [C.c_char_p, vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    l.sa_parse_fasta.restype = C.c_int64
    l.sa_render_affine.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, vp, C.c_uint32, C.c_char_p, C.c_size_t]
    l.sa_render_affine.restype = C.c_int64
    l.sa_render_linear_hit.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_char_p, C.c_size_t]
    l.sa_render_linear_hit.restype = C.c_int64
    l.sa_wfa_reference_stdout.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_int32)]
    l.sa_wfa_reference_stdout.restype = C.c_int64
    l.sa_linear_all_hits.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.POINTER(Scheme), C.c_uint64,
                                     C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
    l.sa_linear_all_hits.restype = C.c_int64
    l.sa_pack_2bit.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
    l.sa_pack_2bit.restype = C.c_int
    l.sa_pack_2bit_mt.argtypes = [vp, C.c_uint64, vp, C.c_int]
    l.sa_pack_2bit_mt.restype = C.c_int
    l.sa_parse_fasta_packed.argtypes = [C.c_char_p, vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t),
                                        vp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    l.sa_parse_fasta_packed.restype = C.c_int64
    l.sa_affine_all_alignments.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(Scheme), C.c_uint64,
                                           C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    l.sa_affine_all_alignments.restype = C.c_int64
    l.sa_affine_count_cooptimal.argtypes = [vp, C.POINTER(Scheme), C.POINTER(Batch), vp]
    l.sa_affine_count_cooptimal.restype = C.c_int
    _lib = l
    return l
