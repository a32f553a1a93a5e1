"""sequencealigning_b200 -- B200-native batched pairwise alignment (affine NW hot path).

Host side of the drop-in boundary; all compute lives in _lib/libsa_engine.so (CUDA, sm_100a).
"""
from ._capi import (ALIGNMENT_OMITTED, ALGO_NW_AFFINE, ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD, MODE_GLOBAL, MODE_LOCAL, MODE_SEMIGLOBAL,
                    NOT_IMPLEMENTED, OK, OP_D, OP_I, OP_M, REF_NO_CONVERGENCE, REF_NO_OUTPUT, REF_PANIC,
                    REF_PANIC_EARLY)
from .engine import (AlignResult, Engine, EngineError, PairBatch, Record, ResidentBatch, parse_fasta,
                     render_affine, render_linear_hit, status_name)

__all__ = [
    "Engine", "EngineError", "PairBatch", "Record", "AlignResult", "ResidentBatch", "parse_fasta",
    "render_affine", "render_linear_hit", "ALGO_NW_AFFINE", "ALGO_NW_LINEAR", "ALGO_WFA", "ALGO_WFA_STANDARD", "MODE_GLOBAL", "MODE_LOCAL",
    "MODE_SEMIGLOBAL", "OK", "REF_PANIC", "REF_NO_CONVERGENCE", "NOT_IMPLEMENTED", "REF_PANIC_EARLY",
    "REF_NO_OUTPUT", "ALIGNMENT_OMITTED", "status_name", "OP_M", "OP_I", "OP_D",
]
