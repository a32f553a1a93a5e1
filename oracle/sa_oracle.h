/*
 * sa_oracle.h -- CPU ORACLE for the SequenceAligning hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This directory is a literal CPU restatement of the reference's three aligners
 * (Qw11111111111/SequenceAligning, crate a_star_align 0.1.0).  It exists so the CUDA
 * engine can be checked bit-for-bit.  Nothing under sequencealigning_b200/ may link,
 * import or call it: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * PARITY PIN STATUS
 *   - The reference is Rust; no Rust toolchain exists in the build image, so the
 *     reference itself cannot be executed here (oracle/_ref is therefore absent).
 *   - affine NW  (src/needleman_wunsch_affine.rs): the reference's own tests are empty
 *     (:458-470)  ->  "parity unpinned" by reference vectors.  It is pinned instead by
 *     (a) hand-derived known-answer vectors in tests/golden/, and (b) a second,
 *     independent, object-graph-literal Python transliteration (oracle/literal_model.py)
 *     that is cross-checked against this C code on thousands of random pairs.
 *   - linear NW  (src/needleman_wunsch.rs): no reference tests -> "parity unpinned",
 *     same two-restatement scheme.
 *   - WFA        (src/wfa.rs): pinned on the reference's `test_initial` (:1104-1186),
 *     `recurrance_eq` (:1002-1102), `test_wavefront_tensor_new_all_none` (:994-1000),
 *     `test_iteration` (:1268-1286) and `test_converge` (:1288-1294) expectations,
 *     committed as tests/golden/wfa_reference_tests.json.
 *   - FASTA parser (src/parse.rs:54-99): pinned on the four reference tests (:166-251).
 *
 * Conventions (src/main.rs:61-66): seq1 = QUERY record, seq2 = DB record,
 * n1 = len(seq1), n2 = len(seq2).  All arithmetic is i32.
 */
#ifndef SA_ORACLE_H
#define SA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-pair outcome of "what would the reference binary do on this pair". */
enum {
  SAO_OK = 0,                 /* reference runs to completion                                   */
  SAO_REF_PANIC = 1,          /* reference panics AFTER printing >=1 alignment (canonical valid) */
  SAO_REF_NO_CONVERGENCE = 2, /* WFA: loop never terminates (wfa.rs:28 with :189)               */
  SAO_NOT_IMPLEMENTED = 3,    /* Err("not implemented") (nw_affine:433-434, wfa.rs:26)          */
  SAO_REF_PANIC_EARLY = 4,    /* reference panics BEFORE printing anything (no canonical)       */
  SAO_REF_NO_OUTPUT = 5       /* completes but prints no alignment (sentinel dead ends only)    */
};

/* CIGAR op codes, one per alignment column, named explicitly:
 *   SAO_OP_M : diagonal column  seq1[y-1] / seq2[x-1]   (state InM, nw_affine:292-296)
 *   SAO_OP_I : seq1[y-1] / '-'                          (state InI, nw_affine:302-306)
 *   SAO_OP_D : '-' / seq2[x-1]                          (state InD, nw_affine:297-301)
 * Run-length packed as (len << 2) | op, in alignment order (first column first). */
enum { SAO_OP_M = 0, SAO_OP_I = 1, SAO_OP_D = 2 };

typedef struct {
  int32_t match_, mismatch, gap_opening, gap_extension;
} sao_scheme_t;

/* nw_affine.rs:15-20 */
static const sao_scheme_t SAO_AFFINE_SCHEME = {5, -4, -8, -6};

typedef struct {
  int32_t status;          /* SAO_* */
  int32_t score;           /* max(I,D,M)[n2][n1]  (nw_affine:247-250; never printed by the ref) */
  int32_t end_m, end_i, end_d; /* the three end-cell scores */
  int32_t any_panic;       /* 1 iff ANY co-optimal path reaches a boundary-chain cell           */
  int64_t n_cooptimal;     /* #complete paths the DFS would print absent a panic (saturating)   */
  uint32_t cigar_len;      /* number of run-length words written                                 */
  uint32_t n_columns;      /* alignment columns of the canonical alignment                       */
} sao_affine_result_t;

/* Affine-gap global NW, literal restatement of needleman_wunsch_affine.rs:169-334.
 * cigar: caller buffer of >= n1+n2+1 words (may be NULL to skip the canonical path).
 * Returns 0, or -1 on allocation failure. */
int sao_affine_align(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                     const sao_scheme_t* scheme, sao_affine_result_t* out, uint32_t* cigar);

/* Score only, O(n1) memory, same recurrences and sentinels (for long pairs / CPU baseline). */
int32_t sao_affine_score(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                         const sao_scheme_t* scheme);

/* Literal LIFO-DFS enumeration (nw_affine:246-329) writing the reference's stdout text
 * (Appendix A.1 of SURVEY.md, without the trailing Duration line) into buf.
 * Stops after max_alignments printed alignments or when the reference would panic.
 * *n_printed gets the number of alignments printed; *panicked is 1 if the DFS hit a panic.
 * Returns the number of bytes that the full text needs (snprintf-style); -1 on OOM. */
int64_t sao_affine_print_all(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                             const sao_scheme_t* scheme, uint64_t max_alignments, char* buf,
                             size_t buf_cap, uint64_t* n_printed, int32_t* panicked);

/* Dump of the three score matrices, row-major [(n2+1) x (n1+1)], for white-box tests. */
int sao_affine_matrices(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                        const sao_scheme_t* scheme, int32_t* m, int32_t* i, int32_t* d,
                        uint8_t* parents /* 7 bits per cell, may be NULL */);

/* Batch helper used by tests/bench: pairs are (off,len) into one residue buffer.
 * cigar_pool holds per-pair slots of stride `cigar_stride` words. n_threads<=1 = serial. */
int sao_affine_batch(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                     const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs,
                     const sao_scheme_t* scheme, int32_t* score, uint8_t* status,
                     uint32_t* cigar_len, uint32_t* cigar_pool, uint32_t cigar_stride,
                     int n_threads);

/* ------------------------------------------------------------------------------------- */
/* Linear ("pseudo-affine", single matrix + gap flags) NW: needleman_wunsch.rs:36-117,180-272.
 * NOTE the transposed geometry: rows i walk seq1 (query), columns j walk seq2 (db).
 * Moves (needleman_wunsch.rs:16-20): Down = seq1 char / '-', Right = '-' / seq2 char.
 * CIGAR ops use the same meaning as the affine aligner: Down -> SAO_OP_I, Right -> SAO_OP_D. */
typedef struct {
  int32_t status;       /* SAO_OK always for global mode (no panic sites reachable)          */
  int32_t score;        /* scores[n1][n2] (global) or the matrix maximum (local)             */
  int64_t n_hits;       /* #hits the recursion would print (saturating)                       */
  uint32_t cigar_len;   /* canonical = FIRST printed hit: priority Down > Right > Diag        */
  uint32_t n_columns;
  uint32_t start1, start2; /* "start in seq1/seq2" lines of the first hit (:171-175)          */
  uint32_t end1, end2;     /* the cell the first hit's recursion starts from: (n1, n2) in global
                              mode, the first argmax cell in row-major order in local mode (:107-111) */
} sao_linear_result_t;

int sao_linear_align(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                     int local, sao_linear_result_t* out, uint32_t* cigar);
/* scores matrix dump, row-major [(n1+1) x (n2+1)], plus 3-bit move sets and gap flags */
int sao_linear_matrices(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                        int local, int32_t* scores, uint8_t* moves, uint8_t* gaps);
int sao_linear_batch(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                     const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs,
                     int32_t* score, uint8_t* status, uint32_t* cigar_len, uint32_t* cigar_pool,
                     uint32_t cigar_stride, int n_threads);
/* the same with the mode switch of n_w_align (:180) and the start cell of every pair (end1/end2 may be NULL) */
int sao_linear_batch_ex(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                        const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs, int local,
                        int32_t* score, uint8_t* status, uint32_t* cigar_len, uint32_t* cigar_pool,
                        uint32_t cigar_stride, uint32_t* end1, uint32_t* end2, int n_threads);
/* Literal recursion of backtrace / get_next (:106-116, :205-254) writing the reference's stdout for
 * the hits of one pair -- "\nHit: {hit}\n\n" per hit, `Hit` as Display :155-178 -- in the order
 * the reference prints them (start cells row-major, moves in stored order Down, Right, Diag).
 * Stops after max_hits.  snprintf-style: returns the bytes the text needs; -1 on OOM. */
int64_t sao_linear_print_hits(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, int local,
                              uint64_t max_hits, char* buf, size_t buf_cap, uint64_t* n_printed);

/* ------------------------------------------------------------------------------------- */
/* parse_fasta, parse.rs:54-99.  Returns number of records (>=0), or -1 for FastaError
 * (bad extension / unreadable).  Records are written as (name_off,name_len,seq_off,seq_len)
 * quadruples into `index` (capacity index_cap records) over `out` (capacity >= file size).
 * err_chars receives the rejected bytes in order (CharError, parse.rs:92-96). */
int64_t sao_parse_fasta_path(const char* path, uint8_t* out, size_t out_cap, uint64_t* index,
                             size_t index_cap, uint8_t* err_chars, size_t err_cap,
                             size_t* n_err);
int64_t sao_parse_fasta_mem(const uint8_t* contents, size_t n, uint8_t* out, size_t out_cap,
                            uint64_t* index, size_t index_cap, uint8_t* err_chars,
                            size_t err_cap, size_t* n_err);

#ifdef __cplusplus
}
#endif
#endif
