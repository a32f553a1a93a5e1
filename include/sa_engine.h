/*
 * sa_engine.h -- C ABI of the B200-native batched pairwise-alignment engine.
 *
 * Drop-in boundary for the data-parallel hot path of Qw11111111111/SequenceAligning
 * (crate a_star_align): the `for d in db { for q in query { match algo {..} } }` loop of
 * src/main.rs:61-79, i.e. the calls
 *     n_w_align(q, d, verbose, mode)   src/needleman_wunsch_affine.rs:424   (SA_ALGO_NW_AFFINE)
 *     n_w_align(q, d, verbose, local)  src/needleman_wunsch.rs:180          (SA_ALGO_NW_LINEAR)
 *     wfa_align(q, d, mode)            src/wfa.rs:23                        (SA_ALGO_WFA)
 * batched over many independent (query, db) pairs.  The reference returns nothing (results
 * exist only as stdout text); this ABI returns, per pair, what that text is made of: a
 * status, the score, and the FIRST alignment the reference prints (as a run-length CIGAR).
 *
 * Plain C: opaque handle, plain pointers and sizes, no C++/torch types, no exceptions.
 * A Rust `extern "C"` block / bindgen, cgo, or ctypes binds exactly these symbols; see
 * INTEGRATION.md for the Rust shim a maintainer of the reference would add.
 *
 * There is NO CPU fallback: every entry point that computes requires a CUDA device of
 * compute capability 10.0 (B200) and fails with SA_E_CUDA otherwise.
 *
 * Conventions (src/main.rs:61-66): seq1 = QUERY record, seq2 = DB record,
 * n1 = len(seq1), n2 = len(seq2).  Residues are raw bytes compared with `==`
 * (needleman_wunsch_affine.rs:220), so 'N' == 'N' matches and 'N' != 'A'.
 */
#ifndef SA_ENGINE_H
#define SA_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SA_ABI_VERSION 3

typedef struct sa_engine sa_engine_t; /* opaque; one per process and device SET (sa_engine_create_multi) */

/* parse.rs:36-42 `enum Algo` (A* is out of scope for the GPU path) */
typedef enum {
  SA_ALGO_NW_AFFINE = 0,
  SA_ALGO_NW_LINEAR = 1,
  SA_ALGO_WFA = 2,          /* wfa_align exactly as the reference executes it (score + status) */
  SA_ALGO_WFA_STANDARD = 3  /* extension: textbook gap-affine WFA cost with the same penalties;
                               the reference's own WFA panics or never converges on inputs
                               beyond a few dozen residues (wfa.rs:189, :577, :603)            */
} sa_algo_t;
/* parse.rs:44-50 `enum Mode` */
typedef enum { SA_MODE_GLOBAL = 0, SA_MODE_LOCAL = 1, SA_MODE_SEMIGLOBAL = 2 } sa_mode_t;

/* Per-pair status (>= 0, stored in sa_result_t.status) and per-call errors (< 0). */
typedef enum {
  SA_OK = 0,                 /* the reference completes; score and alignment valid            */
  SA_REF_PANIC = 1,          /* the reference prints >= 1 alignment, then panics
                                (index 0-1 at nw_affine:299/:303); score and alignment valid  */
  SA_REF_NO_CONVERGENCE = 2, /* WFA: `while is_converged().is_none()` never ends (wfa.rs:28)  */
  SA_NOT_IMPLEMENTED = 3,    /* Err(AlignmentError("not implemented")) nw_affine:433-434,
                                wfa.rs:26 -- non-global modes                                 */
  SA_REF_PANIC_EARLY = 4,    /* the reference panics before printing anything: the first
                                alignment in DFS order starts with a gap; score valid,
                                alignment empty                                               */
  SA_REF_NO_OUTPUT = 5,      /* completes but prints no alignment (only possible once the
                                finite -32768 sentinel leaks, n1+n2 >~ 5.4k)                  */
  SA_ALIGNMENT_OMITTED = 0x80, /* flag ORed into a per-pair status: the pair went through the long-pair
                                  kernel and its traceback matrix did not fit the scratch budget;
                                  score and status are exact, cigar_len is 0                      */
  SA_E_CUDA = -1,            /* CUDA failure (sa_last_error has the string)                   */
  SA_E_ARG = -2,             /* bad argument                                                  */
  SA_E_NOMEM = -3,           /* host or device allocation failed                              */
  SA_E_CIGAR_CAPACITY = -4,  /* cigar_capacity too small; sa_result_t.cigar_used = needed     */
  SA_E_UNSUPPORTED = -5      /* pair shape/scheme outside what the kernels implement          */
} sa_status_t;

/* Scoring scheme; defaults are the reference's compile-time consts
 * (needleman_wunsch_affine.rs:15-20, needleman_wunsch.rs:181-186): 5, -4, -8, -6.
 * For SA_ALGO_WFA the fields are penalties mismatch=4, gap_open=2, gap_ext=6, match=0
 * (wfa.rs:17-21). Pass NULL for the reference's constants. */
typedef struct {
  int32_t match, mismatch, gap_open, gap_ext;
} sa_scheme_t;

/* CIGAR: 32-bit words (len << 2) | op in alignment order (first column first).
 *   SA_OP_M  diagonal column  seq1[y-1] / seq2[x-1]   (match or mismatch; state InM)
 *   SA_OP_I  seq1[y-1] / '-'   consumes a QUERY residue (state InI, nw_affine:302-306)
 *   SA_OP_D  '-' / seq2[x-1]   consumes a DB residue    (state InD, nw_affine:297-301) */
enum { SA_OP_M = 0, SA_OP_I = 1, SA_OP_D = 2 };

/* One batch of pairs, caller-owned HOST memory (pinned via sa_alloc_pinned for overlap;
 * pageable works but copies synchronously).  Pair p aligns
 *   seq1 = residues[q_off[p] .. q_off[p]+q_len[p])   (query)
 *   seq2 = residues[d_off[p] .. d_off[p]+d_len[p])   (db)
 * Offsets may alias (1 query x N db) and need not be ordered; throughput is best when
 * consecutive pairs have similar lengths (length-bucketed, which sa_pack_* produce). */
typedef struct {
  const uint8_t* residues;
  uint64_t residues_len;
  const uint64_t* q_off;
  const uint32_t* q_len;
  const uint64_t* d_off;
  const uint32_t* d_len;
  uint64_t n_pairs;
  uint32_t packing; /* 0 = one byte per residue, offsets in bytes (raw bytes are compared, so
                       'N' == 'N').  1 = 2-bit codes A=0 C=1 G=2 T=3, four per byte, residue i of
                       the buffer at bits 2*(i&3) of byte i>>2; q_off/d_off count RESIDUES,
                       residues_len counts BYTES (see sa_pack_2bit).  A/C/G/T only.           */
} sa_batch_t;

/* Results, caller-allocated HOST memory, engine fills.  Arrays have n_pairs entries. */
typedef struct {
  int32_t* score;      /* affine: max(I,D,M)[n2][n1] (nw_affine:247-250); linear: scores[n1][n2];
                          wfa: the printed `converged with score` value (wfa.rs:31-36)      */
  uint8_t* status;     /* sa_status_t >= 0                                                   */
  uint64_t* cigar_off; /* offset (words) of pair p's CIGAR in `cigar`; monotone in p         */
  uint32_t* cigar_len; /* words; 0 when the reference prints no alignment                    */
  uint32_t* cigar;     /* pool of run-length words; may be NULL with cigar_capacity = 0 to
                          skip traceback (score + status only)                               */
  uint64_t cigar_capacity; /* words available in `cigar`                                     */
  uint64_t cigar_used;     /* OUT: words written (or needed, on SA_E_CIGAR_CAPACITY)         */
  uint32_t* end1;      /* optional (may be NULL): the cell the traceback STARTS from = the end of
                          the alignment, as (residues of seq1, residues of seq2) consumed up to
                          and including it.  Global modes: (n1, n2).  SA_MODE_LOCAL (linear NW):
                          the first cell, in row-major order, that holds the matrix maximum
                          (needleman_wunsch.rs:107-111, :256-272); the alignment covers
                          seq1[end1 - I - M, end1) and seq2[end2 - D - M, end2), the residues
                          the CIGAR's I+M and D+M columns consume                            */
  uint32_t* end2;
} sa_result_t;

/* Time breakdown of the last call: device milliseconds from CUDA events on the launching streams,
 * except wall_ms.  On a multi-device engine: the slowest device's times, sums of the counters. */
typedef struct {
  double kernels_ms;      /* first to last kernel of the call                                 */
  double fill_ms;         /* packed fill kernels alone (nw_affine_fill_s16), summed over segments */
  double long_fwd_ms;     /* tiled long-pair forward launches (nw_long_fwd), summed over waves */
  double long_back_ms;    /* their classification + traceback kernels (nw_long_back)          */
  double wall_ms;         /* host wall clock of sa_align_batch (copies included)              */
  uint64_t cells;         /* sum over pairs of n1*n2                                          */
  uint64_t kernel_launches;
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t pairs_rerun;   /* pairs re-filled without the panic-detection bonus                */
  uint64_t pairs_fallback; /* long pairs handed from the tiled path to the literal kernel     */
  uint64_t wfa_cells;     /* SA_ALGO_WFA_STANDARD: wavefront cells computed    */
  uint64_t wfa_extended;  /* .. and residues passed by the extend step                        */
} sa_timing_t;

/* -- lifecycle ------------------------------------------------------------------------ */
sa_status_t sa_engine_create(int device_id, sa_engine_t** out);
/* One engine over several GPUs of the box.  sa_align_batch then shards ONE pair
 * list -- what the loop of src/main.rs:61-62 iterates -- over the devices, balanced on n1*n2,
 * with one worker thread, one set of streams and one scratch budget per device, and puts every
 * result back in input order; there is no collective: pairs are independent.  The same device id
 * may be listed more than once (several pipelines on one GPU).  The device-resident entry
 * points (sa_batch_upload ..) stay single-device and return SA_E_UNSUPPORTED here. */
sa_status_t sa_engine_create_multi(const int* device_ids, int n_devices, sa_engine_t** out);
int sa_engine_device_count(const sa_engine_t* e); /* 1 for sa_engine_create */
sa_status_t sa_engine_destroy(sa_engine_t* e);
const char* sa_last_error(const sa_engine_t* e); /* never NULL; "" when none */
int sa_abi_version(void);

/* -- the hot path --------------------------------------------------------------------- */
/* Align every pair of `batch` (host buffers in, host buffers out).  Blocking.
 * On a multi-device engine the pair list is sharded over the devices inside this call
 * (sa_plan_shards says how) and the results come back in input order: per-pair arrays exactly as
 * from one device; the CIGAR pool holds one region per device, so cigar_off stays monotone in p
 * but may skip words between regions, and cigar_used is the end of the last region.  When a
 * region overflows the call returns SA_E_CIGAR_CAPACITY with cigar_used = a capacity that fits. */
sa_status_t sa_align_batch(sa_engine_t* e, sa_algo_t algo, sa_mode_t mode,
                           const sa_scheme_t* scheme, const sa_batch_t* batch,
                           sa_result_t* result);

/* Device-resident variant: upload once, align many times (benchmarks, repeated scoring). */
typedef struct sa_resident sa_resident_t;
sa_status_t sa_batch_upload(sa_engine_t* e, const sa_batch_t* batch, sa_resident_t** out);
sa_status_t sa_batch_free(sa_engine_t* e, sa_resident_t* r);
/* Runs the kernels on the resident batch; results stay on the device until
 * sa_resident_download.  Asynchronous w.r.t. the host: returns after enqueueing. */
sa_status_t sa_align_resident(sa_engine_t* e, sa_algo_t algo, sa_mode_t mode,
                              const sa_scheme_t* scheme, sa_resident_t* r, int want_cigar);
sa_status_t sa_resident_download(sa_engine_t* e, sa_resident_t* r, sa_result_t* result);
sa_status_t sa_engine_synchronize(sa_engine_t* e);
/* cudaStream_t the engine launches on, as void* (for CUDA-event timing by the caller). */
void* sa_engine_stream(sa_engine_t* e);

sa_status_t sa_last_timing(const sa_engine_t* e, sa_timing_t* out);

/* -- host helpers --------------------------------------------------------------------- */
void* sa_alloc_pinned(size_t bytes);
void sa_free_pinned(void* p);
/* Page-locks memory the caller already owns (a parser's output, an mmap) so that sa_align_batch
 * streams from / to it like from sa_alloc_pinned memory; undo with sa_host_unregister. */
sa_status_t sa_host_register(void* p, size_t bytes);
sa_status_t sa_host_unregister(void* p);

/* How a multi-device call was sharded (one entry per device of the engine, in creation order). */
typedef struct {
  int32_t device;      /* CUDA device id                                                      */
  uint32_t contiguous; /* 1: the shard is the pair range [first_pair, first_pair + pairs)     */
  uint64_t first_pair, pairs, cells;
  uint64_t h2d_bytes, d2h_bytes, kernel_launches;
  double device_ms;    /* first to last kernel of the shard (CUDA events)                     */
  double host_ms;      /* wall clock of the shard's worker, submit to results on the host     */
} sa_shard_info_t;
sa_status_t sa_last_shards(const sa_engine_t* e, sa_shard_info_t* out, int cap, int* n_out);

/* The shard plan of a multi-device call, pure host code.  Pairs are weighted by n1*n2 + 1.
 * begin[0..n_parts] : the cell-balanced CONTIGUOUS split (part k = pairs [begin[k], begin[k+1])),
 *                     used when no part exceeds the mean weight by more than 2 % -- always the
 *                     case for many small pairs; it needs no gather and keeps pair order.
 * part[0..n_pairs)  : otherwise (few, uneven pairs) the greedy LPT assignment of
 *                     sa_partition_lpt; *contiguous says which of the two a call would use.
 * part may be NULL (then only begin and *contiguous are produced). */
sa_status_t sa_plan_shards(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs, int n_parts,
                           uint64_t* begin, int32_t* part, int* contiguous);

/* Greedy longest-processing-time partition of pairs over n_parts GPUs by n1*n2 cells
 * (SURVEY 8e).  part[p] in [0, n_parts).  Deterministic.  Pure host code. */
sa_status_t sa_partition_lpt(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs,
                             int n_parts, int32_t* part);

/* parse_fasta (src/parse.rs:54-99), byte-for-byte semantics incl. the quirks pinned by
 * parse.rs:166-251.  Returns #records (>= 0) or SA_E_ARG for the reference's FastaError.
 * index holds (name_off, name_len, seq_off, seq_len) per record into `out`. */
int64_t sa_parse_fasta(const char* path, uint8_t* out, size_t out_cap, uint64_t* index,
                       size_t index_cap, uint8_t* err_chars, size_t err_cap, size_t* n_err);

/* EVERY co-optimal alignment of one pair, in the reference's order and text: the LIFO DFS of
 * needleman_wunsch_affine.rs:246-329 over the per-cell parent lists (:96-153), which the device
 * computes (7 bits per cell); "alignment found\n\nseq1: ..\n      ..\nseq2: ..\n" per path.
 * Stops after max_alignments, or where the reference panics (*panicked = 1; the text so far is
 * what the reference had printed).  snprintf-style: returns the bytes needed, < 0 on error. */
int64_t sa_affine_all_alignments(sa_engine_t* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2,
                                 uint32_t n2, const sa_scheme_t* scheme, uint64_t max_alignments,
                                 char* buf, size_t cap, uint64_t* n_printed, int32_t* panicked);

/* Number of co-optimal alignments per pair = how many alignments the reference's traceback
 * (needleman_wunsch_affine.rs:246-329) prints for the pair when nothing panics: the number of
 * parent-list paths (:96-153) from the best end states (:247-280) to the origin.  Paths into the
 * boundary chains (where the reference panics, :299/:303) are not counted; saturates at
 * INT64_MAX / 4.  counts: host array of n_pairs.  Exact 32-bit recurrences, one thread per pair:
 * a side API (about 20x the cost of sa_align_batch), not the hot path. */
sa_status_t sa_affine_count_cooptimal(sa_engine_t* e, const sa_scheme_t* scheme, const sa_batch_t* batch,
                                      int64_t* counts);

/* Packer for packing = 1: appends n residues (A/C/G/T) to dst starting at residue index dst_pos.
 * SA_E_ARG at the first other byte.  Pure host code. */
sa_status_t sa_pack_2bit(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_pos);
/* The same for a whole buffer on n_threads host threads (0 = all): src[0, n) -> dst[0, (n+3)/4). */
sa_status_t sa_pack_2bit_mt(const uint8_t* src, uint64_t n, uint8_t* dst, int n_threads);

/* parse_fasta with the packer fused in (the north star's "parse.rs gains a packer"): as
 * sa_parse_fasta, and `packed` (>= (out bytes + 3) / 4 bytes; pinned memory for the engine to
 * stream from) receives the 2-bit codes of the WHOLE output buffer -- byte i of `out` at bits
 * 2*(i&3) of packed[i>>2], header bytes and 'N' coding as 0 -- so the (seq_off, seq_len) of
 * `index` address both formats.  *all_acgt = 1 when no record sequence holds an 'N': only then
 * may a batch built on `packed` use packing = 1.  *out_len = bytes written to `out`.
 * Chunk-parallel like sa_parse_fasta (SA_HOST_THREADS, default all cores, at most 32). */
int64_t sa_parse_fasta_packed(const char* path, uint8_t* out, size_t out_cap, uint64_t* index, size_t index_cap,
                              uint8_t* err_chars, size_t err_cap, size_t* n_err, uint8_t* packed, size_t packed_cap,
                              int* all_acgt, uint64_t* out_len);

/* The text the reference prints for one alignment (needleman_wunsch_affine.rs:283-286 and
 * Display :390-411): "alignment found\n\nseq1: ..\n      ..\nseq2: ..\n".  snprintf-style:
 * returns the bytes needed; SA_E_ARG if the CIGAR does not fit the sequences. */
int64_t sa_render_affine(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                         const uint32_t* cigar, uint32_t cigar_len, char* buf, size_t cap);

/* EVERY hit the reference's linear aligner prints for one pair, in its order and text (needleman_wunsch.rs:106-116,
 * :205-254): start cells = (n1, n2), or in local mode every cell holding the matrix maximum in row-major order
 * (:256-272); per start cell the pre-order walk over the stored moves (Down, Right, Diag); a hit at (0, 0) or at a
 * cell without moves, rendered as by sa_render_linear_hit.  The device fills the matrix (scores and move sets), the
 * host walks it.  Stops after max_hits.  snprintf-style: returns the bytes needed, < 0 on error. */
int64_t sa_linear_all_hits(sa_engine_t* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, int local,
                           const sa_scheme_t* scheme, uint64_t max_hits, char* buf, size_t cap, uint64_t* n_printed);

/* The reference's COMPLETE stdout for one pair under `-a wfa` (wfa_align, wfa.rs:23-42; SURVEY.md
 * App. A.2), from a traced run of the literal kernel on the device: one `lo: .., hi: ..` line per
 * created wavefront (:251); when the loop converges, `converged with score N: ` (:36), the `huhu`
 * block with the converged element (:650, Debug :104-116), the `yeah / well shit / extend / open /
 * huh` lines of rec_tr (:653-853) and the two prints of the (always empty) Alignment (:38-39).
 * *status = SA_OK, SA_REF_PANIC (the text is what had been printed before the panic of
 * :577/:603) or SA_REF_NO_CONVERGENCE (the first lines of a never-ending output).
 * snprintf-style: returns the bytes needed, < 0 on error. */
int64_t sa_wfa_reference_stdout(sa_engine_t* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                                char* buf, size_t cap, int32_t* status);

/* The text the reference's linear aligner prints for one hit (needleman_wunsch.rs:207/:211
 * `println!("\nHit: {}\n", hit)`, Display for Hit :155-178, start_in_query/db :215-216):
 * "\nHit: \nseq1: ..\n      ..\nseq2: ..\nstart in seq1: a\nstart in seq2: b\n\n\n\n".
 * (end1, end2) = sa_result_t.end1/end2 of the pair.  snprintf-style; SA_E_ARG on a misfit. */
int64_t sa_render_linear_hit(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                             const uint32_t* cigar, uint32_t cigar_len, uint32_t end1, uint32_t end2,
                             char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SA_ENGINE_H */
