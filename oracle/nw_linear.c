/*
 * nw_linear.c -- ORACLE (test infrastructure, never linked into the product).
 *
 * Literal CPU restatement of /root/reference/src/needleman_wunsch.rs (the single-matrix
 * "linear / pseudo-affine" Needleman-Wunsch; dead code at the reference commit, main.rs:4,14).
 * Line citations are into that file.
 *
 * Geometry (:38): scores[i][j], i in [0,n1] walks seq1 (query) = rows, j in [0,n2] walks
 * seq2 (db) = columns -- the TRANSPOSE of the affine file.
 */
#include "sa_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define GAP_OPENING (-8) /* :181-186 */
#define GAP_EXTENSION (-6)
#define MISMATCH (-4)
#define MATCH 5

#define MV_DOWN 1  /* push order Down, Right, Diag (:92-100) */
#define MV_RIGHT 2
#define MV_DIAG 4

typedef struct {
  uint32_t n1, n2;
  size_t w; /* n2 + 1 */
  int32_t* scores;
  uint8_t* moves; /* 3-bit move set */
  uint8_t* gaps;  /* Gap::Gap = 1 */
} lin_t;

static int lin_fill(lin_t* t, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                    int local) {
  t->n1 = n1;
  t->n2 = n2;
  t->w = (size_t)n2 + 1;
  const size_t cells = ((size_t)n1 + 1) * t->w;
  t->scores = (int32_t*)calloc(cells, sizeof(int32_t)); /* init :36-41 */
  t->moves = (uint8_t*)calloc(cells, 1);
  t->gaps = (uint8_t*)calloc(cells, 1);
  if (!t->scores || !t->moves || !t->gaps) return -1;
  const size_t w = t->w;
  if (!local) { /* :44-65; note scores[0][0] receives BOTH increments -> -16 */
    for (uint32_t j = 0; j <= n2; ++j) {
      t->scores[j] += (int32_t)j * GAP_EXTENSION + GAP_OPENING;
      t->moves[j] |= MV_RIGHT;
      t->gaps[j] = 1;
    }
    for (uint32_t i = 0; i <= n1; ++i) {
      t->scores[(size_t)i * w] += (int32_t)i * GAP_EXTENSION + GAP_OPENING;
      t->moves[(size_t)i * w] |= MV_DOWN;
      t->gaps[(size_t)i * w] = 1;
    }
  }
  for (uint32_t i = 1; i <= n1; ++i) { /* :66-103 */
    for (uint32_t j = 1; j <= n2; ++j) {
      const size_t c = (size_t)i * w + j;
      const int32_t diag = t->scores[c - w - 1] + (seq1[i - 1] == seq2[j - 1] ? MATCH : MISMATCH);
      const int32_t down = t->scores[c - w] + (t->gaps[c - w] ? GAP_EXTENSION : GAP_OPENING);
      const int32_t right = t->scores[c - 1] + (t->gaps[c - 1] ? GAP_EXTENSION : GAP_OPENING);
      int32_t mx = down > right ? down : right;
      if (diag > mx) mx = diag;
      if (mx == down || mx == right) t->gaps[c] = 1; /* :85-87, even when diag ties */
      if (local && mx < 0) {
        t->moves[c] = 0; /* :88-89; the score stays 0, the gap flag stays as set above */
      } else {
        t->scores[c] = mx;
        uint8_t mv = 0;
        if (mx == down) mv |= MV_DOWN;
        if (mx == right) mv |= MV_RIGHT;
        if (mx == diag) mv |= MV_DIAG;
        t->moves[c] = mv;
      }
    }
  }
  return 0;
}

static void lin_free(lin_t* t) {
  free(t->scores);
  free(t->moves);
  free(t->gaps);
}

static int64_t sat_add64(int64_t a, int64_t b) {
  int64_t r = a + b;
  return (r < a || r > (INT64_MAX / 4)) ? (INT64_MAX / 4) : r;
}

int sao_linear_align(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                     int local, sao_linear_result_t* out, uint32_t* cigar) {
  lin_t t;
  memset(&t, 0, sizeof(t));
  memset(out, 0, sizeof(*out));
  if (lin_fill(&t, seq1, n1, seq2, n2, local) != 0) {
    lin_free(&t);
    return -1;
  }
  const size_t w = t.w;
  /* #hits per start cell: get_next (:205-254) prints at (0,0) or at a cell with no moves */
  int64_t* cnt = (int64_t*)calloc(((size_t)n1 + 1) * w, sizeof(int64_t));
  if (!cnt) {
    lin_free(&t);
    return -1;
  }
  for (uint32_t i = 0; i <= n1; ++i)
    for (uint32_t j = 0; j <= n2; ++j) {
      const size_t c = (size_t)i * w + j;
      const uint8_t mv = t.moves[c];
      if ((i == 0 && j == 0) || mv == 0) {
        cnt[c] = 1;
        continue;
      }
      int64_t v = 0;
      if (mv & MV_DOWN) v = sat_add64(v, cnt[c - w]);
      if (mv & MV_RIGHT) v = sat_add64(v, cnt[c - 1]);
      if (mv & MV_DIAG) v = sat_add64(v, cnt[c - w - 1]);
      cnt[c] = v;
    }
  uint32_t si = n1, sj = n2;
  if (!local) {
    out->score = t.scores[(size_t)n1 * w + n2];
    out->n_hits = cnt[(size_t)n1 * w + n2];
  } else { /* argmax :256-272, row-major, every tie is a start (:107-115) */
    int32_t best = INT32_MIN;
    int first = 1;
    for (uint32_t i = 0; i <= n1; ++i)
      for (uint32_t j = 0; j <= n2; ++j) {
        const int32_t v = t.scores[(size_t)i * w + j];
        if (v > best) {
          best = v;
          out->n_hits = cnt[(size_t)i * w + j];
          si = i;
          sj = j;
          first = 0;
        } else if (v == best && !first) {
          out->n_hits = sat_add64(out->n_hits, cnt[(size_t)i * w + j]);
        }
      }
    out->score = best;
  }
  out->status = SAO_OK;
  /* first printed hit: at every cell take the first stored move (Down, Right, Diag) */
  uint32_t i = si, j = sj, n = 0, cols = 0;
  while (!(i == 0 && j == 0)) {
    const uint8_t mv = t.moves[(size_t)i * w + j];
    if (mv == 0) break;
    out->start1 = (i > 1 ? i : 1) - 1; /* :215-216 */
    out->start2 = (j > 1 ? j : 1) - 1;
    int op;
    if (mv & MV_DOWN) {
      op = SAO_OP_I;
      --i;
    } else if (mv & MV_RIGHT) {
      op = SAO_OP_D;
      --j;
    } else {
      op = SAO_OP_M;
      --i;
      --j;
    }
    ++cols;
    if (cigar) {
      if (n > 0 && (int)(cigar[n - 1] & 3u) == op) cigar[n - 1] += 4u;
      else cigar[n++] = 4u | (uint32_t)op;
    }
  }
  if (cigar)
    for (uint32_t a = 0, b = n ? n - 1 : 0; a < b; ++a, --b) {
      uint32_t tmp = cigar[a];
      cigar[a] = cigar[b];
      cigar[b] = tmp;
    }
  out->cigar_len = n;
  out->n_columns = cols;
  out->end1 = si;
  out->end2 = sj;
  free(cnt);
  lin_free(&t);
  return 0;
}

int sao_linear_matrices(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                        int local, int32_t* scores, uint8_t* moves, uint8_t* gaps) {
  lin_t t;
  memset(&t, 0, sizeof(t));
  if (lin_fill(&t, seq1, n1, seq2, n2, local) != 0) {
    lin_free(&t);
    return -1;
  }
  const size_t cells = ((size_t)n1 + 1) * t.w;
  if (scores) memcpy(scores, t.scores, cells * sizeof(int32_t));
  if (moves) memcpy(moves, t.moves, cells);
  if (gaps) memcpy(gaps, t.gaps, cells);
  lin_free(&t);
  return 0;
}

typedef struct {
  const uint8_t* residues;
  const uint64_t *q_off, *d_off;
  const uint32_t *q_len, *d_len;
  int32_t* score;
  uint8_t* status;
  uint32_t* cigar_len;
  uint32_t* cigar_pool;
  uint32_t cigar_stride;
  uint64_t lo, hi;
  int rc;
  int local;
  uint32_t *end1, *end2;
} lin_job_t;

static void* lin_worker(void* arg) {
  lin_job_t* j = (lin_job_t*)arg;
  uint32_t* tmp = NULL;
  size_t cap = 0;
  for (uint64_t p = j->lo; p < j->hi; ++p) {
    const uint32_t n1 = j->q_len[p], n2 = j->d_len[p];
    sao_linear_result_t r;
    uint32_t* cig = NULL;
    if (j->cigar_pool) {
      if ((size_t)n1 + n2 + 1 > cap) {
        free(tmp);
        cap = (size_t)n1 + n2 + 1;
        tmp = (uint32_t*)malloc(cap * sizeof(uint32_t));
        if (!tmp) { j->rc = -1; return NULL; }
      }
      cig = tmp;
    }
    if (sao_linear_align(j->residues + j->q_off[p], n1, j->residues + j->d_off[p], n2, j->local, &r,
                         cig) != 0) {
      j->rc = -1;
      break;
    }
    j->score[p] = r.score;
    if (j->status) j->status[p] = (uint8_t)r.status;
    if (j->cigar_len) j->cigar_len[p] = r.cigar_len;
    if (j->end1) j->end1[p] = r.end1;
    if (j->end2) j->end2[p] = r.end2;
    if (j->cigar_pool) {
      uint32_t n = r.cigar_len < j->cigar_stride ? r.cigar_len : j->cigar_stride;
      memcpy(j->cigar_pool + (size_t)p * j->cigar_stride, cig, n * sizeof(uint32_t));
    }
  }
  free(tmp);
  return NULL;
}

int sao_linear_batch(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                     const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs,
                     int32_t* score, uint8_t* status, uint32_t* cigar_len, uint32_t* cigar_pool,
                     uint32_t cigar_stride, int n_threads) {
  return sao_linear_batch_ex(residues, q_off, q_len, d_off, d_len, n_pairs, 0, score, status, cigar_len,
                             cigar_pool, cigar_stride, NULL, NULL, n_threads);
}

int sao_linear_batch_ex(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                        const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs, int local,
                        int32_t* score, uint8_t* status, uint32_t* cigar_len, uint32_t* cigar_pool,
                        uint32_t cigar_stride, uint32_t* end1, uint32_t* end2, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > n_pairs) n_threads = n_pairs ? (int)n_pairs : 1;
  lin_job_t* jobs = (lin_job_t*)calloc((size_t)n_threads, sizeof(lin_job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  if (!jobs || !th) { free(jobs); free(th); return -1; }
  int rc = 0;
  for (int k = 0; k < n_threads; ++k) {
    jobs[k] = (lin_job_t){residues, q_off, d_off, q_len, d_len, score, status, cigar_len,
                          cigar_pool, cigar_stride, n_pairs * (uint64_t)k / (uint64_t)n_threads,
                          n_pairs * (uint64_t)(k + 1) / (uint64_t)n_threads, 0, local, end1, end2};
    if (n_threads == 1) lin_worker(&jobs[k]);
    else pthread_create(&th[k], NULL, lin_worker, &jobs[k]);
  }
  for (int k = 0; k < n_threads; ++k) {
    if (n_threads > 1) pthread_join(th[k], NULL);
    if (jobs[k].rc) rc = -1;
  }
  free(jobs);
  free(th);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * The reference's stdout for the hits of one pair (:106-116, :155-178, :205-254), literally:
 * an explicit-stack version of the recursion get_next, same visiting order.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  char* buf;
  size_t cap, len;
} txt_t;

static void txt_put(txt_t* t, const char* s, size_t n) {
  for (size_t k = 0; k < n; ++k) {
    if (t->buf && t->len + 1 < t->cap) t->buf[t->len] = s[k];
    t->len++;
  }
}
static void txt_str(txt_t* t, const char* s) { txt_put(t, s, strlen(s)); }
static void txt_u(txt_t* t, uint32_t v) {
  char tmp[16];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) txt_put(t, &tmp[--n], 1);
}

/* println!("\nHit: {}\n", hit) with Display for Hit (:155-178); q/d hold the columns in push
 * order (end of the alignment first), the Display reverses them */
static void print_hit(txt_t* t, const char* q, const char* d, size_t n, uint32_t s1, uint32_t s2) {
  txt_str(t, "\nHit: ");
  txt_str(t, "\nseq1: ");
  for (size_t k = n; k-- > 0;) txt_put(t, &q[k], 1);
  txt_str(t, "\n      ");
  for (size_t k = n; k-- > 0;) txt_put(t, q[k] == d[k] ? "|" : " ", 1);
  txt_str(t, "\nseq2: ");
  for (size_t k = n; k-- > 0;) txt_put(t, &d[k], 1);
  txt_str(t, "\n");
  txt_str(t, "start in seq1: ");
  txt_u(t, s1);
  txt_str(t, "\nstart in seq2: ");
  txt_u(t, s2);
  txt_str(t, "\n\n"); /* writeln!(.."\n") of the Display */
  txt_str(t, "\n\n"); /* the "\n" after {} and println's newline */
}

int64_t sao_linear_print_hits(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, int local,
                              uint64_t max_hits, char* buf, size_t buf_cap, uint64_t* n_printed) {
  lin_t t;
  memset(&t, 0, sizeof(t));
  if (lin_fill(&t, seq1, n1, seq2, n2, local) != 0) {
    lin_free(&t);
    return -1;
  }
  const size_t w = t.w;
  txt_t out = {buf, buf_cap, 0};
  uint64_t printed = 0;
  /* start cells (:107-111) */
  size_t n_starts = 0;
  uint32_t* starts = (uint32_t*)malloc(2 * sizeof(uint32_t) * ((size_t)n1 + 1) * w);
  const size_t depth = (size_t)n1 + n2 + 2;
  char* q = (char*)malloc(depth);
  char* d = (char*)malloc(depth);
  /* frame: the cell and the index of the next move to try (0 Down, 1 Right, 2 Diag) */
  uint32_t* fi = (uint32_t*)malloc(depth * sizeof(uint32_t));
  uint32_t* fj = (uint32_t*)malloc(depth * sizeof(uint32_t));
  uint8_t* fk = (uint8_t*)malloc(depth);
  if (!starts || !q || !d || !fi || !fj || !fk) {
    free(starts); free(q); free(d); free(fi); free(fj); free(fk);
    lin_free(&t);
    return -1;
  }
  if (!local) {
    starts[0] = n1;
    starts[1] = n2;
    n_starts = 1;
  } else {
    int32_t best = INT32_MIN;
    for (uint32_t i = 0; i <= n1; ++i)
      for (uint32_t j = 0; j <= n2; ++j) {
        const int32_t v = t.scores[(size_t)i * w + j];
        if (v > best) {
          best = v;
          n_starts = 0;
        }
        if (v == best) {
          starts[2 * n_starts] = i;
          starts[2 * n_starts + 1] = j;
          ++n_starts;
        }
      }
  }
  for (size_t s = 0; s < n_starts && printed < max_hits; ++s) {
    uint32_t hs1 = 0, hs2 = 0; /* Hit::default() per start cell (:113) */
    size_t sp = 0, cols = 0;
    fi[0] = starts[2 * s];
    fj[0] = starts[2 * s + 1];
    fk[0] = 0;
    sp = 1;
    while (sp && printed < max_hits) {
      const uint32_t i = fi[sp - 1], j = fj[sp - 1];
      const uint8_t mv = t.moves[(size_t)i * w + j];
      if (fk[sp - 1] == 0 && ((i == 0 && j == 0) || mv == 0)) { /* :206-213 */
        print_hit(&out, q, d, cols, hs1, hs2);
        ++printed;
        --sp;
        if (sp) --cols; /* the caller's hit.query.pop() / hit.db.pop() (:251-252) */
        continue;
      }
      /* next stored move, in push order Down, Right, Diag (:92-100) */
      int k = fk[sp - 1];
      while (k < 3 && !(mv & (1u << k))) ++k;
      if (k == 3) {
        --sp;
        if (sp) --cols;
        continue;
      }
      fk[sp - 1] = (uint8_t)(k + 1);
      hs1 = (i > 1 ? i : 1) - 1; /* :215-216 */
      hs2 = (j > 1 ? j : 1) - 1;
      uint32_t ni = i, nj = j;
      if (k == 0) { q[cols] = (char)seq1[i - 1]; d[cols] = '-'; --ni; }
      else if (k == 1) { q[cols] = '-'; d[cols] = (char)seq2[j - 1]; --nj; }
      else { q[cols] = (char)seq1[i - 1]; d[cols] = (char)seq2[j - 1]; --ni; --nj; }
      ++cols;
      fi[sp] = ni;
      fj[sp] = nj;
      fk[sp] = 0;
      ++sp;
    }
  }
  if (out.buf && out.cap) out.buf[out.len < out.cap ? out.len : out.cap - 1] = 0;
  if (n_printed) *n_printed = printed;
  free(starts); free(q); free(d); free(fi); free(fj); free(fk);
  lin_free(&t);
  return (int64_t)out.len;
}
