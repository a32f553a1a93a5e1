/*
 * nw_affine.c -- ORACLE (test infrastructure, never linked into the product).
 *
 * Literal CPU restatement of the reference's affine-gap global Needleman-Wunsch,
 * /root/reference/src/needleman_wunsch_affine.rs.  Line citations below are into that file.
 *
 * Geometry (:67-74, :428): matrices are [x][y] with x in [0,n2] walking seq2 (db) and
 * y in [0,n1] walking seq1 (query).  Three states: M (diagonal), I (consumes seq1[y-1]
 * against '-'), D (consumes seq2[x-1] against '-').  "-inf" is the finite i16::MIN (-32768).
 *
 * What the reference does with a pair is a PROCESS (fill, then a LIFO DFS that prints every
 * co-optimal alignment and may panic on an out-of-range index).  This file models that
 * process exactly, including the panics, by computing for every cell, during the fill:
 *   parents   the parent LIST of each state's cell as a 7-bit set, in the reference's push
 *             order (:96-153)
 *   fe        "first event" of the DFS subtree rooted at the cell: NONE / PRINT / PANIC
 *   taint     whether ANY cell of the subtree panics
 * from which status, the first printed ("canonical") alignment and the panic flag follow
 * without enumerating the (possibly exponential) set of paths.
 */
#include "sa_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF ((int32_t)-32768) /* i16::MIN as i32, :174 */

/* parent-set bits; the numeric order inside each state is the reference's PUSH order */
#define P_M_FROM_M 0x01 /* :122-131 */
#define P_M_FROM_I 0x02 /* :132-141 */
#define P_M_FROM_D 0x04 /* :142-151 */
#define P_I_FROM_I 0x08 /* :110-112  (extension) */
#define P_I_FROM_M 0x10 /* :113-117  (opening)   */
#define P_D_FROM_D 0x20 /* :98-100   (extension) */
#define P_D_FROM_M 0x40 /* :101-105  (opening)   */

enum { ST_M = 0, ST_D = 1, ST_I = 2 }; /* enum State :366-371 */
enum { FE_NONE = 0, FE_PRINT = 1, FE_PANIC = 2 };

static inline int32_t max2(int32_t a, int32_t b) { return a > b ? a : b; }

typedef struct {
  uint32_t n1, n2;
  size_t w; /* n1 + 1 */
  int32_t *m, *i, *d;
  uint8_t* par;   /* 7-bit parent sets */
  uint8_t* fe;    /* bits 0-1: M, 2-3: D, 4-5: I */
  uint8_t* taint; /* bit0 M, bit1 D, bit2 I */
} tensor_t;

static void tensor_free(tensor_t* t) {
  free(t->m);
  free(t->i);
  free(t->d);
  free(t->par);
  free(t->fe);
  free(t->taint);
}

static int tensor_alloc(tensor_t* t, uint32_t n1, uint32_t n2, int with_dfs_info) {
  memset(t, 0, sizeof(*t));
  t->n1 = n1;
  t->n2 = n2;
  t->w = (size_t)n1 + 1;
  size_t cells = ((size_t)n2 + 1) * t->w;
  t->m = (int32_t*)malloc(cells * sizeof(int32_t));
  t->i = (int32_t*)malloc(cells * sizeof(int32_t));
  t->d = (int32_t*)malloc(cells * sizeof(int32_t));
  t->par = (uint8_t*)calloc(cells, 1);
  if (with_dfs_info) {
    t->fe = (uint8_t*)calloc(cells, 1);
    t->taint = (uint8_t*)calloc(cells, 1);
  }
  if (!t->m || !t->i || !t->d || !t->par || (with_dfs_info && (!t->fe || !t->taint))) {
    tensor_free(t);
    return -1;
  }
  return 0;
}

static inline int fe_get(const tensor_t* t, size_t c, int st) { return (t->fe[c] >> (2 * st)) & 3; }
static inline void fe_set(tensor_t* t, size_t c, int st, int v) {
  t->fe[c] = (uint8_t)((t->fe[c] & ~(3 << (2 * st))) | (v << (2 * st)));
}
static inline int taint_get(const tensor_t* t, size_t c, int st) { return (t->taint[c] >> st) & 1; }

/* Would the DFS loop body (:287-328) panic when it expands (state, x, y)?  It indexes
 * seq1[y-1] in states InM/InI (:293,:303) and seq2[x-1] in states InM/InD (:294,:299) with
 * usize arithmetic, which panics (debug: overflow, release: out of bounds) when the index
 * underflows.  The body only runs when the cell has at least one parent. */
static inline int expand_panics(int st, uint32_t x, uint32_t y) {
  if (st == ST_M) return x == 0 || y == 0;
  if (st == ST_D) return x == 0;
  return y == 0;
}

/* DFS bookkeeping for one (state, cell): visit parents in REVERSE push order, because the
 * stack is LIFO (:281 `queue.pop()`), so the last pushed parent's subtree is explored first. */
static void dfs_info(tensor_t* t, int st, uint32_t x, uint32_t y) {
  size_t c = (size_t)x * t->w + y;
  uint8_t p = t->par[c];
  int fe = FE_NONE, taint = 0, has_parents;
  /* parent coordinates and (state, mask) in reverse push order */
  int ord_st[3];
  size_t ord_c[3];
  int n = 0;
  if (st == ST_M) {
    has_parents = (p & (P_M_FROM_M | P_M_FROM_I | P_M_FROM_D)) != 0;
    if (has_parents && !expand_panics(st, x, y)) {
      size_t pc = (size_t)(x - 1) * t->w + (y - 1);
      if (p & P_M_FROM_D) { ord_st[n] = ST_D; ord_c[n++] = pc; }
      if (p & P_M_FROM_I) { ord_st[n] = ST_I; ord_c[n++] = pc; }
      if (p & P_M_FROM_M) { ord_st[n] = ST_M; ord_c[n++] = pc; }
    }
  } else if (st == ST_I) {
    has_parents = (p & (P_I_FROM_I | P_I_FROM_M)) != 0;
    if (has_parents && !expand_panics(st, x, y)) {
      size_t pc = c - 1;
      if (p & P_I_FROM_M) { ord_st[n] = ST_M; ord_c[n++] = pc; }
      if (p & P_I_FROM_I) { ord_st[n] = ST_I; ord_c[n++] = pc; }
    }
  } else {
    has_parents = (p & (P_D_FROM_D | P_D_FROM_M)) != 0;
    if (has_parents && !expand_panics(st, x, y)) {
      size_t pc = c - t->w;
      if (p & P_D_FROM_M) { ord_st[n] = ST_M; ord_c[n++] = pc; }
      if (p & P_D_FROM_D) { ord_st[n] = ST_D; ord_c[n++] = pc; }
    }
  }
  if (x == 0 && y == 0) fe = FE_PRINT; /* :283-286, checked before the parent loop */
  if (has_parents && expand_panics(st, x, y)) {
    if (fe == FE_NONE) fe = FE_PANIC;
    taint = 1;
  } else {
    for (int k = 0; k < n; ++k) {
      int pfe = fe_get(t, ord_c[k], ord_st[k]);
      if (fe == FE_NONE && pfe != FE_NONE) fe = pfe;
      taint |= taint_get(t, ord_c[k], ord_st[k]);
    }
  }
  fe_set(t, c, st, fe);
  if (taint) t->taint[c] |= (uint8_t)(1 << st);
}

/* ScoreTensor::fill, Global arm, :169-237 */
static void fill(tensor_t* t, const uint8_t* seq1, const uint8_t* seq2, const sao_scheme_t* s) {
  const uint32_t n1 = t->n1, n2 = t->n2;
  const size_t w = t->w;
  const int dfs = t->fe != NULL;
  /* :172-182 */
  t->m[0] = 0;
  t->d[0] = NEG_INF;
  t->i[0] = NEG_INF;
  t->par[0] = 0;
  if (dfs) {
    dfs_info(t, ST_M, 0, 0);
    dfs_info(t, ST_D, 0, 0);
    dfs_info(t, ST_I, 0, 0);
  }
  /* :183-199  row x = 0: the boundary gap is stored in D and costs one EXTRA extension;
   * its single parent is d_scores[0][i-1] (same row!), state InD. */
  for (uint32_t y = 1; y <= n1; ++y) {
    t->m[y] = NEG_INF;
    t->i[y] = NEG_INF;
    t->d[y] = ((int32_t)y + 1) * s->gap_extension + s->gap_opening;
    t->par[y] = P_D_FROM_D;
    if (dfs) {
      dfs_info(t, ST_M, 0, y);
      dfs_info(t, ST_I, 0, y);
      /* chain cell: has a parent and x == 0 -> expanding it panics (:299) */
      fe_set(t, y, ST_D, FE_PANIC);
      t->taint[y] |= (uint8_t)(1 << ST_D);
    }
  }
  /* :200-216  column y = 0: boundary gap stored in I, parent i_scores[i-1][0], state InI. */
  for (uint32_t x = 1; x <= n2; ++x) {
    size_t c = (size_t)x * w;
    t->m[c] = NEG_INF;
    t->i[c] = s->gap_opening + ((int32_t)x + 1) * s->gap_extension;
    t->d[c] = NEG_INF;
    t->par[c] = P_I_FROM_I;
    if (dfs) {
      dfs_info(t, ST_M, x, 0);
      dfs_info(t, ST_D, x, 0);
      fe_set(t, c, ST_I, FE_PANIC); /* :303 */
      t->taint[c] |= (uint8_t)(1 << ST_I);
    }
  }
  /* :217-236  main loop; M, then I, then D per cell. */
  for (uint32_t x = 1; x <= n2; ++x) {
    const uint8_t b2 = seq2[x - 1];
    for (uint32_t y = 1; y <= n1; ++y) {
      const size_t c = (size_t)x * w + y;
      const size_t diag = c - w - 1, left = c - 1, up = c - w;
      const int32_t sub = (seq1[y - 1] == b2) ? s->match_ : s->mismatch; /* :220 raw byte == */
      uint8_t p = 0;
      /* m_score :76-86, m_pointer :120-153 */
      const int32_t mm = max2(max2(t->m[diag], t->i[diag]), t->d[diag]) + sub;
      if (mm == t->m[diag] + sub) p |= P_M_FROM_M;
      if (mm == t->i[diag] + sub) p |= P_M_FROM_I;
      if (mm == t->d[diag] + sub) p |= P_M_FROM_D;
      t->m[c] = mm;
      /* i_score :91-94, i_pointer :108-119 */
      const int32_t ii = max2(t->m[left] + s->gap_opening, t->i[left]) + s->gap_extension;
      if (ii == t->i[left] + s->gap_extension) p |= P_I_FROM_I;
      if (ii == t->m[left] + s->gap_opening + s->gap_extension) p |= P_I_FROM_M;
      t->i[c] = ii;
      /* d_score :87-90, d_pointer :96-107 */
      const int32_t dd = max2(t->m[up] + s->gap_opening, t->d[up]) + s->gap_extension;
      if (dd == t->d[up] + s->gap_extension) p |= P_D_FROM_D;
      if (dd == t->m[up] + s->gap_opening + s->gap_extension) p |= P_D_FROM_M;
      t->d[c] = dd;
      t->par[c] = p;
      if (dfs) {
        dfs_info(t, ST_M, x, y);
        dfs_info(t, ST_I, x, y);
        dfs_info(t, ST_D, x, y);
      }
    }
  }
}

/* #complete paths below each cell (saturating), rolling rows; only used for n_cooptimal. */
static int64_t sat_add(int64_t a, int64_t b) {
  int64_t r = a + b;
  return (r < a || r > (INT64_MAX / 4)) ? (INT64_MAX / 4) : r;
}

static int64_t count_paths(const tensor_t* t, int32_t max_val) {
  const size_t w = t->w;
  /* cnt[state][y] for previous and current row */
  int64_t* buf = (int64_t*)calloc(6 * w, sizeof(int64_t));
  if (!buf) return -1;
  int64_t *pm = buf, *pd = buf + w, *pi = buf + 2 * w, *cm = buf + 3 * w, *cd = buf + 4 * w,
          *ci = buf + 5 * w;
  int64_t total = 0;
  for (uint32_t x = 0; x <= t->n2; ++x) {
    for (uint32_t y = 0; y <= t->n1; ++y) {
      size_t c = (size_t)x * w + y;
      uint8_t p = t->par[c];
      int64_t vm = 0, vd = 0, vi = 0;
      if (x == 0 && y == 0) {
        vm = vd = vi = 1; /* any state popped at (0,0) prints (:283) */
      } else if (x > 0 && y > 0) {
        if (p & P_M_FROM_M) vm = sat_add(vm, pm[y - 1]);
        if (p & P_M_FROM_I) vm = sat_add(vm, pi[y - 1]);
        if (p & P_M_FROM_D) vm = sat_add(vm, pd[y - 1]);
        if (p & P_I_FROM_I) vi = sat_add(vi, ci[y - 1]);
        if (p & P_I_FROM_M) vi = sat_add(vi, cm[y - 1]);
        if (p & P_D_FROM_D) vd = sat_add(vd, pd[y]);
        if (p & P_D_FROM_M) vd = sat_add(vd, pm[y]);
      } /* boundary cells: chain cells panic, sentinel cells have no parents -> 0 */
      cm[y] = vm;
      cd[y] = vd;
      ci[y] = vi;
      if (x == t->n2 && y == t->n1) {
        if (t->i[c] == max_val) total = sat_add(total, vi);
        if (t->m[c] == max_val) total = sat_add(total, vm);
        if (t->d[c] == max_val) total = sat_add(total, vd);
      }
    }
    int64_t* tmp;
    tmp = pm; pm = cm; cm = tmp;
    tmp = pd; pd = cd; cd = tmp;
    tmp = pi; pi = ci; ci = tmp;
  }
  free(buf);
  return total;
}

static uint32_t rle_push(uint32_t* cigar, uint32_t n, int op) {
  /* cigar is being built BACKWARDS (end of alignment first); merge equal neighbours */
  if (n > 0 && (int)(cigar[n - 1] & 3u) == op) {
    cigar[n - 1] += 4u;
    return n;
  }
  cigar[n] = (1u << 2) | (uint32_t)op;
  return n + 1;
}

int sao_affine_align(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                     const sao_scheme_t* scheme, sao_affine_result_t* out, uint32_t* cigar) {
  tensor_t t;
  if (!scheme) scheme = &SAO_AFFINE_SCHEME;
  if (tensor_alloc(&t, n1, n2, 1) != 0) return -1;
  fill(&t, seq1, seq2, scheme);
  const size_t end = (size_t)n2 * t.w + n1;
  memset(out, 0, sizeof(*out));
  out->end_m = t.m[end];
  out->end_i = t.i[end];
  out->end_d = t.d[end];
  const int32_t max_val = max2(max2(t.i[end], t.d[end]), t.m[end]); /* :247-250 */
  out->score = max_val;
  /* start states are pushed I, M, D (:251-280) and popped D, M, I */
  int start_st[3], ns = 0;
  if (max_val == t.d[end]) start_st[ns++] = ST_D;
  if (max_val == t.m[end]) start_st[ns++] = ST_M;
  if (max_val == t.i[end]) start_st[ns++] = ST_I;
  int fe = FE_NONE, first = -1;
  for (int k = 0; k < ns; ++k) {
    if (fe == FE_NONE && fe_get(&t, end, start_st[k]) != FE_NONE) {
      fe = fe_get(&t, end, start_st[k]);
      first = start_st[k];
    }
    out->any_panic |= taint_get(&t, end, start_st[k]);
  }
  if (fe == FE_PRINT)
    out->status = out->any_panic ? SAO_REF_PANIC : SAO_OK;
  else if (fe == FE_PANIC)
    out->status = SAO_REF_PANIC_EARLY;
  else
    out->status = SAO_REF_NO_OUTPUT;
  out->n_cooptimal = count_paths(&t, max_val);

  if (cigar && fe == FE_PRINT) {
    /* first printed alignment: follow, at every cell, the first parent in reverse push
     * order whose subtree produces an event (it is necessarily a PRINT). */
    uint32_t x = n2, y = n1, n = 0, cols = 0;
    int st = first;
    while (!(x == 0 && y == 0)) {
      size_t c = (size_t)x * t.w + y;
      uint8_t p = t.par[c];
      int nst = -1;
      if (st == ST_M) {
        size_t pc = c - t.w - 1;
        if ((p & P_M_FROM_D) && fe_get(&t, pc, ST_D)) nst = ST_D;
        else if ((p & P_M_FROM_I) && fe_get(&t, pc, ST_I)) nst = ST_I;
        else if ((p & P_M_FROM_M) && fe_get(&t, pc, ST_M)) nst = ST_M;
        n = rle_push(cigar, n, SAO_OP_M);
        --x; --y;
      } else if (st == ST_I) {
        size_t pc = c - 1;
        if ((p & P_I_FROM_M) && fe_get(&t, pc, ST_M)) nst = ST_M;
        else if ((p & P_I_FROM_I) && fe_get(&t, pc, ST_I)) nst = ST_I;
        n = rle_push(cigar, n, SAO_OP_I);
        --y;
      } else {
        size_t pc = c - t.w;
        if ((p & P_D_FROM_M) && fe_get(&t, pc, ST_M)) nst = ST_M;
        else if ((p & P_D_FROM_D) && fe_get(&t, pc, ST_D)) nst = ST_D;
        n = rle_push(cigar, n, SAO_OP_D);
        --x;
      }
      ++cols;
      st = nst; /* never -1 while fe == PRINT */
      if (st < 0) break;
    }
    /* reverse into alignment order */
    for (uint32_t a = 0, b = n ? n - 1 : 0; a < b; ++a, --b) {
      uint32_t tmp = cigar[a];
      cigar[a] = cigar[b];
      cigar[b] = tmp;
    }
    out->cigar_len = n;
    out->n_columns = cols;
  }
  tensor_free(&t);
  return 0;
}

int sao_affine_matrices(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                        const sao_scheme_t* scheme, int32_t* m, int32_t* i, int32_t* d,
                        uint8_t* parents) {
  tensor_t t;
  if (!scheme) scheme = &SAO_AFFINE_SCHEME;
  if (tensor_alloc(&t, n1, n2, 0) != 0) return -1;
  fill(&t, seq1, seq2, scheme);
  size_t cells = ((size_t)n2 + 1) * t.w;
  memcpy(m, t.m, cells * sizeof(int32_t));
  memcpy(i, t.i, cells * sizeof(int32_t));
  memcpy(d, t.d, cells * sizeof(int32_t));
  if (parents) memcpy(parents, t.par, cells);
  tensor_free(&t);
  return 0;
}

/* Same recurrences, rolling rows, no parents: the score the reference computes (:247-250). */
int32_t sao_affine_score(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                         const sao_scheme_t* s) {
  if (!s) s = &SAO_AFFINE_SCHEME;
  const size_t w = (size_t)n1 + 1;
  int32_t* buf = (int32_t*)malloc(6 * w * sizeof(int32_t));
  if (!buf) return INT32_MIN;
  int32_t *pm = buf, *pd = buf + w, *pi = buf + 2 * w;
  int32_t *cm = buf + 3 * w, *cd = buf + 4 * w, *ci = buf + 5 * w;
  pm[0] = 0; /* :172-182 */
  pd[0] = NEG_INF;
  pi[0] = NEG_INF;
  for (uint32_t y = 1; y <= n1; ++y) { /* :183-199 */
    pm[y] = NEG_INF;
    pi[y] = NEG_INF;
    pd[y] = ((int32_t)y + 1) * s->gap_extension + s->gap_opening;
  }
  for (uint32_t x = 1; x <= n2; ++x) {
    const uint8_t b2 = seq2[x - 1];
    cm[0] = NEG_INF; /* :200-216 */
    cd[0] = NEG_INF;
    ci[0] = s->gap_opening + ((int32_t)x + 1) * s->gap_extension;
    for (uint32_t y = 1; y <= n1; ++y) { /* :217-236 */
      const int32_t sub = (seq1[y - 1] == b2) ? s->match_ : s->mismatch;
      cm[y] = max2(max2(pm[y - 1], pi[y - 1]), pd[y - 1]) + sub;
      ci[y] = max2(cm[y - 1] + s->gap_opening, ci[y - 1]) + s->gap_extension;
      cd[y] = max2(pm[y] + s->gap_opening, pd[y]) + s->gap_extension;
    }
    int32_t* tmp;
    tmp = pm; pm = cm; cm = tmp;
    tmp = pd; pd = cd; cd = tmp;
    tmp = pi; pi = ci; ci = tmp;
  }
  const int32_t r = max2(max2(pi[n1], pd[n1]), pm[n1]);
  free(buf);
  return r;
}

/* ---------------------------------------------------------------------------------------
 * Literal DFS (:246-329) with the reference's stdout text (Display :390-411).
 * ------------------------------------------------------------------------------------- */
typedef struct col_s {
  uint8_t c1, c2;
  int64_t next; /* index of the column to the right, -1 = end */
} col_t;

typedef struct {
  int st;
  uint32_t x, y;
  int64_t cols; /* head of the partial alignment (leftmost column so far) */
} frame_t;

typedef struct {
  char* buf;
  size_t cap;
  int64_t len;
} sink_t;

static void sink_put(sink_t* s, const char* p, size_t n) {
  for (size_t k = 0; k < n; ++k) {
    if ((size_t)s->len + 1 < s->cap) s->buf[s->len] = p[k];
    s->len++;
  }
}
static void sink_puts(sink_t* s, const char* p) { sink_put(s, p, strlen(p)); }

int64_t sao_affine_print_all(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                             const sao_scheme_t* scheme, uint64_t max_alignments, char* buf,
                             size_t buf_cap, uint64_t* n_printed, int32_t* panicked) {
  tensor_t t;
  if (!scheme) scheme = &SAO_AFFINE_SCHEME;
  if (tensor_alloc(&t, n1, n2, 0) != 0) return -1;
  fill(&t, seq1, seq2, scheme);
  sink_t sink = {buf, buf_cap, 0};
  size_t fcap = 64, fn = 0, ccap = 256, cn = 0;
  frame_t* stack = (frame_t*)malloc(fcap * sizeof(frame_t));
  col_t* cols = (col_t*)malloc(ccap * sizeof(col_t));
  uint64_t printed = 0;
  int32_t pan = 0;
  if (!stack || !cols) {
    free(stack);
    free(cols);
    tensor_free(&t);
    return -1;
  }
  const size_t end = (size_t)n2 * t.w + n1;
  const int32_t max_val = max2(max2(t.i[end], t.d[end]), t.m[end]);
  /* push order I, M, D (:251-280) */
  if (max_val == t.i[end]) stack[fn++] = (frame_t){ST_I, n2, n1, -1};
  if (max_val == t.m[end]) stack[fn++] = (frame_t){ST_M, n2, n1, -1};
  if (max_val == t.d[end]) stack[fn++] = (frame_t){ST_D, n2, n1, -1};
  while (fn > 0 && !pan) {
    frame_t e = stack[--fn];
    if (e.x == 0 && e.y == 0) { /* :283-286 */
      if (printed >= max_alignments) break;
      sink_puts(&sink, "alignment found\n");
      sink_puts(&sink, "\nseq1: ");
      for (int64_t k = e.cols; k >= 0; k = cols[k].next) sink_put(&sink, (const char*)&cols[k].c1, 1);
      sink_puts(&sink, "\n      ");
      for (int64_t k = e.cols; k >= 0; k = cols[k].next)
        sink_puts(&sink, cols[k].c1 == cols[k].c2 ? "|" : " ");
      sink_puts(&sink, "\nseq2: ");
      for (int64_t k = e.cols; k >= 0; k = cols[k].next) sink_put(&sink, (const char*)&cols[k].c2, 1);
      sink_puts(&sink, "\n");
      ++printed;
    }
    const size_t c = (size_t)e.x * t.w + e.y;
    const uint8_t p = t.par[c];
    /* parents in PUSH order */
    int pst[3];
    int np = 0;
    if (e.st == ST_M) {
      if (p & P_M_FROM_M) pst[np++] = ST_M;
      if (p & P_M_FROM_I) pst[np++] = ST_I;
      if (p & P_M_FROM_D) pst[np++] = ST_D;
    } else if (e.st == ST_I) {
      if (p & P_I_FROM_I) pst[np++] = ST_I;
      if (p & P_I_FROM_M) pst[np++] = ST_M;
    } else {
      if (p & P_D_FROM_D) pst[np++] = ST_D;
      if (p & P_D_FROM_M) pst[np++] = ST_M;
    }
    for (int k = 0; k < np; ++k) {
      if (expand_panics(e.st, e.x, e.y)) { /* :293-294, :299, :303 */
        pan = 1;
        break;
      }
      if (cn == ccap) {
        ccap *= 2;
        col_t* nc = (col_t*)realloc(cols, ccap * sizeof(col_t));
        if (!nc) { pan = -1; break; }
        cols = nc;
      }
      uint32_t x = e.x, y = e.y;
      col_t col;
      col.next = e.cols;
      if (e.st == ST_M) {
        col.c1 = seq1[y - 1]; col.c2 = seq2[x - 1]; --x; --y;
      } else if (e.st == ST_D) {
        col.c1 = '-'; col.c2 = seq2[x - 1]; --x;
      } else {
        col.c1 = seq1[y - 1]; col.c2 = '-'; --y;
      }
      cols[cn] = col;
      if (fn == fcap) {
        fcap *= 2;
        frame_t* ns = (frame_t*)realloc(stack, fcap * sizeof(frame_t));
        if (!ns) { pan = -1; break; }
        stack = ns;
      }
      stack[fn++] = (frame_t){pst[k], x, y, (int64_t)cn};
      ++cn;
    }
  }
  if (sink.cap) sink.buf[(size_t)sink.len < sink.cap ? (size_t)sink.len : sink.cap - 1] = 0;
  if (n_printed) *n_printed = printed;
  if (panicked) *panicked = pan > 0;
  free(stack);
  free(cols);
  tensor_free(&t);
  return pan < 0 ? -1 : sink.len;
}

/* --------------------------------------------------------------------------------------- */
typedef struct {
  const uint8_t* residues;
  const uint64_t *q_off, *d_off;
  const uint32_t *q_len, *d_len;
  const sao_scheme_t* scheme;
  int32_t* score;
  uint8_t* status;
  uint32_t* cigar_len;
  uint32_t* cigar_pool;
  uint32_t cigar_stride;
  uint64_t lo, hi;
  int rc;
} batch_job_t;

static void* batch_worker(void* arg) {
  batch_job_t* j = (batch_job_t*)arg;
  uint32_t* tmp = NULL;
  size_t tmp_cap = 0;
  for (uint64_t p = j->lo; p < j->hi; ++p) {
    const uint32_t n1 = j->q_len[p], n2 = j->d_len[p];
    sao_affine_result_t r;
    uint32_t* cig = NULL;
    if (j->cigar_pool) {
      size_t need = (size_t)n1 + n2 + 1;
      if (need > tmp_cap) {
        free(tmp);
        tmp = (uint32_t*)malloc(need * sizeof(uint32_t));
        tmp_cap = need;
        if (!tmp) { j->rc = -1; return NULL; }
      }
      cig = tmp;
    }
    if (sao_affine_align(j->residues + j->q_off[p], n1, j->residues + j->d_off[p], n2, j->scheme,
                         &r, cig) != 0) {
      j->rc = -1;
      break;
    }
    j->score[p] = r.score;
    if (j->status) j->status[p] = (uint8_t)r.status;
    if (j->cigar_len) j->cigar_len[p] = r.cigar_len;
    if (j->cigar_pool) {
      uint32_t n = r.cigar_len < j->cigar_stride ? r.cigar_len : j->cigar_stride;
      memcpy(j->cigar_pool + (size_t)p * j->cigar_stride, cig, n * sizeof(uint32_t));
    }
  }
  free(tmp);
  return NULL;
}

int sao_affine_batch(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                     const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs,
                     const sao_scheme_t* scheme, int32_t* score, uint8_t* status,
                     uint32_t* cigar_len, uint32_t* cigar_pool, uint32_t cigar_stride,
                     int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > n_pairs) n_threads = n_pairs ? (int)n_pairs : 1;
  batch_job_t* jobs = (batch_job_t*)calloc((size_t)n_threads, sizeof(batch_job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  if (!jobs || !th) { free(jobs); free(th); return -1; }
  int rc = 0;
  for (int k = 0; k < n_threads; ++k) {
    jobs[k] = (batch_job_t){residues, q_off, d_off, q_len, d_len, scheme, score, status,
                            cigar_len, cigar_pool, cigar_stride,
                            n_pairs * (uint64_t)k / (uint64_t)n_threads,
                            n_pairs * (uint64_t)(k + 1) / (uint64_t)n_threads, 0};
    if (n_threads == 1) batch_worker(&jobs[k]);
    else pthread_create(&th[k], NULL, batch_worker, &jobs[k]);
  }
  for (int k = 0; k < n_threads; ++k) {
    if (n_threads > 1) pthread_join(th[k], NULL);
    if (jobs[k].rc) rc = -1;
  }
  free(jobs);
  free(th);
  return rc;
}
