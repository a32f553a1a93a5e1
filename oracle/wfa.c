/*
 * wfa.c -- ORACLE (test infrastructure, never linked into the product).
 *
 * (1) sao_wfa_literal: literal CPU restatement of the reference's gap-affine wavefront aligner,
 *     /root/reference/src/wfa.rs (wfa_align :23-42, WaveFrontElement :84-102, WaveFront::expand
 *     :127-139, WaveFrontTensor::new :225-420, Ocean::{global,expand,trim,is_converged}
 *     :449-632), including its defects, reported as statuses instead of crashes:
 *       - wavefront 0 is never extended (:467-483)
 *       - convergence is tested at (n2-1, n1-1) (:189), so many pairs never converge
 *       - trim (:490-623) panics in Vec::rotate_left when I or D lies more than `len`
 *         diagonals below the trimmed M (:577, :603)
 *       - the printed score is wfs.len(), i.e. the index of the converged wavefront + 1
 *     Release-profile integer semantics (the README runs `cargo run --release`): i32/usize
 *     arithmetic wraps, `Vec::truncate(len - t)` with t > len is a no-op.
 *     Pinned by the reference's own tests wfa.rs:994-1186,1268-1294 through the literal Python
 *     model (oracle/literal_model.py) and tests/golden/wfa_golden.json.
 * (2) sao_wfa_gotoh_cost: textbook gap-affine COST dynamic programme (min-cost, mismatch x,
 *     gap of length L costs o + L*e), the ground truth for the standard-mode WFA kernel on
 *     inputs where the reference has no answer.
 * (3) sao_wfa_standard: textbook gap-affine WFA (Marco-Sola et al. 2021, score only), a
 *     second way to the same number.
 */
#include "sa_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WFA_X 4 /* SCHEME :17-21 */
#define WFA_O 2
#define WFA_E 6
#define MINLENGTH 5 /* :14 */
#define MAXDIFF 20  /* :15 */

enum { S_M = 0, S_D = 1, S_I = 2 };

typedef struct {
  int32_t offset;
  uint8_t some;
  uint8_t state;
  uint8_t nparents;
  uint8_t parents[3];
} el_t;

typedef struct {
  int32_t hi, lo;
  el_t* el;
  size_t len, cap;
  int present;
} wf_t;

typedef struct {
  wf_t i, d, m;
  int present;
} tensor_t;

typedef struct {
  int32_t status, printed_score, panic_line, n_wavefronts;
} wfa_result_t;

static int wf_reserve(wf_t* w, size_t n) {
  if (n <= w->cap) return 0;
  size_t nc = w->cap ? w->cap * 2 : 8;
  while (nc < n) nc *= 2;
  el_t* ne = (el_t*)realloc(w->el, nc * sizeof(el_t));
  if (!ne) return -1;
  w->el = ne;
  w->cap = nc;
  return 0;
}
static int wf_push(wf_t* w, const el_t* e) {
  if (wf_reserve(w, w->len + 1)) return -1;
  if (e) w->el[w->len] = *e;
  else memset(&w->el[w->len], 0, sizeof(el_t));
  w->len++;
  return 0;
}
static void wf_remove_front(wf_t* w) {
  if (w->len) {
    memmove(w->el, w->el + 1, (w->len - 1) * sizeof(el_t));
    w->len--;
  }
}
/* get_element :159-163: index (idx - lo) as usize; out of range (incl. negative) -> None */
static const el_t* wf_get(const wf_t* w, int32_t idx) {
  if (!w || !w->present) return NULL;
  const uint64_t k = (uint64_t)(int64_t)(int32_t)(idx - w->lo); /* `as usize` sign-extends */
  if (k >= w->len) return NULL;
  return w->el[k].some ? &w->el[k] : NULL;
}
/* Vec::rotate_left; returns -1 where Rust panics (k > len) */
static int wf_rotate_left(wf_t* w, uint32_t k) {
  if (k > w->len) return -1;
  if (k == 0 || k == w->len) return 0;
  el_t* tmp = (el_t*)malloc(k * sizeof(el_t));
  if (!tmp) return -2;
  memcpy(tmp, w->el, k * sizeof(el_t));
  memmove(w->el, w->el + k, (w->len - k) * sizeof(el_t));
  memcpy(w->el + (w->len - k), tmp, k * sizeof(el_t));
  free(tmp);
  return 0;
}
static uint32_t abs_diff(int32_t a, int32_t b) { return a > b ? (uint32_t)(a - b) : (uint32_t)(b - a); }

static uint64_t el_x(const el_t* e, int32_t diag) { /* :85-87 */
  return (uint64_t)(int64_t)(int32_t)(e->offset - (diag < 0 ? diag : 0));
}
static uint64_t el_y(const el_t* e, int32_t diag) { /* :88-90 */
  return (uint64_t)(int64_t)(int32_t)(e->offset + (diag > 0 ? diag : 0));
}
static int32_t el_distance(const el_t* e, int32_t n1, int32_t n2, int32_t diag) { /* :96-101 */
  const int32_t lv = n1 - e->offset - diag, lh = n2 - e->offset;
  return lv > lh ? lv : lh;
}

static void tensor_free(tensor_t* t) {
  free(t->i.el);
  free(t->d.el);
  free(t->m.el);
  memset(t, 0, sizeof(*t));
}

/* get_parents :201-209 over up to three optional source elements */
static void set_parents(el_t* out, int32_t offset, const el_t* const* src, int n) {
  out->nparents = 0;
  for (int k = 0; k < n; ++k)
    if (src[k] && src[k]->offset == offset) out->parents[out->nparents++] = src[k]->state;
}

/* WaveFrontTensor::new :225-420.  Returns 0 and out->present = 0 for None. */
static int tensor_new(const tensor_t* open_, const tensor_t* ext, const tensor_t* mis, tensor_t* out,
                      int32_t* lo_out, int32_t* hi_out) {
  memset(out, 0, sizeof(*out));
  const wf_t* om = (open_ && open_->present && open_->m.present) ? &open_->m : NULL;
  const wf_t* mm = (mis && mis->present && mis->m.present) ? &mis->m : NULL;
  const wf_t* ei = (ext && ext->present && ext->i.present) ? &ext->i : NULL;
  const wf_t* ed = (ext && ext->present && ext->d.present) ? &ext->d : NULL;
  const wf_t* srcs[4] = {om, mm, ei, ed};
  int any = 0;
  int32_t hi = 0, lo = 0;
  for (int k = 0; k < 4; ++k)
    if (srcs[k]) {
      if (!any || srcs[k]->hi > hi) hi = srcs[k]->hi;
      if (!any || srcs[k]->lo < lo) lo = srcs[k]->lo;
      any = 1;
    }
  if (!any) return 0; /* `.max()?` on an empty iterator :238 */
  hi += 1;
  lo -= 1;
  *lo_out = lo;
  *hi_out = hi;
  wf_t i = {hi, lo, NULL, 0, 0, 1}, d = {hi, lo, NULL, 0, 0, 1}, m = {hi, lo, NULL, 0, 0, 1};
  int32_t i_lo = lo, i_hi = hi, d_lo = lo, d_hi = hi, m_lo = lo, m_hi = hi;
  int i_set = 0, d_set = 0, m_set = 0;
  for (int64_t idx64 = lo; idx64 <= hi; ++idx64) {
    const int32_t idx = (int32_t)idx64;
    /* D :272-311 */
    {
      const el_t* s[2] = {wf_get(om, idx + 1), wf_get(ed, idx + 1)};
      if (s[0] || s[1]) {
        int32_t off = s[0] ? s[0]->offset : s[1]->offset;
        if (s[1] && s[1]->offset > off) off = s[1]->offset;
        el_t e = {off, 1, S_D, 0, {0, 0, 0}};
        set_parents(&e, off, s, 2);
        if (wf_push(&d, &e)) goto oom;
        d_hi = idx;
        if (!d_set) { d_lo = idx; d_set = 1; }
      } else if (wf_push(&d, NULL)) goto oom;
    }
    /* I :313-352 */
    {
      const el_t* s[2] = {wf_get(om, idx - 1), wf_get(ei, idx - 1)};
      if (s[0] || s[1]) {
        int32_t off = s[0] ? s[0]->offset : s[1]->offset;
        if (s[1] && s[1]->offset > off) off = s[1]->offset;
        el_t e = {off + 1, 1, S_I, 0, {0, 0, 0}};
        set_parents(&e, off, s, 2);
        if (wf_push(&i, &e)) goto oom;
        i_hi = idx;
        if (!i_set) { i_lo = idx; i_set = 1; }
      } else if (wf_push(&i, NULL)) goto oom;
    }
    /* M :353-398 */
    {
      const el_t* me = wf_get(mm, idx);
      el_t tmp = {0, 1, S_M, 0, {0, 0, 0}};
      if (me) tmp.offset = me->offset + 1;
      const el_t* ie = wf_get(&i, idx);
      const el_t* de = wf_get(&d, idx);
      const el_t* s[3] = {me ? &tmp : NULL, ie, de};
      int have = 0;
      int32_t off = 0;
      for (int k = 0; k < 3; ++k)
        if (s[k] && (!have || s[k]->offset > off)) { off = s[k]->offset; have = 1; }
      if (have) {
        el_t e = {off, 1, S_M, 0, {0, 0, 0}};
        set_parents(&e, off, s, 3);
        if (wf_push(&m, &e)) goto oom;
        m_hi = idx;
        if (!m_set) { m_lo = idx; m_set = 1; }
      } else if (m_set) {
        if (wf_push(&m, NULL)) goto oom;
      }
    }
  }
  i.lo = i_lo; i.hi = i_hi; d.lo = d_lo; d.hi = d_hi; m.lo = m_lo; m.hi = m_hi; /* :401-403 */
  if (wf_rotate_left(&i, abs_diff(lo, i.lo))) goto oom; /* cannot exceed len here */
  if (i.len > (size_t)abs_diff(i.hi, i.lo) + 1) i.len = (size_t)abs_diff(i.hi, i.lo) + 1;
  if (wf_rotate_left(&d, abs_diff(lo, d.lo))) goto oom;
  if (d.len > (size_t)abs_diff(d.hi, d.lo) + 1) d.len = (size_t)abs_diff(d.hi, d.lo) + 1;
  if (m.len > (size_t)abs_diff(m.hi, m.lo) + 1) m.len = (size_t)abs_diff(m.hi, m.lo) + 1;
  out->present = 1;
  out->i = i; out->i.present = i_set;
  out->d = d; out->d.present = d_set;
  out->m = m; out->m.present = m_set;
  return 0;
oom:
  free(i.el); free(d.el); free(m.el);
  return -1;
}

/* WaveFront::expand :127-139 */
static void wf_extend(wf_t* m, const uint8_t* s1, uint32_t n1, const uint8_t* s2, uint32_t n2) {
  for (size_t k = 0; k < m->len; ++k) {
    el_t* e = &m->el[k];
    if (!e->some) continue;
    const int32_t diag = m->lo + (int32_t)k;
    for (;;) {
      const uint64_t y = el_y(e, diag), x = el_x(e, diag);
      if (!(y < n1 && x < n2 && s1[y] == s2[x])) break;
      e->offset += 1;
    }
  }
}

/* Ocean::trim :490-623.  Returns 0, or the reference line of the panic. */
static int trim(tensor_t* cur, int32_t n1, int32_t n2) {
  wf_t* m = &cur->m;
  if (!m->present) return 0;
  if (abs_diff(m->lo, m->hi) <= MINLENGTH) return 0;
  int32_t min_d = 0; /* :511 starts at 0 */
  for (int32_t diag = m->lo; diag <= m->hi; ++diag) {
    const el_t* e = wf_get(m, diag);
    if (e) {
      const int32_t dd = el_distance(e, n1, n2, diag);
      if (dd < min_d) min_d = dd;
    }
  }
  if (!m->len || !m->el[0].some) return 519;
  int32_t next_d = el_distance(&m->el[0], n1, n2, m->lo);
  while (m->lo < m->hi && abs_diff(next_d, min_d) > MAXDIFF) {
    m->lo += 1;
    wf_remove_front(m);
    while (wf_get(m, m->lo) == NULL) {
      if (m->lo == m->hi) break;
      m->lo += 1;
      wf_remove_front(m);
    }
    if (!m->len || !m->el[0].some) return 537;
    next_d = el_distance(&m->el[0], n1, n2, m->lo);
  }
  if (!m->len || !m->el[m->len - 1].some) return 545;
  next_d = el_distance(&m->el[m->len - 1], n1, n2, m->hi);
  while (m->hi > m->lo && abs_diff(next_d, min_d) > MAXDIFF) {
    m->hi -= 1;
    if (m->len) m->len--;
    while (wf_get(m, m->hi) == NULL) {
      if (m->lo == m->hi) break;
      m->hi -= 1;
      if (m->len) m->len--;
    }
    if (!m->len || !m->el[m->len - 1].some) return 562;
    next_d = el_distance(&m->el[m->len - 1], n1, n2, m->hi);
  }
  wf_t* comps[2] = {&cur->i, &cur->d};
  const int lines[2] = {577, 603};
  for (int c = 0; c < 2; ++c) {
    wf_t* w = comps[c];
    if (!w->present) continue;
    uint64_t t;
    if (w->lo < m->lo) {
      const int rc = wf_rotate_left(w, abs_diff(w->lo, m->lo));
      if (rc == -1) return lines[c];
      if (rc == -2) return -1;
      t = (uint64_t)abs_diff(w->lo, m->lo) + (w->hi > m->hi ? abs_diff(w->hi, m->hi) : 0);
    } else if (w->hi > m->hi) {
      t = abs_diff(w->hi, m->hi);
    } else {
      t = 0;
    }
    if (t <= w->len) w->len -= (size_t)t; /* else: `len - t` wraps in release, truncate no-op */
    if (m->hi < w->hi) w->hi = m->hi;
    if (m->lo > w->lo) w->lo = m->lo;
  }
  return 0;
}

static const el_t* wf_converged(const wf_t* w, uint32_t n1, uint32_t n2) { /* :180-191 */
  if (!w->present) return NULL;
  const uint64_t tx = (uint64_t)n2 - 1, ty = (uint64_t)n1 - 1; /* usize: wraps for empty input */
  for (size_t k = 0; k < w->len; ++k) {
    const el_t* e = &w->el[k];
    if (!e->some) continue;
    const int32_t diag = w->lo + (int32_t)k;
    if (el_x(e, diag) == tx && el_y(e, diag) == ty) return e;
  }
  return NULL;
}

/* Outcome of wfa_align (:23-42) in global mode.  Optional outputs: lo/hi of every created
 * wavefront (the `lo: .., hi: ..` lines), and the converged element. */
int sao_wfa_literal_ex(const uint8_t* s1, uint32_t n1, const uint8_t* s2, uint32_t n2,
                       uint32_t max_score, wfa_result_t* res, int32_t* lohi, uint32_t lohi_cap,
                       uint32_t* n_lohi, int32_t* conv /* offset, state, nparents, p0, p1, p2 */) {
  memset(res, 0, sizeof(*res));
  if (n_lohi) *n_lohi = 0;
  size_t cap = 64, len = 0;
  tensor_t* wfs = (tensor_t*)calloc(cap, sizeof(tensor_t));
  if (!wfs) return -1;
  int rc = 0;
  /* Ocean::global :450-465 */
  wfs[0].present = 1;
  wfs[0].m.present = 1;
  wfs[0].m.hi = wfs[0].m.lo = 0;
  {
    el_t e0 = {0, 1, S_M, 0, {0, 0, 0}};
    if (wf_push(&wfs[0].m, &e0)) { rc = -1; goto done; }
  }
  len = 1;
  for (;;) {
    const tensor_t* last = &wfs[len - 1];
    const el_t* c = NULL;
    if (last->present) { /* :422-439 order I, D, M */
      c = wf_converged(&last->i, n1, n2);
      if (!c) c = wf_converged(&last->d, n1, n2);
      if (!c) c = wf_converged(&last->m, n1, n2);
    }
    if (c) {
      res->status = SAO_OK;
      res->printed_score = (int32_t)len; /* :31-36 */
      if (conv) {
        conv[0] = c->offset; conv[1] = c->state; conv[2] = c->nparents;
        conv[3] = c->parents[0]; conv[4] = c->parents[1]; conv[5] = c->parents[2];
      }
      break;
    }
    if (len > max_score) {
      res->status = SAO_REF_NO_CONVERGENCE;
      break;
    }
    /* Ocean::expand :467-488 */
    const int64_t s = (int64_t)len;
    if (len == cap) {
      cap *= 2;
      tensor_t* nw = (tensor_t*)realloc(wfs, cap * sizeof(tensor_t));
      if (!nw) { rc = -1; goto done; }
      memset(nw + len, 0, (cap - len) * sizeof(tensor_t));
      wfs = nw;
    }
    const tensor_t* t_open = (s - WFA_O - WFA_E >= 0) ? &wfs[s - WFA_O - WFA_E] : NULL;
    const tensor_t* t_ext = (s - WFA_E >= 0) ? &wfs[s - WFA_E] : NULL;
    const tensor_t* t_mis = (s - WFA_X >= 0) ? &wfs[s - WFA_X] : NULL;
    int32_t lo = 0, hi = 0;
    if (tensor_new(t_open, t_ext, t_mis, &wfs[len], &lo, &hi)) { rc = -1; goto done; }
    tensor_t* cur = &wfs[len];
    ++len;
    if (cur->present) {
      if (lohi && n_lohi && *n_lohi < lohi_cap) { lohi[2 * *n_lohi] = lo; lohi[2 * *n_lohi + 1] = hi; }
      if (n_lohi) ++*n_lohi;
      if (cur->m.present) wf_extend(&cur->m, s1, n1, s2, n2);
      const int line = trim(cur, (int32_t)n1, (int32_t)n2);
      if (line < 0) { rc = -1; goto done; }
      if (line > 0) {
        res->status = SAO_REF_PANIC;
        res->panic_line = line;
        break;
      }
    }
  }
  res->n_wavefronts = (int32_t)len;
done:
  for (size_t k = 0; k < len && k < cap; ++k) tensor_free(&wfs[k]);
  free(wfs);
  return rc;
}

int sao_wfa_literal(const uint8_t* s1, uint32_t n1, const uint8_t* s2, uint32_t n2, uint32_t max_score,
                    wfa_result_t* res) {
  return sao_wfa_literal_ex(s1, n1, s2, n2, max_score, res, NULL, 0, NULL, NULL);
}

/* the bound after which the oracle and the engine report REF_NO_CONVERGENCE */
uint32_t sao_wfa_literal_cap(uint32_t n1, uint32_t n2) {
  const uint64_t c = 8ull * ((uint64_t)n1 + n2) + 64;
  return (uint32_t)(c < 2048 ? c : 2048);
}

int sao_wfa_literal_batch(const uint8_t* residues, const uint64_t* q_off, const uint32_t* q_len,
                          const uint64_t* d_off, const uint32_t* d_len, uint64_t n_pairs,
                          int32_t* score, uint8_t* status) {
  for (uint64_t p = 0; p < n_pairs; ++p) {
    wfa_result_t r;
    if (sao_wfa_literal(residues + q_off[p], q_len[p], residues + d_off[p], d_len[p],
                        sao_wfa_literal_cap(q_len[p], d_len[p]), &r))
      return -1;
    score[p] = r.status == SAO_OK ? r.printed_score : 0;
    status[p] = (uint8_t)r.status;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * The reference's stdout for one pair under `-a wfa` (wfa_align :23-42).
 *   :251      "lo: {}, hi: {}"                       per created wavefront
 *   :36       "converged with score {}: "            wfs.len()
 *   :650      "huhu, diag: {}\n{:#?}\nscore: {}"     diag = n1 - n2, Debug of the element :104-116
 *   :662-665  "ret"                                   when diag == 0 && offset == 0
 *   :667-678  per d in [4, 6, 8]: "well shit" when d > len, else "yeah, score: {len - d}";
 *             wfs[len - d] is always a None tensor (len odd, penalties even), so the arms print
 *             only what they print BEFORE looking into the tensor: "extend" when d == 6 and the
 *             element's parents contain D (:710-711), "open" when d == 8 and they contain M (:754-755)
 *   :851      "huh"
 *   :38-39    Display (:950-980) and pretty Debug of the empty Alignment
 * On a panic (status SAO_REF_PANIC) the text holds what had been printed before it.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  char* buf;
  size_t cap, len;
} wtxt_t;
static void wput(wtxt_t* t, const char* s) {
  for (; *s; ++s) {
    if (t->buf && t->len + 1 < t->cap) t->buf[t->len] = *s;
    t->len++;
  }
}
static void wnum(wtxt_t* t, int64_t v) {
  char tmp[32];
  snprintf(tmp, sizeof(tmp), "%lld", (long long)v);
  wput(t, tmp);
}

int64_t sao_wfa_print(const uint8_t* s1, uint32_t n1, const uint8_t* s2, uint32_t n2, uint32_t max_score,
                      char* buf, size_t buf_cap, int32_t* status) {
  const uint32_t cap = max_score + 8;
  int32_t* lohi = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)cap);
  if (!lohi) return -1;
  uint32_t n_lohi = 0;
  int32_t conv[6] = {0, 0, 0, 0, 0, 0};
  wfa_result_t res;
  if (sao_wfa_literal_ex(s1, n1, s2, n2, max_score, &res, lohi, cap, &n_lohi, conv)) {
    free(lohi);
    return -1;
  }
  if (status) *status = res.status;
  wtxt_t t = {buf, buf_cap, 0};
  for (uint32_t k = 0; k < n_lohi && k < cap; ++k) {
    wput(&t, "lo: "); wnum(&t, lohi[2 * k]); wput(&t, ", hi: "); wnum(&t, lohi[2 * k + 1]); wput(&t, "\n");
  }
  free(lohi);
  if (res.status == SAO_OK) {
    static const char* names[3] = {"M", "D", "I"};
    const int64_t len = res.printed_score, diag = (int64_t)n1 - (int64_t)n2;
    wput(&t, "converged with score "); wnum(&t, len); wput(&t, ": \n");
    wput(&t, "huhu, diag: "); wnum(&t, diag); wput(&t, "\n");
    wput(&t, "Element {\n\tstate: "); wput(&t, names[conv[1]]); wput(&t, "\n\toffset: "); wnum(&t, conv[0]); wput(&t, "\n");
    wput(&t, "\tparents: ");
    if (conv[2] == 0) {
      wput(&t, "[]\n");
    } else {
      wput(&t, "[\n");
      for (int k = 0; k < conv[2]; ++k) { wput(&t, "    "); wput(&t, names[conv[3 + k]]); wput(&t, ",\n"); }
      wput(&t, "]\n");
    }
    wput(&t, "}\n"); /* writeln!(f, "}}") */
    wput(&t, "\nscore: "); wnum(&t, len); wput(&t, "\n");
    int has_m = 0, has_d = 0;
    for (int k = 0; k < conv[2]; ++k) {
      if (conv[3 + k] == S_M) has_m = 1;
      if (conv[3 + k] == S_D) has_d = 1;
    }
    if (diag == 0 && conv[0] == 0) {
      wput(&t, "ret\n");
    } else {
      const int64_t ds[3] = {WFA_X, WFA_E, WFA_O + WFA_E};
      for (int k = 0; k < 3; ++k) {
        if (ds[k] > len) { wput(&t, "well shit\n"); continue; }
        wput(&t, "yeah, score: "); wnum(&t, len - ds[k]); wput(&t, "\n");
        if (k == 1 && has_d) wput(&t, "extend\n");
        if (k == 2 && has_m) wput(&t, "open\n");
      }
      wput(&t, "huh\n");
    }
    wput(&t, "\n\n\n");
    wput(&t, "Alignment {\n    seq1: [],\n    seq2: [],\n}\n");
  }
  if (t.buf && t.cap) t.buf[t.len < t.cap ? t.len : t.cap - 1] = 0;
  return (int64_t)t.len;
}

/* ------------------------------------------------------------------------------------------ */
/* Textbook gap-affine COST DP: min cost, mismatch x, a gap of length L costs o + L*e.        */
int64_t sao_wfa_gotoh_cost(const uint8_t* s1, uint32_t n1, const uint8_t* s2, uint32_t n2, int32_t x,
                           int32_t o, int32_t e) {
  const int64_t INF = (int64_t)1 << 60;
  const size_t w = (size_t)n1 + 1;
  int64_t* buf = (int64_t*)malloc(4 * w * sizeof(int64_t));
  if (!buf) return -1;
  int64_t *ph = buf, *pd = buf + w, *ch = buf + 2 * w, *cd = buf + 3 * w;
  ph[0] = 0;
  pd[0] = INF;
  for (uint32_t y = 1; y <= n1; ++y) {
    ph[y] = o + (int64_t)y * e; /* one gap consuming seq1 */
    pd[y] = INF;
  }
  for (uint32_t i = 1; i <= n2; ++i) {
    ch[0] = o + (int64_t)i * e;
    cd[0] = ch[0];
    int64_t ins = INF; /* gap consuming seq1, running along the row */
    for (uint32_t y = 1; y <= n1; ++y) {
      const int64_t a = ch[y - 1] + o + e, b = ins + e;
      ins = a < b ? a : b;
      const int64_t c = ph[y] + o + e, d = pd[y] + e;
      cd[y] = c < d ? c : d;
      int64_t best = ph[y - 1] + (s1[y - 1] == s2[i - 1] ? 0 : x);
      if (ins < best) best = ins;
      if (cd[y] < best) best = cd[y];
      ch[y] = best;
    }
    int64_t* t;
    t = ph; ph = ch; ch = t;
    t = pd; pd = cd; cd = t;
  }
  const int64_t r = ph[n1];
  free(buf);
  return r;
}

/* Textbook gap-affine WFA, score only.  Offsets count consumed seq1 residues (v), diagonal
 * k = v - h with h the consumed seq2 residues.  Returns the optimal cost. */
int64_t sao_wfa_standard(const uint8_t* s1, uint32_t n1, const uint8_t* s2, uint32_t n2, int32_t x,
                         int32_t o, int32_t e) {
  const int32_t NONE = -(1 << 29);
  const int64_t max_s = (int64_t)o * 2 + (int64_t)e * ((int64_t)n1 + n2) + (int64_t)x + 8;
  const int32_t kmin = -(int32_t)n2, kmax = (int32_t)n1;
  const size_t w = (size_t)(kmax - kmin + 1);
  const int ring = o + e + 1 > x + 1 ? o + e + 1 : x + 1;
  int32_t* buf = (int32_t*)malloc((size_t)ring * 3 * w * sizeof(int32_t));
  if (!buf) return -1;
  for (size_t k = 0; k < (size_t)ring * 3 * w; ++k) buf[k] = NONE;
#define WF(comp, s) (buf + (((size_t)((s) % ring) * 3 + (comp)) * w))
  const int32_t kend = (int32_t)n1 - (int32_t)n2;
  int64_t result = -1;
  for (int64_t s = 0; s <= max_s; ++s) {
    int32_t *M = WF(0, s), *I = WF(1, s), *D = WF(2, s);
    for (size_t k = 0; k < w; ++k) M[k] = I[k] = D[k] = NONE;
    if (s == 0) {
      M[0 - kmin] = 0;
    } else {
      const int32_t* Mx = s - x >= 0 ? WF(0, s - x) : NULL;
      const int32_t* Mo = s - o - e >= 0 ? WF(0, s - o - e) : NULL;
      const int32_t* Ie = s - e >= 0 ? WF(1, s - e) : NULL;
      const int32_t* De = s - e >= 0 ? WF(2, s - e) : NULL;
      for (int32_t k = kmin; k <= kmax; ++k) {
        const size_t c = (size_t)(k - kmin);
        int32_t iv = NONE, dv = NONE, mv = NONE;
        if (k - 1 >= kmin) { /* insertion: consumes a seq1 residue, k-1 -> k, v+1 */
          int32_t a = Mo ? Mo[c - 1] : NONE, b = Ie ? Ie[c - 1] : NONE;
          int32_t best = a > b ? a : b;
          if (best > NONE) iv = best + 1;
        }
        if (k + 1 <= kmax) { /* deletion: consumes a seq2 residue, k+1 -> k, v unchanged */
          int32_t a = Mo ? Mo[c + 1] : NONE, b = De ? De[c + 1] : NONE;
          int32_t best = a > b ? a : b;
          if (best > NONE) dv = best;
        }
        if (Mx && Mx[c] > NONE) mv = Mx[c] + 1;
        if (iv > mv) mv = iv;
        if (dv > mv) mv = dv;
        /* discard cells outside the matrix */
        if (iv > NONE && (iv > (int32_t)n1 || iv - k > (int32_t)n2 || iv - k < 0)) iv = NONE;
        if (dv > NONE && (dv > (int32_t)n1 || dv - k > (int32_t)n2 || dv < 0)) dv = NONE;
        if (mv > NONE && (mv > (int32_t)n1 || mv - k > (int32_t)n2 || mv - k < 0 || mv < 0)) mv = NONE;
        I[c] = iv;
        D[c] = dv;
        M[c] = mv;
      }
    }
    for (int32_t k = kmin; k <= kmax; ++k) { /* extend */
      const size_t c = (size_t)(k - kmin);
      int32_t v = M[c];
      if (v <= NONE) continue;
      int32_t h = v - k;
      while (v < (int32_t)n1 && h < (int32_t)n2 && s1[v] == s2[h]) { ++v; ++h; }
      M[c] = v;
    }
    if (M[kend - kmin] == (int32_t)n1) {
      result = s;
      break;
    }
  }
#undef WF
  free(buf);
  return result;
}
