// engine_side.cu -- side APIs of the affine aligner: batched co-optimal counts (nw_count.cuh) and the
// enumeration of EVERY co-optimal alignment of one pair in the reference's order (nw_parents.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstdint>
#include <cstring>
#include <utility>
#include <string>
#include <vector>

#include "engine_internal.h"
#include "engine_util.h"
#include "nw_parents.cuh"
#include "nw_count.cuh"

namespace sa_host {

// Co-optimal alignment counts for a whole batch (nw_count.cuh), in chunks that bound the scratch.
sa_status_t sd_count_cooptimal(sa_engine* e, const sa_scheme_t* scheme, const sa_batch_t* b, int64_t* counts) {
  if (!e || !b) return SA_E_ARG;
  if (b->packing > 1) return fail(e, SA_E_ARG, "packing %u (0 = bytes, 1 = 2-bit)", b->packing);
  const uint64_t n = b->n_pairs;
  if (n == 0) return SA_OK;
  if (!counts || !b->q_off || !b->q_len || !b->d_off || !b->d_len || (!b->residues && b->residues_len))
    return fail(e, SA_E_ARG, "null array");
  if (n >= (1ull << 31)) return fail(e, SA_E_ARG, "n_pairs %llu too large", (unsigned long long)n);
  CUDA_TRY(e, cudaSetDevice(e->device));
  sa_scheme_t sc = sa_scheme_t{5, -4, -8, -6};  // nw_affine.rs:15-20
  if (scheme) sc = *scheme;
  const uint64_t limit = b->packing ? b->residues_len * 4 : b->residues_len;
  uint32_t n1max = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if ((b->q_len[i] && !view_in_bounds(b->q_off[i], b->q_len[i], limit)) || (b->d_len[i] && !view_in_bounds(b->d_off[i], b->d_len[i], limit)))
      return fail(e, SA_E_ARG, "pair %llu reaches past residues_len", (unsigned long long)i);
    n1max = std::max(n1max, b->q_len[i]);
  }
  sa_status_t st;
  if ((st = ensure(e, e->b_res, b->residues_len)) != SA_OK) return st;
  if ((st = ensure(e, e->b_qoff, n * 8)) != SA_OK) return st;
  if ((st = ensure(e, e->b_doff, n * 8)) != SA_OK) return st;
  if ((st = ensure(e, e->b_qlen, n * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->b_dlen, n * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->b_coff, n * 8)) != SA_OK) return st;  // the counts
  cudaStream_t s = e->stream;
  CUDA_TRY(e, cudaMemcpyAsync(e->b_res.p, b->residues, b->residues_len, cudaMemcpyHostToDevice, s));
  CUDA_TRY(e, cudaMemcpyAsync(e->b_qoff.p, b->q_off, n * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(e, cudaMemcpyAsync(e->b_doff.p, b->d_off, n * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(e, cudaMemcpyAsync(e->b_qlen.p, b->q_len, n * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(e, cudaMemcpyAsync(e->b_dlen.p, b->d_len, n * 4, cudaMemcpyHostToDevice, s));
  // rolling rows: 36 bytes per column per pair in flight; at most ~1 GB of scratch
  const uint64_t cols = (uint64_t)n1max + 1;
  uint64_t chunk = std::max<uint64_t>(128, std::min<uint64_t>(n, ((uint64_t)1 << 30) / (36 * cols)));
  chunk = std::min<uint64_t>(chunk, (uint64_t)1 << 20);
  if ((st = ensure(e, e->par_rows, chunk * cols * 12)) != SA_OK) return st;
  if ((st = ensure(e, e->par_bytes, chunk * cols * 24)) != SA_OK) return st;
  sa::CountParams cp{};
  cp.residues = (const uint8_t*)e->b_res.p;
  cp.q_off = (const uint64_t*)e->b_qoff.p;
  cp.q_len = (const uint32_t*)e->b_qlen.p;
  cp.d_off = (const uint64_t*)e->b_doff.p;
  cp.d_len = (const uint32_t*)e->b_dlen.p;
  cp.packing = b->packing;
  cp.match = sc.match;
  cp.mismatch = sc.mismatch;
  cp.open = sc.gap_open;
  cp.ext = sc.gap_ext;
  cp.rows = (int32_t*)e->par_rows.p;
  cp.cnts = (int64_t*)e->par_bytes.p;
  cp.cols = (uint32_t)cols;
  cp.out = (int64_t*)e->b_coff.p;
  for (uint64_t base = 0; base < n; base += chunk) {
    cp.pair_base = (uint32_t)base;
    cp.n_pairs = (uint32_t)std::min<uint64_t>(chunk, n - base);
    sa::nw_affine_count_kernel<<<(cp.n_pairs + 127) / 128, 128, 0, s>>>(cp);
    CUDA_TRY(e, cudaGetLastError());
  }
  CUDA_TRY(e, cudaMemcpyAsync(counts, e->b_coff.p, n * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(e, cudaStreamSynchronize(s));
  return SA_OK;
}

// Every co-optimal alignment of one pair, in the order and text of the reference's traceback
// (needleman_wunsch_affine.rs:246-329, Display :390-411): the device computes the parent sets,
// the host walks them with the reference's LIFO stack.  snprintf-style return (bytes needed).

int64_t sd_all_alignments(sa_engine* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                          const sa_scheme_t* scheme, uint64_t max_alignments, char* buf, size_t cap,
                          uint64_t* n_printed, int32_t* panicked) {
  if (!e || (n1 && !seq1) || (n2 && !seq2)) return SA_E_ARG;
  if (n_printed) *n_printed = 0;
  if (panicked) *panicked = 0;
  sa_scheme_t sc{5, -4, -8, -6};
  if (scheme) sc = *scheme;
  if (cudaSetDevice(e->device) != cudaSuccess) return fail(e, SA_E_CUDA, "cudaSetDevice failed");
  const uint64_t cells = (uint64_t)n1 * n2;
  if (cells > ((uint64_t)1 << 32)) return fail(e, SA_E_UNSUPPORTED, "pair too large for full parent sets");
  sa_status_t st;
  if ((st = ensure(e, e->par_bytes, cells + 16)) != SA_OK) return st;
  if ((st = ensure(e, e->par_rows, (size_t)6 * (n1 + 1) * 4 + 64)) != SA_OK) return st;
  if ((st = ensure(e, e->par_in, (size_t)n1 + n2 + 256)) != SA_OK) return st;
  // layout of par_in: [meta 128 B][seq1][seq2]
  struct Meta {
    uint64_t q_off, d_off, par_off;
    uint32_t q_len, d_len;
    int32_t end[3];
  } meta{0, n1, 0, n1, n2, {0, 0, 0}};
  uint8_t* d_in = (uint8_t*)e->par_in.p;
  cudaError_t err = cudaMemcpyAsync(d_in, &meta, sizeof(meta), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess && n1) err = cudaMemcpyAsync(d_in + 128, seq1, n1, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess && n2) err = cudaMemcpyAsync(d_in + 128 + n1, seq2, n2, cudaMemcpyHostToDevice, e->stream);
  if (err != cudaSuccess) return fail(e, SA_E_CUDA, "H2D failed: %s", cudaGetErrorString(err));
  sa::ParentsParams pp{};
  pp.residues = d_in + 128;
  pp.q_off = (const uint64_t*)(d_in + offsetof(Meta, q_off));
  pp.d_off = (const uint64_t*)(d_in + offsetof(Meta, d_off));
  pp.q_len = (const uint32_t*)(d_in + offsetof(Meta, q_len));
  pp.d_len = (const uint32_t*)(d_in + offsetof(Meta, d_len));
  pp.parents_off = (const uint64_t*)(d_in + offsetof(Meta, par_off));
  pp.end_scores = (int32_t*)(d_in + offsetof(Meta, end));
  pp.n_pairs = 1;
  pp.packing = 0;
  pp.match = sc.match;
  pp.mismatch = sc.mismatch;
  pp.open = sc.gap_open;
  pp.ext = sc.gap_ext;
  pp.parents = (uint8_t*)e->par_bytes.p;
  pp.rows = (int32_t*)e->par_rows.p;
  pp.row_stride = n1 + 1;
  sa::nw_affine_parents_kernel<<<1, 64, 0, e->stream>>>(pp);
  if ((err = cudaGetLastError()) != cudaSuccess) return fail(e, SA_E_CUDA, "launch failed: %s", cudaGetErrorString(err));
  std::vector<uint8_t> par(cells);
  int32_t end[3] = {0, 0, 0};
  if (cells) err = cudaMemcpyAsync(par.data(), e->par_bytes.p, cells, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(end, d_in + offsetof(Meta, end), 12, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) return fail(e, SA_E_CUDA, "parents kernel failed: %s", cudaGetErrorString(err));

  // ---- the reference's traceback loop over the device-computed parent lists -------------------
  enum { ST_M = 0, ST_D = 1, ST_I = 2 };
  struct Col {
    uint8_t c1, c2;
    int64_t next;
  };
  struct Frame {
    int st;
    uint32_t x, y;
    int64_t cols;
  };
  std::vector<Col> cols;
  std::vector<Frame> stack;
  std::string text;
  const int32_t em = end[0], ei = end[1], ed = end[2];
  const int32_t mx = std::max(std::max(ei, ed), em);  // :247-250
  if (mx == ei) stack.push_back({ST_I, n2, n1, -1});  // push order I, M, D (:251-280)
  if (mx == em) stack.push_back({ST_M, n2, n1, -1});
  if (mx == ed) stack.push_back({ST_D, n2, n1, -1});
  uint64_t printed = 0;
  bool pan = false;
  uint64_t needed = 0;  // bytes of the whole text; only the first cap - 1 are kept
  const size_t keep = (buf && cap) ? cap - 1 : 0;
  while (!stack.empty() && !pan) {
    const Frame f = stack.back();
    stack.pop_back();
    // columns made after this frame was pushed belong to subtrees that are finished: drop them
    cols.resize((size_t)(f.cols + 1));
    if (f.x == 0 && f.y == 0) {  // :283-286
      if (printed >= max_alignments) break;
      std::string r1, r2;
      for (int64_t k = f.cols; k >= 0; k = cols[k].next) {
        r1.push_back((char)cols[k].c1);
        r2.push_back((char)cols[k].c2);
      }
      std::string bars(r1.size(), ' ');
      for (size_t k = 0; k < r1.size(); ++k)
        if (r1[k] == r2[k]) bars[k] = '|';
      const std::string piece = "alignment found\n\nseq1: " + r1 + "\n      " + bars + "\nseq2: " + r2 + "\n";
      needed += piece.size();
      if (text.size() < keep) text.append(piece, 0, keep - text.size());
      ++printed;
    }
    // the popped cell's parent list, in push order
    int pst[3], np = 0;
    if (f.x >= 1 && f.y >= 1) {
      const uint8_t b = par[(uint64_t)(f.x - 1) * n1 + (f.y - 1)];
      if (f.st == ST_M) {
        if (b & 1) pst[np++] = ST_M;
        if (b & 2) pst[np++] = ST_I;
        if (b & 4) pst[np++] = ST_D;
      } else if (f.st == ST_I) {
        if (b & 8) pst[np++] = ST_I;
        if (b & 16) pst[np++] = ST_M;
      } else {
        if (b & 32) pst[np++] = ST_D;
        if (b & 64) pst[np++] = ST_M;
      }
    } else if (f.x == 0 && f.y >= 1 && f.st == ST_D) {
      pst[np++] = ST_D;  // boundary chain, parent d_scores[0][y-1] (:194-198)
    } else if (f.y == 0 && f.x >= 1 && f.st == ST_I) {
      pst[np++] = ST_I;  // parent i_scores[x-1][0] (:206-210)
    }
    for (int k = 0; k < np; ++k) {
      // the loop body indexes seq1[y-1] (InM, InI) and seq2[x-1] (InM, InD): panic on 0-1
      const bool bad = f.st == ST_M ? (f.x == 0 || f.y == 0) : (f.st == ST_D ? f.x == 0 : f.y == 0);
      if (bad) {
        pan = true;
        break;
      }
      Col c;
      c.next = f.cols;
      uint32_t x = f.x, y = f.y;
      if (f.st == ST_M) { c.c1 = seq1[y - 1]; c.c2 = seq2[x - 1]; --x; --y; }
      else if (f.st == ST_D) { c.c1 = '-'; c.c2 = seq2[x - 1]; --x; }
      else { c.c1 = seq1[y - 1]; c.c2 = '-'; --y; }
      cols.push_back(c);
      stack.push_back({pst[k], x, y, (int64_t)cols.size() - 1});
    }
  }
  if (n_printed) *n_printed = printed;
  if (panicked) *panicked = pan ? 1 : 0;
  if (buf && cap) {
    const size_t n = text.size() < cap - 1 ? text.size() : cap - 1;
    memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return (int64_t)needed;
}

// EVERY hit the reference's linear aligner prints for one pair (needleman_wunsch.rs:106-116, :205-254), as its text:
// the device fills the matrix (scores + move sets), the host finds the start cells -- (n1, n2) in global mode, every
// cell holding the maximum in row-major order in local mode (:256-272) -- and walks the move sets in stored order
// (Down, Right, Diag) with an explicit stack, printing a hit at (0, 0) or at a cell without moves.
int64_t sd_linear_all_hits(sa_engine* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, int local,
                           const sa_scheme_t* scheme, uint64_t max_hits, char* buf, size_t cap, uint64_t* n_printed) {
  if (!e || (n1 && !seq1) || (n2 && !seq2)) return SA_E_ARG;
  if (n_printed) *n_printed = 0;
  sa_scheme_t sc{5, -4, -8, -6};  // needleman_wunsch.rs:181-186
  if (scheme) sc = *scheme;
  if (cudaSetDevice(e->device) != cudaSuccess) return fail(e, SA_E_CUDA, "cudaSetDevice failed");
  const uint64_t w = (uint64_t)n2 + 1, cells = ((uint64_t)n1 + 1) * w;
  if (cells > ((uint64_t)1 << 28)) return fail(e, SA_E_UNSUPPORTED, "pair too large for full move sets");
  sa_status_t st;
  if ((st = ensure(e, e->par_bytes, cells + 2 * w + 16)) != SA_OK) return st;
  if ((st = ensure(e, e->par_rows, cells * 4 + 64)) != SA_OK) return st;
  if ((st = ensure(e, e->par_in, (size_t)n1 + n2 + 64)) != SA_OK) return st;
  uint8_t* d_in = (uint8_t*)e->par_in.p;
  cudaError_t err = cudaSuccess;
  if (n1) err = cudaMemcpyAsync(d_in, seq1, n1, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess && n2) err = cudaMemcpyAsync(d_in + n1, seq2, n2, cudaMemcpyHostToDevice, e->stream);
  if (err != cudaSuccess) return fail(e, SA_E_CUDA, "H2D failed: %s", cudaGetErrorString(err));
  sa::LinearMovesParams mp{};
  mp.seq1 = d_in;
  mp.seq2 = d_in + n1;
  mp.n1 = n1;
  mp.n2 = n2;
  mp.match = sc.match;
  mp.mismatch = sc.mismatch;
  mp.open = sc.gap_open;
  mp.ext = sc.gap_ext;
  mp.local = local ? 1 : 0;
  mp.scores = (int32_t*)e->par_rows.p;
  mp.moves = (uint8_t*)e->par_bytes.p;
  mp.gaps = (uint8_t*)e->par_bytes.p + cells;
  sa::nw_linear_moves_kernel<<<1, 32, 0, e->stream>>>(mp);
  if ((err = cudaGetLastError()) != cudaSuccess) return fail(e, SA_E_CUDA, "launch failed: %s", cudaGetErrorString(err));
  e->timing.kernel_launches++;
  std::vector<int32_t> scores(cells);
  std::vector<uint8_t> moves(cells);
  err = cudaMemcpyAsync(scores.data(), e->par_rows.p, cells * 4, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(moves.data(), e->par_bytes.p, cells, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) return fail(e, SA_E_CUDA, "moves kernel failed: %s", cudaGetErrorString(err));

  // start cells (:107-111)
  std::vector<std::pair<uint32_t, uint32_t>> starts;
  if (!local) {
    starts.emplace_back(n1, n2);
  } else {
    int32_t best = INT32_MIN;
    for (uint32_t i = 0; i <= n1; ++i)
      for (uint32_t j = 0; j <= n2; ++j) {
        const int32_t v = scores[(uint64_t)i * w + j];
        if (v > best) {
          best = v;
          starts.clear();
        }
        if (v == best) starts.emplace_back(i, j);
      }
  }
  std::string text;
  uint64_t needed = 0, printed = 0;
  const size_t keep = (buf && cap) ? cap - 1 : 0;
  auto put = [&](const std::string& t) {
    if (text.size() < keep) text.append(t, 0, std::min(t.size(), keep - text.size()));
    needed += t.size();
  };
  struct Frame {
    uint32_t i, j;
    uint8_t next;  // index of the next move to try: 0 Down, 1 Right, 2 Diag
  };
  std::vector<Frame> stack;
  std::string q, d;  // the hit's columns in push order (end of the alignment first)
  for (size_t s = 0; s < starts.size() && printed < max_hits; ++s) {
    uint32_t hs1 = 0, hs2 = 0;  // Hit::default() per start cell (:113)
    q.clear();
    d.clear();
    stack.assign(1, Frame{starts[s].first, starts[s].second, 0});
    while (!stack.empty() && printed < max_hits) {
      Frame& f = stack.back();
      const uint8_t mv = moves[(uint64_t)f.i * w + f.j];
      if (f.next == 0 && ((f.i == 0 && f.j == 0) || mv == 0)) {  // :206-213
        std::string r1(q.rbegin(), q.rend()), r2(d.rbegin(), d.rend()), bars(q.size(), ' ');
        for (size_t k = 0; k < r1.size(); ++k)
          if (r1[k] == r2[k]) bars[k] = '|';
        put("\nHit: \nseq1: " + r1 + "\n      " + bars + "\nseq2: " + r2 + "\nstart in seq1: " + std::to_string(hs1) +
            "\nstart in seq2: " + std::to_string(hs2) + "\n\n\n\n");
        ++printed;
        stack.pop_back();
        if (!stack.empty()) {  // the caller's hit.query.pop() / hit.db.pop() (:251-252)
          q.pop_back();
          d.pop_back();
        }
        continue;
      }
      int k = f.next;
      while (k < 3 && !(mv & (1u << k))) ++k;
      if (k == 3) {
        stack.pop_back();
        if (!stack.empty()) {
          q.pop_back();
          d.pop_back();
        }
        continue;
      }
      f.next = (uint8_t)(k + 1);
      hs1 = std::max(f.i, 1u) - 1;  // :215-216
      hs2 = std::max(f.j, 1u) - 1;
      uint32_t ni = f.i, nj = f.j;
      if (k == 0) { q.push_back((char)seq1[f.i - 1]); d.push_back('-'); --ni; }
      else if (k == 1) { q.push_back('-'); d.push_back((char)seq2[f.j - 1]); --nj; }
      else { q.push_back((char)seq1[f.i - 1]); d.push_back((char)seq2[f.j - 1]); --ni; --nj; }
      stack.push_back(Frame{ni, nj, 0});  // (invalidates f)
    }
  }
  if (n_printed) *n_printed = printed;
  if (buf && cap) {
    memcpy(buf, text.data(), text.size());
    buf[text.size()] = 0;
  }
  return (int64_t)needed;
}

}  // namespace sa_host
