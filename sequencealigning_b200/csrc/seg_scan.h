// seg_scan.h -- the host's pass over one segment of the pair list (pure C++, no CUDA: tests/test_seg_scan.py builds it
// with g++).  For the pairs [base, base + count) of the reference's loop (/root/reference/src/main.rs:61-62) it finds
// what the engine needs before it can launch the segment: the shape maxima over the pairs the packed kernel takes (the
// others are "long"), the real cells, and -- when the inputs stream from the host -- the residue ranges to upload, one
// per block of 4 096 pairs and side, merged when closer than 64 KB.
//
// The pass costs ~3 ns per pair on the GPU box; only the first segment's (0.2 ms) is time the GPU waits for.  It can
// be split over threads (blocks of 4 096 pairs stay whole, so the result is identical; SA_SCAN_THREADS), but starting
// threads for a 0.2-0.4 ms job measured slower than the one-thread pass, so the default is one.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

namespace sa_host {

// end of a residue view, saturating: a garbage 64-bit offset must not wrap past the bounds check
inline uint64_t view_end(uint64_t off, uint32_t len) { return off + len < off ? ~0ull : off + len; }
inline bool view_in_bounds(uint64_t off, uint32_t len, uint64_t limit) { return off <= limit && len <= limit - off; }

struct SegScanIn {
  const uint32_t* q_len = nullptr;  // whole-list arrays, indexed by pair id
  const uint32_t* d_len = nullptr;
  const uint64_t* q_off = nullptr;  // null: no residue ranges wanted (inputs already resident)
  const uint64_t* d_off = nullptr;
  bool linear = false;              // the linear aligner runs with the roles swapped: columns walk seq2
  uint64_t cols_lim = 0, rows_lim = 0;  // a pair with more columns or rows is "long" (separable test)
};

struct SegScanOut {
  uint32_t n_long = 0;
  uint64_t real = 0;   // cells of the pairs the packed kernel takes
  uint64_t cells = 0;  // cells of all pairs
  uint32_t n1max = 0, n2max = 0, n2min = ~0u;          // columns / rows over the packed kernel's pairs
  uint64_t qlo = ~0ull, qhi = 0, dlo = ~0ull, dhi = 0;  // extent of the residues the segment touches
  std::vector<std::pair<uint64_t, uint64_t>> ranges;   // per block and side, in block order (unmerged)
};

constexpr uint32_t kSegScanBlock = 4096;

// pairs [base + lo, base + hi); lo is a multiple of kSegScanBlock
inline void seg_scan_part(const SegScanIn& in, uint64_t base, uint32_t lo, uint32_t hi, SegScanOut& o) {
  uint32_t n_long = 0, n1m = 0, n2m = 0, n2lo = ~0u;
  uint64_t real = 0, cells = 0, qlo = ~0ull, qhi = 0, dlo = ~0ull, dhi = 0;
  uint64_t bql = ~0ull, bqh = 0, bdl = ~0ull, bdh = 0;  // extents of the current block
  const bool ranges = in.q_off != nullptr;
  auto flush_block = [&] {
    if (bql < bqh) o.ranges.emplace_back(bql, bqh);
    if (bdl < bdh) o.ranges.emplace_back(bdl, bdh);
    qlo = std::min(qlo, bql); qhi = std::max(qhi, bqh);
    dlo = std::min(dlo, bdl); dhi = std::max(dhi, bdh);
    bql = bdl = ~0ull;
    bqh = bdh = 0;
  };
  for (uint32_t i = lo; i < hi; ++i) {
    const uint64_t p = base + i;
    const uint32_t ql = in.q_len[p], dl = in.d_len[p];
    const uint64_t c = (uint64_t)ql * dl;
    cells += c;
    if (ranges) {
      const uint64_t qo = in.q_off[p], dO = in.d_off[p];
      if (ql) {
        bql = std::min(bql, qo);
        bqh = std::max(bqh, view_end(qo, ql));
      }
      if (dl) {
        bdl = std::min(bdl, dO);
        bdh = std::max(bdh, view_end(dO, dl));
      }
      if ((i & (kSegScanBlock - 1)) == kSegScanBlock - 1) flush_block();
    }
    const uint32_t a = in.linear ? dl : ql, b = in.linear ? ql : dl;
    if (a && b && (a > in.cols_lim || b > in.rows_lim)) {
      ++n_long;
      continue;
    }
    real += c;
    n1m = std::max(n1m, a);
    n2m = std::max(n2m, b);
    n2lo = std::min(n2lo, b);
  }
  if (ranges) flush_block();
  o.n_long = n_long; o.real = real; o.cells = cells;
  o.n1max = n1m; o.n2max = n2m; o.n2min = n2lo;
  o.qlo = qlo; o.qhi = qhi; o.dlo = dlo; o.dhi = dhi;
}

// the whole segment on `threads` threads (1 = the caller's); the result does not depend on the thread count
inline SegScanOut seg_scan(const SegScanIn& in, uint64_t base, uint32_t count, int threads) {
  const uint32_t blocks = (count + kSegScanBlock - 1) / kSegScanBlock;
  const int T = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)std::max(threads, 1), blocks));
  std::vector<SegScanOut> part((size_t)T);
  auto run = [&](int k) {
    const uint32_t b0 = (uint32_t)((uint64_t)blocks * k / T), b1 = (uint32_t)((uint64_t)blocks * (k + 1) / T);
    seg_scan_part(in, base, std::min(count, b0 * kSegScanBlock), std::min(count, b1 * kSegScanBlock), part[(size_t)k]);
  };
  if (T == 1) {
    run(0);
  } else {
    std::vector<std::thread> th;
    th.reserve((size_t)T - 1);
    for (int k = 1; k < T; ++k) th.emplace_back(run, k);
    run(0);
    for (auto& t : th) t.join();
  }
  SegScanOut o = std::move(part[0]);
  for (int k = 1; k < T; ++k) {
    const SegScanOut& s = part[(size_t)k];
    o.n_long += s.n_long; o.real += s.real; o.cells += s.cells;
    o.n1max = std::max(o.n1max, s.n1max); o.n2max = std::max(o.n2max, s.n2max); o.n2min = std::min(o.n2min, s.n2min);
    o.qlo = std::min(o.qlo, s.qlo); o.qhi = std::max(o.qhi, s.qhi);
    o.dlo = std::min(o.dlo, s.dlo); o.dhi = std::max(o.dhi, s.dhi);
    o.ranges.insert(o.ranges.end(), s.ranges.begin(), s.ranges.end());
  }
  return o;
}

// overlapping or closer than `gap` bytes (a copy has a fixed cost): a record-ordered buffer becomes one interval
inline void merge_ranges(std::vector<std::pair<uint64_t, uint64_t>>& r, uint64_t gap) {
  std::sort(r.begin(), r.end());
  size_t w = 0;
  for (size_t k = 0; k < r.size(); ++k) {
    if (w && r[k].first <= r[w - 1].second + gap)
      r[w - 1].second = std::max(r[w - 1].second, r[k].second);
    else
      r[w++] = r[k];
  }
  r.resize(w);
}

}  // namespace sa_host
