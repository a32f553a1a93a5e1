// CPU check of sequencealigning_b200/csrc/seg_scan.h (built and run by tests/test_seg_scan.py): the threaded scan of a
// segment equals the one-thread scan and a plain restatement, for random pair lists with empty sides, long pairs,
// unordered offsets, segments that do not end on a block, and both roles of the sequences.
#include <cstdio>
#include <random>

#include "../../sequencealigning_b200/csrc/seg_scan.h"

using namespace sa_host;

static bool same(const SegScanOut& a, const SegScanOut& b) {
  return a.n_long == b.n_long && a.real == b.real && a.cells == b.cells && a.n1max == b.n1max && a.n2max == b.n2max &&
         a.n2min == b.n2min && a.qlo == b.qlo && a.qhi == b.qhi && a.dlo == b.dlo && a.dhi == b.dhi && a.ranges == b.ranges;
}

// plain restatement: per-pair loop for the numbers, per-block loop for the ranges
static SegScanOut plain(const SegScanIn& in, uint64_t base, uint32_t count) {
  SegScanOut o;
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t ql = in.q_len[base + i], dl = in.d_len[base + i];
    const uint32_t a = in.linear ? dl : ql, b = in.linear ? ql : dl;
    o.cells += (uint64_t)ql * dl;
    const bool is_long = a != 0 && b != 0 && (a > in.cols_lim || b > in.rows_lim);
    if (is_long) { ++o.n_long; continue; }
    o.real += (uint64_t)ql * dl;
    if (a > o.n1max) o.n1max = a;
    if (b > o.n2max) o.n2max = b;
    if (b < o.n2min) o.n2min = b;
  }
  if (!in.q_off) return o;
  for (uint32_t b0 = 0; b0 < count; b0 += 4096) {
    uint64_t ql_ = ~0ull, qh = 0, dl_ = ~0ull, dh = 0;
    for (uint32_t i = b0; i < count && i < b0 + 4096; ++i) {
      const uint64_t p = base + i;
      if (in.q_len[p]) { ql_ = std::min(ql_, in.q_off[p]); qh = std::max(qh, view_end(in.q_off[p], in.q_len[p])); }
      if (in.d_len[p]) { dl_ = std::min(dl_, in.d_off[p]); dh = std::max(dh, view_end(in.d_off[p], in.d_len[p])); }
    }
    if (ql_ < qh) o.ranges.emplace_back(ql_, qh);
    if (dl_ < dh) o.ranges.emplace_back(dl_, dh);
    o.qlo = std::min(o.qlo, ql_); o.qhi = std::max(o.qhi, qh);
    o.dlo = std::min(o.dlo, dl_); o.dhi = std::max(o.dhi, dh);
  }
  return o;
}

int main() {
  std::mt19937_64 rng(20261019);
  const uint32_t n = 300000;
  std::vector<uint32_t> ql(n), dl(n);
  std::vector<uint64_t> qo(n), dO(n);
  int checks = 0;
  for (int flavour = 0; flavour < 3; ++flavour) {
    uint64_t pos = 0;
    for (uint32_t i = 0; i < n; ++i) {
      ql[i] = 100 + (uint32_t)(rng() % 200);
      dl[i] = 100 + (uint32_t)(rng() % 200);
      if (rng() % 500 == 0) ql[i] = 0;
      if (rng() % 700 == 0) dl[i] = 0;
      if (rng() % 3000 == 0) ql[i] = 4000 + (uint32_t)(rng() % 1000);
      if (rng() % 3000 == 0) dl[i] = 2500;
      if (flavour == 0) {  // record order
        qo[i] = pos; pos += ql[i];
        dO[i] = pos; pos += dl[i];
      } else if (flavour == 1) {  // one query region, scattered db
        qo[i] = (rng() % 64) * 300;
        dO[i] = 1000000 + rng() % 50000000;
      } else {  // garbage offsets near the top of the range (saturating ends)
        qo[i] = ~0ull - rng() % 1000;
        dO[i] = rng();
      }
    }
    for (int linear = 0; linear < 2; ++linear)
      for (int with_ranges = 0; with_ranges < 2; ++with_ranges) {
        SegScanIn in;
        in.q_len = ql.data(); in.d_len = dl.data();
        in.q_off = with_ranges ? qo.data() : nullptr;
        in.d_off = with_ranges ? dO.data() : nullptr;
        in.linear = linear != 0;
        in.cols_lim = 3600; in.rows_lim = 1700;
        const uint64_t bases[] = {0, 5, 65536, 131077};
        const uint32_t counts[] = {1, 4095, 4096, 4097, 65536, 100001, 131072};
        for (uint64_t base : bases)
          for (uint32_t count : counts) {
            if (base + count > n) continue;
            const SegScanOut ref = plain(in, base, count);
            for (int t : {1, 2, 3, 4, 7, 64}) {
              const SegScanOut got = seg_scan(in, base, count, t);
              ++checks;
              if (!same(ref, got)) {
                printf("MISMATCH flavour %d linear %d ranges %d base %llu count %u threads %d\n", flavour, linear, with_ranges,
                       (unsigned long long)base, count, t);
                return 1;
              }
            }
          }
      }
  }
  // merge_ranges: overlapping, adjacent within the gap, apart
  std::vector<std::pair<uint64_t, uint64_t>> r = {{100, 200}, {150, 180}, {90000, 90010}, {210, 300}, {300000, 300001}};
  merge_ranges(r, 65536);
  if (r != std::vector<std::pair<uint64_t, uint64_t>>{{100, 300}, {90000, 90010}, {300000, 300001}}) {
    // {100,300} and {90000,..}: 90000 > 300 + 65536 -> apart
    printf("MERGE MISMATCH\n");
    return 1;
  }
  printf("ok %d checks\n", checks);
  return 0;
}
