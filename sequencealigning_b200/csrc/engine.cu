// engine.cu -- C ABI (include/sa_engine.h) over the CUDA kernels.  sm_100a only, no CPU path.
//
// Replaces the reference's per-pair dispatch loop (/root/reference/src/main.rs:61-79) with
// batched launches.  The pair list is cut into SEGMENTS (up to 128 Ki pairs, fewer when the
// packed traceback matrices would not fit the scratch budget).  Per segment:
//     copy-in stream : offsets/lengths of the segment + the residue ranges it touches
//     compute stream : fill (panic bonus on) -> classify/count walk -> [host reads the number of
//                      pairs whose end cell carries the bonus] -> clean refill + count walk of
//                      those pairs -> scan of CIGAR lengths -> write walks
//     copy-out stream: score / status / cigar_len / cigar_off of the segment
// The copy-in of segment i+1 is enqueued before the host waits on segment i, so transfers,
// kernels and result copies of neighbouring segments overlap; the CIGAR pool is copied once at
// the end (its size is only known then).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "engine_internal.h"
#include "engine_util.h"
#include "nw_affine_s16.cuh"
#include "nw_walk.cuh"
#include "nw_general.cuh"
#include "nw_long.cuh"
#include "nw_order.cuh"

namespace sa_host {

sa_status_t fail(sa_engine* e, sa_status_t st, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return st;
}

}  // namespace sa_host

namespace {

using sa_host::fail;
using sa_host::ensure;
using sa_host::view_end;
using sa_host::SegScanIn;
using sa_host::SegScanOut;
using sa_host::seg_scan;
using sa_host::merge_ranges;
using sa_host::kSegScanBlock;
using sa_host::view_in_bounds;
using sa_host::run_wfa;
using sa_host::run_linear_local;

struct Geometry {
  int K = 8;  // columns per lane per row (8, 13, 16 or 19)
  int G = 4;  // lanes per pair-of-pairs
  bool single = false;  // K*G covers every column: one pass, no boundary column
  uint32_t w = 1;       // traceback words (uint2) per strip row
  uint32_t ng = 8, ppt = 16;
  uint32_t nstrips_pad = 0, n1pad = 0, tb_rows = 0;
  uint64_t tile_stride = 0;  // uint2 per tile
  size_t smem_bytes = 0;
};

// dynamic shared memory of one fill warp: the db residue panel (u16 per row per pair-of-pairs)
// and, for multi-pass launches, the boundary column (uint2 per row per pair-of-pairs)
inline uint64_t fill_smem_bytes(uint64_t rows, uint64_t ng, bool single = false) {
  return (rows * ng * (single ? 2 : 10) + 15) & ~(uint64_t)15;
}

Geometry make_geometry(int K, int G, uint32_t n1max, uint32_t n2max) {
  Geometry g;
  g.K = K;
  g.G = G;
  g.w = (uint32_t)(K + 7) / 8;
  g.ng = 32 / G;
  g.ppt = 2 * g.ng;
  const uint32_t nstrips = (n1max + K - 1) / K;
  const uint32_t npass = std::max(1u, (nstrips + G - 1) / G);
  g.single = (K != 8) && npass == 1;  // the 8-column kernels keep the general multi-pass form
  g.nstrips_pad = npass * G;
  g.n1pad = g.nstrips_pad * K;
  g.tb_rows = std::max(1u, n2max);
  g.tile_stride = (uint64_t)g.nstrips_pad * g.tb_rows * g.ng * g.w;
  g.smem_bytes = (size_t)fill_smem_bytes(g.tb_rows, g.ng, g.single);
  return g;
}

// The (K, G) forms that are compiled.  K = 13, 16 and 19 exist only as single-pass kernels
// (100/200 bp, 125/250 bp and 150/300 bp reads).
struct Form {
  int K, G, regs;  // regs: registers per thread of the instantiation (cuobjdump -res-usage)
};
constexpr Form kForms[] = {{8, 1, 98},  {8, 2, 98},  {8, 4, 98},   {8, 8, 98},   {8, 16, 98},
                           {8, 32, 98}, {13, 8, 100}, {13, 16, 100}, {16, 8, 115}, {16, 16, 115}, {19, 8, 134}, {19, 16, 134}};

// Cost model for one shape class, in issue slots per pair:
//   steps = passes x (rows + G - 1 ramp rows), 16 instructions per column + ~24 per row step,
//   divided by the fraction of issue slots w resident warps keep busy (1 - 0.71^(w/4): fitted
//   to ncu at 14 and 16 warps per SM; the fill is bound by the latency of its tie-bit predicates).
template <class Fits>
Geometry choose_geometry(const sa_engine* e, uint32_t n1max, uint32_t n2max, bool linear, Fits&& fits) {
  Geometry best;
  best.G = 0;
  // a forced form (SA_FORCE_K / SA_FORCE_G) that cannot take the shape is ignored
  for (int forced = (e->force_g || e->force_k) ? 1 : 0; forced >= 0 && !best.G; --forced) {
    double best_cost = 1e300;
    for (const Form& f : kForms) {
      if (forced && e->force_g && f.G != e->force_g) continue;
      if (forced && e->force_k && f.K != e->force_k) continue;
      if (linear && f.K != 8) continue;
      const Geometry g = make_geometry(f.K, f.G, n1max, n2max);
      if (f.K != 8 && !g.single) continue;
      if (g.smem_bytes > e->smem_optin || !fits(g.n1pad, n2max)) continue;
      const double by_smem = std::floor(228.0 * 1024 / (double)(g.smem_bytes + 1024));
      const double by_regs = std::floor(65536.0 / (32.0 * ((f.regs + 7) / 8 * 8)));
      const double warps = std::min(32.0, std::min(by_smem, by_regs));
      if (warps < 1) continue;
      const double npass = g.nstrips_pad / f.G;
      const double steps = npass * (n2max + f.G - 1);
      const double instr = 16.0 * f.K + (g.single ? 21.0 : 25.0);
      const double busy = 1.0 - std::pow(0.71, warps / 4.0);
      // (multi-pass forms measure ~6 % slower than this count says: boundary-column traffic)
      const double cost = steps * instr / (double)g.ppt / busy * (g.single ? 1.0 : 1.06);
      if (cost < best_cost - 1e-12) {
        best_cost = cost;
        best = g;
      }
    }
  }
  return best;
}

template <int K, int G, uint32_t ORMASK, int ALGO, bool SINGLE, int MINB = 1>
sa_status_t launch_fill_m(sa_engine* e, const sa::AffineS16Params& p, const Geometry& g,
                          uint32_t n_tiles, cudaStream_t stream) {
  auto kern = sa::nw_affine_fill_s16<K, G, ORMASK, ALGO, SINGLE, MINB>;
  size_t& configured = e->smem_configured[(const void*)kern];  // per engine = per device; only grows
  if (g.smem_bytes > configured) {
    CUDA_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)std::min(e->smem_optin, std::max<size_t>(g.smem_bytes, 48 * 1024))));
    configured = std::max<size_t>(g.smem_bytes, 48 * 1024);
  }
  kern<<<n_tiles, 32, g.smem_bytes, stream>>>(p);
  CUDA_TRY(e, cudaGetLastError());
  e->timing.kernel_launches++;
  return SA_OK;
}

// 8-column multi-pass forms: every lane-group width, the linear aligner, the pipe-split knob
template <int G>
sa_status_t launch_fill8(sa_engine* e, const sa::AffineS16Params& p, const Geometry& g,
                         uint32_t n_tiles, cudaStream_t stream, int algo) {
  if (algo == SA_ALGO_NW_LINEAR) return launch_fill_m<8, G, 0x0F, sa::kLinear, false>(e, p, g, n_tiles, stream);
  switch (e->ormask) {
    case 0x00: return launch_fill_m<8, G, 0x00, sa::kAffine, false>(e, p, g, n_tiles, stream);
    case 0x0F: return launch_fill_m<8, G, 0x0F, sa::kAffine, false>(e, p, g, n_tiles, stream);
  }
  return fail(e, SA_E_ARG, "SA_ORMASK 0x%x has no instantiation", e->ormask);
}

sa_status_t launch_fill_g(sa_engine* e, const sa::AffineS16Params& p, const Geometry& g,
                          uint32_t n_tiles, cudaStream_t stream, int algo = SA_ALGO_NW_AFFINE) {
  if (g.K == 8) {
    switch (g.G) {
      case 1: return launch_fill8<1>(e, p, g, n_tiles, stream, algo);
      case 2: return launch_fill8<2>(e, p, g, n_tiles, stream, algo);
      case 4: return launch_fill8<4>(e, p, g, n_tiles, stream, algo);
      case 8: return launch_fill8<8>(e, p, g, n_tiles, stream, algo);
      case 16: return launch_fill8<16>(e, p, g, n_tiles, stream, algo);
      case 32: return launch_fill8<32>(e, p, g, n_tiles, stream, algo);
    }
  } else if (g.single && algo == SA_ALGO_NW_AFFINE) {
    // single-pass forms
    if (g.K == 13 && g.G == 8) return launch_fill_m<13, 8, 0x00, sa::kAffine, true>(e, p, g, n_tiles, stream);
    if (g.K == 13 && g.G == 16) return launch_fill_m<13, 16, 0x00, sa::kAffine, true>(e, p, g, n_tiles, stream);
    if (g.K == 16 && g.G == 8) return launch_fill_m<16, 8, 0x00, sa::kAffine, true>(e, p, g, n_tiles, stream);
    if (g.K == 16 && g.G == 16) return launch_fill_m<16, 16, 0x00, sa::kAffine, true>(e, p, g, n_tiles, stream);
    // K = 19 wants ~147 registers (3 warps per scheduler); the 128-register build (4 per scheduler; its few spills are
    // outside the steady-state rows) is the faster one for G = 8 since the steady-state rows lost the end-cell test
    // (150 bp fill 3 190 -> 3 350 GCUPS) and the slower one for G = 16 (300 bp: 2 890 vs 2 930).  SA_FILL_MINB=1 / 16 force one.
    const uint32_t minb = e->fill_minb ? e->fill_minb : (g.G == 8 ? 16u : 1u);
    if (g.K == 19 && g.G == 8 && minb == 16 && e->ormask == 0xFF) return launch_fill_m<19, 8, 0x00, sa::kAffine, true, 16>(e, p, g, n_tiles, stream);
    if (g.K == 19 && g.G == 8 && minb == 16 && e->ormask == 0x1201) return launch_fill_m<19, 8, 0x1201, sa::kAffine, true, 16>(e, p, g, n_tiles, stream);
    if (g.K == 19 && g.G == 8 && minb == 16 && e->ormask == 0x01) return launch_fill_m<19, 8, 0x01, sa::kAffine, true, 16>(e, p, g, n_tiles, stream);
    if (g.K == 19 && g.G == 8 && minb == 16) return launch_fill_m<19, 8, 0x11201, sa::kAffine, true, 16>(e, p, g, n_tiles, stream);
    if (g.K == 19 && g.G == 16 && minb == 16) return launch_fill_m<19, 16, 0x00, sa::kAffine, true, 16>(e, p, g, n_tiles, stream);
    // Which tie-bit sets go to the alu pipe as LOP3 instead of the fma-heavy pipe as VIADD (SA_ORMASK).  With every
    // set a VIADD the fma-heavy pipe is the busier one (80 % vs 71 %, ncu); one of the eight sets of a cell pair as a
    // LOP3 in EVERY column tips it the other way (alu 80 %, fma-heavy 74 %); in every second column the pipes are
    // level.  Fill alone, 262 144 pairs x 150 bp after the row ordering: 0xFF (none) 3 102, 0x01 3 146, even columns
    // 3 159, odd columns 3 190 GCUPS.  (K = 16: no difference between the variants.)
    if (g.K == 19 && g.G == 8 && e->ormask == 0xFF) return launch_fill_m<19, 8, 0x00, sa::kAffine, true>(e, p, g, n_tiles, stream);
    if (g.K == 19 && g.G == 8 && e->ormask == 0x1201) return launch_fill_m<19, 8, 0x1201, sa::kAffine, true>(e, p, g, n_tiles, stream);  // even columns
    if (g.K == 19 && g.G == 8 && e->ormask == 0x01) return launch_fill_m<19, 8, 0x01, sa::kAffine, true>(e, p, g, n_tiles, stream);
    if (g.K == 19 && g.G == 8) return launch_fill_m<19, 8, 0x11201, sa::kAffine, true>(e, p, g, n_tiles, stream);  // odd columns
    if (g.K == 19 && g.G == 16) return launch_fill_m<19, 16, 0x00, sa::kAffine, true>(e, p, g, n_tiles, stream);
  }
  return fail(e, SA_E_ARG, "no fill kernel for K %d G %d", g.K, g.G);
}

uint32_t pack2(uint32_t v) { return v | (v << 16); }

// Byte ranges of the residue buffer already resident on the device (sorted, disjoint).
struct Coverage {
  std::vector<std::pair<uint64_t, uint64_t>> iv;
  // calls f(lo, hi) for every missing sub-range of [lo, hi) and marks it covered
  template <class F>
  void request(uint64_t lo, uint64_t hi, F&& f) {
    if (lo >= hi) return;
    std::vector<std::pair<uint64_t, uint64_t>> out;
    uint64_t cur = lo;
    uint64_t nlo = lo, nhi = hi;
    for (auto& p : iv) {
      if (p.second < lo || p.first > hi) {
        out.push_back(p);
        continue;
      }
      if (p.first > cur) f(cur, std::min(p.first, hi));
      cur = std::max(cur, p.second);
      nlo = std::min(nlo, p.first);
      nhi = std::max(nhi, p.second);
    }
    if (cur < hi) f(cur, hi);
    out.emplace_back(nlo, nhi);
    std::sort(out.begin(), out.end());
    iv.swap(out);
  }
};

struct Scheme2 {
  sa_scheme_t sc;
  int pen, openp, extp;
  int algo = SA_ALGO_NW_AFFINE;
};

// Launch plan of the literal long-pair kernels (nw_general.cuh) for a list of pairs.
struct LitPlan {
  std::vector<uint32_t> ids;
  std::vector<uint64_t> meta;   // per pair: tb offset (bytes, ~0 = none), runs end, checkpoint offset
                                // (records, ~0 = none), block offset (bytes)
  std::vector<uint32_t> waves;  // indices into ids where a new wave (reusing the tb words) starts
  uint64_t tb_total = 0, runs_total = 0;
  uint32_t n1max = 0, n2max = 0;
};

// Launch plan of the tiled long-pair path (nw_long.cuh).
struct FastPlan {
  std::vector<uint32_t> ids;
  std::vector<uint64_t> row_off, col_off, ck_off, runs_end;  // per pair (int2 units / words)
  struct Wave {
    uint32_t lo = 0, hi = 0;          // range of ids
    uint32_t tr_max = 0, tc_max = 0;  // tiles per dimension of the wave's largest pair
  };
  std::vector<Wave> waves;
  uint64_t edges_total = 0;  // int2 units (the largest wave)
  uint64_t runs_total = 0;   // words
  uint32_t R = 1024, S = 4;
  uint32_t back_warps = 0;
};

struct Segment {
  uint64_t base = 0;
  uint32_t n = 0;
  uint32_t n1max = 0, n2max = 0;
  std::vector<uint32_t> order;  // launch index -> pair id, sorted by shape; empty = identity
  bool dev_sort = false;        // one shape class whose rows differ a little: the DEVICE orders the pairs by rows (nw_order.cuh)
  uint32_t n2min = 0;           // fewest rows among the packed-kernel pairs
  // A ragged (hence sorted) segment is launched as a few shape classes, each with its own
  // lane-group width, tile size and traceback stride; a uniform segment is one class.
  struct Sub {
    uint32_t off = 0, cnt = 0;  // range of launch indices
    Geometry g;
    uint64_t tb_off = 0;        // uint2 offset of the class's first tile in the slot's scratch
  };
  std::vector<Sub> subs;
  uint64_t tb_total = 0;        // uint2 for the whole segment
  uint32_t n_short = 0;         // pairs handled by the packed kernel (= order.size() if explicit)
  // Pairs outside the packed 16-bit range.  Affine: the tiled 32-bit path (nw_long.cuh), with the
  // literal kernel (nw_general.cuh) as the fallback for pairs whose traceback can meet a dead end.
  // Linear: the literal kernel.
  LitPlan lit;
  FastPlan fast;
  uint32_t n_long = 0;
  uint64_t qlo = ~0ull, qhi = 0, dlo = ~0ull, dhi = 0;
  // residue ranges to upload (merged intervals of the per-4096-pair extents): tighter than [qlo, qhi) and
  // [dlo, dhi) when queries and db sequences live in separate regions of the buffer
  std::vector<std::pair<uint64_t, uint64_t>> res_ranges;
  uint64_t cells = 0;  // sum of n1*n2 over the segment
  Geometry g;
};

// The whole affine path over device views.  `in`/`out` non-null: stream inputs from / results
// to host buffers segment by segment; null: everything is already resident / stays resident.
struct Plan {
  std::vector<Segment> segs;
};

sa_status_t run_affine(sa_engine* e, DeviceBatch& db, uint64_t n, const uint32_t* h_q_len,
                       const uint32_t* h_d_len, const Scheme2& s2, bool want_cigar,
                       const sa_batch_t* in, sa_result_t* out, uint64_t* used_out,
                       uint64_t* pool_sent_out, Plan* plan = nullptr, bool plan_valid = false) {
  const sa_scheme_t& sc = s2.sc;
  sa_status_t st;
  *used_out = 0;
  *pool_sent_out = db.pool_base;

  // ---- scratch budget ---------------------------------------------------------------------
  size_t budget = e->tb_budget;
  if (!budget) {
    // cudaMemGetInfo is slow enough to show up per call; look once and keep the figure while
    // the scratch already held covers it
    const size_t held = e->slot[0].tb.cap + e->slot[1].tb.cap + e->tb2.cap;
    if (!e->budget_cached || held > e->budget_cached) {
      size_t free_b = 0, total_b = 0;
      CUDA_TRY(e, cudaMemGetInfo(&free_b, &total_b));
      e->budget_cached = std::min<size_t>((size_t)((double)(free_b + held) * 0.6), (size_t)96 << 30);
    }
    budget = e->budget_cached;
  }
  // two segment buffers and one refill region
  const size_t budget_main = budget * 2 / 5, budget_re = budget / 5;
  if ((st = ensure(e, e->misc, 256)) != SA_OK) return st;
  uint32_t* d_counts = (uint32_t*)e->misc.p;  // [2] refill queue lengths

  const bool linear = s2.algo == SA_ALGO_NW_LINEAR;
  sa::AffineS16Params fp{};
  fp.residues = db.residues;
  fp.packing = db.packing;
  fp.pen2 = pack2((uint32_t)s2.pen);
  fp.zero = 0;
  uint32_t row0_clean;
  if (linear) {
    // needleman_wunsch.rs: rows walk seq1, columns walk seq2 -> swap the kernel's roles.
    // S' = 2S - match*(i+j): a gap step costs match - 2*open (flag clear) or match - 2*ext.
    fp.q_off = db.d_off;
    fp.q_len = db.d_len;
    fp.d_off = db.q_off;
    fp.d_len = db.q_len;
    fp.open2 = pack2((uint32_t)(sc.match - 2 * sc.gap_open));
    fp.ext2 = pack2((uint32_t)(2 * (sc.gap_ext - sc.gap_open)));
    fp.step2 = pack2((uint32_t)s2.extp);
    fp.origin = pack2(sa::kBias - (uint32_t)(2 * s2.openp));  // S[0][0] = 2*open (:45-64)
    row0_clean = sa::kBias - (uint32_t)s2.openp;               // S[0][j] = open + j*ext
  } else {
    fp.q_off = db.q_off;
    fp.q_len = db.q_len;
    fp.d_off = db.d_off;
    fp.d_len = db.d_len;
    // V' = 2V - 2*ext*(x+y) + bias: extensions free, diagonal +cm, boundaries constant
    const uint32_t bias = sa::s16_affine_bias(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext);
    fp.open2 = pack2((uint32_t)s2.openp);
    fp.cm2 = pack2((uint32_t)(2 * sc.match - 4 * sc.gap_ext));
    fp.origin = pack2(bias);
    row0_clean = bias - (uint32_t)(s2.openp + (-2 * sc.gap_ext));  // 2*(open + (y+1)*ext) - 2*ext*y
  }
  // kernel columns / rows: (seq1, seq2) for affine, (seq2, seq1) for linear
  const uint32_t* h_cols = linear ? h_d_len : h_q_len;
  const uint32_t* h_rows = linear ? h_q_len : h_d_len;

  sa::WalkParams wp{};
  wp.q_len = db.q_len;
  wp.d_len = db.d_len;
  wp.match = sc.match;
  wp.bias = linear ? (int32_t)sa::kBias : (int32_t)sa::s16_affine_bias(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext);
  wp.diag2 = linear ? sc.match : 2 * sc.gap_ext;
  wp.open = sc.gap_open;
  wp.ext = sc.gap_ext;
  wp.score = db.score;
  wp.status = db.status;
  wp.cigar_len = db.cigar_len;
  wp.cigar_off = db.cigar_off;
  wp.pool = db.pool;
  wp.pool_cap = db.pool_cap;
  wp.pf = e->walk_pf;

  // Segment size: at least two segments (alternate streams overlap each other's tails), at
  // most seg_pairs; when inputs stream from the host the first segment is small.
  const uint64_t seg_max = e->seg_pairs_forced ? e->seg_pairs
                                                : std::min<uint64_t>(e->seg_pairs, std::max<uint64_t>(32768, (n + 1) / 2));
  uint64_t seg_target = (in && !e->seg_pairs_forced) ? std::min<uint64_t>(seg_max, e->seg_head) : seg_max;

  Coverage cov;
  // Scans the next segment on the host: extent, shape maxima, residue ranges, geometry.
  auto fits_packed = [&](uint32_t n1pad, uint32_t rows) {
    return sa::s16_min_value_bound(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext, std::max(1u, n1pad), rows) + 64 <= sa::kBias &&
           (linear || sa::s16_affine_in_range(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext, std::max(1u, n1pad), rows));
  };
  // kernel form for a class: the cheapest compiled (K, G) whose padded shape stays in range and
  // in shared memory.  G == 0: none.
  auto pick_geometry = [&](uint32_t cols, uint32_t rows) -> Geometry {
    return choose_geometry(e, cols, rows, linear, fits_packed);
  };
  // cheap per-pair test: no lane-group width keeps the pair inside the packed range and inside
  // one SM's shared memory (same arithmetic as make_geometry / s16_min_value_bound, no structs)
  const uint32_t bound0 = sa::s16_min_value_bound(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext, 0, 0) + 64;
  const uint32_t per_step = sa::s16_min_value_bound(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext, 1, 0) + 64 - bound0;
  const uint64_t max_sum = per_step ? (sa::kBias - std::min<uint32_t>(bound0, sa::kBias)) / per_step : 0;
  // fast path in three compares: with G = 32 (least shared memory, at most 255 columns of
  // padding) the pair is inside the sentinel guard, the affine range and one SM's shared memory
  const uint64_t cm_aff = (uint64_t)(2 * sc.match - 4 * sc.gap_ext);
  const uint64_t bias_aff = sa::s16_affine_bias(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext);
  const uint64_t nmin_fast = linear ? ~0ull : (cm_aff && bias_aff + 2 + cm_aff <= 0xFFFFull ? (0xFFFFull - bias_aff - 2) / cm_aff - 1 : 0);
  const uint64_t rows_fast = e->smem_optin >= 16 ? (e->smem_optin - 15) / 10 : 0;  // fill_smem_bytes(rows, 1) fits
  // The test is SEPARABLE (one limit per dimension), so that a shape class, whose column and row
  // maxima come from different pairs, fits a compiled form whenever each of its pairs does:
  // cols_lim + 256 + rows_lim <= max_sum, rows_lim <= rows_fast, rows_lim <= nmin_fast.  Skewed
  // pairs beyond a limit (3000 x 150, say) take the long-pair path.
  const uint64_t rows_lim = std::min<uint64_t>(std::min<uint64_t>(rows_fast, nmin_fast), max_sum > 256 ? (max_sum - 256) / 2 : 0);
  const uint64_t cols_lim = max_sum > 256 + rows_lim ? max_sum - 256 - rows_lim : 0;
  auto is_long = [&](uint32_t cols, uint32_t rows) -> bool {
    if (!cols || !rows) return false;
    return cols > cols_lim || rows > rows_lim;
  };
  // Literal long-pair kernels (nw_general.cuh): traceback bytes (one per cell), the pairs launched in
  // WAVES that each fit `room` (a wave's kernel also walks, so the next wave can reuse the bytes);
  // beyond 128 MB per pair the CHECKPOINTED form (right edge of every column pass + one pass-wide
  // block); only a pair that fits neither goes without (SA_ALIGNMENT_OMITTED: score and status exact).
  auto plan_literal = [&](const std::vector<uint32_t>& ids, uint64_t room) -> LitPlan {
    LitPlan lp;
    uint64_t wave_used = 0;
    for (uint32_t id : ids) lp.n1max = std::max(lp.n1max, h_cols[id]);
    const uint64_t pass_cols = (uint64_t)(lp.n1max >= 8192 ? sa::kGeneralThreadsWide : sa::kGeneralThreads) * sa::kGeneralCols;
    for (uint32_t id : ids) {
      const uint32_t a = h_cols[id], b = h_rows[id];
      const uint64_t words = (uint64_t)a * b;
      uint64_t off = ~0ull, ck_off = ~0ull, blk_off = ~0ull;
      auto take = [&](uint64_t bytes) -> uint64_t {  // 16-byte aligned room in the current wave
        bytes = (bytes + 15) & ~(uint64_t)15;
        if (wave_used + bytes > room) {
          lp.waves.push_back((uint32_t)lp.ids.size());
          wave_used = 0;
        }
        const uint64_t at = wave_used;
        wave_used += bytes;
        lp.tb_total = std::max(lp.tb_total, wave_used);
        return at;
      };
      const uint64_t npass = (a + pass_cols - 1) / pass_cols;
      const uint64_t ck_bytes = npass * ((uint64_t)b + 2) * 16, blk_bytes = (uint64_t)b * pass_cols;
      const bool can_ckpt = !linear && ck_bytes + blk_bytes <= room / 2;
      if (can_ckpt && (e->long_ckpt_always || words > ((uint64_t)128 << 20))) {
        const uint64_t at = take(ck_bytes + blk_bytes);
        ck_off = at / 16;
        blk_off = at + ck_bytes;
      } else if (words <= room / 8) {  // (at least 8 pairs per wave: one block per pair)
        off = take(words);
      }
      lp.runs_total += (uint64_t)a + b + 1;
      lp.ids.push_back(id);
      lp.meta.push_back(off);
      lp.meta.push_back(lp.runs_total);
      lp.meta.push_back(ck_off);
      lp.meta.push_back(blk_off);
      lp.n2max = std::max(lp.n2max, b);
    }
    return lp;
  };
  // Tiled long-pair path (nw_long.cuh): tile shape by the number of pairs (few pairs -> small tiles,
  // more of them in flight), edge storage per pair, waves that each fit `room`.
  auto plan_fast = [&](const std::vector<uint32_t>& ids, uint64_t room) -> FastPlan {
    FastPlan fp_;
    const size_t nl = ids.size();
    // (measured, 256 x 50 kbp: S = 2 is ~10 % faster than S = 4; S = 4 halves the kept column edges)
    fp_.S = e->long_s ? e->long_s : (nl >= 16 ? 2u : 1u);
    fp_.R = e->long_r ? e->long_r : (nl >= 64 ? 1024u : (nl >= 16 ? 512u : 256u));
    // backward warps: two CTAs of four per SM (shared memory bound), no more than pairs
    fp_.back_warps = (uint32_t)std::min<uint64_t>(((uint64_t)nl + 3) / 4 * 4, (uint64_t)e->sm_count * 8);
    auto edges_of = [&](uint32_t a, uint32_t b, uint32_t S, uint64_t* row, uint64_t* col, uint64_t* ck) {
      const uint64_t n1pad = ((uint64_t)a + sa::kLongStrip - 1) / sa::kLongStrip * sa::kLongStrip;
      const uint64_t tc = ((uint64_t)a + (uint64_t)S * sa::kLongStrip - 1) / ((uint64_t)S * sa::kLongStrip);
      *row = n1pad;
      *col = (tc - 1) * ((uint64_t)b + 1);
      *ck = (((uint64_t)b - 1) / sa::kLongMr) * n1pad;
    };
    // wider tiles (fewer kept column edges): to S = 4 when that lets all pairs share one wave,
    // beyond when even one pair would not fit otherwise
    for (;;) {
      uint64_t worst = 0, total = 0;
      for (uint32_t id : ids) {
        uint64_t r, c, k;
        edges_of(h_cols[id], h_rows[id], fp_.S, &r, &c, &k);
        worst = std::max(worst, (r + c + k) * 8);
        total += (r + c + k) * 8;
      }
      const uint64_t back = (uint64_t)fp_.back_warps * fp_.S * 32 * sa::kLongMr * 8;
      const bool fits = fp_.S < 4 ? total + back <= room : worst + back <= room;
      if (fits || fp_.S >= 16 || e->long_s) break;
      fp_.S *= 2;
    }
    const uint64_t back = (uint64_t)fp_.back_warps * fp_.S * 32 * sa::kLongMr * 8;
    const uint64_t room_edges = room > back + (1u << 20) ? (room - back) / 8 : (1u << 17);  // int2 units
    FastPlan::Wave w;
    uint64_t used = 0;
    for (uint32_t id : ids) {
      const uint32_t a = h_cols[id], b = h_rows[id];
      uint64_t r, c, k;
      edges_of(a, b, fp_.S, &r, &c, &k);
      if (used && used + r + c + k > room_edges) {
        w.hi = (uint32_t)fp_.ids.size();
        fp_.waves.push_back(w);
        w = FastPlan::Wave{};
        w.lo = (uint32_t)fp_.ids.size();
        used = 0;
      }
      fp_.row_off.push_back(used);
      fp_.col_off.push_back(used + r);
      fp_.ck_off.push_back(used + r + c);
      used += r + c + k;
      fp_.edges_total = std::max(fp_.edges_total, used);
      fp_.runs_total += (uint64_t)a + b + 1;
      fp_.runs_end.push_back(fp_.runs_total);
      fp_.ids.push_back(id);
      w.tr_max = std::max<uint32_t>(w.tr_max, (b + fp_.R - 1) / fp_.R);
      w.tc_max = std::max<uint32_t>(w.tc_max, (uint32_t)(((uint64_t)a + (uint64_t)fp_.S * sa::kLongStrip - 1) / ((uint64_t)fp_.S * sa::kLongStrip)));
    }
    w.hi = (uint32_t)fp_.ids.size();
    if (w.hi > w.lo) fp_.waves.push_back(w);
    return fp_;
  };
  auto prepare_fresh = [&](uint64_t base, Segment& sg) -> sa_status_t {
    sg = Segment{};
    sg.base = base;
    uint32_t cn = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(seg_target, 1u << 24), n - base);
    if (in && !e->seg_pairs_forced) {
      // Streaming from the host: segment sizes ramp up so the first copy-in is short, and the
      // batch ends with a ~128 K and a ~64 K segment so that little work (stage B of the last two
      // segments, their copy-out) is left when the last fill ends.  (Measured and dropped: a geometric tail
      // down to 16 Ki pairs and a 16 Ki first segment -- the small segments' host round trips and partial waves
      // cost more than they hide: 2 720 -> 2 565 GCUPS resident, 2 305 -> 2 277 end to end.)
      const uint64_t rem = n - base, last = 65536, second_last = 131072;
      if (rem <= last + last / 2)
        cn = (uint32_t)rem;
      else if (rem <= last + second_last + last / 2)
        cn = (uint32_t)(rem - last);
      else
        cn = (uint32_t)std::min<uint64_t>(cn, rem - last - second_last);
    }
    if (seg_target < seg_max) seg_target = std::min<uint64_t>(seg_max, seg_target * 2);
    // ONE pass over the segment's pairs (this scan is host time the GPU may be waiting on):
    // shape maxima over the pairs the packed kernel can take (the others are "long"), real
    // cells, and -- when streaming from the host -- the residue ranges to upload
    uint32_t n_long = 0;
    uint64_t real = 0;
    // (seg_scan.h; SA_SCAN_THREADS > 1 splits the scan of a streamed call's first segments over threads)
    auto scan = [&](uint32_t count) {
      SegScanIn si;
      si.q_len = h_q_len;
      si.d_len = h_d_len;
      si.q_off = in ? in->q_off : nullptr;
      si.d_off = in ? in->d_off : nullptr;
      si.linear = linear;
      si.cols_lim = cols_lim;
      si.rows_lim = rows_lim;
      const int threads = (in && base < e->scan_mt_pairs && count >= 4 * kSegScanBlock) ? e->scan_threads : 1;
      SegScanOut so = seg_scan(si, base, count, threads);
      n_long = so.n_long;
      real = so.real;
      sg.res_ranges.swap(so.ranges);
      if (in) merge_ranges(sg.res_ranges, 65536);
      sg.n1max = so.n1max;
      sg.n2max = so.n2max;
      sg.n2min = so.n2min == ~0u ? 0u : so.n2min;
      sg.qlo = so.qlo; sg.qhi = so.qhi; sg.dlo = so.dlo; sg.dhi = so.dhi;
      sg.cells = so.cells;
    };
    scan(cn);
    sg.g = pick_geometry(sg.n1max, sg.n2max);
    if (!sg.g.G) sg.g = make_geometry(8, 1, sg.n1max, sg.n2max);  // only possible when the segment has no short pair at all
    const size_t tile_bytes = (size_t)sg.g.tile_stride * 8;
    const uint64_t tiles_fit = std::max<uint64_t>(1, budget_main / tile_bytes);
    if ((uint64_t)cn > tiles_fit * sg.g.ppt) {  // the traceback scratch cannot hold the segment: shorten it
      cn = (uint32_t)(tiles_fit * sg.g.ppt);
      const Geometry keep = sg.g;
      scan(cn);
      sg.g = keep;
    }
    sg.n = cn;
    sg.n_long = n_long;
    if (n_long) {
      // explicit order without the long pairs; the long ones get their own launches
      sg.order.reserve(cn - n_long);
      std::vector<uint32_t> long_ids;
      long_ids.reserve(n_long);
      for (uint32_t i = 0; i < cn; ++i) {
        if (is_long(h_cols[base + i], h_rows[base + i])) long_ids.push_back((uint32_t)(base + i));
        else sg.order.push_back((uint32_t)(base + i));
      }
      // a segment of long pairs only does not need the packed kernel's scratch: use its share
      const uint64_t room = (n_long == cn) ? (uint64_t)budget_main + budget_re : (uint64_t)budget_re;
      if (linear || e->long_literal) sg.lit = plan_literal(long_ids, room);
      else sg.fast = plan_fast(long_ids, room);
    }
    const uint32_t ns = n_long ? (uint32_t)sg.order.size() : cn;
    sg.n_short = ns;
    auto id_at = [&](uint32_t i) -> uint32_t { return sg.order.empty() ? (uint32_t)(base + i) : sg.order[i]; };
    if (ns) {
      // Ragged segment: the 16-64 pairs of a warp tile all run to the tile's largest shape, so
      // bucket pairs by (rows, columns) when the padded work of the given order exceeds the real
      // work by more than ~15 % (uniform read sets skip this; the sort is host time).
      // Cheap bound first: if even padding every pair to the segment's largest shape stays within
      // 15 % of the real cells, the per-tile padding does too and the tile scan is skipped.
      const uint32_t ppt = sg.g.ppt;
      uint64_t padded = (uint64_t)sg.n1max * sg.n2max * ns;
      if (e->sort_mode == 0 && padded > real + real / 7 && ns > ppt) {
        padded = 0;
        for (uint32_t t0 = 0; t0 < ns; t0 += ppt) {
          uint32_t a = 0, b = 0;
          const uint32_t t1 = std::min(ns, t0 + ppt);
          for (uint32_t i = t0; i < t1; ++i) {
            const uint32_t id = id_at(i);
            a = std::max(a, h_cols[id]);
            b = std::max(b, h_rows[id]);
          }
          padded += (uint64_t)a * b * (t1 - t0);
        }
      }
      if (e->sort_mode == 1 || (e->sort_mode == 0 && padded > real + real / 7 && ns > ppt)) {
        // stable counting sort by columns, then by rows (LSD): O(n + longest sequence)
        std::vector<uint32_t> src_ids(ns), tmp(ns), cnt;
        for (uint32_t i = 0; i < ns; ++i) src_ids[i] = id_at(i);
        sg.order.resize(ns);
        cnt.assign((size_t)sg.n1max + 2, 0);
        for (uint32_t i = 0; i < ns; ++i) cnt[h_cols[src_ids[i]] + 1]++;
        for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
        for (uint32_t i = 0; i < ns; ++i) tmp[cnt[h_cols[src_ids[i]]]++] = src_ids[i];
        cnt.assign((size_t)sg.n2max + 2, 0);
        for (uint32_t i = 0; i < ns; ++i) cnt[h_rows[tmp[i]] + 1]++;
        for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
        for (uint32_t i = 0; i < ns; ++i) sg.order[cnt[h_rows[tmp[i]]]++] = tmp[i];
      }
    }
    {
      sg.subs.clear();
      auto add_class = [&](uint32_t off, uint32_t cnt, uint32_t cmax, uint32_t rmax) -> sa_status_t {
        if (!cnt) return SA_OK;
        Segment::Sub sub;
        sub.off = off;
        sub.cnt = cnt;
        sub.g = pick_geometry(cmax, rmax);
        if (!sub.g.G) {
          if (cmax && rmax)
            return fail(e, SA_E_UNSUPPORTED, "pair shape %u x %u fits no packed-kernel configuration", cmax, rmax);
          sub.g = make_geometry(8, 1, cmax, rmax);  // a class of empty pairs: nothing to fill
        }
        sub.tb_off = sg.tb_total;
        sg.tb_total += (uint64_t)((cnt + sub.g.ppt - 1) / sub.g.ppt) * sub.g.tile_stride;
        sg.subs.push_back(sub);
        return SA_OK;
      };
      if (sg.order.empty()) {
        sa_status_t rr = add_class(0, ns, sg.n1max, sg.n2max);
        if (rr != SA_OK) return rr;
        // one class, identity order, rows not all alike: let the device order the pairs by rows
        sg.dev_sort = e->sort_mode != 2 && !sg.subs.empty() && ns > 4 * sg.subs[0].g.ppt && sg.n2min < sg.n2max && sg.n2max < 11000;  // (bins in shared memory)
      } else {
        // (possibly) sorted by rows, then columns: cut where the rows have grown by more than a
        // quarter since the class began (at most 8 classes, at least 2048 pairs each)
        uint32_t start = 0, r_lo = ~0u, c_hi = 0, r_hi = 0;
        for (uint32_t i = 0; i < ns; ++i) {
          const uint32_t id = sg.order[i];
          const uint32_t c = h_cols[id], r = h_rows[id];
          const bool grow = i - start >= 2048 && sg.subs.size() < 7 &&
                            ((uint64_t)std::max(r_hi, r) * 4 > (uint64_t)std::max(r_lo, 16u) * 5 + 32);
          if (grow) {
            sa_status_t rr = add_class(start, i - start, c_hi, r_hi);
            if (rr != SA_OK) return rr;
            start = i;
            r_lo = ~0u;
            c_hi = r_hi = 0;
          }
          r_lo = std::min(r_lo, r);
          c_hi = std::max(c_hi, c);
          r_hi = std::max(r_hi, r);
        }
        sa_status_t rr = add_class(start, ns - start, c_hi, r_hi);
        if (rr != SA_OK) return rr;
      }
    }
    if (in) e->timing.cells += sg.cells;  // (a resident batch knows its cells from the upload)
    if (in) {
      const uint64_t limit = in->packing ? in->residues_len * 4 : in->residues_len;
      if (sg.qhi > limit || sg.dhi > limit)
        return fail(e, SA_E_ARG, "a pair in [%llu, %llu) reaches past residues_len",
                    (unsigned long long)base, (unsigned long long)(base + cn));
    }
    return SA_OK;
  };
  // Resident batches keep their segment plan between calls (the host scan is then skipped).
  size_t plan_pos = 0;
  auto prepare = [&](uint64_t base, Segment& sg) -> sa_status_t {
    if (plan && plan_valid && plan_pos < plan->segs.size() && plan->segs[plan_pos].base == base) {
      sg = plan->segs[plan_pos++];
      return SA_OK;
    }
    const sa_status_t r = prepare_fresh(base, sg);
    if (r == SA_OK && plan && !plan_valid) plan->segs.push_back(sg);
    return r;
  };
  // Enqueues the segment's inputs on the copy-in stream.
  auto upload = [&](const Segment& sg) -> sa_status_t {
    if (!in) return SA_OK;
    cudaError_t err = cudaSuccess;
    auto cp = [&](void* dst, const void* src, size_t bytes) {
      if (err == cudaSuccess && bytes) {
        err = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->s_in);
        e->timing.h2d_bytes += bytes;
      }
    };
    cp(db.q_off + sg.base, in->q_off + sg.base, (size_t)sg.n * 8);
    cp(db.d_off + sg.base, in->d_off + sg.base, (size_t)sg.n * 8);
    cp(db.q_len + sg.base, in->q_len + sg.base, (size_t)sg.n * 4);
    cp(db.d_len + sg.base, in->d_len + sg.base, (size_t)sg.n * 4);
    auto range = [&](uint64_t lo, uint64_t hi) { cp(db.residues + lo, in->residues + lo, hi - lo); };
    const int sh = in->packing ? 2 : 0;  // residue index -> byte index
    for (const auto& r : sg.res_ranges) cov.request(r.first >> sh, (r.second + (sh ? 3 : 0)) >> sh, range);
    if (err != cudaSuccess) return fail(e, SA_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(err));
    CUDA_TRY(e, cudaEventRecord(e->ev_in, e->s_in));
    return SA_OK;
  };
  auto set_geometry = [&](const Geometry& g) {
    fp.tb_tile_stride = g.tile_stride;
    fp.tb_rows = g.tb_rows;
    fp.smem_bnd_rows = g.tb_rows;
    wp.tb_tile_stride = g.tile_stride;
    wp.tb_rows = g.tb_rows;
    wp.ng = g.ng;
    wp.k = (uint32_t)g.K;
    wp.k_inv = (uint32_t)((0x100000000ull + (uint32_t)g.K - 1) / (uint32_t)g.K);
    wp.w = g.w;
  };
  // Streams the CIGAR pool to the host as segments finish.  After segment i's write walks a
  // copy of the running total is queued (slot i&1); one segment later the host reads it (by then
  // the walks are complete) and queues pool[pool_sent, total) on the copy-out stream.
  uint64_t pool_sent = db.pool_base;
  int seg_index = 0;
  auto send_pool_upto = [&](int slot) -> cudaError_t {
    cudaError_t er = cudaEventSynchronize(e->ev_carry[slot]);
    if (er != cudaSuccess) return er;
    uint64_t total;
    memcpy(&total, e->h_count + 8 + 2 * slot, 8);
    const uint64_t hi = std::min<uint64_t>(std::min<uint64_t>(total, db.pool_cap), out->cigar_capacity);
    if (hi > pool_sent) {
      er = cudaMemcpyAsync(out->cigar + pool_sent, db.pool + pool_sent, (hi - pool_sent) * 4,
                           cudaMemcpyDeviceToHost, e->s_out);
      e->timing.d2h_bytes += (hi - pool_sent) * 4;
      pool_sent = hi;
    }
    return er;
  };
  // Launches the literal long-pair kernels for a plan: one thread block per pair, one launch per wave.
  auto launch_literal = [&](const LitPlan& lp, sa_engine::LitBufs& gb, cudaStream_t sx) -> sa_status_t {
    sa_status_t r;
    const uint32_t nl = (uint32_t)lp.ids.size();
    const bool wide = lp.n1max >= 8192;  // more lanes per pair when the pairs are few and long
    // per pair 6 * stride ints: the affine kernel's edge column (4 ints per row), the linear
    // kernel's rolling row (ceil(columns / threads) * threads entries); even, for 16-byte alignment
    const uint32_t stride = (std::max(lp.n1max, lp.n2max) + 2 + sa::kGeneralThreadsWide + 1) & ~1u;
    if ((r = ensure(e, gb.ids, (size_t)nl * 4)) != SA_OK) return r;
    if ((r = ensure(e, gb.meta, (size_t)nl * (32 + sizeof(sa::LongWalkState)))) != SA_OK) return r;
    if ((r = ensure(e, gb.tb, (size_t)lp.tb_total + 256)) != SA_OK) return r;
    if ((r = ensure(e, gb.rows, (size_t)nl * 6 * stride * 4)) != SA_OK) return r;
    if ((r = ensure(e, gb.info, (size_t)nl * 4 * stride)) != SA_OK) return r;
    if ((r = ensure(e, gb.runs, (size_t)lp.runs_total * 4 + 256)) != SA_OK) return r;
    std::vector<uint64_t> meta((size_t)4 * nl);  // [tb_off | runs_end | ck_off | blk_off]
    bool any_ckpt = false;
    for (uint32_t t = 0; t < nl; ++t) {
      for (int f = 0; f < 4; ++f) meta[(size_t)f * nl + t] = lp.meta[4 * t + f];
      any_ckpt |= lp.meta[4 * t + 2] != ~0ull;
    }
    CUDA_TRY(e, cudaMemcpyAsync(gb.ids.p, lp.ids.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(gb.meta.p, meta.data(), (size_t)nl * 32, cudaMemcpyHostToDevice, sx));
    sa::GeneralParams gp{};
    gp.residues = db.residues;
    gp.q_off = db.q_off;
    gp.q_len = db.q_len;
    gp.d_off = db.d_off;
    gp.d_len = db.d_len;
    gp.ids = (const uint32_t*)gb.ids.p;
    gp.n_ids = nl;
    gp.packing = db.packing;
    gp.match = sc.match;
    gp.mismatch = sc.mismatch;
    gp.open = sc.gap_open;
    gp.ext = sc.gap_ext;
    gp.tb = want_cigar ? (uint8_t*)gb.tb.p : nullptr;
    gp.ck = want_cigar ? (int4*)gb.tb.p : nullptr;  // checkpoints and blocks share the wave's scratch
    gp.ck_off = (const uint64_t*)gb.meta.p + 2 * (size_t)nl;
    gp.blk = (uint8_t*)gb.tb.p;
    gp.blk_off = (const uint64_t*)gb.meta.p + 3 * (size_t)nl;
    gp.ws = (sa::LongWalkState*)((uint64_t*)gb.meta.p + 4 * (size_t)nl);
    gp.tb_off = (const uint64_t*)gb.meta.p;
    gp.rows = (int32_t*)gb.rows.p;
    gp.info = (uint8_t*)gb.info.p;
    gp.row_stride = stride;
    gp.runs = (uint32_t*)gb.runs.p;
    gp.runs_end = (const uint64_t*)gb.meta.p + (size_t)nl;
    gp.score = db.score;
    gp.status = db.status;
    gp.cigar_len = db.cigar_len;
    // one block per pair, one launch per wave (waves share the traceback words: stream order)
    for (size_t wv = 0; wv <= lp.waves.size(); ++wv) {
      const uint32_t lo = wv ? lp.waves[wv - 1] : 0u;
      const uint32_t hi = wv < lp.waves.size() ? lp.waves[wv] : nl;
      if (lo >= hi) continue;
      sa::GeneralParams gw = gp;
      gw.ids = gp.ids + lo;
      gw.n_ids = hi - lo;
      gw.tb_off = gp.tb_off + lo;
      gw.runs_end = gp.runs_end + lo;
      gw.rows = gp.rows + (uint64_t)lo * 6 * stride;
      gw.info = gp.info + (uint64_t)lo * 4 * stride;
      gw.ck_off = gp.ck_off + lo;
      gw.blk_off = gp.blk_off + lo;
      gw.ws = gp.ws + lo;
      if (linear && wide)
        sa::nw_linear_general_kernel<sa::kGeneralThreadsWide><<<gw.n_ids, sa::kGeneralThreadsWide, 0, sx>>>(gw);
      else if (linear)
        sa::nw_linear_general_kernel<sa::kGeneralThreads><<<gw.n_ids, sa::kGeneralThreads, 0, sx>>>(gw);
      else if (wide)
        sa::nw_affine_general_kernel<sa::kGeneralThreadsWide><<<gw.n_ids, sa::kGeneralThreadsWide, 0, sx>>>(gw);
      else
        sa::nw_affine_general_kernel<sa::kGeneralThreads><<<gw.n_ids, sa::kGeneralThreads, 0, sx>>>(gw);
      if (!linear && any_ckpt && want_cigar) {  // second kernel of a checkpointed traceback (idle blocks leave at once)
        if (wide)
          sa::nw_affine_general_back<sa::kGeneralThreadsWide><<<gw.n_ids, sa::kGeneralThreadsWide, 0, sx>>>(gw);
        else
          sa::nw_affine_general_back<sa::kGeneralThreads><<<gw.n_ids, sa::kGeneralThreads, 0, sx>>>(gw);
        e->timing.kernel_launches++;
      }
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches++;
    }
    return SA_OK;
  };
  // runs of the literal kernels' pairs -> pool (after the scan)
  auto literal_runs_to_pool = [&](const LitPlan& lp, sa_engine::LitBufs& gb, cudaStream_t sx) -> sa_status_t {
    const uint32_t nl = (uint32_t)lp.ids.size();
    if (!nl) return SA_OK;
    sa::general_runs_to_pool<<<nl, 128, 0, sx>>>((const uint32_t*)gb.ids.p, nl, (const uint32_t*)gb.runs.p,
                                                  (const uint64_t*)gb.meta.p + nl, db.cigar_len, db.cigar_off, db.pool, db.pool_cap);
    CUDA_TRY(e, cudaGetLastError());
    e->timing.kernel_launches++;
    return SA_OK;
  };
  // Tiled long-pair path (nw_long.cuh): per wave one forward launch per tile anti-diagonal, the
  // classification, the backward (traceback) kernel.  fb_count: the slot's fallback counter.
  auto launch_fast = [&](const FastPlan& fq, sa_engine::Slot& sl, uint32_t* fb_count, cudaStream_t sx) -> sa_status_t {
    sa_status_t r;
    const uint32_t nl = (uint32_t)fq.ids.size();
    // meta: [ids u32 | fb_ids u32 | end_h i32 | flag u8 (padded to 4)] [row_off | col_off | ck_off | runs_end] u64, [next_back u32 x waves]
    const size_t m32 = (size_t)nl * 4;
    const size_t meta_bytes = 4 * m32 + 4 * (size_t)nl * 8 + (fq.waves.size() + 1) * 4 + 64;
    if ((r = ensure(e, sl.f_meta, meta_bytes)) != SA_OK) return r;
    if ((r = ensure(e, sl.f_edges, (size_t)fq.edges_total * 8 + 256)) != SA_OK) return r;
    if ((r = ensure(e, sl.f_tb, (size_t)fq.back_warps * fq.S * 32 * sa::kLongMr * 8)) != SA_OK) return r;
    if ((r = ensure(e, sl.f_runs, (size_t)fq.runs_total * 4 + 256)) != SA_OK) return r;
    uint8_t* mb = (uint8_t*)sl.f_meta.p;
    uint32_t* d_ids = (uint32_t*)mb;
    uint32_t* d_fb = (uint32_t*)(mb + m32);
    int32_t* d_endh = (int32_t*)(mb + 2 * m32);
    uint8_t* d_flag = mb + 3 * m32;
    uint64_t* d_off = (uint64_t*)(mb + 4 * m32);
    uint32_t* d_next = (uint32_t*)(mb + 4 * m32 + 4 * (size_t)nl * 8);
    CUDA_TRY(e, cudaMemcpyAsync(d_ids, fq.ids.data(), m32, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(d_off, fq.row_off.data(), (size_t)nl * 8, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(d_off + nl, fq.col_off.data(), (size_t)nl * 8, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(d_off + 2 * (size_t)nl, fq.ck_off.data(), (size_t)nl * 8, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(d_off + 3 * (size_t)nl, fq.runs_end.data(), (size_t)nl * 8, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemsetAsync(d_next, 0, (fq.waves.size() + 1) * 4, sx));
    const size_t smem_f = (size_t)sa::kLongWarps * sa::long_smem_per_warp(fq.R, fq.S > 1);
    const size_t smem_b = (size_t)sa::kLongWarps * sa::long_smem_per_warp(sa::kLongMr, true);
    const bool minb5 = e->long_minb == 5;
    const void* fwd_fn = minb5 ? (const void*)sa::nw_long_fwd<5> : e->long_cell == 2 ? (const void*)sa::nw_long_fwd<4, 2>
                         : e->long_cell == 1 ? (const void*)sa::nw_long_fwd<4, 1> : (const void*)sa::nw_long_fwd<4>;
    for (auto kv : {std::make_pair(fwd_fn, smem_f), std::make_pair((const void*)sa::nw_long_back, smem_b)}) {
      size_t& configured = e->smem_configured[kv.first];
      if (kv.second > configured) {
        CUDA_TRY(e, cudaFuncSetAttribute(kv.first, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)std::min(e->smem_optin, std::max<size_t>(kv.second, 48 * 1024))));
        configured = std::max<size_t>(kv.second, 48 * 1024);
      }
    }
    sa::LongParams lp{};
    lp.residues = db.residues;
    lp.q_off = db.q_off;
    lp.q_len = db.q_len;
    lp.d_off = db.d_off;
    lp.d_len = db.d_len;
    lp.packing = db.packing;
    lp.sc = sa::make_long_scheme(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext);
    lp.ext = sc.gap_ext;
    lp.R = fq.R;
    lp.S = fq.S;
    lp.edges = (int2*)sl.f_edges.p;
    lp.fb_count = fb_count;
    lp.fb_ids = d_fb;
    lp.tb = (uint2*)sl.f_tb.p;
    lp.runs = want_cigar ? (uint32_t*)sl.f_runs.p : nullptr;
    lp.want_runs = want_cigar ? 1 : 0;
    lp.score = db.score;
    lp.status = db.status;
    lp.cigar_len = db.cigar_len;
    for (size_t wv = 0; wv < fq.waves.size(); ++wv) {
      const FastPlan::Wave& w = fq.waves[wv];
      const uint32_t cnt = w.hi - w.lo;
      if (!cnt) continue;
      sa::LongParams lw = lp;
      lw.ids = d_ids + w.lo;
      lw.n_ids = cnt;
      lw.row_off = d_off + w.lo;
      lw.col_off = d_off + nl + w.lo;
      lw.ck_off = d_off + 2 * (size_t)nl + w.lo;
      lw.runs_end = d_off + 3 * (size_t)nl + w.lo;
      lw.end_h = d_endh + w.lo;
      lw.flag = d_flag + w.lo;
      lw.next_back = d_next + wv;
      const uint32_t ndiag = w.tr_max + w.tc_max - 1;
      if (wv == 0) CUDA_TRY(e, cudaEventRecord(sl.ev_l0, sx));
      for (uint32_t d = 0; d < ndiag; ++d) {
        const uint32_t i_lo = d >= w.tc_max ? d - (w.tc_max - 1) : 0u, i_hi = std::min(d, w.tr_max - 1);
        const uint32_t tiles = i_hi - i_lo + 1;
        lw.diag = d;
        // pairs in grid.y (at most 65535 per launch)
        for (uint32_t y0 = 0; y0 < cnt; y0 += 65535) {
          sa::LongParams ly = lw;
          const uint32_t ny = std::min<uint32_t>(65535, cnt - y0);
          ly.ids = lw.ids + y0;
          ly.row_off = lw.row_off + y0;
          ly.col_off = lw.col_off + y0;
          ly.ck_off = lw.ck_off + y0;
          ly.end_h = lw.end_h + y0;
          const dim3 grid((tiles + sa::kLongWarps - 1) / sa::kLongWarps, ny);
          if (minb5) sa::nw_long_fwd<5><<<grid, 32 * sa::kLongWarps, smem_f, sx>>>(ly);
          else if (e->long_cell == 2) sa::nw_long_fwd<4, 2><<<grid, 32 * sa::kLongWarps, smem_f, sx>>>(ly);
          else if (e->long_cell == 1) sa::nw_long_fwd<4, 1><<<grid, 32 * sa::kLongWarps, smem_f, sx>>>(ly);
          else sa::nw_long_fwd<4><<<grid, 32 * sa::kLongWarps, smem_f, sx>>>(ly);
          e->timing.kernel_launches++;
        }
      }
      CUDA_TRY(e, cudaGetLastError());
      if (wv + 1 == fq.waves.size()) CUDA_TRY(e, cudaEventRecord(sl.ev_l1, sx));  // (several waves: the last one's split)
      sa::nw_long_classify<<<(cnt + 127) / 128, 128, 0, sx>>>(lw);
      sa::nw_long_back<<<(std::min(fq.back_warps, (cnt + 3) / 4 * 4) + sa::kLongWarps - 1) / sa::kLongWarps, 32 * sa::kLongWarps, smem_b, sx>>>(lw);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches += 2;
    }
    CUDA_TRY(e, cudaEventRecord(sl.ev_l2, sx));
    return SA_OK;
  };
  auto fast_runs_to_pool = [&](const FastPlan& fq, sa_engine::Slot& sl, cudaStream_t sx) -> sa_status_t {
    const uint32_t nl = (uint32_t)fq.ids.size();
    if (!nl) return SA_OK;
    const uint64_t* d_off = (const uint64_t*)((uint8_t*)sl.f_meta.p + 4 * (size_t)nl * 4);
    sa::general_runs_to_pool<<<nl, 128, 0, sx>>>((const uint32_t*)sl.f_meta.p, nl, (const uint32_t*)sl.f_runs.p, d_off + 3 * (size_t)nl,
                                                  db.cigar_len, db.cigar_off, db.pool, db.pool_cap);
    CUDA_TRY(e, cudaGetLastError());
    e->timing.kernel_launches++;
    return SA_OK;
  };
  // Stage A of a segment: fill with the panic bonus on, classify + count walk, queue length.
  auto stage_a = [&](const Segment& sg, int k) -> sa_status_t {
    sa_engine::Slot& sl = e->slot[k];
    const uint32_t cn = sg.n;
    sa_status_t r;
    if ((r = ensure(e, sl.tb, (size_t)sg.tb_total * 8 + 256)) != SA_OK) return r;
    if ((r = ensure(e, sl.end, (size_t)cn * 4 + 256)) != SA_OK) return r;
    if ((r = ensure(e, sl.rerun_ids, (size_t)cn * 4 + 256)) != SA_OK) return r;
    if (want_cigar && (r = ensure(e, sl.tmp_runs, (size_t)cn * sa::kTmpRuns * 4)) != SA_OK) return r;
    cudaStream_t sx = sl.stream;
    CUDA_TRY(e, cudaStreamWaitEvent(sx, sl.ev_bdone, 0));  // the slot's previous user is done
    if (in) CUDA_TRY(e, cudaStreamWaitEvent(sx, e->ev_in, 0));
    const uint32_t* d_order = nullptr;
    if (!sg.order.empty()) {
      if ((r = ensure(e, sl.order, sg.order.size() * 4)) != SA_OK) return r;
      // pageable source: the copy is staged by the runtime before the call returns
      CUDA_TRY(e, cudaMemcpyAsync(sl.order.p, sg.order.data(), sg.order.size() * 4, cudaMemcpyHostToDevice, sx));
      d_order = (const uint32_t*)sl.order.p;
    } else if (sg.dev_sort) {
      const uint32_t nbins = sg.n2max + 2;
      if ((r = ensure(e, sl.order, (size_t)cn * 4)) != SA_OK) return r;
      sa::order_window<<<(cn + sa::kOrderWindow - 1) / sa::kOrderWindow, sa::kOrderWindow, (size_t)nbins * 4, sx>>>(
          fp.d_len, (uint32_t)sg.base, cn, nbins, (uint32_t*)sl.order.p);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches += 1;
      d_order = (const uint32_t*)sl.order.p;
    }
    CUDA_TRY(e, cudaMemsetAsync(d_counts + k, 0, 4, sx));
    CUDA_TRY(e, cudaMemsetAsync(d_counts + 2 + k, 0, 4, sx));  // fallback queue of the tiled long-pair path
    CUDA_TRY(e, cudaEventRecord(sl.ev_f0, sx));
    CUDA_TRY(e, cudaStreamWaitEvent(sl.fill_stream, sl.ev_f0, 0));
    for (const Segment::Sub& sub : sg.subs) {
      set_geometry(sub.g);
      fp.pair_ids = d_order ? d_order + sub.off : nullptr;
      fp.pair_base = (uint32_t)sg.base + sub.off;
      fp.n_launch_pairs = sub.cnt;
      fp.tb = (uint2*)sl.tb.p + sub.tb_off;
      fp.end = (uint32_t*)sl.end.p + sub.off;
      fp.row0 = pack2(row0_clean + (linear ? 0u : 1u));  // affine: panic bonus on
      fp.rerun_ids = linear ? nullptr : (uint32_t*)sl.rerun_ids.p;  // the fill queues the pairs whose end cell carries the bonus
      fp.rerun_count = linear ? nullptr : d_counts + k;
      if ((r = launch_fill_g(e, fp, sub.g, (sub.cnt + sub.g.ppt - 1) / sub.g.ppt, sl.fill_stream, s2.algo)) != SA_OK) return r;
    }
    fp.rerun_ids = nullptr;
    fp.rerun_count = nullptr;
    CUDA_TRY(e, cudaEventRecord(sl.ev_f1, sl.fill_stream));
    CUDA_TRY(e, cudaStreamWaitEvent(sx, sl.ev_f1, 0));
    // the refill queue is complete when the fill is: its length goes to the host BEFORE the walks, so that stage B's
    // refill (main stream) runs beside the classify + count walk of the untainted pairs instead of after it
    CUDA_TRY(e, cudaMemcpyAsync(e->h_count + 4 + k, d_counts + k, 4, cudaMemcpyDeviceToHost, sx));
    CUDA_TRY(e, cudaEventRecord(sl.ev_count, sx));
    for (const Segment::Sub& sub : sg.subs) {
      set_geometry(sub.g);
      wp.pair_ids = d_order ? d_order + sub.off : nullptr;
      wp.pair_base = (uint32_t)sg.base + sub.off;
      wp.n_launch_pairs = sub.cnt;
      wp.n_launch_dev = nullptr;
      wp.tb = (const uint2*)sl.tb.p + sub.tb_off;
      wp.end = (const uint32_t*)sl.end.p + sub.off;
      wp.phase = 0;
      wp.tmp_runs = want_cigar ? (uint32_t*)sl.tmp_runs.p : nullptr;
      wp.tmp_base = (uint32_t)sg.base;
      if (linear)
        sa::nw_linear_walk<0><<<(sub.cnt + 127) / 128, 128, 0, sx>>>(wp);
      else
        sa::nw_affine_walk<0><<<(sub.cnt + 127) / 128, 128, 0, sx>>>(wp);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches++;
    }
    if (!sg.lit.ids.empty() && (r = launch_literal(sg.lit, sl.lit, sx)) != SA_OK) return r;
    if (!sg.fast.ids.empty() && (r = launch_fast(sg.fast, sl, d_counts + 2 + k, sx)) != SA_OK) return r;
    if (!sg.fast.ids.empty()) CUDA_TRY(e, cudaMemcpyAsync(e->h_count + 6 + k, d_counts + 2 + k, 4, cudaMemcpyDeviceToHost, sx));
    CUDA_TRY(e, cudaEventRecord(sl.ev_w0, sx));  // walks (and long-pair kernels) of the segment done
    return SA_OK;
  };
  // Stage B: clean refill of the queued pairs, scan of the lengths, write walks, results out.
  auto stage_b = [&](const Segment& sg, int k) -> sa_status_t {
    sa_engine::Slot& sl = e->slot[k];
    const Geometry& g = sg.g;
    const uint32_t cn = sg.n;
    const size_t tile_bytes = (size_t)g.tile_stride * 8;
    const uint32_t ctiles = (cn + g.ppt - 1) / g.ppt;
    sa_status_t r;
    // (the fallback count of tiled long pairs is only known after their kernels, at the end of stage A)
    CUDA_TRY(e, cudaEventSynchronize(sg.fast.ids.empty() ? sl.ev_count : sl.ev_w0));
    if (getenv("SA_TRACE")) fprintf(stderr, "[sa trace]   count of segment at %llu arrived\n", (unsigned long long)sg.base);
    CUDA_TRY(e, cudaStreamWaitEvent(e->stream, sg.fast.ids.empty() ? sl.ev_count : sl.ev_w0, 0));
    const uint32_t n_re = e->h_count[4 + k];
    e->timing.pairs_rerun += n_re;
    // tiled long pairs whose traceback can meet a dead end (provenance class 2): literal kernel, now
    const uint32_t n_fb = sg.fast.ids.empty() ? 0u : e->h_count[6 + k];
    LitPlan fbp;
    if (n_fb) {
      std::vector<uint32_t> ids(n_fb);
      CUDA_TRY(e, cudaMemcpyAsync(ids.data(), (const uint8_t*)sl.f_meta.p + sg.fast.ids.size() * 4, (size_t)n_fb * 4,
                                  cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      std::sort(ids.begin(), ids.end());
      if (getenv("SA_TRACE")) fprintf(stderr, "[sa trace]   %u long pair(s) handed to the literal kernel\n", n_fb);
      e->timing.pairs_fallback += n_fb;
      fbp = plan_literal(ids, budget_re);
      if ((r = launch_literal(fbp, e->fb_lit, e->stream)) != SA_OK) return r;
    }
    float fms = 0;
    if (cudaEventElapsedTime(&fms, sl.ev_f0, sl.ev_f1) == cudaSuccess) e->timing.fill_ms += fms;
    if (getenv("SA_TRACE")) {  // device-side timeline: when the segment's fill started and ended, from the call's first event
      float t0 = 0, t1 = 0;
      if (cudaEventElapsedTime(&t0, e->ev_t0, sl.ev_f0) == cudaSuccess && cudaEventElapsedTime(&t1, e->ev_t0, sl.ev_f1) == cudaSuccess)
        fprintf(stderr, "[sa trace]   device: fill of segment at %llu (%u pairs) ran %.0f .. %.0f us\n", (unsigned long long)sg.base, sg.n,
                t0 * 1e3, t1 * 1e3);
    }
    if (!sg.fast.ids.empty()) {
      if (cudaEventElapsedTime(&fms, sl.ev_l0, sl.ev_l1) == cudaSuccess) e->timing.long_fwd_ms += fms;
      if (cudaEventElapsedTime(&fms, sl.ev_l1, sl.ev_l2) == cudaSuccess) e->timing.long_back_ms += fms;
    }
    cudaGetLastError();
    const uint64_t tiles_re = std::max<uint64_t>(1, std::min<uint64_t>(ctiles, std::max<uint64_t>(1, budget_re / tile_bytes)));
    const uint32_t sb = (cn + sa::kScanBlock - 1) / sa::kScanBlock;
    if (n_re) {
      if ((r = ensure(e, e->tb2, (size_t)tiles_re * tile_bytes)) != SA_OK) return r;
      if ((r = ensure(e, e->end2, (size_t)tiles_re * g.ppt * 4)) != SA_OK) return r;
    }
    if ((r = ensure(e, e->block_sums, (size_t)sb * 8)) != SA_OK) return r;
    set_geometry(g);
    const uint32_t re_chunk = (uint32_t)(tiles_re * g.ppt);
    struct ReLaunch {
      uint32_t off, cnt;
    };
    std::vector<ReLaunch> re_launches;
    for (uint32_t off = 0; off < n_re; off += re_chunk)
      re_launches.push_back({off, std::min(re_chunk, n_re - off)});
    // The refill region may be smaller than the queue: refill and count in slices; the write
    // pass (after the scan, which needs every length of the segment) repeats the fill of a
    // slice unless there was only one.
    auto refill = [&](const ReLaunch& rl, bool do_fill) -> sa_status_t {
      fp.pair_ids = (const uint32_t*)sl.rerun_ids.p + rl.off;
      fp.pair_base = 0;
      fp.n_launch_pairs = rl.cnt;
      fp.tb = (uint2*)e->tb2.p;
      fp.end = (uint32_t*)e->end2.p;
      fp.row0 = pack2(row0_clean);
      if (do_fill) {
        sa_status_t r2 = launch_fill_g(e, fp, g, (rl.cnt + g.ppt - 1) / g.ppt, e->stream);
        if (r2 != SA_OK) return r2;
      }
      wp.pair_ids = fp.pair_ids;
      wp.pair_base = 0;
      wp.n_launch_pairs = rl.cnt;
      wp.n_launch_dev = nullptr;
      wp.tb = (const uint2*)e->tb2.p;
      wp.end = (const uint32_t*)e->end2.p;
      wp.phase = 1;
      wp.tmp_runs = want_cigar ? (uint32_t*)sl.tmp_runs.p : nullptr;
      wp.tmp_base = (uint32_t)sg.base;
      return SA_OK;
    };
    for (const ReLaunch& rl : re_launches) {
      if ((r = refill(rl, true)) != SA_OK) return r;
      sa::nw_affine_walk<0><<<(rl.cnt + 127) / 128, 128, 0, e->stream>>>(wp);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches++;
    }
    CUDA_TRY(e, cudaStreamWaitEvent(e->stream, sl.ev_w0, 0));  // every length of the segment is known
    // offsets of this segment (continuing from the previous segments' total)
    sa::scan_block_sums<<<sb, sa::kScanBlock, 0, e->stream>>>(db.cigar_len + sg.base, (uint64_t*)e->block_sums.p, cn);
    sa::scan_block_offsets<<<1, sa::kScanBlock, 0, e->stream>>>((uint64_t*)e->block_sums.p, sb, db.carry);
    sa::scan_apply<<<sb, sa::kScanBlock, 0, e->stream>>>(db.cigar_len + sg.base, (const uint64_t*)e->block_sums.p, db.cigar_off + sg.base, cn);
    CUDA_TRY(e, cudaGetLastError());
    e->timing.kernel_launches += 3;
    if (want_cigar) {
      // pool fill: a gather from the runs parked by the count passes, then a second walk for
      // the (rare) pairs with more than kTmpRuns runs: main region, then each refill slice.
      // Writes past pool_cap are dropped and detected by the caller from the final total.
      sa::cigar_gather<<<(cn * 8 + 255) / 256, 256, 0, e->stream>>>(
          (const uint32_t*)sl.tmp_runs.p, db.cigar_len, db.cigar_off, db.pool, db.pool_cap, (uint32_t)sg.base, cn);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches++;
      if ((r = literal_runs_to_pool(sg.lit, sl.lit, e->stream)) != SA_OK) return r;
      if ((r = fast_runs_to_pool(sg.fast, sl, e->stream)) != SA_OK) return r;
      if (n_fb && (r = literal_runs_to_pool(fbp, e->fb_lit, e->stream)) != SA_OK) return r;
      wp.tmp_runs = nullptr;
      for (const Segment::Sub& sub : sg.subs) {
        set_geometry(sub.g);
        wp.pair_ids = (sg.order.empty() && !sg.dev_sort) ? nullptr : (const uint32_t*)sl.order.p + sub.off;
        wp.pair_base = (uint32_t)sg.base + sub.off;
        wp.n_launch_pairs = sub.cnt;
        wp.tb = (const uint2*)sl.tb.p + sub.tb_off;
        wp.end = (const uint32_t*)sl.end.p + sub.off;
        wp.phase = 0;
        if (linear)
          sa::nw_linear_walk<1><<<(sub.cnt + 127) / 128, 128, 0, e->stream>>>(wp);
        else
          sa::nw_affine_walk<1><<<(sub.cnt + 127) / 128, 128, 0, e->stream>>>(wp);
        CUDA_TRY(e, cudaGetLastError());
        e->timing.kernel_launches++;
      }
      set_geometry(g);
      for (const ReLaunch& rl : re_launches) {
        if ((r = refill(rl, re_launches.size() > 1)) != SA_OK) return r;
        sa::nw_affine_walk<1><<<(rl.cnt + 127) / 128, 128, 0, e->stream>>>(wp);
        CUDA_TRY(e, cudaGetLastError());
        e->timing.kernel_launches++;
      }
    }
    CUDA_TRY(e, cudaEventRecord(sl.ev_bdone, e->stream));  // the slot's scratch is free again
    if (out && want_cigar) {
      const int cs = seg_index & 1;
      CUDA_TRY(e, cudaMemcpyAsync(e->h_count + 8 + 2 * cs, db.carry, 8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(e, cudaEventRecord(e->ev_carry[cs], e->stream));
      if (seg_index > 0) CUDA_TRY(e, send_pool_upto(cs ^ 1));  // the previous segment's words
    }
    ++seg_index;
    if (out) {  // results of this segment to the host
      CUDA_TRY(e, cudaEventRecord(e->ev_done, e->stream));
      CUDA_TRY(e, cudaStreamWaitEvent(e->s_out, e->ev_done, 0));
      cudaError_t err = cudaSuccess;
      auto cp = [&](void* dst, const void* src, size_t bytes) {
        if (err == cudaSuccess && dst && bytes) {
          err = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->s_out);
          e->timing.d2h_bytes += bytes;
        }
      };
      cp(out->score ? out->score + sg.base : nullptr, db.score + sg.base, (size_t)cn * 4);
      cp(out->status ? out->status + sg.base : nullptr, db.status + sg.base, (size_t)cn);
      cp(out->cigar_len ? out->cigar_len + sg.base : nullptr, db.cigar_len + sg.base, (size_t)cn * 4);
      cp(out->cigar_off ? out->cigar_off + sg.base : nullptr, db.cigar_off + sg.base, (size_t)cn * 8);
      if (err != cudaSuccess) return fail(e, SA_E_CUDA, "D2H copy failed: %s", cudaGetErrorString(err));
    }
    return SA_OK;
  };

  CUDA_TRY(e, cudaMemsetAsync(db.carry, 0, 16, e->stream));
  if (db.pool_base) {  // offsets continue from the slice's start
    memcpy(e->h_count + 12, &db.pool_base, 8);
    CUDA_TRY(e, cudaMemcpyAsync(db.carry, e->h_count + 12, 8, cudaMemcpyHostToDevice, e->stream));
  }
  CUDA_TRY(e, cudaEventRecord(e->ev_t0, e->stream));
  for (int k = 0; k < 2; ++k) {  // order the slot streams after everything queued so far
    CUDA_TRY(e, cudaEventRecord(e->slot[k].ev_bdone, e->stream));
  }

  // Software pipeline over segments: A(0); then for each i: A(i+1) is queued BEFORE the host
  // waits for segment i's refill count, so the GPU never idles across the host round trip.
  // SA_TRACE=1: host-side timeline of the pipeline on stderr (developer aid)
  const bool trace = getenv("SA_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto mark = [&](const char* what, uint64_t a) {
    if (trace)
      fprintf(stderr, "[sa trace] %8.1f us  %s %llu\n",
              std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_start).count(), what,
              (unsigned long long)a);
  };
  // The host scan (and copy-in) of segment i+2 is done before the host blocks on segment i's
  // count, so stage A of i+2 can be queued the moment stage B of i is.
  Segment seg[3];
  bool have[3] = {false, false, false};
  if ((st = prepare(0, seg[0])) != SA_OK) return st;
  have[0] = true;
  mark("prepared", seg[0].n);
  if ((st = upload(seg[0])) != SA_OK) return st;
  if ((st = stage_a(seg[0], 0)) != SA_OK) return st;
  mark("queued A", seg[0].n);
  auto prepare_ahead = [&](int cur, int nxt) -> sa_status_t {  // segment after seg[cur] into seg[nxt]
    have[nxt] = false;
    const uint64_t nb = seg[cur].base + seg[cur].n;
    if (nb >= n) return SA_OK;
    sa_status_t r = prepare(nb, seg[nxt]);
    if (r != SA_OK) return r;
    have[nxt] = true;
    mark("prepared", seg[nxt].n);
    return upload(seg[nxt]);
  };
  if ((st = prepare_ahead(0, 1)) != SA_OK) return st;
  for (int i = 0;; ++i) {
    const int cur = i % 3, nxt = (i + 1) % 3, nn = (i + 2) % 3;
    if (have[nxt]) {
      if ((st = stage_a(seg[nxt], (i + 1) & 1)) != SA_OK) return st;
      mark("queued A", seg[nxt].n);
      if ((st = prepare_ahead(nxt, nn)) != SA_OK) return st;
    }
    if ((st = stage_b(seg[cur], i & 1)) != SA_OK) return st;
    mark("queued B", seg[cur].n);
    if (!have[nxt]) break;
  }
  CUDA_TRY(e, cudaEventRecord(e->ev_t1, e->stream));
  if (out || want_cigar) {
    // total CIGAR words (also the overflow check)
    CUDA_TRY(e, cudaMemcpyAsync(e->h_count + 2, db.carry, 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    mark("main stream drained", 0);
    memcpy(used_out, e->h_count + 2, 8);
    if (out && want_cigar) {
      CUDA_TRY(e, send_pool_upto((seg_index - 1) & 1));  // the last segment's words
      *pool_sent_out = pool_sent;
    }
  }
  return SA_OK;
}

sa_status_t resolve_scheme(sa_engine* e, const sa_scheme_t* scheme, Scheme2& s2) {
  s2.sc = sa_scheme_t{5, -4, -8, -6};  // nw_affine.rs:15-20
  if (scheme) s2.sc = *scheme;
  const sa_scheme_t& sc = s2.sc;
  if (!(sc.match > sc.mismatch) || sc.gap_open > 0 || sc.gap_ext > 0 || sc.match < 0)
    return fail(e, SA_E_UNSUPPORTED, "scheme (%d,%d,%d,%d) outside the packed kernel's domain",
                sc.match, sc.mismatch, sc.gap_open, sc.gap_ext);
  s2.pen = 2 * (sc.match - sc.mismatch);
  s2.openp = -2 * sc.gap_open;
  s2.extp = sc.match - 2 * sc.gap_ext;
  if (s2.pen > 128 || s2.extp <= 0)
    return fail(e, SA_E_UNSUPPORTED, "scheme magnitudes exceed the packed kernel's range");
  return SA_OK;
}

// nw_affine:433-434, wfa.rs:26: every pair returns Err("not implemented") in non-global modes.
sa_status_t fill_not_implemented(sa_engine* e, DeviceBatch& db, uint64_t n) {
  CUDA_TRY(e, cudaMemsetAsync(db.status, SA_NOT_IMPLEMENTED, n, e->stream));
  CUDA_TRY(e, cudaMemsetAsync(db.score, 0, n * 4, e->stream));
  CUDA_TRY(e, cudaMemsetAsync(db.cigar_len, 0, n * 4, e->stream));
  CUDA_TRY(e, cudaMemsetAsync(db.cigar_off, 0, n * 8, e->stream));
  CUDA_TRY(e, cudaMemsetAsync(db.carry, 0, 16, e->stream));
  return SA_OK;
}

sa_status_t check_algo(sa_engine* e, sa_algo_t algo, sa_mode_t mode, bool* not_impl) {
  *not_impl = false;
  if (mode != SA_MODE_GLOBAL && mode != SA_MODE_LOCAL && mode != SA_MODE_SEMIGLOBAL)
    return fail(e, SA_E_ARG, "mode %d", (int)mode);
  if (mode != SA_MODE_GLOBAL) {
    if (algo == SA_ALGO_NW_AFFINE || algo == SA_ALGO_WFA || algo == SA_ALGO_WFA_STANDARD) {
      *not_impl = true;
      return SA_OK;
    }
    // needleman_wunsch.rs implements exactly one non-global mode: `local` (:88-89, :107-111)
    if (algo == SA_ALGO_NW_LINEAR && mode == SA_MODE_LOCAL) return SA_OK;
    return fail(e, SA_E_UNSUPPORTED, "mode %d for algo %d does not exist in the reference (needleman_wunsch.rs:180 takes `local: bool`)", (int)mode, (int)algo);
  }
  if (algo != SA_ALGO_NW_AFFINE && algo != SA_ALGO_NW_LINEAR && algo != SA_ALGO_WFA &&
      algo != SA_ALGO_WFA_STANDARD)
    return fail(e, SA_E_ARG, "algo %d", (int)algo);
  return SA_OK;
}

}  // namespace

namespace sa_host {

sa_status_t sd_create(int device_id, sa_engine** out) {
  if (!out) return SA_E_ARG;
  *out = nullptr;
  sa_engine* e = new (std::nothrow) sa_engine();
  if (!e) return SA_E_NOMEM;
  *out = e;  // returned even on failure so the caller can read sa_last_error
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0)
    return fail(e, SA_E_CUDA, "no CUDA device: %s (this engine has no CPU fallback)",
                cudaGetErrorString(err));
  if (device_id < 0 || device_id >= n) return fail(e, SA_E_ARG, "device %d of %d", device_id, n);
  e->device = device_id;
  CUDA_TRY(e, cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CUDA_TRY(e, cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10)
    return fail(e, SA_E_CUDA, "device %s is sm_%d%d; this build is sm_100a only", prop.name,
                prop.major, prop.minor);
  e->sm_count = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  int prio_least = 0, prio_greatest = 0;
  CUDA_TRY(e, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUDA_TRY(e, cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_greatest));
  CUDA_TRY(e, cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
  CUDA_TRY(e, cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    CUDA_TRY(e, cudaStreamCreateWithPriority(&e->slot[k].stream, cudaStreamNonBlocking, prio_greatest));
    CUDA_TRY(e, cudaStreamCreateWithPriority(&e->slot[k].fill_stream, cudaStreamNonBlocking, prio_least));
  }
  for (cudaEvent_t* ev : {&e->ev_in, &e->ev_done, &e->ev_carry[0], &e->ev_carry[1], &e->slot[0].ev_count, &e->slot[1].ev_count,
                          &e->slot[0].ev_bdone, &e->slot[1].ev_bdone, &e->slot[0].ev_w0, &e->slot[1].ev_w0})
    CUDA_TRY(e, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  for (cudaEvent_t* ev : {&e->ev_t0, &e->ev_t1, &e->slot[0].ev_f0, &e->slot[0].ev_f1, &e->slot[1].ev_f0, &e->slot[1].ev_f1,
                          &e->slot[0].ev_l0, &e->slot[0].ev_l1, &e->slot[0].ev_l2, &e->slot[1].ev_l0, &e->slot[1].ev_l1, &e->slot[1].ev_l2})
    CUDA_TRY(e, cudaEventCreate(ev));
  CUDA_TRY(e, cudaMallocHost((void**)&e->h_count, 64));
  if (const char* s = getenv("SA_FORCE_G")) e->force_g = atoi(s);
  if (const char* s = getenv("SA_FORCE_K")) e->force_k = atoi(s);
  if (const char* s = getenv("SA_LONG_CKPT")) e->long_ckpt_always = atoi(s) != 0;
  if (const char* s = getenv("SA_LONG_LITERAL")) e->long_literal = atoi(s) != 0;
  if (const char* s = getenv("SA_LONG_S")) e->long_s = (uint32_t)std::max(0, atoi(s));
  if (const char* s = getenv("SA_LONG_R")) e->long_r = (uint32_t)std::max(0, atoi(s));
  if (const char* s = getenv("SA_LONG_CELL")) e->long_cell = (uint32_t)std::max(0, atoi(s));
  if (const char* s = getenv("SA_LONG_MINB")) e->long_minb = (uint32_t)std::max(0, atoi(s));
  if (const char* s = getenv("SA_SEG_HEAD")) e->seg_head = (uint32_t)std::max(1024, atoi(s));
  if (const char* s = getenv("SA_FILL_MINB")) e->fill_minb = (uint32_t)std::max(0, atoi(s));
  if (const char* s = getenv("SA_SCAN_THREADS")) e->scan_threads = std::max(1, atoi(s));
  if (const char* s = getenv("SA_WALK_PF")) e->walk_pf = (uint32_t)std::max(0, atoi(s));
  if (const char* s = getenv("SA_ORMASK")) e->ormask = (uint32_t)strtoul(s, nullptr, 0);
  if (const char* s = getenv("SA_TB_BUDGET_MB")) e->tb_budget = (size_t)atoll(s) << 20;
  if (const char* s = getenv("SA_SORT")) e->sort_mode = atoi(s);
  if (const char* s = getenv("SA_SEG_PAIRS")) {
    e->seg_pairs = std::max(1, atoi(s));
    e->seg_pairs_forced = true;
  }
  return SA_OK;
}

sa_status_t sd_destroy(sa_engine* e) {
  if (!e) return SA_OK;
  if (e->stream) {
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    std::vector<DevBuf*> bufs = {&e->tb2, &e->end2, &e->misc, &e->wfa_scratch, &e->par_bytes, &e->par_rows, &e->par_in, &e->block_sums,
                                 &e->b_res, &e->b_qoff, &e->b_doff, &e->b_qlen, &e->b_dlen, &e->b_score, &e->b_status, &e->b_clen,
                                 &e->b_coff, &e->b_pool, &e->b_carry};
    auto add_lit = [&](sa_engine::LitBufs& g) {
      for (DevBuf* b : {&g.ids, &g.meta, &g.tb, &g.rows, &g.info, &g.runs}) bufs.push_back(b);
    };
    add_lit(e->fb_lit);
    for (int k = 0; k < 2; ++k) {
      sa_engine::Slot& sl = e->slot[k];
      add_lit(sl.lit);
      for (DevBuf* b : {&sl.tb, &sl.end, &sl.rerun_ids, &sl.tmp_runs, &sl.order, &sl.f_meta, &sl.f_edges, &sl.f_tb, &sl.f_runs})
        bufs.push_back(b);
    }
    for (DevBuf* b : bufs)
      if (b->p) cudaFree(b->p);
    for (cudaEvent_t ev : {e->ev_in, e->ev_done, e->ev_carry[0], e->ev_carry[1], e->ev_t0, e->ev_t1, e->slot[0].ev_count,
                           e->slot[0].ev_f0, e->slot[0].ev_f1, e->slot[1].ev_count, e->slot[1].ev_f0,
                           e->slot[1].ev_f1, e->slot[0].ev_bdone, e->slot[1].ev_bdone, e->slot[0].ev_l0, e->slot[0].ev_l1,
                           e->slot[0].ev_l2, e->slot[1].ev_l0, e->slot[1].ev_l1, e->slot[1].ev_l2, e->slot[0].ev_w0, e->slot[1].ev_w0})
      if (ev) cudaEventDestroy(ev);
    if (e->h_count) cudaFreeHost(e->h_count);
    cudaStreamDestroy(e->stream);
    for (int k = 0; k < 2; ++k) {
      if (e->slot[k].stream) cudaStreamDestroy(e->slot[k].stream);
      if (e->slot[k].fill_stream) cudaStreamDestroy(e->slot[k].fill_stream);
    }
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
  }
  delete e;
  return SA_OK;
}

sa_status_t sd_synchronize(sa_engine* e) {
  if (!e) return SA_E_ARG;
  for (int k = 0; k < 2; ++k) CUDA_TRY(e, cudaStreamSynchronize(e->slot[k].stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->s_in));
  CUDA_TRY(e, cudaStreamSynchronize(e->s_out));
  return SA_OK;
}

sa_status_t sd_batch_free(sa_engine* e, sa_resident_t* r) {
  if (!r) return SA_OK;
  if (e) {
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
  }
  for (void* p : {(void*)r->d.residues, (void*)r->d.q_off, (void*)r->d.d_off, (void*)r->d.q_len,
                  (void*)r->d.d_len, (void*)r->d.score, (void*)r->d.status, (void*)r->d.cigar_len,
                  (void*)r->d.cigar_off, (void*)r->d.pool, (void*)r->d.carry, (void*)r->d.end1, (void*)r->d.end2})
    if (p) cudaFree(p);
  delete r;
  return SA_OK;
}

sa_status_t sd_batch_upload(sa_engine* e, const sa_batch_t* b, sa_resident_t** out) {
  if (!e || !b || !out) return SA_E_ARG;
  *out = nullptr;
  if (b->packing > 1) return fail(e, SA_E_ARG, "packing %u (0 = bytes, 1 = 2-bit)", b->packing);
  if (b->n_pairs >= (1ull << 31)) return fail(e, SA_E_ARG, "n_pairs %llu too large", (unsigned long long)b->n_pairs);
  if (b->n_pairs && (!b->q_off || !b->q_len || !b->d_off || !b->d_len || (!b->residues && b->residues_len)))
    return fail(e, SA_E_ARG, "null input array");
  CUDA_TRY(e, cudaSetDevice(e->device));
  sa_resident* r = new (std::nothrow) sa_resident();
  if (!r) return SA_E_NOMEM;
  const uint64_t n = b->n_pairs;
  r->n_pairs = n;
  r->residues_len = b->residues_len;
  r->d.packing = b->packing;
  r->h_q_len.assign(b->q_len, b->q_len + n);
  r->h_d_len.assign(b->d_len, b->d_len + n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t limit = b->packing ? b->residues_len * 4 : b->residues_len;
    if (!view_in_bounds(b->q_off[i], b->q_len[i], limit) || !view_in_bounds(b->d_off[i], b->d_len[i], limit)) {
      delete r;
      return fail(e, SA_E_ARG, "pair %llu reaches past residues_len", (unsigned long long)i);
    }
    r->cells += (uint64_t)b->q_len[i] * b->d_len[i];
  }
  const size_t n1 = std::max<uint64_t>(n, 1);
  auto alloc = [&](void** p, size_t bytes) { return cudaMalloc(p, std::max<size_t>(bytes, 16)); };
  cudaError_t err = cudaSuccess;
  if (err == cudaSuccess) err = alloc((void**)&r->d.residues, b->residues_len);
  if (err == cudaSuccess) err = alloc((void**)&r->d.q_off, n1 * 8);
  if (err == cudaSuccess) err = alloc((void**)&r->d.d_off, n1 * 8);
  if (err == cudaSuccess) err = alloc((void**)&r->d.q_len, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->d.d_len, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->d.score, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->d.status, n1);
  if (err == cudaSuccess) err = alloc((void**)&r->d.cigar_len, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->d.cigar_off, n1 * 8);
  if (err == cudaSuccess) err = alloc((void**)&r->d.carry, 16);
  if (err != cudaSuccess) {
    cudaGetLastError();
    sd_batch_free(e, r);
    return fail(e, SA_E_NOMEM, "device allocation for the batch failed: %s", cudaGetErrorString(err));
  }
  if (b->residues_len)
    cudaMemcpyAsync(r->d.residues, b->residues, b->residues_len, cudaMemcpyHostToDevice, e->stream);
  if (n) {
    cudaMemcpyAsync(r->d.q_off, b->q_off, n * 8, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(r->d.d_off, b->d_off, n * 8, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(r->d.q_len, b->q_len, n * 4, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(r->d.d_len, b->d_len, n * 4, cudaMemcpyHostToDevice, e->stream);
  }
  cudaError_t last = cudaStreamSynchronize(e->stream);
  if (last != cudaSuccess) {
    sd_batch_free(e, r);
    return fail(e, SA_E_CUDA, "upload failed: %s", cudaGetErrorString(last));
  }
  *out = r;
  return SA_OK;
}

sa_status_t sd_align_resident(sa_engine* e, sa_algo_t algo, sa_mode_t mode,
                              const sa_scheme_t* scheme, sa_resident_t* r, int want_cigar) {
  if (!e || !r) return SA_E_ARG;
  CUDA_TRY(e, cudaSetDevice(e->device));
  const uint64_t n = r->n_pairs;
  r->want_cigar = want_cigar != 0;
  r->used = 0;
  e->timing = sa_timing_t{};
  e->timing.cells = r->cells;
  bool not_impl = false;
  sa_status_t st = check_algo(e, algo, mode, &not_impl);
  if (st != SA_OK) return st;
  if (n == 0) {
    r->aligned = true;
    return SA_OK;
  }
  if (not_impl) {
    st = fill_not_implemented(e, r->d, n);
    r->aligned = st == SA_OK;
    return st;
  }
  if (algo == SA_ALGO_WFA || algo == SA_ALGO_WFA_STANDARD) {
    st = run_wfa(e, r->d, n, r->h_q_len.data(), r->h_d_len.data(), scheme, algo == SA_ALGO_WFA, nullptr, nullptr);
    r->aligned = st == SA_OK;
    r->want_cigar = false;
    return st;
  }
  Scheme2 s2;
  if ((st = resolve_scheme(e, scheme, s2)) != SA_OK) return st;
  s2.algo = (int)algo;
  const bool local = mode == SA_MODE_LOCAL;  // (check_algo: linear NW only)
  r->local = local;
  if (local && !r->d.end1) {
    cudaError_t err = cudaMalloc((void**)&r->d.end1, n * 4);
    if (err == cudaSuccess) err = cudaMalloc((void**)&r->d.end2, n * 4);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(e, SA_E_NOMEM, "end-cell arrays: %s", cudaGetErrorString(err));
    }
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (want_cigar && !r->d.pool) {
      r->d.pool_cap = std::max<uint64_t>(std::max<uint64_t>(1024, n * 24), r->used + 1024);
      cudaError_t err = cudaMalloc((void**)&r->d.pool, r->d.pool_cap * 4);
      if (err != cudaSuccess) {
        cudaGetLastError();
        r->d.pool = nullptr;
        return fail(e, SA_E_NOMEM, "cigar pool allocation failed");
      }
    }
    uint64_t used = 0;
    uint64_t sent = 0;
    // plan key: everything the segment plan depends on besides the (immutable) lengths
    const uint64_t key = 0x9e3779b97f4a7c15ull ^ ((uint64_t)algo << 56) ^ ((uint64_t)(uint32_t)s2.sc.match << 40) ^
                         ((uint64_t)(uint32_t)(s2.sc.mismatch & 0xffff) << 24) ^ ((uint64_t)(uint32_t)(s2.sc.gap_open & 0xfff) << 12) ^
                         (uint64_t)(uint32_t)(s2.sc.gap_ext & 0xfff) ^ ((uint64_t)e->tb_budget << 1) ^ ((uint64_t)e->seg_pairs << 20) ^
                         ((uint64_t)e->force_g << 8) ^ ((uint64_t)e->sort_mode << 4);
    Plan* plan = (Plan*)r->plan;
    bool valid = plan && r->plan_key == key && e->budget_cached == r->plan_budget;
    if (!valid) {
      if (r->plan && r->plan_free) r->plan_free(r->plan);
      plan = new (std::nothrow) Plan();
      r->plan = plan;
      r->plan_free = [](void* p) { delete (Plan*)p; };
      r->plan_key = key;
    }
    if (local)
      st = run_linear_local(e, r->d, n, r->h_q_len.data(), r->h_d_len.data(), s2.sc, want_cigar != 0, nullptr, nullptr, &used);
    else
      st = run_affine(e, r->d, n, r->h_q_len.data(), r->h_d_len.data(), s2, want_cigar != 0, nullptr,
                      nullptr, &used, &sent, plan, valid);
    r->plan_budget = e->budget_cached;
    if (st != SA_OK && plan) plan->segs.clear(), r->plan_key = 0;
    if (st != SA_OK) return st;
    r->used = used;
    if (!want_cigar || used <= r->d.pool_cap) break;
    // the pool was too small: writes past its end were dropped; grow to the exact size, redo
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    cudaFree(r->d.pool);
    r->d.pool = nullptr;
  }
  r->aligned = true;
  return SA_OK;
}

sa_status_t sd_resident_download(sa_engine* e, sa_resident_t* r, sa_result_t* res) {
  if (!e || !r || !res) return SA_E_ARG;
  if (!r->aligned) return fail(e, SA_E_ARG, "sa_align_resident has not run on this batch");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const uint64_t n = r->n_pairs;
  const uint64_t used = r->used;
  if (n) {
    if (res->score) CUDA_TRY(e, cudaMemcpyAsync(res->score, r->d.score, n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (res->status) CUDA_TRY(e, cudaMemcpyAsync(res->status, r->d.status, n, cudaMemcpyDeviceToHost, e->stream));
    if (res->cigar_len) CUDA_TRY(e, cudaMemcpyAsync(res->cigar_len, r->d.cigar_len, n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (res->cigar_off) CUDA_TRY(e, cudaMemcpyAsync(res->cigar_off, r->d.cigar_off, n * 8, cudaMemcpyDeviceToHost, e->stream));
    if (r->local && r->d.end1) {
      if (res->end1) CUDA_TRY(e, cudaMemcpyAsync(res->end1, r->d.end1, n * 4, cudaMemcpyDeviceToHost, e->stream));
      if (res->end2) CUDA_TRY(e, cudaMemcpyAsync(res->end2, r->d.end2, n * 4, cudaMemcpyDeviceToHost, e->stream));
    } else {  // global alignments end at (n1, n2)
      if (res->end1) memcpy(res->end1, r->h_q_len.data(), n * 4);
      if (res->end2) memcpy(res->end2, r->h_d_len.data(), n * 4);
    }
  }
  res->cigar_used = used;
  sa_status_t rc = SA_OK;
  if (res->cigar && r->want_cigar && used) {
    if (used > res->cigar_capacity)
      rc = fail(e, SA_E_CIGAR_CAPACITY, "cigar pool needs %llu words, capacity is %llu",
                (unsigned long long)used, (unsigned long long)res->cigar_capacity);
    else
      CUDA_TRY(e, cudaMemcpyAsync(res->cigar, r->d.pool, used * 4, cudaMemcpyDeviceToHost, e->stream));
  }
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return rc;
}

sa_status_t sd_align_batch(sa_engine* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                           const sa_batch_t* b, sa_result_t* res, uint64_t pool_base) {
  if (!e || !b || !res) return SA_E_ARG;
  if (b->packing > 1) return fail(e, SA_E_ARG, "packing %u (0 = bytes, 1 = 2-bit)", b->packing);
  if (b->n_pairs >= (1ull << 31)) return fail(e, SA_E_ARG, "n_pairs %llu too large", (unsigned long long)b->n_pairs);
  if (b->n_pairs && (!b->q_off || !b->q_len || !b->d_off || !b->d_len || (!b->residues && b->residues_len)))
    return fail(e, SA_E_ARG, "null input array");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const uint64_t n = b->n_pairs;
  const auto t_call = std::chrono::steady_clock::now();
  e->timing = sa_timing_t{};
  res->cigar_used = 0;
  bool not_impl = false;
  sa_status_t st = check_algo(e, algo, mode, &not_impl);
  if (st != SA_OK) return st;
  if (n == 0) return SA_OK;
  if (not_impl) {
    if (res->status) memset(res->status, SA_NOT_IMPLEMENTED, n);
    if (res->score) memset(res->score, 0, n * 4);
    if (res->cigar_len) memset(res->cigar_len, 0, n * 4);
    if (res->cigar_off) memset(res->cigar_off, 0, n * 8);
    if (res->end1) memset(res->end1, 0, n * 4);
    if (res->end2) memset(res->end2, 0, n * 4);
    return SA_OK;
  }
  const bool local = mode == SA_MODE_LOCAL;  // (check_algo: linear NW only)
  if (!local) {  // global alignments end at (n1, n2)
    if (res->end1) memcpy(res->end1, b->q_len, n * 4);
    if (res->end2) memcpy(res->end2, b->d_len, n * 4);
  }
  const bool is_wfa = algo == SA_ALGO_WFA || algo == SA_ALGO_WFA_STANDARD;
  Scheme2 s2;
  if (!is_wfa && (st = resolve_scheme(e, scheme, s2)) != SA_OK) return st;
  s2.algo = (int)algo;
  const bool want_cigar = !is_wfa && res->cigar != nullptr && res->cigar_capacity > 0;

  // engine-owned staging, grow-only: no allocation on the steady-state path
  if ((st = ensure(e, e->b_res, b->residues_len)) != SA_OK) return st;
  if ((st = ensure(e, e->b_qoff, n * 8)) != SA_OK) return st;
  if ((st = ensure(e, e->b_doff, n * 8)) != SA_OK) return st;
  if ((st = ensure(e, e->b_qlen, n * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->b_dlen, n * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->b_score, n * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->b_status, n)) != SA_OK) return st;
  if ((st = ensure(e, e->b_clen, n * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->b_coff, n * 8)) != SA_OK) return st;
  if ((st = ensure(e, e->b_carry, 16)) != SA_OK) return st;
  if (local) {
    if ((st = ensure(e, e->b_end1, n * 4)) != SA_OK) return st;
    if ((st = ensure(e, e->b_end2, n * 4)) != SA_OK) return st;
  }
  if (is_wfa) {
    DeviceBatch db;
    db.residues = (uint8_t*)e->b_res.p;
    db.q_off = (uint64_t*)e->b_qoff.p;
    db.d_off = (uint64_t*)e->b_doff.p;
    db.q_len = (uint32_t*)e->b_qlen.p;
    db.d_len = (uint32_t*)e->b_dlen.p;
    db.score = (int32_t*)e->b_score.p;
    db.status = (uint8_t*)e->b_status.p;
    db.cigar_len = (uint32_t*)e->b_clen.p;
    db.cigar_off = (uint64_t*)e->b_coff.p;
    db.carry = (uint64_t*)e->b_carry.p;
    db.packing = b->packing;
    st = run_wfa(e, db, n, b->q_len, b->d_len, scheme, algo == SA_ALGO_WFA, b, res);
    float ms = 0;
    if (st == SA_OK && cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1) == cudaSuccess) e->timing.kernels_ms = ms;
    e->timing.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
    cudaGetLastError();
    return st;
  }
  if (want_cigar && res->cigar_capacity < pool_base) return fail(e, SA_E_ARG, "pool slice ends before it starts");
  uint64_t want_pool = want_cigar ? std::max<uint64_t>(res->cigar_capacity - pool_base, 1024) : 0;
  uint64_t used = 0, sent = pool_base;  // absolute word positions in the caller's pool
  uint32_t* dev_pool = nullptr;          // device buffer shifted down by pool_base
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (want_cigar && (st = ensure(e, e->b_pool, want_pool * 4)) != SA_OK) return st;
    DeviceBatch db;
    db.residues = (uint8_t*)e->b_res.p;
    db.q_off = (uint64_t*)e->b_qoff.p;
    db.d_off = (uint64_t*)e->b_doff.p;
    db.q_len = (uint32_t*)e->b_qlen.p;
    db.d_len = (uint32_t*)e->b_dlen.p;
    db.score = (int32_t*)e->b_score.p;
    db.status = (uint8_t*)e->b_status.p;
    db.cigar_len = (uint32_t*)e->b_clen.p;
    db.cigar_off = (uint64_t*)e->b_coff.p;
    dev_pool = want_cigar ? (uint32_t*)e->b_pool.p - pool_base : nullptr;
    db.pool = dev_pool;
    db.pool_cap = want_cigar ? pool_base + e->b_pool.cap / 4 : 0;
    db.pool_base = want_cigar ? pool_base : 0;
    db.carry = (uint64_t*)e->b_carry.p;
    db.packing = b->packing;
    e->timing.h2d_bytes = 0;
    e->timing.d2h_bytes = 0;
    e->timing.cells = 0;
    sent = pool_base;
    if (getenv("SA_TRACE"))
      fprintf(stderr, "[sa trace] sd_align_batch: staging ready %.3f ms after entry\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
    if (local) {
      db.end1 = (uint32_t*)e->b_end1.p;
      db.end2 = (uint32_t*)e->b_end2.p;
      st = run_linear_local(e, db, n, b->q_len, b->d_len, s2.sc, want_cigar, b, res, &used);
    } else {
      st = run_affine(e, db, n, b->q_len, b->d_len, s2, want_cigar, b, res, &used, &sent);
    }
    if (st != SA_OK) {
      cudaStreamSynchronize(e->s_in);
      cudaStreamSynchronize(e->s_out);
      cudaStreamSynchronize(e->stream);
      return st;
    }
    if (!want_cigar || used <= db.pool_cap) break;
    // The caller cannot take it either if it exceeds cigar_capacity; otherwise redo with room.
    if (used > res->cigar_capacity) break;
    want_pool = used - pool_base + 1024;
  }
  if (want_cigar && used < pool_base) used = pool_base;
  res->cigar_used = want_cigar ? used : 0;
  sa_status_t rc = SA_OK;
  if (want_cigar && used > pool_base) {
    if (used > res->cigar_capacity) {
      rc = fail(e, SA_E_CIGAR_CAPACITY, "cigar pool needs %llu words, capacity is %llu",
                (unsigned long long)(used - pool_base), (unsigned long long)(res->cigar_capacity - pool_base));
    } else if (sent < used) {  // (normally everything was streamed out already)
      CUDA_TRY(e, cudaMemcpyAsync(res->cigar + sent, dev_pool + sent, (used - sent) * 4, cudaMemcpyDeviceToHost, e->s_out));
      e->timing.d2h_bytes += (used - sent) * 4;
    }
  }
  if (getenv("SA_TRACE"))
    fprintf(stderr, "[sa trace] sd_align_batch: pipeline returned %.3f ms after entry\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
  CUDA_TRY(e, cudaStreamSynchronize(e->s_out));
  if (getenv("SA_TRACE"))
    fprintf(stderr, "[sa trace] sd_align_batch: copy-out stream drained %.3f ms after entry\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1) == cudaSuccess) e->timing.kernels_ms = ms;
  cudaGetLastError();
  e->timing.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
  if (getenv("SA_TRACE"))
    fprintf(stderr, "[sa trace] sd_align_batch on device %d: %llu pairs, wall %.3f ms, kernels %.3f ms, h2d %.1f MB, d2h %.1f MB\n", e->device,
            (unsigned long long)n, e->timing.wall_ms, e->timing.kernels_ms, e->timing.h2d_bytes / 1e6, e->timing.d2h_bytes / 1e6);
  return rc;
}

}  // namespace sa_host
