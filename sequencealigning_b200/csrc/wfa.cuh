// wfa.cuh -- gap-affine wavefront alignment on the GPU (score only), two modes.
//
// STANDARD mode  (wfa_standard_kernel): the textbook gap-affine WFA recurrences the reference's
//   wfa.rs is modelled on (WaveFrontTensor::new :225-420 is the "next" step, WaveFront::expand
//   :127-139 the "extend" step), with the reference's penalties x=4, o=2, e=6 (:17-21) as
//   defaults, WITHOUT the reference's defects.  One warp per pair: lanes own diagonals, the
//   extend step compares 4 residues per iteration on word-packed sequences staged in shared
//   memory, wavefronts of the last max(x, o+e)+1 scores live in a per-warp ring in global
//   memory (L2 resident).  Result = optimal gap-affine cost; validated against a Gotoh cost DP.
//   This is what BASELINE.json configs[3] (1-10 kbp pairs) can actually exercise, because the
//   reference itself panics or never converges on inputs of that size.
//
// LITERAL mode (wfa_literal_kernel): wfa_align (:23-42) exactly as the reference executes it,
//   defects included -- wavefront 0 never extended (:467-483), the x()/y() geometry of :85-90,
//   convergence tested at (n2-1, n1-1) (:189), the trim heuristic (:490-623) and its
//   Vec::rotate_left panics (:577/:603) -- one thread per pair, reported as per-pair status
//   OK (score = wfs.len(), :31-36) / REF_PANIC / REF_NO_CONVERGENCE.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_affine_s16.cuh"

namespace sa {

#ifndef SA_STATUS_CODES
#define SA_STATUS_CODES
enum : uint8_t { kOk = 0, kRefPanic = 1, kRefNoConv = 2, kNotImpl = 3, kRefPanicEarly = 4, kRefNoOutput = 5 };
#endif

constexpr int32_t kWfNone = -(1 << 29);
constexpr int32_t kWfLive = -(1 << 28);  // standard mode: offsets above this are real, everything below is "None"

struct WfaParams {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  uint32_t pair_base, n_launch_pairs;
  int32_t x, o, e;        // penalties
  int32_t* __restrict__ score;
  uint8_t* __restrict__ status;
  int32_t* __restrict__ scratch;   // per warp (standard) / per thread (literal) wavefront storage
  uint64_t scratch_stride;         // int32 per warp / thread
  uint32_t width;                  // standard: diagonals per component array (max n1+n2+1)
  uint32_t* __restrict__ next_pair;  // standard: dynamic work counter
  uint32_t smem_seq_bytes;         // standard: bytes available per warp for staged sequences
  uint32_t lit_wcap;               // literal: element capacity per component
  uint32_t packing;                // input format (see load_residue)
  const uint32_t* __restrict__ order;  // standard: hand-out order (longest pairs first), or nullptr
  int32_t s_step;                  // standard: gcd of the penalties -- only these scores can hold a wavefront
  unsigned long long* __restrict__ work;  // standard: [0] wavefront cells computed, [1] residues passed by extend (or nullptr)
  int32_t* __restrict__ trace;     // literal, single-pair launches only (else nullptr): what the reference's stdout
                                   // is made of -- [0] number of `lo: .., hi: ..` lines (wfa.rs:251), [1] offset,
                                   // [2] state (0 M, 1 D, 2 I), [3] number of parents, [4..6] parents of the converged
                                   // element (:634-651), [8 + 2k], [9 + 2k] = lo, hi of the k-th created wavefront
  uint32_t trace_cap;              // lines the trace can hold
  int32_t ring_dm, ring_de;        // standard: ring depth of the M component (max(x, o+e) / s_step + 1) and of I / D (e / s_step + 1)
};

// ---------------------------------------------------------------------------------------------
// STANDARD MODE
// ---------------------------------------------------------------------------------------------
constexpr int kWfRing = 16;  // most scores the ring may have to keep: max(x, o+e) / s_step + 1 <= kWfRing

// Extend along a diagonal on word-packed sequences: BITS = 8 (any byte alphabet, 4 residues per
// comparison) or BITS = 2 (A/C/G/T codes, 16 residues per comparison).  `sa`, `sb` are 4-byte
// aligned copies with at least one zero word of padding after the end.
template <int BITS>
__device__ __forceinline__ int32_t wf_extend(const uint32_t* sa, const uint32_t* sb, int32_t v,
                                             int32_t h, int32_t n1, int32_t n2) {
  constexpr int PER = 32 / BITS;          // residues per word
  constexpr int SH = BITS == 8 ? 2 : 4;   // log2(PER)
  for (;;) {
    const int32_t room = min(n1 - v, n2 - h);
    if (room <= 0) return v;
    const uint32_t a0 = sa[v >> SH], a1 = sa[(v >> SH) + 1];
    const uint32_t b0 = sb[h >> SH], b1 = sb[(h >> SH) + 1];
    const uint32_t wa = __funnelshift_r(a0, a1, (v & (PER - 1)) * BITS);
    const uint32_t wb = __funnelshift_r(b0, b1, (h & (PER - 1)) * BITS);
    const uint32_t x = wa ^ wb;
    const int32_t same = x ? ((__ffs(x) - 1) / BITS) : PER;  // matching leading residues
    const int32_t adv = min(same, room);
    v += adv;
    h += adv;
    if (adv < PER) return v;
  }
}

// One warp per pair, pairs handed out dynamically.  blockDim.x = 32 * warps.
__global__ void __launch_bounds__(128) wfa_standard_kernel(const WfaParams p) {
  extern __shared__ uint32_t smem_w[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
  uint32_t* sq = smem_w + (size_t)warp * (p.smem_seq_bytes >> 2);
  // The ring keeps only what the recurrences read: M of the last max(x, o+e) / s_step + 1 scores, I and D of
  // the last e / s_step + 1 (13 arrays for the reference's 4 / 2 / 6 instead of 3 x 16): the live set of a warp
  // is then small enough for the resident warps' rings to stay in L2.
  int32_t* ring = p.scratch + (uint64_t)gwarp * p.scratch_stride;  // [ring_dm] M, then [ring_de] I, then [ring_de] D; width each
  const int32_t DM = p.ring_dm, DE = p.ring_de;
  const int32_t W = (int32_t)p.width;
  // per-slot diagonal range, kept in shared memory (uniform per warp)
  __shared__ int32_t s_lo[4][kWfRing], s_hi[4][kWfRing];
  int32_t* lo_ = s_lo[warp];
  int32_t* hi_ = s_hi[warp];

  for (;;) {
    uint32_t li = 0;
    if (lane == 0) li = atomicAdd(p.next_pair, 1u);
    li = __shfl_sync(0xffffffffu, li, 0);
    if (li >= p.n_launch_pairs) break;
    const uint32_t id = p.order ? p.order[li] : p.pair_base + li;
    const int32_t n1 = (int32_t)p.q_len[id], n2 = (int32_t)p.d_len[id];
    const uint64_t o1 = p.q_off[id], o2 = p.d_off[id];
    // Stage both sequences as words: 2-bit codes (16 residues per word) when every residue is
    // A/C/G/T -- always true for packing 1 -- else bytes (4 per word).  They stay in shared
    // memory when they fit, else in a global scratch area behind the ring.
    bool wide = false;
    if (p.packing == 0) {
      for (int32_t pos = lane; pos < n1; pos += 32) {
        const uint32_t c = p.residues[o1 + pos];
        wide |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
      }
      for (int32_t pos = lane; pos < n2; pos += 32) {
        const uint32_t c = p.residues[o2 + pos];
        wide |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
      }
      wide = __any_sync(0xffffffffu, wide);
    }
    const int per = wide ? 4 : 16;
    const uint32_t w1 = (uint32_t)(n1 + 2 * per) / per, w2 = (uint32_t)(n2 + 2 * per) / per;
    uint32_t* s1w;
    uint32_t* s2w;
    if ((w1 + w2) * 4 <= p.smem_seq_bytes) {
      s1w = sq;
      s2w = sq + w1;
    } else {
      s1w = reinterpret_cast<uint32_t*>(ring + (size_t)(DM + 2 * DE) * W);
      s2w = s1w + w1;
    }
    __syncwarp();
    for (int which = 0; which < 2; ++which) {
      uint32_t* dst = which ? s2w : s1w;
      const uint32_t nw = which ? w2 : w1;
      const int32_t len = which ? n2 : n1;
      const uint64_t off = which ? o2 : o1;
      for (uint32_t k = lane; k < nw; k += 32) {
        uint32_t v = 0;
        for (int b = 0; b < per; ++b) {
          const int32_t pos = (int32_t)(k * per + b);
          if (pos < len) {
            uint32_t c = load_residue(p.residues, off + pos, p.packing);
            if (!wide && p.packing == 0) c = (c >> 1) & 3u;  // 'A' 0x41, 'C' 0x43, 'G' 0x47, 'T' 0x54 -> 0,1,3,2
            v |= c << (wide ? 8 * b : 2 * b);
          }
        }
        dst[k] = v;
      }
    }
    __syncwarp();
    __threadfence_block();

    const int32_t kend = n1 - n2;
    const int32_t koff = n2;  // array index = k + n2
    int32_t result = -1;
    uint32_t w_cells = 0, w_ext = 0;  // this lane's share of the pair's work (SURVEY.md 8d units)
    // score 0
    if (lane < kWfRing) {
      lo_[lane] = 1;
      hi_[lane] = 0;  // empty
    }
    __syncwarp();
    {
      const int32_t v0 = wide ? wf_extend<8>(s1w, s2w, 0, 0, n1, n2) : wf_extend<2>(s1w, s2w, 0, 0, n1, n2);
      if (lane == 0) {
        ring[koff] = v0;                                  // M, slot 0
        ring[(size_t)DM * W + koff] = kWfNone;            // I, slot 0
        ring[(size_t)(DM + DE) * W + koff] = kWfNone;     // D, slot 0
        lo_[0] = 0;
        hi_[0] = 0;
      }
      if (kend == 0 && v0 == n1) result = 0;
    }
    __syncwarp();
    const int32_t max_s = 2 * p.o + p.e * (n1 + n2) + p.x + 8;
    const int32_t step = p.s_step, bx = p.x / step, bo = (p.o + p.e) / step, be = p.e / step;  // look-backs in ring steps
    int slot = 0, slot_e = 0;  // (s / step) mod depth, kept by increment-and-wrap (no division in the score loop)
    for (int32_t s = step; result < 0 && s <= max_s; s += step) {
      // slot of score s; the range arrays follow the M ring
      slot = slot + 1 == DM ? 0 : slot + 1;
      slot_e = slot_e + 1 == DE ? 0 : slot_e + 1;
      int sx = slot - bx, so = slot - bo, se = slot - be, se_e = slot_e - be;  // look-backs are < depth
      sx += sx < 0 ? DM : 0;
      so += so < 0 ? DM : 0;
      se += se < 0 ? DM : 0;
      se_e += se_e < 0 ? DE : 0;
      const bool hx = s - p.x >= 0 && lo_[sx] <= hi_[sx];
      const bool ho = s - p.o - p.e >= 0 && lo_[so] <= hi_[so];
      const bool he = s - p.e >= 0 && lo_[se] <= hi_[se];
      int32_t lo = 1 << 30, hi = -(1 << 30);
      if (hx) { lo = min(lo, lo_[sx]); hi = max(hi, hi_[sx]); }
      if (ho) { lo = min(lo, lo_[so] - 1); hi = max(hi, hi_[so] + 1); }
      if (he) { lo = min(lo, lo_[se] - 1); hi = max(hi, hi_[se] + 1); }
      __syncwarp();
      if (lo > hi) {
        if (lane == 0) { lo_[slot] = 1; hi_[slot] = 0; }
        __syncwarp();
        continue;
      }
      lo = max(lo, -n2);
      hi = min(hi, n1);
      int32_t* Mc = ring + (size_t)slot * W + koff;
      int32_t* Ic = ring + (size_t)(DM + slot_e) * W + koff;
      int32_t* Dc = ring + (size_t)(DM + DE + slot_e) * W + koff;
      const int32_t* Mx = ring + (size_t)sx * W + koff;
      const int32_t* Mo = ring + (size_t)so * W + koff;
      const int32_t* Ie = ring + (size_t)(DM + se_e) * W + koff;
      const int32_t* De = ring + (size_t)(DM + DE + se_e) * W + koff;
      const int32_t xlo = hx ? lo_[sx] : 1, xhi = hx ? hi_[sx] : 0;
      const int32_t olo = ho ? lo_[so] : 1, ohi = ho ? hi_[so] : 0;
      const int32_t elo = he ? lo_[se] : 1, ehi = he ? hi_[se] : 0;
      bool found = false;
      // Source windows as (first diagonal, width): ONE unsigned compare says whether a neighbour exists; an
      // absent source matches nothing.  "None" is any value below kWfLive: kWfNone + (a few steps) stays far
      // below it, so None + 1 needs no special case and the recurrences are plain max / add.
      const int32_t oa = ho ? olo : (1 << 30), ea = he ? elo : (1 << 30), xa = hx ? xlo : (1 << 30);
      const uint32_t ow = ho ? (uint32_t)(ohi - olo) : 0u, ew = he ? (uint32_t)(ehi - elo) : 0u, xw = hx ? (uint32_t)(xhi - xlo) : 0u;
      // kUnroll diagonals per lane per trip: all source loads of the trip are issued before
      // the first use, so one L2 round trip covers kUnroll cells instead of one
      constexpr int kUnroll = 8;
      for (int32_t k0 = lo + lane; k0 <= hi; k0 += 32 * kUnroll) {
        int32_t am[kUnroll], bi[kUnroll], ap[kUnroll], bd[kUnroll], mx[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int32_t k = k0 + 32 * u;  // (k > hi: nothing matches or the values are never used)
          am[u] = bi[u] = ap[u] = bd[u] = mx[u] = kWfNone;
          if ((uint32_t)(k - 1 - oa) <= ow) am[u] = Mo[k - 1];
          if ((uint32_t)(k - 1 - ea) <= ew) bi[u] = Ie[k - 1];
          if ((uint32_t)(k + 1 - oa) <= ow) ap[u] = Mo[k + 1];
          if ((uint32_t)(k + 1 - ea) <= ew) bd[u] = De[k + 1];
          if ((uint32_t)(k - xa) <= xw) mx[u] = Mx[k];
        }
        w_cells += (uint32_t)min(kUnroll, (hi - k0) / 32 + 1);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int32_t k = k0 + 32 * u;
          if (k > hi) break;
          // a cell of diagonal k exists for offsets up to min(n1, n2 + k); the lower ends hold by construction
          // (an insertion keeps h = v - k, a deletion keeps v, and every source is inside the matrix)
          const int32_t vmax = min(n1, n2 + k);
          int32_t iv = max(am[u], bi[u]) + 1;  // insertion: consumes a seq1 residue: (k-1) -> k, v + 1
          int32_t dv = max(ap[u], bd[u]);      // deletion: consumes a seq2 residue: (k+1) -> k, v unchanged
          int32_t mv = mx[u] + 1;              // mismatch
          iv = iv > vmax ? kWfNone : iv;
          dv = dv > vmax ? kWfNone : dv;
          mv = mv > vmax ? kWfNone : mv;
          mv = max(mv, max(iv, dv));
          if (mv > kWfLive) {
            const int32_t before = mv;
            mv = wide ? wf_extend<8>(s1w, s2w, mv, mv - k, n1, n2) : wf_extend<2>(s1w, s2w, mv, mv - k, n1, n2);
            w_ext += (uint32_t)(mv - before);
          }
          Mc[k] = mv;
          Ic[k] = iv;
          Dc[k] = dv;
          if (k == kend && mv == n1) found = true;
        }
      }
      if (lane == 0) { lo_[slot] = lo; hi_[slot] = hi; }
      __syncwarp();
      if (__any_sync(0xffffffffu, found)) result = s;
    }
    if (p.work) {
      w_cells = __reduce_add_sync(0xffffffffu, w_cells);
      w_ext = __reduce_add_sync(0xffffffffu, w_ext);
    }
    if (lane == 0) {
      p.score[id] = result;
      p.status[id] = 0;
      if (p.work) {
        atomicAdd(p.work, (unsigned long long)w_cells);
        atomicAdd(p.work + 1, (unsigned long long)w_ext);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// LITERAL MODE: one thread per pair, the reference's algorithm step by step.
// A wavefront component is a Vec<Option<element>> (:118-124); only the offsets matter for
// status and score, so an element is an int32 offset (kWfNone = None).  Each component keeps
// (present, lo, hi, len) and its elements at data[0..len) with element k on diagonal lo + k
// -- exactly the indexing of get_element (:159-163), including its behaviour once trim has
// made lo/hi inconsistent with the vector.
// ---------------------------------------------------------------------------------------------
struct LitComp {
  int32_t present, lo, hi, len;
  int32_t* data;
};
constexpr int kLitRing = 9;  // scores s-8 .. s

__device__ __forceinline__ int32_t lit_get(const LitComp* w, int32_t idx) {  // get_offset :165-171
  if (!w || !w->present) return kWfNone;
  const int64_t k = (int64_t)idx - w->lo;
  if (k < 0 || k >= w->len) return kWfNone;
  return w->data[k];
}
__device__ __forceinline__ uint64_t lit_x(int32_t off, int32_t diag) {  // :85-87, `as usize`
  return (uint64_t)(int64_t)(off - min(diag, 0));
}
__device__ __forceinline__ uint64_t lit_y(int32_t off, int32_t diag) {  // :88-90
  return (uint64_t)(int64_t)(off + max(diag, 0));
}
__device__ __forceinline__ int32_t lit_distance(int32_t off, int32_t n1, int32_t n2, int32_t diag) {
  return max(n1 - off - diag, n2 - off);  // get_distance :96-101
}
__device__ __forceinline__ uint32_t lit_absdiff(int32_t a, int32_t b) {
  return a > b ? (uint32_t)(a - b) : (uint32_t)(b - a);
}
// Vec::rotate_left; false where Rust panics (k > len)
__device__ bool lit_rotate_left(LitComp* w, uint32_t k) {
  if (k > (uint32_t)w->len) return false;
  if (k == 0 || k == (uint32_t)w->len) return true;
  // reverse three times (in place)
  auto rev = [&](int a, int b) {
    for (; a < b; ++a, --b) {
      const int32_t t = w->data[a];
      w->data[a] = w->data[b];
      w->data[b] = t;
    }
  };
  rev(0, (int)k - 1);
  rev((int)k, w->len - 1);
  rev(0, w->len - 1);
  return true;
}
__device__ __forceinline__ void lit_remove_front(LitComp* w) {
  if (w->len) {
    w->data += 1;  // the slot's buffer has room: data only ever moves forward within one score
    w->len -= 1;
  }
}

__global__ void __launch_bounds__(64) wfa_literal_kernel(const WfaParams p) {
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= p.n_launch_pairs) return;
  const uint32_t id = p.pair_base + li;
  const int32_t n1 = (int32_t)p.q_len[id], n2 = (int32_t)p.d_len[id];
  const uint64_t o1 = p.q_off[id], o2 = p.d_off[id];
  const int32_t wcap = (int32_t)p.lit_wcap;
  int32_t* base = p.scratch + (uint64_t)li * p.scratch_stride;  // [kLitRing][3][wcap]
  LitComp ring[kLitRing][3];  // 0 = I, 1 = D, 2 = M
  int32_t tensor_present[kLitRing];
  for (int r = 0; r < kLitRing; ++r) {
    tensor_present[r] = 0;
    for (int c = 0; c < 3; ++c) ring[r][c] = LitComp{0, 0, 0, 0, base + ((size_t)r * 3 + c) * wcap};
  }
  const uint64_t cap64 = 8ull * ((uint64_t)n1 + (uint64_t)n2) + 64;
  const int32_t max_score = (int32_t)(cap64 < 2048 ? cap64 : 2048);
  // Ocean::global :450-465: wavefront 0 = M{lo=hi=0, offset 0}; it is NOT extended
  tensor_present[0] = 1;
  ring[0][2].present = 1;
  ring[0][2].lo = ring[0][2].hi = 0;
  ring[0][2].len = 1;
  ring[0][2].data[0] = 0;
  int32_t len = 1;  // wfs.len()
  int32_t status = -1, score = 0;
  const uint64_t tx = (uint64_t)(int64_t)n2 - 1, ty = (uint64_t)(int64_t)n1 - 1;  // :189 (usize)
  while (status < 0) {
    // is_converged on the last tensor, components in order I, D, M (:422-439)
    const int cur = (len - 1) % kLitRing;
    if (tensor_present[cur]) {
      for (int c = 0; c < 3 && status < 0; ++c) {
        const LitComp& w = ring[cur][c];
        if (!w.present) continue;
        for (int32_t k = 0; k < w.len; ++k) {
          const int32_t off = w.data[k];
          if (off == kWfNone) continue;
          const int32_t diag = w.lo + k;
          if (lit_x(off, diag) == tx && lit_y(off, diag) == ty) {
            status = kOk;
            score = len;  // printed `converged with score {wfs.len()}` (:31-36)
            if (p.trace) {
              // The element's parent list (get_parents :201-209) is a function of the source wavefronts
              // of tensor s = len - 1, which are still in the ring and were final when s was created:
              // recompute it here instead of carrying parents for every element.
              const int32_t s = len - 1;
              auto src = [&](int32_t back, int comp) -> const LitComp* {
                if (s - back < 0 || !tensor_present[(s - back) % kLitRing]) return nullptr;
                const LitComp* q = &ring[(s - back) % kLitRing][comp];
                return q->present ? q : nullptr;
              };
              const LitComp *om = src(p.o + p.e, 2), *mm = src(p.x, 2), *ei = src(p.e, 0), *ed = src(p.e, 1);
              int32_t np = 0, par[3] = {0, 0, 0};
              const int32_t dsm = lit_get(om, diag + 1), dsd = lit_get(ed, diag + 1), dv = max(dsm, dsd);
              const int32_t ism = lit_get(om, diag - 1), isi = lit_get(ei, diag - 1), ib = max(ism, isi);
              const int32_t iv = ib != kWfNone ? ib + 1 : kWfNone;
              if (s > 0 && c == 1) {          // D: sources in the order (M, D), :272-311
                if (dsm != kWfNone && dsm == dv) par[np++] = 0;
                if (dsd != kWfNone && dsd == dv) par[np++] = 1;
              } else if (s > 0 && c == 0) {   // I: (M, I), compared before the increment, :313-352
                if (ism != kWfNone && ism == ib) par[np++] = 0;
                if (isi != kWfNone && isi == ib) par[np++] = 2;
              } else if (s > 0) {             // M: (M, I, D) against the un-extended maximum, :353-398
                const int32_t mb = lit_get(mm, diag);
                const int32_t m0 = mb != kWfNone ? mb + 1 : kWfNone;
                const int32_t mv = max(m0, max(iv, dv));
                if (m0 != kWfNone && m0 == mv) par[np++] = 0;
                if (iv != kWfNone && iv == mv) par[np++] = 2;
                if (dv != kWfNone && dv == mv) par[np++] = 1;
              }
              p.trace[1] = off;
              p.trace[2] = c == 2 ? 0 : (c == 1 ? 1 : 2);
              p.trace[3] = np;
              p.trace[4] = par[0]; p.trace[5] = par[1]; p.trace[6] = par[2];
            }
            break;
          }
        }
      }
      if (status >= 0) break;
    }
    if (len > max_score) {
      status = kRefNoConv;
      break;
    }
    // Ocean::expand :467-488
    const int32_t s = len;
    const int slot = s % kLitRing;
    const LitComp* om = (s - p.o - p.e >= 0 && tensor_present[(s - p.o - p.e) % kLitRing]) ? &ring[(s - p.o - p.e) % kLitRing][2] : nullptr;
    const LitComp* mm = (s - p.x >= 0 && tensor_present[(s - p.x) % kLitRing]) ? &ring[(s - p.x) % kLitRing][2] : nullptr;
    const LitComp* ei = (s - p.e >= 0 && tensor_present[(s - p.e) % kLitRing]) ? &ring[(s - p.e) % kLitRing][0] : nullptr;
    const LitComp* ed = (s - p.e >= 0 && tensor_present[(s - p.e) % kLitRing]) ? &ring[(s - p.e) % kLitRing][1] : nullptr;
    if (om && !om->present) om = nullptr;
    if (mm && !mm->present) mm = nullptr;
    if (ei && !ei->present) ei = nullptr;
    if (ed && !ed->present) ed = nullptr;
    LitComp* ci = &ring[slot][0];
    LitComp* cd = &ring[slot][1];
    LitComp* cm = &ring[slot][2];
    for (int c = 0; c < 3; ++c) {
      ring[slot][c].present = 0;
      ring[slot][c].len = 0;
      ring[slot][c].data = base + ((size_t)slot * 3 + c) * wcap;
    }
    tensor_present[slot] = 0;
    ++len;
    // WaveFrontTensor::new :225-420
    bool any = false;
    int32_t hi = 0, lo = 0;
    const LitComp* srcs[4] = {om, mm, ei, ed};
    for (int k = 0; k < 4; ++k)
      if (srcs[k]) {
        if (!any || srcs[k]->hi > hi) hi = srcs[k]->hi;
        if (!any || srcs[k]->lo < lo) lo = srcs[k]->lo;
        any = true;
      }
    if (!any) continue;  // the tensor is None (:238)
    hi += 1;
    lo -= 1;
    if (p.trace) {  // println!("lo: {}, hi: {}", lo, hi) (:251)
      const int32_t k = p.trace[0];
      if ((uint32_t)k < p.trace_cap) { p.trace[8 + 2 * k] = lo; p.trace[9 + 2 * k] = hi; }
      p.trace[0] = k + 1;
    }
    if ((int64_t)hi - lo + 1 > wcap) {  // cannot happen below max_score; defensive
      status = kRefNoConv;
      break;
    }
    int32_t i_lo = lo, i_hi = hi, d_lo = lo, d_hi = hi, m_lo = lo, m_hi = hi;
    bool i_set = false, d_set = false, m_set = false;
    // every component is first built over [lo, hi] (element k on diagonal lo + k), M without
    // its leading Nones (:396-398)
    ci->lo = lo; cd->lo = lo;
    int32_t m_first = 0;
    for (int32_t idx = lo; idx <= hi; ++idx) {
      const int32_t k = idx - lo;
      const int32_t dv = max(lit_get(om, idx + 1), lit_get(ed, idx + 1));  // :272-311
      cd->data[k] = dv;
      if (dv != kWfNone) { d_hi = idx; if (!d_set) { d_lo = idx; d_set = true; } }
      const int32_t ib = max(lit_get(om, idx - 1), lit_get(ei, idx - 1));  // :313-352
      const int32_t iv = ib != kWfNone ? ib + 1 : kWfNone;
      ci->data[k] = iv;
      if (iv != kWfNone) { i_hi = idx; if (!i_set) { i_lo = idx; i_set = true; } }
      const int32_t mb = lit_get(mm, idx);                                 // :353-398
      int32_t mv = mb != kWfNone ? mb + 1 : kWfNone;
      mv = max(mv, max(iv, dv));
      if (mv != kWfNone) {
        if (!m_set) { m_lo = idx; m_set = true; m_first = k; }
        m_hi = idx;
      }
      cm->data[k] = mv;
    }
    // rotate_left + truncate of :401-409 leave each component's own [lo, hi] window
    ci->data += (i_lo - lo); ci->lo = i_lo; ci->hi = i_hi; ci->len = (int32_t)lit_absdiff(i_hi, i_lo) + 1; ci->present = i_set;
    cd->data += (d_lo - lo); cd->lo = d_lo; cd->hi = d_hi; cd->len = (int32_t)lit_absdiff(d_hi, d_lo) + 1; cd->present = d_set;
    cm->data += m_first;     cm->lo = m_lo; cm->hi = m_hi; cm->len = (int32_t)lit_absdiff(m_hi, m_lo) + 1; cm->present = m_set;
    if (!m_set) cm->len = 0;
    tensor_present[slot] = 1;
    // WaveFrontTensor::expand: extend M only (:219-223, :127-139)
    if (cm->present) {
      for (int32_t k = 0; k < cm->len; ++k) {
        int32_t off = cm->data[k];
        if (off == kWfNone) continue;
        const int32_t diag = cm->lo + k;
        for (;;) {
          const uint64_t y = lit_y(off, diag), x = lit_x(off, diag);
          if (!(y < (uint64_t)n1 && x < (uint64_t)n2 &&
                load_residue(p.residues, o1 + y, p.packing) == load_residue(p.residues, o2 + x, p.packing)))
            break;
          ++off;
        }
        cm->data[k] = off;
      }
    }
    // Ocean::trim :490-623
    if (cm->present && lit_absdiff(cm->lo, cm->hi) > 5) {
      int32_t min_d = 0;  // :511
      for (int32_t diag = cm->lo; diag <= cm->hi; ++diag) {
        const int32_t off = lit_get(cm, diag);
        if (off != kWfNone) min_d = min(min_d, lit_distance(off, n1, n2, diag));
      }
      bool panic = false;
      if (!cm->len || cm->data[0] == kWfNone) panic = true;
      if (!panic) {
        int32_t next_d = lit_distance(cm->data[0], n1, n2, cm->lo);
        while (cm->lo < cm->hi && lit_absdiff(next_d, min_d) > 20) {
          cm->lo += 1;
          lit_remove_front(cm);
          while (lit_get(cm, cm->lo) == kWfNone) {
            if (cm->lo == cm->hi) break;
            cm->lo += 1;
            lit_remove_front(cm);
          }
          if (!cm->len || cm->data[0] == kWfNone) { panic = true; break; }
          next_d = lit_distance(cm->data[0], n1, n2, cm->lo);
        }
      }
      if (!panic && (!cm->len || cm->data[cm->len - 1] == kWfNone)) panic = true;
      if (!panic) {
        int32_t next_d = lit_distance(cm->data[cm->len - 1], n1, n2, cm->hi);
        while (cm->hi > cm->lo && lit_absdiff(next_d, min_d) > 20) {
          cm->hi -= 1;
          if (cm->len) cm->len -= 1;
          while (lit_get(cm, cm->hi) == kWfNone) {
            if (cm->lo == cm->hi) break;
            cm->hi -= 1;
            if (cm->len) cm->len -= 1;
          }
          if (!cm->len || cm->data[cm->len - 1] == kWfNone) { panic = true; break; }
          next_d = lit_distance(cm->data[cm->len - 1], n1, n2, cm->hi);
        }
      }
      for (int c = 0; c < 2 && !panic; ++c) {  // clip I then D to M's range (:574-622)
        LitComp* w = &ring[slot][c];
        if (!w->present) continue;
        uint64_t t;
        if (w->lo < cm->lo) {
          if (!lit_rotate_left(w, lit_absdiff(w->lo, cm->lo))) { panic = true; break; }  // :577 / :603
          t = (uint64_t)lit_absdiff(w->lo, cm->lo) + (w->hi > cm->hi ? lit_absdiff(w->hi, cm->hi) : 0);
        } else if (w->hi > cm->hi) {
          t = lit_absdiff(w->hi, cm->hi);
        } else {
          t = 0;
        }
        if (t <= (uint64_t)w->len) w->len -= (int32_t)t;  // else `len - t` wraps (release): no-op
        w->hi = min(w->hi, cm->hi);
        w->lo = max(w->lo, cm->lo);
      }
      if (panic) {
        status = kRefPanic;
        break;
      }
    }
  }
  p.score[id] = status == kOk ? score : 0;
  p.status[id] = (uint8_t)status;
}

}  // namespace sa
