// nw_affine_s16.cuh -- affine-gap global NW fill for short pairs, packed u16x2, sm_100a.
//
// Computes what ScoreTensor::fill (Global) computes in the reference
// (/root/reference/src/needleman_wunsch_affine.rs:169-237): the three-state M/I/D recurrences
//     M[x][y] = max(M,I,D)[x-1][y-1] + (seq1[y-1]==seq2[x-1] ? match : mismatch)      :76-86
//     I[x][y] = max(M[x][y-1] + open, I[x][y-1]) + ext                                :91-94
//     D[x][y] = max(M[x-1][y] + open, D[x-1][y]) + ext                                :87-90
// with the reference's boundary rows (:183-216: the boundary gap costs one extra extension)
// and, per cell, the four tie bits that determine the FIRST alignment the reference's LIFO
// DFS prints (:246-329).  x in [0,n2] walks seq2 (db), y in [0,n1] walks seq1 (query).
//
// Mapping to the hardware (see DESIGN.md section 4 for the derivation and the roofline):
//   * Two pairs per register: every 32-bit register holds the same quantity for pair A (low
//     half) and pair B (high half) as biased unsigned 16-bit numbers, so one VIMNMX.U16x2
//     does two max operations and returns BOTH "a >= b" predicates, which are exactly the
//     traceback tie bits.  Plain 32-bit IADD does two subtractions because no half can
//     borrow (range proven below).
//   * Scores are stored as V' = 2*V - 2*ext*(x+y) + bias.  Every comparison in the
//     recurrence is between values of the same cell, so the per-cell offset cancels, and a gap
//     EXTENSION becomes free: I'[x][y+1] = max(M' - open', I') with no add, likewise D'.  The
//     diagonal step is + (2*match - 4*ext) for a match and 2*(match-mismatch) less for a
//     mismatch: one XOR + one VIMNMX.U16x2 (min(q^d, penalty)) + one three-operand IADD3.
//     A cell therefore costs 2 adds, 5 VIMNMX, 1 XOR and 8 predicated bit-sets = 16
//     instructions per two cells; values stay within bias + (2*match-4*ext)*min(x,y).
//   * The factor 2 leaves the low bit free: boundary-chain cells get +1 ("panic bonus"),
//     the bonus can only break ties, and it survives to the end cell iff some co-optimal
//     path starts with a gap -- the condition under which the reference panics
//     (nw_affine:299/:303).  Status is therefore exact with zero extra instructions.
//   * G lanes share one pair-of-pairs: lane j owns column strips j, j+G, j+2G.. (K columns
//     each, in registers) and runs one row behind lane j-1; the strip's right-edge H and E
//     go to lane j+1 by __shfl_up_sync and from the last lane of the group to shared memory
//     for the next pass.  Sequence words are staged once per tile in shared memory.
//   * Traceback bits (4 per cell) leave the SM as coalesced 8-byte stores per lane per
//     row-strip: [strip][row][W][group] uint2 = {pair A's 8 cells, pair B's 8 cells}.
//   * K (columns per lane) and G are chosen per shape class.  When K*G covers the whole query
//     (150 bp: K = 19, G = 8) the launch is SINGLE-pass: no boundary column, no shared memory
//     beyond the db panel, and the per-row overhead (shuffles, loads, stores) is spread over
//     19 columns instead of 8.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sa {

constexpr uint32_t kBias = 0xFF00u;  // linear aligner (S' = 2S - match*(i+j) + kBias, S' <= kBias)
// Sentinel guard, in doubled score units: twice the most negative true score of any state of
// any cell (a path of gaps only, plus one mismatch, one open and one extension of look-ahead)
// plus the anti-diagonal term.  While it stays below kBias no score comes near the reference's
// -32768 sentinel (nw_affine:183-216), so the packed kernels may ignore the sentinel; pairs
// beyond it go to the literal 32-bit kernel.  The linear aligner's 16-bit range is the same bound.
__host__ __device__ inline uint32_t s16_min_value_bound(int match, int mismatch, int open, int ext,
                                                        uint32_t n1pad, uint32_t n2) {
  const int extp = match - 2 * ext;                  // > 0 magnitude per boundary step
  const int openp = -2 * open;                       // > 0
  const int pen = 2 * (match - mismatch);            // > 0
  // worst case: every step is the costliest of (ext', pen/2 per unit of x+y)
  const int per_step = extp > (pen + 1) / 2 ? extp : (pen + 1) / 2;
  return (uint32_t)(2 * (openp + extp) + per_step * (int)(n1pad + n2 + 2) + pen + openp + extp);
}

// Range of the affine transform: every state of every cell lies in
//   [bias - (2*open' + ext'' + max(0, pen - cm)), bias + cm*min(x,y) + 1]
// with open' = -2*open, ext'' = -2*ext, cm = 2*match - 4*ext, pen = 2*(match-mismatch): the
// lower end because D'(x,y) >= D'(1,y) = H'(0,y) = bias - (open' + ext'') (extensions are free),
// M' = H'diag + cm - {0,pen} and the look-ahead M' - open'; the upper end because a diagonal
// step adds at most cm and gap steps add nothing (+1: panic bonus).
__host__ __device__ inline uint32_t s16_affine_bias(int match, int mismatch, int open, int ext) {
  const int cm = 2 * match - 4 * ext, pen = 2 * (match - mismatch);
  return (uint32_t)(-4 * open - 2 * ext + (pen > cm ? pen - cm : 0) + 16);
}
__host__ __device__ inline bool s16_affine_in_range(int match, int mismatch, int open, int ext,
                                                    uint32_t n1pad, uint32_t n2) {
  const uint64_t cm = (uint64_t)(2 * match - 4 * ext);
  const uint64_t nmin = n1pad < n2 ? n1pad : n2;
  return s16_affine_bias(match, mismatch, open, ext) + cm * (nmin + 1) + 2 <= 0xFFFFull;
}

// Residue access for both input formats of sa_batch_t: packing 0 = one byte per residue, offsets
// in bytes; packing 1 = 2-bit codes (A=0 C=1 G=2 T=3), four per byte, residue i of the buffer at
// bits 2*(i&3) of byte i>>2, offsets in RESIDUES.  Only equality of residues is ever used.
__device__ __forceinline__ uint32_t load_residue(const uint8_t* __restrict__ residues, uint64_t pos,
                                                 uint32_t packing) {
  if (packing == 0) return residues[pos];
  return ((uint32_t)residues[pos >> 2] >> (2 * (uint32_t)(pos & 3))) & 3u;
}

struct AffineS16Params {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  const uint32_t* __restrict__ pair_ids;  // launch index -> pair id, or nullptr for identity+base
  uint32_t pair_base;
  uint32_t n_launch_pairs;
  uint2* __restrict__ tb;        // traceback scratch, tile-major
  uint64_t tb_tile_stride;       // uint2 per tile
  uint32_t tb_rows;              // row stride (launch-wide max n2)
  uint32_t* __restrict__ end;    // per launch index: H'(16) | start_state << 16 | valid << 31
  uint32_t smem_bnd_rows;        // rows of the boundary column that precedes the db panel (multi-pass)
  // scheme in transformed units (all positive magnitudes)
  uint32_t pen2;    // 2*(match-mismatch), packed in both halves
  uint32_t open2;   // -2*open, packed
  uint32_t ext2;    // linear aligner only: what a set gap flag saves, 2*(ext-open), packed
  uint32_t cm2;     // affine: diagonal constant 2*match - 4*ext, packed
  uint32_t row0;    // affine: bias + 2*(open+ext) + bonus, packed = H'[0][y] = H'[x][0] for all x,y >= 1
                    // linear: kBias + open', H'[0][y] = row0 - y*step2
  uint32_t origin;  // bias packed: H'[0][0]
  uint32_t zero;    // always 0; opaque to ptxas so that `or` bit-sets stay LOP3 (alu pipe)
  uint32_t step2;   // linear aligner: boundary step magnitude match - 2*ext, packed
  uint32_t packing; // input format (see load_residue)
  // affine fill with the panic bonus on: a pair whose end cell carries the bonus (low bit of H') is queued
  // here for the clean refill the moment its end cell is computed, so the refill does not wait for the walk
  uint32_t* __restrict__ rerun_ids;
  uint32_t* __restrict__ rerun_count;
};

// min of two packed u16 pairs on whole 32-bit registers.  (__vminu2 takes its operands apart into 16-bit
// halves; with a kernel parameter as one operand ptxas then rebuilds the packed constant with a PRMT
// in front of every use: 12 of them per 19-column row step.)
__device__ __forceinline__ uint32_t vmin_u16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// max of two packed u16 pairs that also records, per half, whether the FIRST operand won or
// tied: bit BIT of acc_lo (low half = pair A) / acc_hi (high half = pair B) is set when
// a >= b.  ptxas folds the setp.eq pair into the predicate outputs of one VIMNMX.U16x2, and
// each conditional add becomes one predicated integer add (either integer pipe), so a tie
// bit costs one issue slot instead of a SEL plus a merge.
template <uint32_t BIT, bool LO_OR, bool HI_OR>
__device__ __forceinline__ uint32_t vmax_tie(uint32_t a, uint32_t b, uint32_t& acc_lo,
                                             uint32_t& acc_hi) {
  uint32_t r;
  // The two conditional bit-sets can issue on either integer pipe: `or` becomes a predicated
  // LOP3 (alu pipe), `add` a predicated VIADD (fma-heavy pipe on sm_100).  Which of the eight
  // bit-sets of a cell go where is a template mask so the pipes can be balanced against the
  // VIMNMX (alu, half rate) and IMAD.IADD (fma-heavy) work of the recurrence.
  if (LO_OR && HI_OR) {
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
        "max.u16x2 %0, %3, %4;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
        "@pl or.b32 %1, %1, %5;\n\t@ph or.b32 %2, %2, %5;\n\t}"
        : "=r"(r), "+r"(acc_lo), "+r"(acc_hi) : "r"(a), "r"(b), "n"(BIT));
  } else if (LO_OR) {
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
        "max.u16x2 %0, %3, %4;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
        "@pl or.b32 %1, %1, %5;\n\t@ph add.u32 %2, %2, %5;\n\t}"
        : "=r"(r), "+r"(acc_lo), "+r"(acc_hi) : "r"(a), "r"(b), "n"(BIT));
  } else if (HI_OR) {
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
        "max.u16x2 %0, %3, %4;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
        "@pl add.u32 %1, %1, %5;\n\t@ph or.b32 %2, %2, %5;\n\t}"
        : "=r"(r), "+r"(acc_lo), "+r"(acc_hi) : "r"(a), "r"(b), "n"(BIT));
  } else {
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
        "max.u16x2 %0, %3, %4;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
        "@pl add.u32 %1, %1, %5;\n\t@ph add.u32 %2, %2, %5;\n\t}"
        : "=r"(r), "+r"(acc_lo), "+r"(acc_hi) : "r"(a), "r"(b), "n"(BIT));
  }
  return r;
}

// ORMASK: bit (2*k + half) set -> the tie bit of comparison k (0 pI, 1 pD, 2 pE, 3 pF) for
// that half is set with `or` (alu pipe), else with `add` (fma-heavy pipe).
// The traceback nibble of column c lives in word c/8 of the pair's accumulators.
// CAP: also keep M' and the incoming I' of every column (only the rarely executed copy of the
// row that holds a pair's end cell needs them).
template <int K, int C, uint32_t ORMASK, bool CAP>
struct StripCells {
  static constexpr int W = (K + 7) / 8;
  static __device__ __forceinline__ void run(uint32_t (&Hrow)[K], uint32_t (&F)[K],
                                             const uint32_t (&q)[K], uint32_t d, uint32_t hdiag,
                                             uint32_t& E, uint32_t pen2, uint32_t open2,
                                             uint32_t cm2, uint32_t zero, uint32_t (&acc_a)[W], uint32_t (&acc_b)[W],
                                             uint32_t (&Mv)[CAP ? K : 1], uint32_t (&Ev)[CAP ? K : 1]) {
    constexpr int c = C;
    constexpr int w = c / 8;
    constexpr uint32_t sh = 4 * (c % 8);
    // ORMASK bits 8-11 = period P, bits 12-15 = count N, bits 16-19 = offset O (P = 0: every column): the `or` selection of
    // the low eight bits applies to the columns with (c + O) % P < N only, so the share of tie-bit sets that change pipe is adjustable in
    // steps of one column
    constexpr uint32_t OP = (ORMASK >> 8) & 15u, ON = (ORMASK >> 12) & 15u, OO = (ORMASK >> 16) & 15u;  // (.. + offset)
    constexpr uint32_t OM = (OP != 0u && ((uint32_t)c + OO) % (OP ? OP : 1u) >= ON) ? 0u : (ORMASK & 0xFFu);
    // A new accumulator word starts from the previous one (AND 0): one dependency chain for all
    // tie-bit sets keeps ptxas from hoisting the VIMNMXs of later columns, whose predicates it
    // would otherwise have to spill (there are only seven predicate registers).
    if (c > 0 && c % 8 == 0) {
      acc_a[w] = acc_a[w - 1] & zero;
      acc_b[w] = acc_b[w - 1] & zero;
    }
    const uint32_t hup = Hrow[c];
    const uint32_t m = vmin_u16x2(q[c] ^ d, pen2);  // 0 if equal, penalty otherwise (per half)
    const uint32_t M = hdiag + cm2 - m;           // M'[x][y]: one IADD3; no half leaves [0, 65535]
    if (CAP) {
      Mv[CAP ? c : 0] = M;
      Ev[CAP ? c : 0] = E;
    }
    const uint32_t t = vmax_tie<(1u << sh), (OM >> 0) & 1, (OM >> 1) & 1>(E, M, acc_a[w], acc_b[w]);     // I >= M
    const uint32_t H = vmax_tie<(2u << sh), (OM >> 2) & 1, (OM >> 3) & 1>(F[c], t, acc_a[w], acc_b[w]);  // D >= max(I,M)
    const uint32_t Mo = M - open2;
    E = vmax_tie<(4u << sh), (OM >> 4) & 1, (OM >> 5) & 1>(Mo, E, acc_a[w], acc_b[w]);        // open ties/wins: I'[x][y+1]
    F[c] = vmax_tie<(8u << sh), (OM >> 6) & 1, (OM >> 7) & 1>(Mo, F[c], acc_a[w], acc_b[w]);  // open ties/wins: D'[x+1][y]
    Hrow[c] = H;
    StripCells<K, C + 1, ORMASK, CAP>::run(Hrow, F, q, d, hup, E, pen2, open2, cm2, zero, acc_a, acc_b, Mv, Ev);
  }
};
template <int K, uint32_t ORMASK, bool CAP>
struct StripCells<K, K, ORMASK, CAP> {
  static constexpr int W = (K + 7) / 8;
  static __device__ __forceinline__ void run(uint32_t (&)[K], uint32_t (&)[K], const uint32_t (&)[K],
                                             uint32_t, uint32_t, uint32_t&, uint32_t, uint32_t,
                                             uint32_t, uint32_t, uint32_t (&)[W], uint32_t (&)[W],
                                             uint32_t (&)[CAP ? K : 1], uint32_t (&)[CAP ? K : 1]) {}
};

// ---------------------------------------------------------------------------------------------
// Single-matrix "linear / pseudo-affine" NW of /root/reference/src/needleman_wunsch.rs:66-103:
//     diag  = S[i-1][j-1] + (eq ? match : mismatch)
//     down  = S[i-1][j]   + (gaps[i-1][j] ? ext : open)        (consumes seq1[i-1])
//     right = S[i][j-1]   + (gaps[i][j-1] ? ext : open)        (consumes seq2[j-1])
//     S[i][j] = max3; gaps[i][j] = (S == down || S == right); moves pushed Down, Right, Diag
// Rows walk seq1 and columns seq2 in the reference (:38); the kernel is launched with the
// sequence roles swapped to match.  The first printed hit follows the FIRST stored move, i.e.
// Down if S == down, else Right if S == right, else Diag -- which two predicates decide:
//     p1 = down >= right,  p2 = max(down,right) >= diag   (p2 is also the gap flag)
// so the traceback needs 2 bits per cell (kept in the low half of the affine nibble).
// The per-cell gap cost is carried as a packed magnitude g = open' - p2*(open' - ext').
// ---------------------------------------------------------------------------------------------
enum { kAffine = 0, kLinear = 1 };

// max that records the tie bit and, when the first operand wins or ties, also subtracts DELTA
// from the matching half of g (the gap-cost magnitude of this cell)
template <uint32_t BIT>
__device__ __forceinline__ uint32_t vmax_tie_g(uint32_t a, uint32_t b, uint32_t& acc_lo,
                                               uint32_t& acc_hi, uint32_t& g, uint32_t delta_lo,
                                               uint32_t delta_hi) {
  uint32_t r;
  asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
      "max.u16x2 %0, %4, %5;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %4;\n\t"
      "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
      "@pl or.b32 %1, %1, %6;\n\t@ph or.b32 %2, %2, %6;\n\t"
      "@pl sub.u32 %3, %3, %7;\n\t@ph sub.u32 %3, %3, %8;\n\t}"
      : "=r"(r), "+r"(acc_lo), "+r"(acc_hi), "+r"(g)
      : "r"(a), "r"(b), "n"(BIT), "r"(delta_lo), "r"(delta_hi));
  return r;
}

template <int K, int C>
struct LinearCells {
  // Hrow: S' of the row above; Gm: gap-cost magnitude of the cell above; (sl, gl): S' and gap
  // magnitude of the cell to the left (in: left boundary, out: last column of the strip)
  static __device__ __forceinline__ void run(uint32_t (&Hrow)[K], uint32_t (&Gm)[K],
                                             const uint32_t (&q)[K], uint32_t d, uint32_t sdiag,
                                             uint32_t& sl, uint32_t& gl, uint32_t pen2,
                                             uint32_t gopen2, uint32_t delta_lo, uint32_t delta_hi,
                                             uint32_t& acc_a, uint32_t& acc_b) {
    constexpr int c = C;
    const uint32_t hup = Hrow[c];
    const uint32_t m = vmin_u16x2(q[c] ^ d, pen2);
    const uint32_t diag = sdiag - m;
    const uint32_t down = hup - Gm[c];
    const uint32_t right = sl - gl;
    const uint32_t t = vmax_tie<(1u << (4 * c)), true, true>(down, right, acc_a, acc_b);  // p1
    uint32_t g = gopen2;
    const uint32_t S = vmax_tie_g<(2u << (4 * c))>(t, diag, acc_a, acc_b, g, delta_lo, delta_hi);  // p2
    Hrow[c] = S;
    Gm[c] = g;
    sl = S;
    gl = g;
    LinearCells<K, C + 1>::run(Hrow, Gm, q, d, hup, sl, gl, pen2, gopen2, delta_lo, delta_hi, acc_a, acc_b);
  }
};
template <int K>
struct LinearCells<K, K> {
  static __device__ __forceinline__ void run(uint32_t (&)[K], uint32_t (&)[K], const uint32_t (&)[K],
                                             uint32_t, uint32_t, uint32_t&, uint32_t&, uint32_t,
                                             uint32_t, uint32_t, uint32_t, uint32_t&, uint32_t&) {}
};

// one residue byte per pair (pair A in byte 0, pair B in byte 1) -> (byte << 8) in each 16-bit
// half: a single PRMT.  Different residues then XOR to >= 256 > penalty, equal ones to 0.
__device__ __forceinline__ uint32_t widen(uint32_t v) {
  return __byte_perm(v, 0u, 0x1404u);
}

// start state of the traceback at the end cell (nw_affine:251-280: pushed I, M, D; popped
// D, M, I): D if D == max, else M if M == max, else I.   codes: 0 = M, 1 = I, 2 = D
__device__ __forceinline__ uint32_t end_word(uint32_t H, uint32_t M, uint32_t E, bool d_wins) {
  const uint32_t st = d_wins ? 2u : (M >= E ? 0u : 1u);
  return H | (st << 16) | 0x80000000u;
}

// Per-lane state of one pass over a strip (all in registers).
template <int K>
struct StripState {
  uint32_t Hrow[K], F[K], q[K];
  uint32_t hd_prev;      // H'[x-1][y0]: the diagonal input of the strip's first column
  uint32_t out_h, out_e; // right boundary of the row just computed: H'[x][yK], E'[x][yK+1]
  uint32_t x;            // row this lane computes next (1-based)
  const uint16_t* dptr;  // shared: db residues of row x (both pairs)
  uint2* bptr;           // shared: boundary column entry of row x (multi-pass launches only)
  uint2* tptr;           // global: traceback words of (strip, row x)
  uint32_t capx_a, capx_b;  // row of pair A's / B's end cell if it lies in this strip, else 0
};

// One row of the lane's strip.  CHECKED = the lane may be outside [1, n2t] (ramp rows of a
// pass, where the G lanes of a group are not all active yet / any more).  SINGLE = the launch
// has one pass: the left edge of lane 0 is column 0 (a constant in V') and nothing is handed
// to a next pass, so no boundary column exists at all.
// NOCAP = no lane of the warp computes a pair's end cell in this row (decided per warp for a whole run
// of rows), so the row carries neither the two compares nor the branch over the end-cell copy.
template <int K, int G, uint32_t ORMASK, bool CHECKED, int ALGO, bool SINGLE, bool NOCAP = false>
__device__ __forceinline__ void row_step(StripState<K>& st, const AffineS16Params& p, int j,
                                         uint32_t n2t, uint32_t pen2, uint32_t open2,
                                         uint32_t ext2, uint32_t zero, uint32_t la, uint32_t lb,
                                         uint32_t ca_, uint32_t cb_) {
  constexpr int NG = 32 / G;
  constexpr int W = (K + 7) / 8;
  uint32_t rh = 0, re = 0;
  if (G > 1) {
    rh = __shfl_up_sync(0xffffffffu, st.out_h, 1);
    re = __shfl_up_sync(0xffffffffu, st.out_e, 1);
  }
  const bool active = !CHECKED || (st.x >= 1 && st.x <= n2t);
  if (active) {
    if (j == 0) {  // left edge of the group
      if (SINGLE) {
        rh = re = p.row0;  // H'[x][0] = I'[x][0], and I'[x][1] extends it for free (:200-216)
      } else {
        const uint2 b = *st.bptr;  // column 0 (staged in the prologue) or the previous pass's last strip
        rh = b.x;
        re = b.y;
      }
    }
    const uint32_t d = widen(*st.dptr);
    uint32_t E = re;
    uint32_t acc_a[W], acc_b[W];
    acc_a[0] = acc_b[0] = zero;  // the other words are started inside StripCells
    if constexpr (ALGO == kLinear) {
      // open2 = open' magnitude, ext2 = open' - ext' (what a set gap flag saves), per half
      uint32_t sl = rh;
      LinearCells<K, 0>::run(st.Hrow, st.F, st.q, d, st.hd_prev, sl, E, pen2, open2, ext2 & 0xffffu,
                             ext2 & 0xffff0000u, acc_a[0], acc_b[0]);
    } else if (NOCAP || (st.x != st.capx_a && st.x != st.capx_b)) {
      uint32_t Mv[1], Ev[1];
      StripCells<K, 0, ORMASK, false>::run(st.Hrow, st.F, st.q, d, st.hd_prev, E, pen2, open2, ext2,
                                           zero, acc_a, acc_b, Mv, Ev);
    } else {  // rare: this row holds a pair's end cell -- same cells, M' and I' kept
      uint32_t Mv[K], Ev[K];
      StripCells<K, 0, ORMASK, true>::run(st.Hrow, st.F, st.q, d, st.hd_prev, E, pen2, open2, ext2,
                                          zero, acc_a, acc_b, Mv, Ev);
      if (st.x == st.capx_a) {
        uint32_t H = 0, M = 0, Ei = 0, dw = 0;
#pragma unroll
        for (int c = 0; c < K; ++c)
          if ((uint32_t)c == ca_) {
            H = st.Hrow[c] & 0xffffu; M = Mv[c] & 0xffffu; Ei = Ev[c] & 0xffffu;
            dw = (acc_a[c / 8] >> (4 * (c % 8) + 1)) & 1u;
          }
        p.end[la] = end_word(H, M, Ei, dw);
        if (p.rerun_count && (H & 1u)) p.rerun_ids[atomicAdd(p.rerun_count, 1u)] = p.pair_ids ? p.pair_ids[la] : p.pair_base + la;
      }
      if (st.x == st.capx_b) {
        uint32_t H = 0, M = 0, Ei = 0, dw = 0;
#pragma unroll
        for (int c = 0; c < K; ++c)
          if ((uint32_t)c == cb_) {
            H = st.Hrow[c] >> 16; M = Mv[c] >> 16; Ei = Ev[c] >> 16;
            dw = (acc_b[c / 8] >> (4 * (c % 8) + 1)) & 1u;
          }
        p.end[lb] = end_word(H, M, Ei, dw);
        if (p.rerun_count && (H & 1u)) p.rerun_ids[atomicAdd(p.rerun_count, 1u)] = p.pair_ids ? p.pair_ids[lb] : p.pair_base + lb;
      }
    }
    if (ALGO == kLinear && !NOCAP) if (st.x == st.capx_a || st.x == st.capx_b) {
      // the linear aligner's end word is just S' (its walk needs no start state)
      if (st.x == st.capx_a) {
        uint32_t H = 0;
#pragma unroll
        for (int c = 0; c < K; ++c)
          if ((uint32_t)c == ca_) H = st.Hrow[c] & 0xffffu;
        p.end[la] = end_word(H, 0, 0, false);
      }
      if (st.x == st.capx_b) {
        uint32_t H = 0;
#pragma unroll
        for (int c = 0; c < K; ++c)
          if ((uint32_t)c == cb_) H = st.Hrow[c] >> 16;
        p.end[lb] = end_word(H, 0, 0, false);
      }
    }
    st.hd_prev = rh;
    st.out_h = st.Hrow[K - 1];
    st.out_e = E;
    if (!SINGLE && j == G - 1) *st.bptr = make_uint2(st.out_h, st.out_e);
#pragma unroll
    for (int w = 0; w < W; ++w) st.tptr[w * NG] = make_uint2(acc_a[w], acc_b[w]);
  }
  // the cursor advances whether or not the row was in range, so that x == t - j always
  st.x += 1;
  st.dptr += NG;
  if (!SINGLE) st.bptr += NG;
  st.tptr += NG * W;
}

// Traceback layout: [tile][strip][row][W][group] uint2 = {pair A's word w, pair B's word w};
// word w of a strip holds the nibbles of its columns 8w .. 8w+7.  The groups of a warp that
// share a strip (same lane index j) write one contiguous run of NG * 8 bytes per word: whole
// 32-byte sectors for NG >= 4.
template <int K, int G, uint32_t ORMASK, int ALGO = kAffine, bool SINGLE = false, int MINB = 1>
__global__ void __launch_bounds__(32, MINB) nw_affine_fill_s16(const AffineS16Params p) {
  constexpr int NG = 32 / G;       // pair-of-pairs per warp tile
  constexpr int PPT = 2 * NG;      // pairs per tile
  constexpr int W = (K + 7) / 8;   // traceback words per pair per strip row
  static_assert(ALGO == kAffine || (K == 8 && !SINGLE), "the linear aligner uses the 8-column multi-pass form");
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x;
  const int grp = lane / G, j = lane % G;
  const uint32_t tile = blockIdx.x;

  // ---- this lane's two pairs -------------------------------------------------------------
  const uint32_t la = tile * PPT + 2 * grp, lb = la + 1;  // launch indices
  uint32_t n1a = 0, n2a = 0, n1b = 0, n2b = 0;
  uint64_t qoa = 0, doa = 0, qob = 0, dob = 0;
  if (la < p.n_launch_pairs) {
    const uint32_t id = p.pair_ids ? p.pair_ids[la] : p.pair_base + la;
    n1a = p.q_len[id]; n2a = p.d_len[id]; qoa = p.q_off[id]; doa = p.d_off[id];
  }
  if (lb < p.n_launch_pairs) {
    const uint32_t id = p.pair_ids ? p.pair_ids[lb] : p.pair_base + lb;
    n1b = p.q_len[id]; n2b = p.d_len[id]; qob = p.q_off[id]; dob = p.d_off[id];
  }
  // pairs with an empty side have no interior cells; the walk kernel handles them in closed form
  if (n1a == 0 || n2a == 0) n1a = n2a = 0;
  if (n1b == 0 || n2b == 0) n1b = n2b = 0;
  uint32_t n1t = max(n1a, n1b), n2t = max(n2a, n2b);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n1t = max(n1t, __shfl_xor_sync(0xffffffffu, n1t, o));
    n2t = max(n2t, __shfl_xor_sync(0xffffffffu, n2t, o));
  }
  if (n1t == 0 || n2t == 0) return;
  const uint32_t nstrips = (n1t + K - 1) / K;
  const uint32_t npass = SINGLE ? 1u : (nstrips + G - 1) / G;

  // shared memory: [boundary column, multi-pass only: uint2 [rows][NG]] [db panel: u16 [rows][NG]]
  // the panel holds one byte per pair per row: low byte pair A, high byte pair B
  uint2* bnd = reinterpret_cast<uint2*>(smem);
  uint16_t* dp = reinterpret_cast<uint16_t*>(bnd + (SINGLE ? 0 : (size_t)p.smem_bnd_rows * NG));

  // ext2: the per-cell constant of the recurrence (affine: diagonal constant; linear: flag saving)
  // pen2 as an opaque 32-bit register value.  Read straight from the parameter bank, ptxas
  // keeps its two 16-bit halves in uniform registers and rebuilds the packed word (2 moves + PRMT) in front of
  // most VIMNMX.U16x2 that use it -- 14 extra instructions per 19-column row step.
  const uint32_t pen2 = p.pen2 + (n1t & p.zero);  // (n1t: a per-lane value from memory; p.zero: always 0, unknown to ptxas)
  const uint32_t open2 = p.open2, ext2 = (ALGO == kLinear) ? p.ext2 : p.cm2, zero = p.zero;

  // ---- stage the db residues (one per row, read by every strip); they are widened to
  //      (byte << 8) per 16-bit half when read, so the XOR of two different residues is >= 256 >
  //      pen2 and the XOR of equal residues is 0.  Query residues are read once per strip,
  //      straight from global memory into registers: shared memory bounds the warps per SM ----
  for (uint32_t x = j; x < n2t; x += G) {
    const uint32_t a = (x < n2a) ? load_residue(p.residues, doa + x, p.packing) : 0u;
    const uint32_t b = (x < n2b) ? load_residue(p.residues, dob + x, p.packing) : 0u;
    dp[x * NG + grp] = (uint16_t)(a | (b << 8));
    if (!SINGLE) {
      // column 0 as the "previous pass" of pass 0.  Affine: H'[x][0] = I'[x][0] (:200-216), and
      // I'[x][1] extends it (M[x][0] + open is the sentinel).  Linear: S'[i][0] with its gap flag
      // set (needleman_wunsch.rs:55-64), i.e. the cell to its right pays an extension.
      if (ALGO == kLinear) {
        bnd[x * NG + grp] = make_uint2(p.row0 - (x + 1) * p.step2, open2 - ext2);
      } else {
        bnd[x * NG + grp] = make_uint2(p.row0, p.row0);  // extensions are free in V'
      }
    }
  }
  __syncwarp();

  // end-cell capture coordinates (strip, column-in-strip) per half
  const uint32_t sa_ = n1a ? (n1a - 1) / K : 0xffffffffu, ca_ = n1a ? (n1a - 1) % K : 0;
  const uint32_t sb_ = n1b ? (n1b - 1) / K : 0xffffffffu, cb_ = n1b ? (n1b - 1) % K : 0;

  uint2* tb_tile = p.tb + (uint64_t)tile * p.tb_tile_stride;

  for (uint32_t pass = 0; pass < npass; ++pass) {
    const uint32_t s = pass * G + j;  // this lane's strip
    const uint32_t y0 = s * K;        // columns to the left of the strip
    StripState<K> st;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const uint32_t y = y0 + c + 1;
      if (ALGO == kLinear) {
        st.Hrow[c] = p.row0 - y * p.step2;  // S'[0][j], gap flag set (needleman_wunsch.rs:45-54)
        st.F[c] = open2 - ext2;             // -> the cell below pays an extension
      } else {
        st.Hrow[c] = p.row0;  // H'[0][y] = D'[0][y]  (nw_affine:194-198), constant in V'
        st.F[c] = p.row0;     // D'[1][y] extends D[0][y]; M[0][y]+open is the sentinel
      }
      const uint32_t qa = (y <= n1a) ? load_residue(p.residues, qoa + y - 1, p.packing) : 0u;
      const uint32_t qb = (y <= n1b) ? load_residue(p.residues, qob + y - 1, p.packing) : 0u;
      st.q[c] = widen(qa | (qb << 8));
    }
    if (ALGO == kLinear)
      st.hd_prev = (y0 == 0) ? p.origin : p.row0 - y0 * p.step2;  // S'[0][j0]; S[0][0] = 2*open
    else
      st.hd_prev = (y0 == 0) ? p.origin : p.row0;  // H'[0][y0]
    st.out_h = 0;
    st.out_e = 0;
    st.capx_a = (s == sa_) ? n2a : 0u;
    st.capx_b = (s == sb_) ? n2b : 0u;
    // lane j runs j rows behind lane 0: its cursor starts at row 1 - j
    st.x = 1u - (uint32_t)j;
    st.dptr = dp + grp - (ptrdiff_t)j * NG;
    st.bptr = bnd + grp - (ptrdiff_t)j * NG;
    st.tptr = tb_tile + ((uint64_t)s * p.tb_rows - (ptrdiff_t)j) * (NG * W) + grp;

    uint32_t t = 1;
    // ramp-up: lanes j >= t are not active yet
    for (; t < (uint32_t)G && t <= n2t + G - 1; ++t)
      row_step<K, G, ORMASK, true, ALGO, SINGLE>(st, p, j, n2t, pen2, open2, ext2, zero, la, lb, ca_, cb_);
    // steady state: every lane is inside [1, n2t].  Lane j computes row x in step x + j, so the first step in which
    // any lane reaches an end cell is known up front; the steps before it (nearly all of them when the tile's
    // pairs have about the same number of rows, which the segment's ordering by rows arranges) skip that test.
    uint32_t t_cap = 0xffffffffu;
    if (st.capx_a) t_cap = min(t_cap, st.capx_a + (uint32_t)j);
    if (st.capx_b) t_cap = min(t_cap, st.capx_b + (uint32_t)j);
    t_cap = min(__reduce_min_sync(0xffffffffu, t_cap), n2t + 1);
#pragma unroll 2
    for (; t < t_cap; ++t)
      row_step<K, G, ORMASK, false, ALGO, SINGLE, true>(st, p, j, n2t, pen2, open2, ext2, zero, la, lb, ca_, cb_);
    for (; t <= n2t; ++t)
      row_step<K, G, ORMASK, false, ALGO, SINGLE>(st, p, j, n2t, pen2, open2, ext2, zero, la, lb, ca_, cb_);
    // ramp-down
    for (; t <= n2t + G - 1; ++t)
      row_step<K, G, ORMASK, true, ALGO, SINGLE>(st, p, j, n2t, pen2, open2, ext2, zero, la, lb, ca_, cb_);
    __syncwarp();
  }
}

}  // namespace sa
