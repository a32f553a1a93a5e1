// nw_long.cuh -- affine-gap global NW for LONG pairs (n1 + n2 beyond the packed 16-bit kernel,
// up to 100 kbp and more), tiled over the whole GPU, sm_100a.
//
// Same recurrences, boundary rows and tie-breaking as the packed kernel (nw_affine_s16.cuh), i.e.
// ScoreTensor::fill + traceback of /root/reference/src/needleman_wunsch_affine.rs:169-329, but in
// 32-bit integers with the reference's LITERAL finite "minus infinity" -32768 (:174-215), which at
// these lengths lies far above the real scores near the borders and leaks into the matrix.
//
//   V' = 4*V - 4*ext*(x+y) + b        b in {0,1,2}: provenance bonus in the two low bits
//
// The transform makes gap extensions free (as in the packed kernel); a cell costs 6 issue slots in
// the score-only forward pass (LOP3, VIMNMX, IADD3, VIMNMX3, 2 x VIADDMNMX) and 12 with the four
// tie bits.  The bonus rides on maxima and only ever breaks ties:
//   b = 1  boundary-chain cells D[0][y>=1], I[x>=1][0]   (reaching one = the reference panics,
//          :299/:303)
//   b = 2  the parentless sentinel cells of row 0 / column 0 (reaching one = a DEAD END of the
//          reference's depth-first traceback: nothing printed for that branch)
// so the end cell carries max(b) over the sources of ALL co-optimal paths:
//   0: every co-optimal path starts at the origin -> status OK, first alignment = greedy walk;
//   1: some path starts with a gap, none at a dead end -> the walk decides REF_PANIC / _EARLY;
//   2: a dead end is reachable -> the pair is handed to the literal kernel (nw_general.cuh), which
//      models the DFS cell by cell.  That needs a gap of > 5.4 k residues to be optimal; it is
//      a fallback, not a path real data takes.
//
// Tiling ("intra-pair anti-diagonal tiling across SMs", BASELINE.json configs[4]): a pair's matrix is
// cut into tiles of R rows x (S * 512) columns.  One WARP computes one tile: 32 lanes x 16 columns
// in registers, lane j one row behind lane j-1 (the systolic order of the packed kernel), strip
// after strip.  Tile (i, j) needs (i-1, j) and (i, j-1): the tiles of one anti-diagonal i + j = d
// of ALL pairs of a wave form one launch, launches follow in stream order -- no flags, no spinning.
// Tile edges live in global memory: a rolling row edge per pair (H', F' per column), a column edge
// per tile column (H', E' per row, kept: it is also the traceback checkpoint), a row-edge
// checkpoint every Mr rows.
//
// Traceback: the forward pass stores NO per-cell bits (100 kbp x 100 kbp x 4 bits = 5 GB per pair).
// nw_long_back, one warp per pair, goes from the end cell to the origin region by region (a region
// = the part of a (Mr rows x tile column) block above/left of the walk): recomputes it WITHOUT
// the bonus from the checkpointed edges (edge values & ~3 are exactly the bonus-free values) with
// the four tie bits per cell into a per-warp block, then lane 0 walks through it.  About
// (n1 + n2) * Mr / 2 cells are computed twice: 2 % at 100 kbp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "nw_affine_s16.cuh"
#include "nw_walk.cuh"

namespace sa {

constexpr int kLongK = 16;                  // columns per lane
constexpr int kLongStrip = 32 * kLongK;     // columns per warp strip
constexpr uint32_t kLongMr = 2048;          // rows between row-edge checkpoints
constexpr int kLongWarps = 4;               // warps per CTA (independent tiles / pairs)

struct LongScheme {       // transformed units (all magnitudes > 0 except chain and sent)
  int32_t cm;             // 4*match - 8*ext: a diagonal step on a match
  int32_t pen;            // 4*(match - mismatch)
  int32_t open;           // -4*open
  int32_t ext4;           // -4*ext: what a fixed score gains per step of x + y
  int32_t chain;          // 4*(open + ext): D'[0][y] = I'[x][0] for all x, y >= 1 (:194-198, :206-210)
  int32_t sent;           // 4 * -32768
  int32_t one;            // 1, as a value ptxas cannot see: multiplier of the adds that go to the fma pipe (IMAD)
};

__host__ __device__ inline LongScheme make_long_scheme(int match, int mismatch, int open, int ext) {
  LongScheme s;
  s.cm = 4 * match - 8 * ext;
  s.pen = 4 * (match - mismatch);
  s.open = -4 * open;
  s.ext4 = -4 * ext;
  s.chain = 4 * (open + ext);
  s.sent = 4 * -32768;
  s.one = 1;
  return s;
}

// Border cells in V' (z = x + y, one of them 0): best state H' and the gap candidate handed to the
// first interior cell (E' for column 0, F' for row 0: "open from the sentinel M, or extend the
// chain"; extensions are free).  B: bonus on.
template <bool B>
__device__ __forceinline__ int32_t long_edge_h(const LongScheme& sc, uint32_t z) {
  const int32_t s = sc.sent + sc.ext4 * (int32_t)z;
  return max(s + (B ? 2 : 0), sc.chain + (B ? 1 : 0));
}
template <bool B>
__device__ __forceinline__ int32_t long_edge_g(const LongScheme& sc, uint32_t z) {
  const int32_t s = sc.sent + sc.ext4 * (int32_t)z;
  return max(s + (B ? 2 : 0) - sc.open, sc.chain + (B ? 1 : 0));
}

struct LongParams {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  uint32_t packing;
  const uint32_t* __restrict__ ids;   // launch index k -> pair id
  uint32_t n_ids;
  LongScheme sc;
  int32_t ext;                        // raw, for the score
  uint32_t R, S;                      // tile: R rows x S strips of 512 columns
  uint32_t diag;                      // forward: the tile anti-diagonal of this launch
  int2* __restrict__ edges;           // all edge storage of the wave
  const uint64_t* __restrict__ row_off;   // per k (int2 units): rolling row edge, n1pad entries
  const uint64_t* __restrict__ col_off;   // per k: (TC - 1) column edges of n2 + 1 entries
  const uint64_t* __restrict__ ck_off;    // per k: (n2 - 1) / Mr checkpoint rows of n1pad entries
  int32_t* __restrict__ end_h;        // per k: H' of the end cell (forward)
  uint8_t* __restrict__ flag;         // per k: 0 clean, 1 chain-tainted, 2 handed to the literal kernel
  uint32_t* __restrict__ fb_count;    // fallback queue: pair ids
  uint32_t* __restrict__ fb_ids;
  // backward
  uint2* __restrict__ tb;             // per backward warp: S * 32 * Mr words
  uint32_t* __restrict__ next_back;   // hand-out counter
  uint32_t* __restrict__ runs;        // runs of pair k written back to front, ending at runs_end[k]
  const uint64_t* __restrict__ runs_end;
  int want_runs;
  int32_t* __restrict__ score;
  uint8_t* __restrict__ status;
  uint32_t* __restrict__ cigar_len;
};

// per-warp shared memory: [left ring 64 int2][right ring 64 int2][tmpcol (rows_cap + 1) int2][panel rows_cap bytes]
__host__ __device__ inline uint32_t long_smem_per_warp(uint32_t rows_cap, bool tmpcol) {
  return 1024u + (tmpcol ? (rows_cap + 2u) * 8u : 0u) + ((rows_cap + 15u) & ~15u);
}

// max(a, b) that also sets `bit` in acc when the FIRST operand wins or ties: VIMNMX + ISETP + one
// predicated add (ptxas folds the compare into the VIMNMX only for the packed 16-bit form).
__device__ __forceinline__ int32_t long_max_tie(int32_t a, int32_t b, uint32_t& acc, uint32_t bit) {
  int32_t r;
  asm("{\n\t.reg .pred p;\n\tmax.s32 %0, %2, %3;\n\tsetp.eq.s32 p, %0, %2;\n\t@p add.u32 %1, %1, %4;\n\t}"
      : "=r"(r), "+r"(acc)
      : "r"(a), "r"(b), "r"(bit));
  return r;
}

// a + b as an IMAD (fma pipe) instead of an IADD3 (alu pipe, where every VIMNMX of the cell already sits)
__device__ __forceinline__ int32_t fma_pipe_add(int32_t a, int32_t one, int32_t b) {
  int32_t r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
  return r;
}

// TB: with tie bits, bonus off.  CAP: also keep M' and the incoming I' of column cap_c.
// CELL (score-only forward cell): 0 = LOP3, VIMNMX, IADD3, VIMNMX3, 2 x VIADDMNMX: 6 issue slots, but the fused
// three-input ops hold the alu pipe for two slots each (int_peak: 3.2 against 6.4 per clock and SM): 9 alu slots per
// cell.  1 = the two gap candidates from M - open computed once on the fma pipe: LOP3, VIMNMX, IADD3, VIMNMX3, IMAD,
// 2 x VIMNMX: 7 issue slots, 7 alu.  2 = the diagonal add on the fma pipe as well: 8 issue slots, 6 alu.
// Measured, 1 000 x 100 kbp: 0: 4 821 ms, 1: 4 912 ms, 2: 4 670 ms (the default).
template <bool TB, bool CAP, int CELL = 0>
__device__ __forceinline__ void long_cells(int32_t (&H)[kLongK], int32_t (&F)[kLongK], const uint32_t (&q)[kLongK],
                                           uint32_t d, int32_t hdiag, int32_t& E, const LongScheme& sc,
                                           uint32_t& acc0, uint32_t& acc1, int cap_c, int32_t& capM, int32_t& capE) {
#pragma unroll
  for (int c = 0; c < kLongK; ++c) {
    const int32_t hup = H[c];
    const int32_t m = (int32_t)min(q[c] ^ d, (uint32_t)sc.pen);  // residues sit in bits 16+: unequal -> >= 65536 > pen
    const int32_t M = (!TB && CELL == 2) ? fma_pipe_add(hdiag, sc.one, sc.cm) - m : hdiag + sc.cm - m;
    if (CAP && c == cap_c) {
      capM = M;
      capE = E;
    }
    if (TB) {
      uint32_t& acc = c < 8 ? acc0 : acc1;
      const uint32_t sh = 4u * (c & 7);
      const int32_t t = long_max_tie(E, M, acc, 1u << sh);          // I >= M
      const int32_t Hn = long_max_tie(F[c], t, acc, 2u << sh);      // D >= max(I, M)
      const int32_t Mo = M - sc.open;
      E = long_max_tie(Mo, E, acc, 4u << sh);                       // opening ties or wins: I'[x][y+1]
      F[c] = long_max_tie(Mo, F[c], acc, 8u << sh);                 // opening ties or wins: D'[x+1][y]
      H[c] = Hn;
    } else if (CELL == 0) {
      const int32_t Hn = __vimax3_s32(F[c], E, M);
      E = __viaddmax_s32(M, -sc.open, E);
      F[c] = __viaddmax_s32(M, -sc.open, F[c]);
      H[c] = Hn;
    } else {
      const int32_t Hn = __vimax3_s32(F[c], E, M);
      const int32_t Mo = fma_pipe_add(M, sc.one, -sc.open);
      E = max(Mo, E);
      F[c] = max(Mo, F[c]);
      H[c] = Hn;
    }
    hdiag = hup;
  }
}

// One strip: columns col0+1 .. col0+512 (lane l owns 16 of them), rows row0+1 .. row0+nrows, by one
// warp.  top: (H', F') per column at row row0, indexed by column - 1, or nullptr for row 0 (closed
// form).  left: (H', E') of column col0 for rows row0 .. row0+nrows (entry 0 = the corner), global
// or shared, or nullptr for column 0.  bottom/bottom2: where the (H', F') of the last row go.
// right: the strip's own right edge in the same format as left (may BE left: the reads run a
// chunk ahead of the writes); right[0] is written only when write_corner.  tbs (TB): tie-bit
// words [lane][tb_stride rows].  Returns the end-cell record when the strip holds (n2, n1):
// forward: H'; TB: the start state of the traceback.
template <bool TB, int CELL = 0>
__device__ __forceinline__ int32_t long_strip(const LongParams& p, uint32_t n1, uint32_t n2, uint64_t qo, uint64_t dof,
                                              uint32_t col0, uint32_t row0, uint32_t nrows, const int2* __restrict__ top,
                                              const int2* left, int2* bottom, int2* bottom2, int2* right, bool write_corner,
                                              uint2* tbs, uint32_t tb_stride, int2* lring, int2* rring, uint8_t* panel) {
  const LongScheme sc = p.sc;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t y0 = col0 + lane * kLongK;  // this lane's columns are y0+1 .. y0+16
  constexpr bool B = !TB;                     // the forward pass carries the provenance bonus
  int32_t H[kLongK], F[kLongK];
  uint32_t q[kLongK];
#pragma unroll
  for (int c = 0; c < kLongK; ++c) {
    const uint32_t y = y0 + c + 1;
    if (top) {
      const int2 v = __ldcg(&top[y - 1]);
      H[c] = TB ? (v.x & ~3) : v.x;
      F[c] = TB ? (v.y & ~3) : v.y;
    } else {
      H[c] = long_edge_h<B>(sc, y);
      F[c] = long_edge_g<B>(sc, y);
    }
    q[c] = y <= n1 ? load_residue(p.residues, qo + y - 1, p.packing) << 16 : 0xffffffffu;
  }
  for (uint32_t r = lane; r < nrows; r += 32) panel[r] = (uint8_t)load_residue(p.residues, dof + row0 + r, p.packing);
  // the cell (row0, y0): diagonal input of this lane's first column in its first row
  int32_t hd = __shfl_up_sync(0xffffffffu, H[kLongK - 1], 1);
  int2 pre = make_int2(0, 0);
  if (left) {
    if (lane == 0) {
      const int2 v = left[0];
      hd = TB ? (v.x & ~3) : v.x;
    }
    if (lane < nrows) pre = left[1 + lane];  // rows row0+1 .. row0+32
  } else if (lane == 0) {
    hd = (row0 == 0) ? 0 : long_edge_h<B>(sc, row0);  // M[0][0] = 0 is the origin (:172)
  }
  __syncwarp();
  if (left) lring[lane] = pre;
  if (right && write_corner && lane == 31) right[0] = make_int2(H[kLongK - 1], 0);
  const int cap_c = (row0 + nrows == n2 && n1 > y0 && n1 <= y0 + kLongK) ? (int)(n1 - y0 - 1) : -1;
  int32_t capM = 0, capE = 0, cap_bits = 0;
  int32_t out_h = 0, out_e = 0;
  uint2* tbp = TB ? tbs + (uint64_t)lane * tb_stride : nullptr;
  __syncwarp();

  // One row of the lane's 16 columns (relative row xr, step t).
  auto row = [&](const uint32_t t, const uint32_t xr, const int32_t rh_n, const int32_t re_n, auto cap_possible) {
    int32_t rh = rh_n, re = re_n;
    if (lane == 0) {
      if (left) {
        const int2 v = lring[(t - 1) & 63u];
        rh = TB ? (v.x & ~3) : v.x;
        re = TB ? (v.y & ~3) : v.y;
      } else {
        rh = long_edge_h<B>(sc, row0 + xr);
        re = long_edge_g<B>(sc, row0 + xr);
      }
    }
    const uint32_t d = (uint32_t)panel[xr - 1] << 16;
    int32_t E = re;
    uint32_t acc0 = 0, acc1 = 0;
    if (TB && decltype(cap_possible)::value && (int)xr == (cap_c >= 0 ? (int)nrows : -1)) {
      long_cells<TB, true, CELL>(H, F, q, d, hd, E, sc, acc0, acc1, cap_c, capM, capE);
      cap_bits = (int32_t)(((cap_c < 8 ? acc0 : acc1) >> (4 * (cap_c & 7))) & 15u);
    } else {
      long_cells<TB, false, CELL>(H, F, q, d, hd, E, sc, acc0, acc1, -1, capM, capE);
    }
    hd = rh;
    out_h = H[kLongK - 1];
    out_e = E;
    if (right && lane == 31) rring[(xr - 1) & 63u] = make_int2(out_h, out_e);
    if (TB) tbp[xr - 1] = make_uint2(acc0, acc1);
  };
  // the right edge leaves in whole chunks of 32 rows (coalesced): rows base+1 .. min(base+32, upto)
  auto flush_right = [&](const uint32_t base, const uint32_t upto) {
    __syncwarp();
    if (base + lane < upto) right[base + 1 + lane] = rring[(base + lane) & 63u];
    __syncwarp();
  };
  // One step with every periodic duty checked: the ramps at both ends of the strip, where not every
  // lane has a row.  Lane l computes row t - l.
  auto slow_step = [&](const uint32_t t) {
    if (left) {  // left edge, one 32-row chunk ahead: loaded at the chunk's first step, parked at its 17th
      const uint32_t ph = (t - 1) & 31u;
      if (ph == 0) {
        const uint32_t r = t + 32 + lane;  // row (relative) this lane fetches for the next chunk
        if (r <= nrows) pre = left[r];
      } else if (ph == 16) {
        lring[((((t - 1) >> 5) + 1) & 1u) * 32 + lane] = pre;
        __syncwarp();
      }
    }
    const int32_t rh_n = __shfl_up_sync(0xffffffffu, out_h, 1);
    const int32_t re_n = __shfl_up_sync(0xffffffffu, out_e, 1);
    const uint32_t xr = t - lane;  // (wraps when t < lane)
    if (xr >= 1 && xr <= nrows) row(t, xr, rh_n, re_n, std::true_type{});
    if (right) {
      const uint32_t x31 = t - 31;  // lane 31's relative row at this step
      if (t >= 32 && ((x31 & 31u) == 0 || x31 == nrows)) flush_right((x31 - 1) & ~31u, x31);
    }
  };
  const uint32_t steps = nrows + 31;
  uint32_t t = 1;
  for (; t <= 32 && t <= steps; ++t) slow_step(t);
  // Steady state in chunks of 32 steps (t = 32c+1 .. 32c+32, all lanes inside the strip, the end
  // cell not among the rows): the periodic duties sit between two branch-free runs of 16 steps.
  while (t + 31 < nrows) {
    if (left) {
      const uint32_t r = t + 32 + lane;
      if (r <= nrows) pre = left[r];
    }
#pragma unroll 2
    for (uint32_t i = 0; i < 16; ++i, ++t) {
      const int32_t rh_n = __shfl_up_sync(0xffffffffu, out_h, 1);
      const int32_t re_n = __shfl_up_sync(0xffffffffu, out_e, 1);
      row(t, t - lane, rh_n, re_n, std::false_type{});
    }
    if (left) {
      lring[((((t - 1) >> 5) + 1) & 1u) * 32 + lane] = pre;
      __syncwarp();
    }
#pragma unroll 2
    for (uint32_t i = 0; i < 16; ++i, ++t) {
      const int32_t rh_n = __shfl_up_sync(0xffffffffu, out_h, 1);
      const int32_t re_n = __shfl_up_sync(0xffffffffu, out_e, 1);
      row(t, t - lane, rh_n, re_n, std::false_type{});
    }
    // lane 31 has now finished row t - 32: the chunk of rows before this one is complete
    if (right) flush_right(t - 1 - 64, t - 1 - 32);
  }
  for (; t <= steps; ++t) slow_step(t);
  __syncwarp();
  if (bottom || bottom2) {
#pragma unroll
    for (int c = 0; c < kLongK; ++c) {
      const int2 v = make_int2(H[c], F[c]);
      if (bottom) __stcg(&bottom[y0 + c], v);
      if (bottom2) __stcg(&bottom2[y0 + c], v);
    }
  }
  int32_t ret = 0;
  if (cap_c >= 0) {
    if (TB) {
      // start state (:251-280: pushed I, M, D; popped D, M, I): D if D == max, else M if M >= I, else I
      ret = (cap_bits & 2) ? 2 : (capM >= capE ? 0 : 1);
    } else {
#pragma unroll
      for (int c = 0; c < kLongK; ++c)
        if (c == cap_c) ret = H[c];
    }
  }
  return ret;
}

// tiles per dimension of pair (n1, n2)
__device__ __forceinline__ uint32_t long_tc(uint32_t n1, uint32_t S) { return (n1 + S * kLongStrip - 1) / (S * kLongStrip); }
__device__ __forceinline__ uint32_t long_n1pad(uint32_t n1) { return (n1 + kLongStrip - 1) / kLongStrip * kLongStrip; }

// Forward pass, one launch per tile anti-diagonal: grid (tiles / 4, pairs), 4 warps per CTA.
// MINB: resident CTAs per SM the register allocation is held to (4 -> 128 registers, 5 -> 102).
template <int MINB, int CELL = 0>
__global__ void __launch_bounds__(32 * kLongWarps, MINB) nw_long_fwd(const LongParams p) {
  extern __shared__ __align__(16) uint8_t long_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t k = blockIdx.y;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint32_t TC = long_tc(n1, p.S), TR = (n2 + p.R - 1) / p.R;
  const uint32_t d = p.diag;
  const uint32_t i_lo = d >= TC ? d - (TC - 1) : 0u, i_hi = min(d, TR - 1);
  const uint32_t idx = blockIdx.x * kLongWarps + warp;
  if (i_lo > i_hi || idx > i_hi - i_lo) return;  // (whole warps leave; the CTA has no barriers)
  const uint32_t i = i_lo + idx, j = d - i;
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const bool tmp = p.S > 1;
  uint8_t* wsm = long_smem + (size_t)warp * long_smem_per_warp(p.R, tmp);
  int2* lring = reinterpret_cast<int2*>(wsm);
  int2* rring = lring + 64;
  int2* tmpcol = rring + 64;
  uint8_t* panel = wsm + 1024 + (tmp ? (p.R + 2u) * 8u : 0u);

  const uint32_t n1pad = long_n1pad(n1);
  int2* rowedge = p.edges + p.row_off[k];
  int2* coledge = p.edges + p.col_off[k];
  const uint32_t row0 = i * p.R, nrows = min(p.R, n2 - row0);
  const uint32_t last_row = row0 + nrows;
  int2* ck = (last_row % kLongMr == 0 && last_row < n2) ? p.edges + p.ck_off[k] + (uint64_t)(last_row / kLongMr - 1) * n1pad : nullptr;
  const uint32_t c_first = j * p.S * kLongStrip;
  int32_t endh = 0;
  bool have_end = false;
  for (uint32_t s = 0; s < p.S; ++s) {
    const uint32_t col0 = c_first + s * kLongStrip;
    if (col0 >= n1) break;
    const bool last_strip = (s + 1 == p.S) || (col0 + kLongStrip >= n1);
    const int2* left = s ? tmpcol : (j ? coledge + (uint64_t)(j - 1) * (n2 + 1) + row0 : nullptr);
    int2* right = last_strip ? ((j + 1 < TC) ? coledge + (uint64_t)j * (n2 + 1) + row0 : nullptr) : tmpcol;
    const int32_t r = long_strip<false, CELL>(p, n1, n2, qo, dof, col0, row0, nrows, i ? rowedge : nullptr, left, rowedge, ck,
                                        right, /*write_corner=*/!last_strip || row0 == 0, nullptr, 0, lring, rring, panel);
    if (last_row == n2 && n1 > col0 && n1 <= col0 + kLongStrip) {
      const uint32_t owner = (n1 - 1 - col0) / kLongK;
      endh = __shfl_sync(0xffffffffu, r, owner);
      have_end = true;
    }
  }
  if (have_end && lane == 0) p.end_h[k] = endh;
}

// After the last forward launch: score, provenance class, fallback queue.
__global__ void __launch_bounds__(128) nw_long_classify(const LongParams p) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_ids) return;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const int32_t h = p.end_h[k];
  const int32_t b = h & 3;
  // V = (V' - b) / 4 + ext * (x + y)
  p.score[id] = (h - b) / 4 + p.ext * (int32_t)(n1 + n2);
  p.cigar_len[id] = 0;
  if (b >= 2) {
    const uint32_t slot = atomicAdd(p.fb_count, 1u);
    p.fb_ids[slot] = id;
    p.flag[k] = 2;
    p.status[id] = kRefNoOutput;  // placeholder; the literal kernel decides
  } else {
    p.flag[k] = (uint8_t)b;
    p.status[id] = b ? kRefPanic : kOk;  // b == 1: refined by the walk (REF_PANIC / REF_PANIC_EARLY)
  }
}

// Backward pass: one warp per pair, regions from the end cell to the origin.
__global__ void __launch_bounds__(32 * kLongWarps) nw_long_back(const LongParams p) {
  extern __shared__ __align__(16) uint8_t long_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint8_t* wsm = long_smem + (size_t)warp * long_smem_per_warp(kLongMr, true);
  int2* lring = reinterpret_cast<int2*>(wsm);
  int2* rring = lring + 64;
  int2* tmpcol = rring + 64;
  uint8_t* panel = wsm + 1024 + (kLongMr + 2u) * 8u;
  const uint32_t gw = blockIdx.x * kLongWarps + warp;
  uint2* tb = p.tb + (uint64_t)gw * p.S * 32u * kLongMr;
  const uint32_t Wc = p.S * kLongStrip;

  for (;;) {
    uint32_t k = 0;
    if (lane == 0) k = atomicAdd(p.next_back, 1u);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= p.n_ids) break;
    const uint32_t fl = p.flag[k];
    if (fl >= 2 || (fl == 0 && !p.want_runs)) continue;
    const uint32_t id = p.ids[k];
    const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
    const uint64_t qo = p.q_off[id], dof = p.d_off[id];
    const uint32_t n1pad = long_n1pad(n1);
    const int2* coledge = p.edges + p.col_off[k];
    const int2* ckrows = p.edges + p.ck_off[k];
    uint32_t x = n2, y = n1;
    int st = -1, pend = -1;
    uint32_t nruns = 0, run_op = 3, run_len = 0;
    uint32_t* out = (p.want_runs && p.runs) ? p.runs + p.runs_end[k] : nullptr;
    while (x > 0 && y > 0) {
      const uint32_t m = (x - 1) / kLongMr, r0 = m * kLongMr, nrows = x - r0;
      const uint32_t j = (y - 1) / Wc, c0 = j * Wc;
      const uint32_t nstrips = (y - c0 + kLongStrip - 1) / kLongStrip;
      int32_t cap = 0;
      for (uint32_t s = 0; s < nstrips; ++s) {
        const uint32_t col0 = c0 + s * kLongStrip;
        const int2* left = s ? tmpcol : (j ? coledge + (uint64_t)(j - 1) * (n2 + 1) + r0 : nullptr);
        int2* right = (s + 1 < nstrips) ? tmpcol : nullptr;
        const int32_t r = long_strip<true>(p, n1, n2, qo, dof, col0, r0, nrows, m ? ckrows + (uint64_t)(m - 1) * n1pad : nullptr,
                                           left, nullptr, nullptr, right, true, tb + (uint64_t)s * 32u * kLongMr, kLongMr, lring,
                                           rring, panel);
        if (st < 0 && x == n2 && n1 > col0 && n1 <= col0 + kLongStrip) cap = __shfl_sync(0xffffffffu, r, (n1 - 1 - col0) / kLongK);
      }
      __syncwarp();
      if (lane == 0) {
        auto nib = [&](uint32_t xx, uint32_t yy) -> uint32_t {
          const uint32_t cc = yy - 1 - c0;
          const uint2 w = __ldcg(&tb[((uint64_t)(cc / kLongStrip) * 32u + (cc % kLongStrip) / kLongK) * kLongMr + (xx - 1 - r0)]);
          return (((cc & 8u) ? w.y : w.x) >> (4u * (cc & 7u))) & 15u;
        };
        auto lut = [](int from, uint32_t nb) -> int {
          const uint32_t l = from == 0 ? kLutM : (from == 1 ? kLutI : kLutD);
          return (int)((l >> (2 * nb)) & 3u);
        };
        if (st < 0) st = cap;
        else if (pend >= 0) st = lut(pend, nib(x, y));
        pend = -1;
        for (;;) {
          if ((uint32_t)st != run_op) {
            if (run_len) {
              ++nruns;
              if (out) *--out = (run_len << 2) | run_op;
            }
            run_op = (uint32_t)st;
            run_len = 0;
          }
          ++run_len;
          x -= (st != 1);   // M and D consume a db residue
          y -= (st != 2);   // M and I consume a query residue
          if (x == 0 || y == 0) break;
          if (x <= r0 || y <= c0) {  // the cell that decides the next state lies in the next region
            pend = st;
            break;
          }
          st = lut(st, nib(x, y));
        }
      }
      x = __shfl_sync(0xffffffffu, x, 0);
      y = __shfl_sync(0xffffffffu, y, 0);
      st = __shfl_sync(0xffffffffu, st, 0);
    }
    if (lane == 0) {
      const bool complete = x == 0 && y == 0;
      if (run_len) {
        ++nruns;
        if (out) *--out = (run_len << 2) | run_op;
      }
      // an untainted pair cannot run into the boundary chain (that path would carry the bonus)
      p.status[id] = complete ? (fl ? kRefPanic : kOk) : kRefPanicEarly;
      p.cigar_len[id] = (complete && p.want_runs) ? nruns : 0u;
    }
    __syncwarp();
  }
}

}  // namespace sa
