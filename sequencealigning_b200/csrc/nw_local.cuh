// nw_local.cuh -- LOCAL mode of the single-matrix ("linear / pseudo-affine") aligner, sm_100a.
//
// The only non-global mode the reference implements (/root/reference/src/needleman_wunsch.rs):
//   fill  :66-103  no border initialisation (:43 `if !local`): row 0 / column 0 are score 0, NoGap,
//                  empty move lists;
//                     diag  = S[i-1][j-1] + (seq1[i-1] == seq2[j-1] ? match : mismatch)
//                     down  = S[i-1][j]   + (gaps[i-1][j] ? ext : open)
//                     right = S[i][j-1]   + (gaps[i][j-1] ? ext : open)
//                     mx = max3;  gaps[i][j] = (mx == down || mx == right)            :85-87
//                     mx < 0: the score stays 0 and the move list stays EMPTY          :88-89
//                     else  : S = mx, moves pushed Down, Right, Diag for every tie     :91-100
//   start :107-111, :256-272  EVERY cell holding the matrix maximum, in row-major order;
//   hit   :205-254  pre-order recursion over the stored moves; a hit is printed at (0,0) or at a
//                  cell with no moves.
// Rows i walk seq1 (query), columns j walk seq2 (db) (:38).  Per pair the kernel returns what the
// FIRST printed hit is made of: the maximum, the first start cell in row-major order (end1, end2)
// and the path that follows the first stored move of every cell (Down > Right > Diag) until a
// cell without moves, as a run-length CIGAR (Down = SA_OP_I, Right = SA_OP_D, Diag = SA_OP_M).
//
// Mapping: one WARP per pair, pairs handed out by an atomic counter (persistent warps).  Lane L owns
// K consecutive columns per pass (in registers: score and the gap cost the next cell pays) and runs
// one row behind lane L-1; the strip's right edge goes to lane L+1 by __shfl_up_sync, and from lane
// 31 to a per-warp boundary column for the next pass of 32*K columns.  The clamp at 0 is absolute, so
// the packed kernels' "per-cell offset cancels" transform does not apply: plain 32-bit integers.
// Two traceback bits per cell (0 Diag, 1 Down, 2 Right, 3 no moves), one word per (row, strip), in
// shared memory when a pair's matrix fits the warp's share, else in the warp's global scratch; the
// warp's lane 0 walks them as soon as the fill ends, so the words never leave the SM / L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_affine_s16.cuh"  // load_residue

namespace sa {

constexpr int kLocalWarps = 4;  // warps per block

struct LocalParams {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  uint32_t packing;
  uint32_t pair_base, n_launch_pairs;
  int32_t match, mismatch, open, ext;
  uint32_t* next_pair;        // atomic hand-out counter (zeroed by the host)
  uint32_t smem_words;        // traceback words (of the form's width) each warp has in shared memory
  uint8_t* tb;                // global traceback scratch, tb_stride bytes per warp of the grid
  uint64_t tb_stride;
  int2* bnd;                  // boundary column per warp: (S, gap cost) of column y0 for every row
  uint64_t bnd_stride;        // entries per warp (>= longest query + 1)
  uint32_t* runs;             // CIGAR staging: pair t of the launch writes backwards from runs_end[t]
  const uint64_t* runs_end;   // (nullptr: no traceback wanted)
  int32_t* score;
  uint8_t* status;
  uint32_t* cigar_len;
  uint32_t *end1, *end2;      // start cell of the traceback = END of the local alignment (may be null)
  uint32_t omit_flag;         // ORed into the status when the traceback was skipped for lack of scratch
};

template <int K>
struct LocalWord {
  using type = uint32_t;
};
template <>
struct LocalWord<5> {
  using type = uint16_t;
};
template <>
struct LocalWord<8> {
  using type = uint16_t;
};

template <int K>
__global__ void __launch_bounds__(32 * kLocalWarps) nw_linear_local_kernel(const LocalParams p) {
  using word_t = typename LocalWord<K>::type;
  extern __shared__ uint32_t local_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * kLocalWarps + warp;  // warp of the grid: owns one scratch slice
  word_t* const tb_sh = reinterpret_cast<word_t*>(local_smem) + (size_t)warp * p.smem_words;
  word_t* const tb_gl = reinterpret_cast<word_t*>(p.tb + (uint64_t)gw * p.tb_stride);
  int2* const bnd = p.bnd + (uint64_t)gw * p.bnd_stride;
  const int32_t match = p.match, mismatch = p.mismatch, open = p.open, ext = p.ext;

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(p.next_pair, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= p.n_launch_pairs) return;
    const uint32_t id = p.pair_base + t;
    const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
    const uint64_t qo = p.q_off[id], dof = p.d_off[id];
    const uint32_t ns = (n2 + K - 1) / K;  // strips = traceback words per row
    const bool want_tb = p.runs_end != nullptr;
    word_t* const tb = ((uint64_t)n1 * ns <= p.smem_words) ? tb_sh : tb_gl;

    // best cell of this lane's strips: larger score, then smaller row (a later pass restarts at row 1),
    // then smaller column.  (0, 0) with score 0 is the first cell of the row-major order (:256-272).
    int32_t best = 0;
    uint32_t bi = 0, bj = 0;

    const uint32_t npass = (ns + 31) / 32;
    for (uint32_t pass = 0; pass < npass && n1; ++pass) {
      const uint32_t s = pass * 32 + lane;  // this lane's strip
      const uint32_t y0 = s * K;            // columns to the left of it
      const int kv = (int)min((uint32_t)K, n2 > y0 ? n2 - y0 : 0u);  // columns of the strip inside the pair
      int32_t H[K], G[K];
      uint32_t q[K];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        H[c] = 0;      // row 0: score 0 ..
        G[c] = open;   // .. NoGap: the cell below pays an opening
        q[c] = (c < kv) ? load_residue(p.residues, dof + y0 + c, p.packing) : 0x100u;  // never equal to a residue
      }
      int32_t hd = 0;              // S[i-1][y0]
      int32_t out_s = 0, out_g = open;
      const uint32_t steps = n1 + 31;
      for (uint32_t step = 1; step <= steps; ++step) {
        int32_t ls = __shfl_up_sync(0xffffffffu, out_s, 1);
        int32_t lg = __shfl_up_sync(0xffffffffu, out_g, 1);
        const uint32_t i = step - (uint32_t)lane;  // row of this lane (1-based)
        if (i >= 1 && i <= n1 && kv > 0) {
          if (lane == 0) {
            if (pass == 0) {
              ls = 0;      // column 0: score 0, NoGap
              lg = open;
            } else {
              const int2 b = bnd[i];
              ls = b.x;
              lg = b.y;
            }
          }
          const uint32_t r = load_residue(p.residues, qo + i - 1, p.packing);
          const int32_t next_hd = ls;
          int32_t d = hd, sl = ls, gl = lg, rowmax = 0;
          uint32_t word = 0;
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const int32_t up = H[c];
            const int32_t diag = d + (q[c] == r ? match : mismatch);
            const int32_t down = up + G[c];
            const int32_t right = sl + gl;
            const int32_t tmax = max(down, right);
            const int32_t mx = max(tmax, diag);
            const int32_t g = (mx == tmax) ? ext : open;  // gaps[i][j]: set even when diag ties, and when mx < 0
            uint32_t code = (mx == down) ? 1u : ((mx == right) ? 2u : 0u);
            code = mx < 0 ? 3u : code;
            word |= code << (2 * c);
            const int32_t S = max(mx, 0);
            H[c] = S;
            G[c] = g;
            d = up;
            sl = S;
            gl = g;
            if (c < kv) rowmax = max(rowmax, S);
          }
          hd = next_hd;
          out_s = sl;
          out_g = gl;
          if (lane == 31) bnd[i] = make_int2(sl, gl);  // left edge of the next pass's lane 0
          if (want_tb) tb[(uint64_t)(i - 1) * ns + s] = (word_t)word;
          if (rowmax > best || (rowmax == best && i < bi)) {
            best = rowmax;
            bi = i;
#pragma unroll
            for (int c = K - 1; c >= 0; --c)
              if (c < kv && H[c] == rowmax) bj = y0 + c + 1;
          }
        }
      }
      __syncwarp();
    }
    // first maximum in row-major order over the lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ob > best || (ob == best && (oi < bi || (oi == bi && oj < bj)))) {
        best = ob;
        bi = oi;
        bj = oj;
      }
    }
    __syncwarp();
    if (lane == 0) {
      uint32_t nruns = 0;
      if (want_tb) {
        // get_next (:205-254) along the first stored move of every cell; border cells hold no moves
        uint32_t i = bi, j = bj, run_op = 3, run_len = 0;
        uint32_t* out = p.runs + p.runs_end[t];
        while (i > 0 && j > 0) {
          const uint32_t w = tb[(uint64_t)(i - 1) * ns + (j - 1) / K];
          const uint32_t code = (w >> (2 * ((j - 1) % K))) & 3u;
          if (code == 3u) break;
          uint32_t op;
          if (code == 1u) { op = 1; --i; }        // Down: seq1[i-1] over '-'
          else if (code == 2u) { op = 2; --j; }   // Right: '-' over seq2[j-1]
          else { op = 0; --i; --j; }
          if (op != run_op) {
            if (run_len) *--out = (run_len << 2) | run_op;
            run_op = op;
            run_len = 0;
            ++nruns;
          }
          ++run_len;
        }
        if (run_len) *--out = (run_len << 2) | run_op;
      }
      p.score[id] = best;
      p.status[id] = (uint8_t)p.omit_flag;  // SA_OK: nothing in this aligner panics
      p.cigar_len[id] = nruns;
      if (p.end1) p.end1[id] = bi;
      if (p.end2) p.end2[id] = bj;
    }
    __syncwarp();
  }
}

// runs of launch pair t (staged backwards from runs_end[t]) -> pool at cigar_off; one warp per pair
__global__ void __launch_bounds__(128) local_runs_to_pool(uint32_t pair_base, uint32_t n, const uint32_t* __restrict__ runs,
                                                          const uint64_t* __restrict__ runs_end,
                                                          const uint32_t* __restrict__ cigar_len,
                                                          const uint64_t* __restrict__ cigar_off,
                                                          uint32_t* __restrict__ pool, uint64_t pool_cap) {
  const uint32_t t = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (t >= n) return;
  const uint32_t id = pair_base + t, len = cigar_len[id];
  const uint64_t off = cigar_off[id];
  const uint32_t* src = runs + runs_end[t] - len;
  for (uint32_t k = threadIdx.x & 31; k < len; k += 32)
    if (off + k < pool_cap) pool[off + k] = src[k];
}

}  // namespace sa
