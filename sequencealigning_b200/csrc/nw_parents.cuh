// nw_parents.cuh -- full parent sets of the affine aligner, for printing EVERY co-optimal
// alignment the way the reference does.
//
// The reference keeps, per cell and state, the LIST of parents that attain the maximum
// (/root/reference/src/needleman_wunsch_affine.rs:96-153) and its traceback is a LIFO DFS over
// those lists (:246-329).  This kernel runs the literal recurrences (:76-94, :169-237) in
// 32-bit integers with the finite -32768 sentinel, one thread per pair, and writes one byte per
// interior cell with the 7 "is a parent" bits in the reference's push order:
//     bit0 M<-M  bit1 M<-I  bit2 M<-D   (:122-151)
//     bit3 I<-I  bit4 I<-M              (:110-117)
//     bit5 D<-D  bit6 D<-M              (:98-105)
// plus the three end-cell scores.  It is an on-demand path (a CLI flag / one API call), not the
// batched hot path: the batched kernels return only the FIRST alignment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_affine_s16.cuh"

namespace sa {

struct ParentsParams {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  uint32_t n_pairs, packing;
  int32_t match, mismatch, open, ext;
  uint8_t* __restrict__ parents;           // per pair: [n2][n1] bytes at parents_off[p]
  const uint64_t* __restrict__ parents_off;
  int32_t* __restrict__ rows;              // scratch: per pair 6 * (n1max + 1) ints
  uint32_t row_stride;                     // n1max + 1
  int32_t* __restrict__ end_scores;        // per pair: M, I, D at (n2, n1)
};

__global__ void __launch_bounds__(64) nw_affine_parents_kernel(const ParentsParams p) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= p.n_pairs) return;
  constexpr int32_t kNegInf = -32768;  // i16::MIN as i32 (:174)
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const uint32_t w = p.row_stride;
  int32_t* base = p.rows + (uint64_t)id * 6 * w;
  int32_t *pm = base, *pi = base + w, *pd = base + 2 * w, *cm = base + 3 * w, *ci = base + 4 * w, *cd = base + 5 * w;
  uint8_t* par = p.parents + p.parents_off[id];
  pm[0] = 0;            // :172-182
  pi[0] = kNegInf;
  pd[0] = kNegInf;
  for (uint32_t y = 1; y <= n1; ++y) {  // :183-199
    pm[y] = kNegInf;
    pi[y] = kNegInf;
    pd[y] = ((int32_t)y + 1) * p.ext + p.open;
  }
  for (uint32_t x = 1; x <= n2; ++x) {
    const uint32_t b2 = load_residue(p.residues, dof + x - 1, p.packing);
    cm[0] = kNegInf;  // :200-216
    cd[0] = kNegInf;
    ci[0] = p.open + ((int32_t)x + 1) * p.ext;
    for (uint32_t y = 1; y <= n1; ++y) {  // :217-236
      const int32_t sub = load_residue(p.residues, qo + y - 1, p.packing) == b2 ? p.match : p.mismatch;
      const int32_t dm = pm[y - 1], di = pi[y - 1], dd = pd[y - 1];
      const int32_t mm = max(max(dm, di), dd) + sub;
      const int32_t ii = max(cm[y - 1] + p.open, ci[y - 1]) + p.ext;
      const int32_t dv = max(pm[y] + p.open, pd[y]) + p.ext;
      uint32_t bits = 0;
      if (mm == dm + sub) bits |= 1u;
      if (mm == di + sub) bits |= 2u;
      if (mm == dd + sub) bits |= 4u;
      if (ii == ci[y - 1] + p.ext) bits |= 8u;
      if (ii == cm[y - 1] + p.open + p.ext) bits |= 16u;
      if (dv == pd[y] + p.ext) bits |= 32u;
      if (dv == pm[y] + p.open + p.ext) bits |= 64u;
      cm[y] = mm;
      ci[y] = ii;
      cd[y] = dv;
      par[(uint64_t)(x - 1) * n1 + (y - 1)] = (uint8_t)bits;
    }
    int32_t* t;
    t = pm; pm = cm; cm = t;
    t = pi; pi = ci; ci = t;
    t = pd; pd = cd; cd = t;
  }
  p.end_scores[3 * id + 0] = pm[n1];
  p.end_scores[3 * id + 1] = pi[n1];
  p.end_scores[3 * id + 2] = pd[n1];
}

// ---------------------------------------------------------------------------------------------
// The same service for the single-matrix ("linear") aligner: the reference prints EVERY hit of every start cell
// (/root/reference/src/needleman_wunsch.rs:106-116, :205-254).  One thread runs the literal fill (:43-103) for one
// pair and writes, per cell of the (n1+1) x (n2+1) matrix, the score and the move set in push order
// (bit0 Down, bit1 Right, bit2 Diag; 0 = no moves); the host finds the start cells and walks the move sets.
// Rows i walk seq1 (query), columns j walk seq2 (db) (:38).  On demand, one pair per call.
// ---------------------------------------------------------------------------------------------
struct LinearMovesParams {
  const uint8_t* __restrict__ seq1;
  const uint8_t* __restrict__ seq2;
  uint32_t n1, n2;
  int32_t match, mismatch, open, ext;
  int32_t local;
  int32_t* __restrict__ scores;  // [(n1+1) * (n2+1)]
  uint8_t* __restrict__ moves;   // [(n1+1) * (n2+1)]
  uint8_t* __restrict__ gaps;    // scratch: two rows of n2 + 1 flags
};

__global__ void nw_linear_moves_kernel(const LinearMovesParams p) {
  if (blockIdx.x || threadIdx.x) return;
  const uint32_t n1 = p.n1, n2 = p.n2, w = n2 + 1;
  uint8_t* gprev = p.gaps;
  uint8_t* gcur = p.gaps + w;
  for (uint32_t j = 0; j <= n2; ++j) {  // row 0 (:44-54; scores[0][0] is initialised twice, :45-64)
    p.scores[j] = p.local ? 0 : (int32_t)j * p.ext + p.open + (j == 0 ? p.open : 0);
    p.moves[j] = p.local ? 0 : (uint8_t)(2u | (j == 0 ? 1u : 0u));
    gprev[j] = p.local ? 0 : 1;
  }
  for (uint32_t i = 1; i <= n1; ++i) {
    int32_t* row = p.scores + (uint64_t)i * w;
    const int32_t* up = row - w;
    uint8_t* mv = p.moves + (uint64_t)i * w;
    row[0] = p.local ? 0 : (int32_t)i * p.ext + p.open;  // column 0 (:55-64)
    mv[0] = p.local ? 0 : 1;
    gcur[0] = p.local ? 0 : 1;
    const uint8_t r1 = p.seq1[i - 1];
    for (uint32_t j = 1; j <= n2; ++j) {  // :66-103
      const int32_t diag = up[j - 1] + (r1 == p.seq2[j - 1] ? p.match : p.mismatch);
      const int32_t down = up[j] + (gprev[j] ? p.ext : p.open);
      const int32_t right = row[j - 1] + (gcur[j - 1] ? p.ext : p.open);
      const int32_t mx = max(max(down, right), diag);
      gcur[j] = (mx == down || mx == right) ? 1 : 0;  // :85-87, also when the score is then dropped
      if (p.local && mx < 0) {                         // :88-89
        row[j] = 0;
        mv[j] = 0;
      } else {
        row[j] = mx;
        mv[j] = (uint8_t)((mx == down ? 1u : 0u) | (mx == right ? 2u : 0u) | (mx == diag ? 4u : 0u));
      }
    }
    uint8_t* t = gprev; gprev = gcur; gcur = t;
  }
}

}  // namespace sa
