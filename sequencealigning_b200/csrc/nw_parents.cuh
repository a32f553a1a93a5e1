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

}  // namespace sa
