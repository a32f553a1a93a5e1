// nw_count.cuh -- number of co-optimal alignments per pair, batched.
//
// The reference prints EVERY co-optimal alignment (needleman_wunsch_affine.rs:246-329): its
// traceback is a LIFO depth-first walk over the per-cell parent lists built at :96-153.  The
// number of alignments it prints when nothing panics is the number of parent-list paths from
// the best end states (:247-280) to the origin (0,0), where any popped state prints (:283-286).
// Boundary-chain cells D[0][y>=1] / I[x>=1][0] panic when expanded (:299/:303) and sentinel
// boundary cells have no parents, so both contribute no complete path.
//
// One thread per pair runs the literal 32-bit recurrences (finite -32768 sentinel, boundary
// gaps with one extra extension, :169-237) over rolling rows and carries, per state of every
// cell, the saturating count of complete paths below it.  Rows live in global scratch laid out
// [column][thread], so neighbouring threads touch neighbouring words.  A side API, not the hot
// path (the parity tests compare it with the CPU restatement's path count).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_affine_s16.cuh"

namespace sa {

struct CountParams {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  uint32_t pair_base, n_pairs, packing;  // this launch handles pairs [pair_base, pair_base + n_pairs)
  int32_t match, mismatch, open, ext;
  int32_t* __restrict__ rows;   // [3][cols][n_pairs]  M, I, D of the row above / being overwritten
  int64_t* __restrict__ cnts;   // [3][cols][n_pairs]
  uint32_t cols;                // longest query of the launch + 1
  int64_t* __restrict__ out;    // indexed by pair id
};

constexpr int64_t kCountCap = INT64_MAX / 4;  // saturation value (documented in sa_engine.h)

__device__ __forceinline__ int64_t count_sat_add(int64_t a, int64_t b) {
  const int64_t r = a + b;
  return (r < a || r > kCountCap) ? kCountCap : r;
}

__global__ void __launch_bounds__(128) nw_affine_count_kernel(const CountParams p) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n_pairs) return;
  constexpr int32_t kNegInf = -32768;  // i16::MIN as i32 (:174)
  const uint32_t id = p.pair_base + t;
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const uint64_t T = p.n_pairs, plane = (uint64_t)p.cols * T;
  int32_t *rm = p.rows + t, *ri = rm + plane, *rd = ri + plane;
  int64_t *cm = p.cnts + t, *ci = cm + plane, *cd = ci + plane;

  // row x = 0 (:172-199): origin, then the boundary chain stored in D
  rm[0] = 0; ri[0] = kNegInf; rd[0] = kNegInf;
  cm[0] = 1; ci[0] = 1; cd[0] = 1;  // any state popped at (0,0) prints (:283)
  for (uint32_t y = 1; y <= n1; ++y) {
    rm[y * T] = kNegInf; ri[y * T] = kNegInf; rd[y * T] = ((int32_t)y + 1) * p.ext + p.open;
    cm[y * T] = 0; ci[y * T] = 0; cd[y * T] = 0;  // chain cells panic, sentinel cells have no parents
  }
  for (uint32_t x = 1; x <= n2; ++x) {
    const uint32_t b2 = load_residue(p.residues, dof + x - 1, p.packing);
    // column 0 (:200-216): boundary chain stored in I
    int32_t dM = rm[0], dI = ri[0], dD = rd[0];
    int64_t dcM = cm[0], dcI = ci[0], dcD = cd[0];
    int32_t lM = kNegInf, lI = p.open + ((int32_t)x + 1) * p.ext;
    int64_t lcM = 0, lcI = 0;
    rm[0] = lM; ri[0] = lI; rd[0] = kNegInf;
    cm[0] = 0; ci[0] = 0; cd[0] = 0;
    for (uint32_t y = 1; y <= n1; ++y) {
      const uint64_t o = (uint64_t)y * T;
      const int32_t uM = rm[o], uI = ri[o], uD = rd[o];
      const int64_t ucM = cm[o], ucI = ci[o], ucD = cd[o];
      const int32_t sub = (load_residue(p.residues, qo + y - 1, p.packing) == b2) ? p.match : p.mismatch;  // :220
      // M (:76-86, parents :120-153)
      const int32_t mm = max(max(dM, dI), dD) + sub;
      int64_t vm = 0;
      if (mm == dM + sub) vm = count_sat_add(vm, dcM);
      if (mm == dI + sub) vm = count_sat_add(vm, dcI);
      if (mm == dD + sub) vm = count_sat_add(vm, dcD);
      // I (:91-94, parents :108-119)
      const int32_t ii = max(lM + p.open, lI) + p.ext;
      int64_t vi = 0;
      if (ii == lI + p.ext) vi = count_sat_add(vi, lcI);
      if (ii == lM + p.open + p.ext) vi = count_sat_add(vi, lcM);
      // D (:87-90, parents :96-107)
      const int32_t dd = max(uM + p.open, uD) + p.ext;
      int64_t vd = 0;
      if (dd == uD + p.ext) vd = count_sat_add(vd, ucD);
      if (dd == uM + p.open + p.ext) vd = count_sat_add(vd, ucM);
      rm[o] = mm; ri[o] = ii; rd[o] = dd;
      cm[o] = vm; ci[o] = vi; cd[o] = vd;
      dM = uM; dI = uI; dD = uD;
      dcM = ucM; dcI = ucI; dcD = ucD;
      lM = mm; lI = ii;
      lcM = vm; lcI = vi;
    }
  }
  // end cell (:247-280): every state that attains the maximum is a start state
  const uint64_t o = (uint64_t)n1 * T;
  const int32_t eM = rm[o], eI = ri[o], eD = rd[o];
  const int32_t best = max(max(eI, eD), eM);
  int64_t total = 0;
  if (eI == best) total = count_sat_add(total, ci[o]);
  if (eM == best) total = count_sat_add(total, cm[o]);
  if (eD == best) total = count_sat_add(total, cd[o]);
  p.out[id] = total;
}

}  // namespace sa
