// nw_general.cuh -- affine NW for pairs OUTSIDE the packed 16-bit kernel's range (long pairs,
// n1 + n2 above ~3.6 k): the literal recurrences in 32-bit integers with the finite -32768
// sentinel, one thread per pair.  Slow by design (it is the completeness path, the batched hot
// path is nw_affine_s16.cuh) but exact in every regime the reference has, including the one
// where the sentinel leaks (n1 + n2 > ~5.4 k) and the traceback meets dead ends:
//
//   per cell and state the kernel keeps what the reference's LIFO DFS
//   (/root/reference/src/needleman_wunsch_affine.rs:246-329) would do below that cell:
//     fe    "first event" of the DFS subtree: NONE (only dead ends), PRINT, PANIC
//     taint some cell of the subtree panics
//   computed in fill order from the parents' values (parents precede the cell), visiting the
//   parent list (:96-153) in REVERSE push order because the stack is LIFO.  The first printed
//   alignment follows, at every cell, the first parent in that order whose fe is not NONE.
//
// Traceback storage: one 16-bit word per interior cell =
//   bits 0-6  parent set (bit0 M<-M, 1 M<-I, 2 M<-D, 3 I<-I, 4 I<-M, 5 D<-D, 6 D<-M)
//   bits 7-8  fe of the M state, 9-10 fe of D, 11-12 fe of I
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_affine_s16.cuh"

namespace sa {

struct GeneralParams {
  const uint8_t* __restrict__ residues;
  const uint64_t* __restrict__ q_off;
  const uint32_t* __restrict__ q_len;
  const uint64_t* __restrict__ d_off;
  const uint32_t* __restrict__ d_len;
  const uint32_t* __restrict__ ids;   // pairs handled by this launch
  uint32_t n_ids, packing;
  int32_t match, mismatch, open, ext;
  uint16_t* __restrict__ tb;               // per launch index: [n2][n1] words at tb_off[k]; may be nullptr
  const uint64_t* __restrict__ tb_off;     // (uint16 units); UINT64_MAX = no room: score/status only
  int32_t* __restrict__ rows;              // per launch index: 6 * row_stride ints
  uint8_t* __restrict__ info;              // per launch index: 2 * row_stride bytes (fe/taint of two rows)
  uint32_t row_stride;                     // max n1 + 1
  uint32_t* __restrict__ runs;             // per launch index: runs written back to front, ending at runs_end[k]
  const uint64_t* __restrict__ runs_end;
  int32_t* __restrict__ score;
  uint8_t* __restrict__ status;
  uint32_t* __restrict__ cigar_len;
};

enum { kFeNone = 0, kFePrint = 1, kFePanic = 2 };
constexpr uint8_t kAlignmentOmitted = 0x80;  // ORed into status: the CIGAR was not materialised

// fe/taint byte of a cell: bits 0-1 fe(M), 2-3 fe(D), 4-5 fe(I), 6 unused; taint kept separately
__device__ __forceinline__ uint32_t fe_of(uint32_t info, int st) { return (info >> (2 * st)) & 3u; }

__global__ void __launch_bounds__(32) nw_affine_general_kernel(const GeneralParams p) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_ids) return;
  enum { ST_M = 0, ST_D = 1, ST_I = 2 };
  constexpr int32_t kNegInf = -32768;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const uint32_t w = p.row_stride;
  int32_t* base = p.rows + (uint64_t)k * 6 * w;
  int32_t *pm = base, *pi = base + w, *pd = base + 2 * w, *cm = base + 3 * w, *ci = base + 4 * w, *cd = base + 5 * w;
  // info rows: low 6 bits fe (M, D, I), bits 6.. taint is kept in a second array
  uint8_t* pinfo = p.info + (uint64_t)k * 4 * w;
  uint8_t* cinfo = pinfo + w;
  uint8_t* ptaint = pinfo + 2 * w;
  uint8_t* ctaint = pinfo + 3 * w;
  const bool keep_tb = p.tb != nullptr && p.tb_off[k] != ~0ull;
  uint16_t* tb = keep_tb ? p.tb + p.tb_off[k] : nullptr;

  // row 0 (:172-199).  M[0][0]: popped at (0,0) -> PRINT.  D[0][0], I[0][0]: sentinels without
  // parents; if ever popped at (0,0) they print too (:283).  D[0][y>=1] is the boundary chain:
  // it has a parent and x == 0 -> expanding it panics (:299).  M[0][y], I[0][y]: dead ends.
  pm[0] = 0; pi[0] = kNegInf; pd[0] = kNegInf;
  pinfo[0] = (uint8_t)(kFePrint | (kFePrint << 2) | (kFePrint << 4));
  ptaint[0] = 0;
  for (uint32_t y = 1; y <= n1; ++y) {
    pm[y] = kNegInf;
    pi[y] = kNegInf;
    pd[y] = ((int32_t)y + 1) * p.ext + p.open;
    pinfo[y] = (uint8_t)(kFePanic << 2);
    ptaint[y] = (uint8_t)(1u << ST_D);
  }
  for (uint32_t x = 1; x <= n2; ++x) {
    const uint32_t b2 = load_residue(p.residues, dof + x - 1, p.packing);
    cm[0] = kNegInf;  // column 0 (:200-216): I[x][0] is the boundary chain (:303)
    cd[0] = kNegInf;
    ci[0] = p.open + ((int32_t)x + 1) * p.ext;
    cinfo[0] = (uint8_t)(kFePanic << 4);
    ctaint[0] = (uint8_t)(1u << ST_I);
    for (uint32_t y = 1; y <= n1; ++y) {
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
      // DFS bookkeeping, parents in reverse push order
      const uint32_t idg = pinfo[y - 1], tdg = ptaint[y - 1];  // (x-1, y-1)
      const uint32_t ilf = cinfo[y - 1], tlf = ctaint[y - 1];  // (x, y-1)
      const uint32_t iup = pinfo[y], tup = ptaint[y];          // (x-1, y)
      uint32_t feM = kFeNone, feI = kFeNone, feD = kFeNone, tM = 0, tI = 0, tD = 0;
      if (bits & 4u) { if (!feM) feM = fe_of(idg, ST_D); tM |= (tdg >> ST_D) & 1u; }
      if (bits & 2u) { if (!feM) feM = fe_of(idg, ST_I); tM |= (tdg >> ST_I) & 1u; }
      if (bits & 1u) { if (!feM) feM = fe_of(idg, ST_M); tM |= (tdg >> ST_M) & 1u; }
      if (bits & 16u) { if (!feI) feI = fe_of(ilf, ST_M); tI |= (tlf >> ST_M) & 1u; }
      if (bits & 8u) { if (!feI) feI = fe_of(ilf, ST_I); tI |= (tlf >> ST_I) & 1u; }
      if (bits & 64u) { if (!feD) feD = fe_of(iup, ST_M); tD |= (tup >> ST_M) & 1u; }
      if (bits & 32u) { if (!feD) feD = fe_of(iup, ST_D); tD |= (tup >> ST_D) & 1u; }
      cm[y] = mm;
      ci[y] = ii;
      cd[y] = dv;
      cinfo[y] = (uint8_t)(feM | (feD << 2) | (feI << 4));
      ctaint[y] = (uint8_t)((tM << ST_M) | (tD << ST_D) | (tI << ST_I));
      if (tb) tb[(uint64_t)(x - 1) * n1 + (y - 1)] = (uint16_t)(bits | (feM << 7) | (feD << 9) | (feI << 11));
    }
    int32_t* t;
    t = pm; pm = cm; cm = t;
    t = pi; pi = ci; ci = t;
    t = pd; pd = cd; cd = t;
    uint8_t* u;
    u = pinfo; pinfo = cinfo; cinfo = u;
    u = ptaint; ptaint = ctaint; ctaint = u;
  }
  // end cell (:246-280): start states pushed I, M, D, popped D, M, I
  const int32_t em = pm[n1], ei = pi[n1], ed = pd[n1];
  const int32_t mx = max(max(ei, ed), em);
  const uint32_t einfo = pinfo[n1], etaint = ptaint[n1];
  int first = -1;
  uint32_t fe = kFeNone, any_panic = 0;
  const int order[3] = {ST_D, ST_M, ST_I};
  const int32_t vals[3] = {ed, em, ei};
  for (int t = 0; t < 3; ++t)
    if (vals[t] == mx) {
      if (fe == kFeNone && fe_of(einfo, order[t]) != kFeNone) { fe = fe_of(einfo, order[t]); first = order[t]; }
      any_panic |= (etaint >> order[t]) & 1u;
    }
  uint8_t status = fe == kFePrint ? (any_panic ? kRefPanic : kOk) : (fe == kFePanic ? kRefPanicEarly : kRefNoOutput);
  uint32_t nruns = 0;
  if (fe == kFePrint && (n1 | n2)) {
    if (!tb) {
      status |= kAlignmentOmitted;
    } else {
      // first printed alignment: at every cell the first parent (reverse push order) with an event
      uint32_t x = n2, y = n1, run_op = 3, run_len = 0;
      int st = first;
      uint32_t* out = p.runs + p.runs_end[k];
      while (!(x == 0 && y == 0) && st >= 0) {
        const uint32_t op = st == ST_M ? 0u : (st == ST_I ? 1u : 2u);
        if (op != run_op) {
          if (run_len) *--out = (run_len << 2) | run_op;
          run_op = op;
          run_len = 0;
          ++nruns;
        }
        ++run_len;
        const uint32_t wd = (x >= 1 && y >= 1) ? tb[(uint64_t)(x - 1) * n1 + (y - 1)] : 0u;
        // fe of a neighbour's state: interior cells from their word, border cells by rule
        auto nfe = [&](uint32_t nx, uint32_t ny, int nst) -> uint32_t {
          if (nx >= 1 && ny >= 1) {
            const uint32_t nw = tb[(uint64_t)(nx - 1) * n1 + (ny - 1)];
            return (nw >> (nst == ST_M ? 7 : (nst == ST_D ? 9 : 11))) & 3u;
          }
          if (nx == 0 && ny == 0) return kFePrint;
          if (nx == 0) return nst == ST_D ? kFePanic : kFeNone;
          return nst == ST_I ? kFePanic : kFeNone;
        };
        int nst = -1;
        if (st == ST_M) {
          if ((wd & 4u) && nfe(x - 1, y - 1, ST_D)) nst = ST_D;
          else if ((wd & 2u) && nfe(x - 1, y - 1, ST_I)) nst = ST_I;
          else if ((wd & 1u) && nfe(x - 1, y - 1, ST_M)) nst = ST_M;
          --x; --y;
        } else if (st == ST_I) {
          if ((wd & 16u) && nfe(x, y - 1, ST_M)) nst = ST_M;
          else if ((wd & 8u) && nfe(x, y - 1, ST_I)) nst = ST_I;
          --y;
        } else {
          if ((wd & 64u) && nfe(x - 1, y, ST_M)) nst = ST_M;
          else if ((wd & 32u) && nfe(x - 1, y, ST_D)) nst = ST_D;
          --x;
        }
        st = nst;
      }
      if (run_len) *--out = (run_len << 2) | run_op;
    }
  }
  p.score[id] = mx;
  p.status[id] = status;
  p.cigar_len[id] = nruns;
}

// moves the runs of the general kernel's pairs from their staging area into the pool
__global__ void __launch_bounds__(128) general_runs_to_pool(const uint32_t* __restrict__ ids, uint32_t n_ids,
                                                            const uint32_t* __restrict__ runs,
                                                            const uint64_t* __restrict__ runs_end,
                                                            const uint32_t* __restrict__ cigar_len,
                                                            const uint64_t* __restrict__ cigar_off,
                                                            uint32_t* __restrict__ pool, uint64_t pool_cap) {
  const uint32_t k = blockIdx.x;
  if (k >= n_ids) return;
  const uint32_t id = ids[k], len = cigar_len[id];
  const uint64_t off = cigar_off[id];
  const uint32_t* src = runs + runs_end[k] - len;
  for (uint32_t t = threadIdx.x; t < len; t += blockDim.x)
    if (off + t < pool_cap) pool[off + t] = src[t];
}

}  // namespace sa
