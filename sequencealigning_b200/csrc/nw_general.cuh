// nw_general.cuh -- affine NW for pairs OUTSIDE the packed 16-bit kernel's range (long pairs,
// n1 + n2 above ~3.6 k): the literal recurrences in 32-bit integers with the finite -32768
// sentinel, one thread block per pair.  It is the completeness path (the batched hot path is
// nw_affine_s16.cuh): exact in every regime the reference has, including the one
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

// One BLOCK of kGeneralThreads per pair.  Thread g owns the columns [g*C + 1, (g+1)*C]
// (C = ceil(n1 / threads)) and runs one row behind thread g-1: at step t it computes row
// x = t - g of its columns, left to right, over a rolling row kept in place in global scratch
// (layout [column-in-thread][thread]: the threads of a warp touch one line per access, and a
// pair's whole row stays in L1/L2).  What a thread needs from its left neighbour -- the cell
// (x, y_lo - 1), and one step later the same record as the diagonal cell of row x + 1 -- goes
// through a double-buffered shared-memory slot, one __syncthreads per step.  Thread 0's left
// neighbour is column 0 (:200-216).
constexpr int kGeneralThreads = 256;       // pairs up to ~8 k columns
constexpr int kGeneralThreadsWide = 1024;  // longer pairs: four times the lanes per pair

template <int THREADS>
__global__ void __launch_bounds__(THREADS) nw_affine_general_kernel(const GeneralParams p) {
  constexpr uint32_t T = THREADS;
  const uint32_t k = blockIdx.x;
  if (k >= p.n_ids) return;
  const uint32_t g = threadIdx.x;
  enum { ST_M = 0, ST_D = 1, ST_I = 2 };
  constexpr int32_t kNegInf = -32768;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const uint32_t w = p.row_stride;  // >= n1 + T: room for C * T entries
  int32_t* base = p.rows + (uint64_t)k * 6 * w;
  int32_t *rm = base + g, *ri = base + w + g, *rd = base + 2 * w + g;  // entry c of this thread at [c * T]
  uint8_t* rinfo = p.info + (uint64_t)k * 4 * w + g;                   // fe bits: M (0-1), D (2-3), I (4-5)
  uint8_t* rtaint = rinfo + w;                                         // taint bits by state
  const bool keep_tb = p.tb != nullptr && p.tb_off[k] != ~0ull;
  uint16_t* tb = keep_tb ? p.tb + p.tb_off[k] : nullptr;

  const uint32_t C = (n1 + T - 1) / T;
  const uint32_t y_lo = g * C + 1, y_hi = min(n1, (g + 1) * C);
  const bool owns = C > 0 && y_lo <= n1;
  const uint32_t owner = C ? (n1 - 1) / C : 0;  // thread that holds column n1

  __shared__ int32_t xM[2][T], xI[2][T], xD[2][T];
  __shared__ uint32_t xP[2][T];

  // row 0 (:172-199) of this thread's columns: D[0][y>=1] is the boundary chain (has a parent and
  // x == 0: expanding it panics, :299); M[0][y], I[0][y] are sentinels without parents.
  if (owns)
    for (uint32_t y = y_lo, c = 0; y <= y_hi; ++y, ++c) {
      rm[c * T] = kNegInf;
      ri[c * T] = kNegInf;
      rd[c * T] = ((int32_t)y + 1) * p.ext + p.open;
      rinfo[c * T] = (uint8_t)(kFePanic << 2);
      rtaint[c * T] = (uint8_t)(1u << ST_D);
    }
  // the cell (0, y_lo - 1): diagonal input of row 1
  int32_t gM, gI, gD;
  uint32_t gInfo, gTaint;
  if (g == 0) {  // origin: M[0][0] = 0 is popped at (0,0) -> PRINT; D/I[0][0] print too if popped (:283)
    gM = 0; gI = kNegInf; gD = kNegInf;
    gInfo = kFePrint | (kFePrint << 2) | (kFePrint << 4);
    gTaint = 0;
  } else {
    gM = kNegInf; gI = kNegInf; gD = ((int32_t)(y_lo - 1) + 1) * p.ext + p.open;
    gInfo = kFePanic << 2;
    gTaint = 1u << ST_D;
  }
  int32_t em = 0, ei = 0, ed = 0;  // end cell (owner thread)
  uint32_t einfo = 0, etaint = 0;

  const uint32_t steps = (n1 && n2) ? n2 + T - 1 : 0;
  for (uint32_t t = 1; t <= steps; ++t) {
    const uint32_t buf = t & 1u;
    const uint32_t x = t - g;  // (wraps for t < g: then x > n2)
    if (owns && x >= 1 && x <= n2) {
      // the cell to the left of this thread's first column, row x (written by thread g-1 at step t-1)
      int32_t lM, lI, lD;
      uint32_t lInfo, lTaint;
      if (g == 0) {  // column 0 (:200-216): I[x][0] is the boundary chain (:303)
        lM = kNegInf; lD = kNegInf; lI = p.open + ((int32_t)x + 1) * p.ext;
        lInfo = kFePanic << 4;
        lTaint = 1u << ST_I;
      } else {
        lM = xM[buf ^ 1u][g - 1]; lI = xI[buf ^ 1u][g - 1]; lD = xD[buf ^ 1u][g - 1];
        const uint32_t inP = xP[buf ^ 1u][g - 1];
        lInfo = inP & 0xffu;
        lTaint = inP >> 8;
      }
      const int32_t nextM = lM, nextI = lI, nextD = lD;  // becomes the diagonal input of row x + 1
      const uint32_t nextInfo = lInfo, nextTaint = lTaint;
      const uint32_t b2 = load_residue(p.residues, dof + x - 1, p.packing);
      int32_t dm = gM, di = gI, dd = gD;  // (x-1, y-1)
      uint32_t idg = gInfo, tdg = gTaint;
      for (uint32_t y = y_lo, c = 0; y <= y_hi; ++y, ++c) {
        const int32_t um = rm[c * T], ui = ri[c * T], ud = rd[c * T];  // (x-1, y)
        const uint32_t iup = rinfo[c * T], tup = rtaint[c * T];
        const int32_t sub = load_residue(p.residues, qo + y - 1, p.packing) == b2 ? p.match : p.mismatch;
        const int32_t mm = max(max(dm, di), dd) + sub;
        const int32_t ii = max(lM + p.open, lI) + p.ext;
        const int32_t dv = max(um + p.open, ud) + p.ext;
        uint32_t bits = 0;
        if (mm == dm + sub) bits |= 1u;
        if (mm == di + sub) bits |= 2u;
        if (mm == dd + sub) bits |= 4u;
        if (ii == lI + p.ext) bits |= 8u;
        if (ii == lM + p.open + p.ext) bits |= 16u;
        if (dv == ud + p.ext) bits |= 32u;
        if (dv == um + p.open + p.ext) bits |= 64u;
        // DFS bookkeeping, parents in reverse push order
        uint32_t feM = kFeNone, feI = kFeNone, feD = kFeNone, tM = 0, tI = 0, tD = 0;
        if (bits & 4u) { if (!feM) feM = fe_of(idg, ST_D); tM |= (tdg >> ST_D) & 1u; }
        if (bits & 2u) { if (!feM) feM = fe_of(idg, ST_I); tM |= (tdg >> ST_I) & 1u; }
        if (bits & 1u) { if (!feM) feM = fe_of(idg, ST_M); tM |= (tdg >> ST_M) & 1u; }
        if (bits & 16u) { if (!feI) feI = fe_of(lInfo, ST_M); tI |= (lTaint >> ST_M) & 1u; }
        if (bits & 8u) { if (!feI) feI = fe_of(lInfo, ST_I); tI |= (lTaint >> ST_I) & 1u; }
        if (bits & 64u) { if (!feD) feD = fe_of(iup, ST_M); tD |= (tup >> ST_M) & 1u; }
        if (bits & 32u) { if (!feD) feD = fe_of(iup, ST_D); tD |= (tup >> ST_D) & 1u; }
        const uint32_t ninfo = feM | (feD << 2) | (feI << 4);
        const uint32_t ntaint = (tM << ST_M) | (tD << ST_D) | (tI << ST_I);
        rm[c * T] = mm;
        ri[c * T] = ii;
        rd[c * T] = dv;
        rinfo[c * T] = (uint8_t)ninfo;
        rtaint[c * T] = (uint8_t)ntaint;
        if (tb) tb[(uint64_t)(x - 1) * n1 + (y - 1)] = (uint16_t)(bits | (feM << 7) | (feD << 9) | (feI << 11));
        dm = um; di = ui; dd = ud;  // the cell above becomes the diagonal of the next column
        idg = iup; tdg = tup;
        lM = mm; lI = ii; lD = dv;  // and this cell its left neighbour
        lInfo = ninfo; lTaint = ntaint;
      }
      xM[buf][g] = lM; xI[buf][g] = lI; xD[buf][g] = lD;
      xP[buf][g] = lInfo | (lTaint << 8);
      gM = nextM; gI = nextI; gD = nextD;
      gInfo = nextInfo; gTaint = nextTaint;
      if (g == owner && x == n2) {
        em = lM; ei = lI; ed = lD;
        einfo = lInfo; etaint = lTaint;
      }
    }
    __syncthreads();
  }
  __threadfence_block();  // the owner's walk reads traceback words written by the other threads
  __syncthreads();
  if (steps == 0) {
    // an empty side: the end cell is a border cell of row 0 / column 0
    if (g != 0) return;
    if (n1 == 0 && n2 == 0) {
      em = 0; ei = kNegInf; ed = kNegInf;
      einfo = kFePrint | (kFePrint << 2) | (kFePrint << 4);
      etaint = 0;
    } else if (n2 == 0) {  // (0, n1): the D chain of row 0
      em = kNegInf; ei = kNegInf; ed = ((int32_t)n1 + 1) * p.ext + p.open;
      einfo = kFePanic << 2;
      etaint = 1u << ST_D;
    } else {  // (n2, 0): the I chain of column 0
      em = kNegInf; ed = kNegInf; ei = p.open + ((int32_t)n2 + 1) * p.ext;
      einfo = kFePanic << 4;
      etaint = 1u << ST_I;
    }
  } else if (g != owner) {
    return;
  }
  // end cell (:246-280): start states pushed I, M, D, popped D, M, I
  const int32_t mx = max(max(ei, ed), em);
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

// ---------------------------------------------------------------------------------------------
// Linear ("pseudo-affine") NW for pairs outside the packed range: the literal recurrences of
// /root/reference/src/needleman_wunsch.rs:36-117 in 32-bit integers, same block-per-pair layout
// as above.  Rows i walk seq1 (query), columns j walk seq2 (db) (:38); thread g owns the columns
// [g*C + 1, (g+1)*C], C = ceil(n2 / threads).  Per column of the rolling row: (S << 1) | gap flag.
//   diag  = S[i-1][j-1] + (eq ? match : mismatch)
//   down  = S[i-1][j]   + (gap[i-1][j] ? ext : open)      Down  = seq1[i-1] over '-'  -> SA_OP_I
//   right = S[i][j-1]   + (gap[i][j-1] ? ext : open)      Right = '-' over seq2[j-1]  -> SA_OP_D
//   S = max3; gap = (S == down || S == right); moves pushed Down, Right, Diag (:92-100)
// Borders (:44-65): S[0][j] = open + j*ext, S[i][0] = open + i*ext, S[0][0] = 2*open, gap set.
// The first hit of get_next (:205-254) follows the first stored move; the traceback word keeps
// that move (1 Down, 2 Right, 0 Diag).  Nothing panics; status is always OK.
// ---------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) nw_linear_general_kernel(const GeneralParams p) {
  constexpr uint32_t T = THREADS;
  const uint32_t k = blockIdx.x;
  if (k >= p.n_ids) return;
  const uint32_t g = threadIdx.x;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const uint32_t w = p.row_stride;  // >= n2 + T
  int32_t* row = p.rows + (uint64_t)k * 6 * w + g;  // entry c of this thread at [c * T]: (S << 1) | gap
  const bool keep_tb = p.tb != nullptr && p.tb_off[k] != ~0ull;
  uint16_t* tb = keep_tb ? p.tb + p.tb_off[k] : nullptr;

  const uint32_t C = (n2 + T - 1) / T;
  const uint32_t j_lo = g * C + 1, j_hi = min(n2, (g + 1) * C);
  const bool owns = C > 0 && j_lo <= n2;
  const uint32_t owner = C ? (n2 - 1) / C : 0;
  __shared__ int32_t xS[2][T];

  if (owns)
    for (uint32_t j = j_lo, c = 0; j <= j_hi; ++j, ++c) row[c * T] = ((p.open + (int32_t)j * p.ext) << 1) | 1;
  // S[0][j_lo - 1], gap set: the diagonal input of row 1
  int32_t dg = g == 0 ? ((2 * p.open) << 1) | 1 : ((p.open + (int32_t)(j_lo - 1) * p.ext) << 1) | 1;
  int32_t end_s = 0;

  const uint32_t steps = (n1 && n2) ? n1 + T - 1 : 0;
  for (uint32_t t = 1; t <= steps; ++t) {
    const uint32_t buf = t & 1u;
    const uint32_t i = t - g;
    if (owns && i >= 1 && i <= n1) {
      int32_t lf = g == 0 ? ((p.open + (int32_t)i * p.ext) << 1) | 1 : xS[buf ^ 1u][g - 1];  // (i, j_lo - 1)
      const int32_t next_dg = lf;
      const uint32_t b1 = load_residue(p.residues, qo + i - 1, p.packing);
      int32_t d = dg;
      for (uint32_t j = j_lo, c = 0; j <= j_hi; ++j, ++c) {
        const int32_t up = row[c * T];
        const int32_t diag = (d >> 1) + (load_residue(p.residues, dof + j - 1, p.packing) == b1 ? p.match : p.mismatch);
        const int32_t down = (up >> 1) + ((up & 1) ? p.ext : p.open);
        const int32_t right = (lf >> 1) + ((lf & 1) ? p.ext : p.open);
        const int32_t mx = max(max(down, right), diag);
        const int32_t gap = (mx == down || mx == right) ? 1 : 0;  // :85-87, even when diag ties
        const uint32_t mv = mx == down ? 1u : (mx == right ? 2u : 0u);
        const int32_t v = (mx << 1) | gap;
        row[c * T] = v;
        if (tb) tb[(uint64_t)(i - 1) * n2 + (j - 1)] = (uint16_t)mv;
        d = up;
        lf = v;
      }
      xS[buf][g] = lf;
      dg = next_dg;
      if (g == owner && i == n1) end_s = lf >> 1;
    }
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  if (steps == 0) {
    if (g != 0) return;
    end_s = (n1 == 0 && n2 == 0) ? 2 * p.open : p.open + (int32_t)(n1 + n2) * p.ext;  // border cell
  } else if (g != owner) {
    return;
  }
  uint32_t nruns = 0;
  if (n1 | n2) {
    if (!tb && n1 && n2) {
      p.score[id] = end_s;
      p.status[id] = (uint8_t)(kOk | kAlignmentOmitted);
      p.cigar_len[id] = 0;
      return;
    }
    uint32_t i = n1, j = n2, run_op = 3, run_len = 0;
    uint32_t* out = p.runs + p.runs_end[k];
    while (i > 0 || j > 0) {
      uint32_t op;
      if (i == 0) { op = 2; --j; }        // row 0 holds [Right]
      else if (j == 0) { op = 1; --i; }   // column 0 holds [Down]
      else {
        const uint32_t mv = tb[(uint64_t)(i - 1) * n2 + (j - 1)];
        if (mv == 1u) { op = 1; --i; }
        else if (mv == 2u) { op = 2; --j; }
        else { op = 0; --i; --j; }
      }
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
  p.score[id] = end_s;
  p.status[id] = kOk;
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
