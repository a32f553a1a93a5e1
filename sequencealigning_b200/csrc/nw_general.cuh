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
// Traceback storage: one BYTE per interior cell: for each state of the cell (M bits 0-1, D bits
// 2-3, I bits 4-5) the parent the first printed alignment continues with -- the first parent in
// reverse push order whose fe is not NONE -- as 0 = none (dead end), 1 = M, 2 = D, 3 = I.  The
// parents precede the cell in fill order, so the choice is known when the cell is computed.
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
  uint8_t* __restrict__ tb;                // per launch index: [n2][n1] bytes at tb_off[k]; may be nullptr
  const uint64_t* __restrict__ tb_off;     // (bytes); UINT64_MAX = no room: score/status only
  int32_t* __restrict__ rows;              // per launch index: 6 * row_stride ints (affine: the edge column, 4 per row)
  uint8_t* __restrict__ info;              // per launch index: 2 * row_stride bytes (fe/taint of two rows)
  uint32_t row_stride;                     // max n1 + 1
  uint32_t* __restrict__ runs;             // per launch index: runs written back to front, ending at runs_end[k]
  const uint64_t* __restrict__ runs_end;
  int32_t* __restrict__ score;
  uint8_t* __restrict__ status;
  uint32_t* __restrict__ cigar_len;
  // Checkpointed traceback (pairs whose [n2][n1] traceback bytes do not fit): the forward kernel
  // keeps the right edge of EVERY column pass (ck: per pair passes x (n2 + 2) records), and the
  // backward kernel recomputes one pass at a time, last to first, into a block of traceback bytes
  // (blk: per pair n2 x pass-width) and walks the alignment through it.
  int4* __restrict__ ck;
  const uint64_t* __restrict__ ck_off;   // (records); UINT64_MAX = not checkpointed
  uint8_t* __restrict__ blk;
  const uint64_t* __restrict__ blk_off;  // (bytes)
  struct LongWalkState* __restrict__ ws;  // per launch index
};

// where the walk of a checkpointed pair stands between the two kernels
struct LongWalkState {
  uint32_t x, y;
  int32_t st;
  uint32_t pending;  // 1: the backward kernel has to produce the alignment
};

enum { kFeNone = 0, kFePrint = 1, kFePanic = 2 };
constexpr uint8_t kAlignmentOmitted = 0x80;  // ORed into status: the CIGAR was not materialised

// fe/taint byte of a cell: bits 0-1 fe(M), 2-3 fe(D), 4-5 fe(I), 6 unused; taint kept separately
__device__ __forceinline__ uint32_t fe_of(uint32_t info, int st) { return (info >> (2 * st)) & 3u; }

// One BLOCK per pair, columns in PASSES of THREADS * kGeneralCols.  In a pass thread g owns
// kGeneralCols consecutive columns, whose rolling row (three scores and the packed fe/taint bits
// per column) it keeps in REGISTERS, and runs one row behind thread g-1: at step t it computes
// row x = t - g, left to right.  What a thread needs from its left neighbour -- the cell
// (x, y0), and one step later the same record as the diagonal cell of row x + 1 -- goes through a
// double-buffered shared-memory slot, one __syncthreads per step.  The last thread of a pass
// leaves its right edge, one record per row, in global scratch (16 bytes per row); thread 0 of
// the next pass reads it back one row ahead of use.  Thread 0 of pass 0 has column 0 to its left
// (:200-216).  Memory traffic is the edge column and (if kept) the traceback words -- not the
// rolling row, which is what made a row longer than L2's share slow.
constexpr int kGeneralThreads = 256;       // pairs of a few thousand residues
constexpr int kGeneralThreadsWide = 512;   // longer pairs: twice the lanes per pair
constexpr int kGeneralCols = 8;

template <int THREADS>
struct GeneralShared {
  int32_t xM[2][THREADS], xI[2][THREADS], xD[2][THREADS];
  uint32_t xP[2][THREADS];
};

struct GeneralEnd {
  int32_t em = 0, ei = 0, ed = 0;
  uint32_t einfo = 0, etaint = 0;
};

// One column pass over rows 1..nrows.  edge_r: the previous pass's right edge (pass > 0);
// edge_w: where this pass leaves its own (or nullptr); tb: traceback bytes of this call's cells,
// row stride tb_stride, column (y - 1 - tb_col0); end: the end cell's record if it lies in this
// pass and row nrows == n2.
template <int THREADS>
__device__ __forceinline__ void general_pass(const GeneralParams& p, GeneralShared<THREADS>& sh, uint32_t n1,
                                             uint32_t n2, uint64_t qo, uint64_t dof, uint32_t pass,
                                             uint32_t nrows, const int4* edge_r, int4* edge_w, uint8_t* tb,
                                             uint32_t tb_stride, uint32_t tb_col0, bool want_end,
                                             GeneralEnd& end) {
  constexpr uint32_t T = THREADS;
  constexpr int C = kGeneralCols;
  enum { ST_M = 0, ST_D = 1, ST_I = 2 };
  constexpr int32_t kNegInf = -32768;
  const uint32_t g = threadIdx.x;
  const uint32_t P = T * C;
  const uint32_t owner = n1 ? ((n1 - 1) % P) / C : 0, owner_c = n1 ? (n1 - 1) % C : 0;
  const uint32_t y0 = pass * P + g * C;  // this thread's columns are y0+1 .. y0+C
  const bool owns = y0 < n1;
  // row 0 (:172-199): D[0][y>=1] is the boundary chain (has a parent and x == 0: expanding it
  // panics, :299); M[0][y], I[0][y] are sentinels without parents.
  int32_t rM[C], rI[C], rD[C];
  uint32_t rP[C], q[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const uint32_t y = y0 + c + 1;
    rM[c] = kNegInf;
    rI[c] = kNegInf;
    rD[c] = ((int32_t)y + 1) * p.ext + p.open;
    rP[c] = (kFePanic << 2) | ((1u << ST_D) << 8);
    q[c] = y <= n1 ? load_residue(p.residues, qo + y - 1, p.packing) : 0xffffffffu;
  }
  // the cell (0, y0): diagonal input of row 1
  int32_t gM, gI, gD;
  uint32_t gP;
  if (y0 == 0) {  // origin: M[0][0] = 0 is popped at (0,0) -> PRINT; D/I[0][0] print too if popped (:283)
    gM = 0; gI = kNegInf; gD = kNegInf;
    gP = kFePrint | (kFePrint << 2) | (kFePrint << 4);
  } else {
    gM = kNegInf; gI = kNegInf; gD = ((int32_t)y0 + 1) * p.ext + p.open;
    gP = (kFePanic << 2) | ((1u << ST_D) << 8);
  }
  // thread 0 of a later pass: the previous pass's right edge, fetched one row ahead
  int4 ahead = make_int4(0, 0, 0, 0);
  if (g == 0 && pass > 0) ahead = __ldcg(&edge_r[1]);

  const uint32_t steps = nrows + T - 1;
  for (uint32_t t = 1; t <= steps; ++t) {
    const uint32_t buf = t & 1u;
    const uint32_t x = t - g;  // (wraps for t < g: then x > nrows)
    if (owns && x >= 1 && x <= nrows) {
      // the cell to the left of this thread's first column, row x
      int32_t lM, lI, lD;
      uint32_t lP;
      if (g == 0) {
        if (pass == 0) {  // column 0 (:200-216): I[x][0] is the boundary chain (:303)
          lM = kNegInf; lD = kNegInf; lI = p.open + ((int32_t)x + 1) * p.ext;
          lP = (kFePanic << 4) | ((1u << ST_I) << 8);
        } else {
          lM = ahead.x; lI = ahead.y; lD = ahead.z; lP = (uint32_t)ahead.w;
          if (x < nrows) ahead = __ldcg(&edge_r[x + 1]);
        }
      } else {
        lM = sh.xM[buf ^ 1u][g - 1]; lI = sh.xI[buf ^ 1u][g - 1]; lD = sh.xD[buf ^ 1u][g - 1];
        lP = sh.xP[buf ^ 1u][g - 1];
      }
      const int32_t nextM = lM, nextI = lI, nextD = lD;  // becomes the diagonal input of row x + 1
      const uint32_t nextP = lP;
      const uint32_t b2 = load_residue(p.residues, dof + x - 1, p.packing);
      int32_t dm = gM, di = gI, dd = gD;  // (x-1, y-1)
      uint32_t pdg = gP;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int32_t um = rM[c], ui = rI[c], ud = rD[c];  // (x-1, y)
        const uint32_t pup = rP[c];
        const int32_t sub = q[c] == b2 ? p.match : p.mismatch;
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
        // DFS bookkeeping (info in bits 0-5, taint in bits 8-10 of the packed records)
        const uint32_t idg = pdg & 0xffu, tdg = pdg >> 8;
        const uint32_t ilf = lP & 0xffu, tlf = lP >> 8;
        const uint32_t iup = pup & 0xffu, tup = pup >> 8;
        // Per state: first event = that of the first parent, in reverse push order, whose own
        // first event is not NONE; that parent is where the first alignment continues (ch =
        // state + 1, 0 = dead end); taint = OR over all parents.  Selects only, no branches:
        // the lanes of a warp sit on unrelated cells.
        const uint32_t mD = (bits & 4u) ? fe_of(idg, ST_D) : 0u;   // M <- D, I, M (:120-153, popped in reverse)
        const uint32_t mI = (bits & 2u) ? fe_of(idg, ST_I) : 0u;
        const uint32_t mM = (bits & 1u) ? fe_of(idg, ST_M) : 0u;
        const uint32_t feM = mD ? mD : (mI ? mI : mM);
        const uint32_t chM = mD ? ST_D + 1u : (mI ? ST_I + 1u : (mM ? ST_M + 1u : 0u));
        const uint32_t tM = (((bits >> 2) & (tdg >> ST_D)) | ((bits >> 1) & (tdg >> ST_I)) | (bits & (tdg >> ST_M))) & 1u;
        const uint32_t iM = (bits & 16u) ? fe_of(ilf, ST_M) : 0u;  // I <- M, I (:108-119)
        const uint32_t iI = (bits & 8u) ? fe_of(ilf, ST_I) : 0u;
        const uint32_t feI = iM ? iM : iI;
        const uint32_t chI = iM ? ST_M + 1u : (iI ? ST_I + 1u : 0u);
        const uint32_t tI = (((bits >> 4) & (tlf >> ST_M)) | ((bits >> 3) & (tlf >> ST_I))) & 1u;
        const uint32_t dM = (bits & 64u) ? fe_of(iup, ST_M) : 0u;  // D <- M, D (:96-107)
        const uint32_t dD = (bits & 32u) ? fe_of(iup, ST_D) : 0u;
        const uint32_t feD = dM ? dM : dD;
        const uint32_t chD = dM ? ST_M + 1u : (dD ? ST_D + 1u : 0u);
        const uint32_t tD = (((bits >> 6) & (tup >> ST_M)) | ((bits >> 5) & (tup >> ST_D))) & 1u;
        const uint32_t np = feM | (feD << 2) | (feI << 4) | (((tM << ST_M) | (tD << ST_D) | (tI << ST_I)) << 8);
        rM[c] = mm;
        rI[c] = ii;
        rD[c] = dv;
        rP[c] = np;
        if (tb && y0 + c < n1)
          tb[(uint64_t)(x - 1) * tb_stride + (y0 + c - tb_col0)] =
              (uint8_t)((chM << (2 * ST_M)) | (chD << (2 * ST_D)) | (chI << (2 * ST_I)));
        dm = um; di = ui; dd = ud;  // the cell above becomes the diagonal of the next column
        pdg = pup;
        lM = mm; lI = ii; lD = dv;  // and this cell its left neighbour
        lP = np;
      }
      sh.xM[buf][g] = lM; sh.xI[buf][g] = lI; sh.xD[buf][g] = lD;
      sh.xP[buf][g] = lP;
      if (g == T - 1 && edge_w) edge_w[x] = make_int4(lM, lI, lD, (int)lP);
      gM = nextM; gI = nextI; gD = nextD;
      gP = nextP;
      if (want_end && g == owner && x == n2) {
#pragma unroll
        for (int c = 0; c < C; ++c)
          if ((uint32_t)c == owner_c) {
            end.em = rM[c]; end.ei = rI[c]; end.ed = rD[c];
            end.einfo = rP[c] & 0xffu; end.etaint = rP[c] >> 8;
          }
      }
    }
    __syncthreads();
  }
  __threadfence_block();  // the pass's edge column and traceback bytes before the next readers
  __syncthreads();
}

// One step of the walk of the first printed alignment through a block of traceback bytes; the
// runs are produced last to first.
struct LongWalk {
  uint32_t x, y;
  int st;
  uint32_t nruns, run_op, run_len;
  uint32_t* out;
  __device__ __forceinline__ void emit() {
    const uint32_t op = st == 0 ? 0u : (st == 2 ? 1u : 2u);  // M -> SA_OP_M, I -> SA_OP_I, D -> SA_OP_D
    if (op != run_op) {
      if (run_len) *--out = (run_len << 2) | run_op;
      run_op = op;
      run_len = 0;
      ++nruns;
    }
    ++run_len;
  }
  __device__ __forceinline__ void step(uint32_t wd) {  // wd: the byte of cell (x, y), 0 on the border
    const int nst = (int)((wd >> (2 * st)) & 3u) - 1;
    if (st == 0) { --x; --y; }
    else if (st == 2) --y;
    else --x;
    st = nst;
  }
  __device__ __forceinline__ bool alive() const { return !(x == 0 && y == 0) && st >= 0; }
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) nw_affine_general_kernel(const GeneralParams p) {
  constexpr uint32_t T = THREADS;
  constexpr int C = kGeneralCols;
  const uint32_t k = blockIdx.x;
  if (k >= p.n_ids) return;
  const uint32_t g = threadIdx.x;
  enum { ST_M = 0, ST_D = 1, ST_I = 2 };
  constexpr int32_t kNegInf = -32768;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  int4* edge = reinterpret_cast<int4*>(p.rows + (uint64_t)k * 6 * p.row_stride);  // [row]: M, I, D, info | taint << 8
  const bool keep_tb = p.tb != nullptr && p.tb_off[k] != ~0ull;
  uint8_t* tb = keep_tb ? p.tb + p.tb_off[k] : nullptr;
  const bool ckpt = !keep_tb && p.ck != nullptr && p.ck_off[k] != ~0ull;
  int4* ck = ckpt ? p.ck + p.ck_off[k] : nullptr;
  const uint64_t S = (uint64_t)n2 + 2;  // records per checkpointed edge

  const uint32_t P = T * C;
  const uint32_t npass = (n1 + P - 1) / P;
  __shared__ GeneralShared<THREADS> sh;
  GeneralEnd end;
  const bool interior = n1 && n2;
  const uint32_t owner = n1 ? ((n1 - 1) % P) / C : 0;
  if (p.ws && g == 0) p.ws[k].pending = 0;

  for (uint32_t pass = 0; pass < npass && interior; ++pass) {
    const int4* er = pass ? (ckpt ? ck + (uint64_t)(pass - 1) * S : edge) : nullptr;
    int4* ew = pass + 1 < npass ? (ckpt ? ck + (uint64_t)pass * S : edge) : nullptr;
    general_pass<THREADS>(p, sh, n1, n2, qo, dof, pass, n2, er, ew, tb, n1, 0, pass + 1 == npass, end);
  }
  int32_t em = end.em, ei = end.ei, ed = end.ed;
  uint32_t einfo = end.einfo, etaint = end.etaint;
  if (!interior) {
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
    if (tb) {
      // first printed alignment: at every cell the parent chosen at fill time
      LongWalk wk{n2, n1, first, 0, 3, 0, p.runs + p.runs_end[k]};
      while (wk.alive()) {
        wk.emit();
        // border cells have no continuing parent: the walk ends there (a dead end or, at (0,0), the print)
        wk.step((wk.x >= 1 && wk.y >= 1) ? tb[(uint64_t)(wk.x - 1) * n1 + (wk.y - 1)] : 0u);
      }
      if (wk.run_len) *--wk.out = (wk.run_len << 2) | wk.run_op;
      nruns = wk.nruns;
    } else if (ckpt && p.ws) {
      p.ws[k].x = n2;
      p.ws[k].y = n1;
      p.ws[k].st = first;
      p.ws[k].pending = 1;  // nw_affine_general_back produces the alignment and cigar_len
    } else {
      status |= kAlignmentOmitted;
    }
  }
  p.score[id] = mx;
  p.status[id] = status;
  p.cigar_len[id] = nruns;
}

// Checkpointed traceback, second kernel: passes last to first.  A pass is recomputed from the
// edge its left neighbour left behind, only down to the row the walk stands on, with its
// traceback bytes going to the pair's block; thread 0 then walks until the path leaves the
// pass on the left (or ends).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) nw_affine_general_back(const GeneralParams p) {
  constexpr uint32_t T = THREADS;
  constexpr int C = kGeneralCols;
  const uint32_t k = blockIdx.x;
  if (k >= p.n_ids || !p.ws || !p.ws[k].pending) return;  // (uniform per block)
  const uint32_t g = threadIdx.x;
  const uint32_t id = p.ids[k];
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint64_t qo = p.q_off[id], dof = p.d_off[id];
  const int4* ck = p.ck + p.ck_off[k];
  uint8_t* blk = p.blk + p.blk_off[k];
  const uint64_t S = (uint64_t)n2 + 2;
  const uint32_t P = T * C;
  const uint32_t npass = (n1 + P - 1) / P;
  __shared__ GeneralShared<THREADS> sh;
  __shared__ LongWalk wk;
  if (g == 0) wk = LongWalk{p.ws[k].x, p.ws[k].y, p.ws[k].st, 0, 3, 0, p.runs + p.runs_end[k]};
  __syncthreads();
  GeneralEnd unused;
  for (int pass = (int)npass - 1; pass >= 0; --pass) {
    const uint32_t y_base = (uint32_t)pass * P;
    const bool here = wk.alive() && wk.y > y_base && wk.x >= 1;  // (shared: uniform)
    if (!here) {
      if (!wk.alive() || wk.x == 0) break;
      continue;
    }
    const uint32_t nrows = wk.x;
    __syncthreads();
    general_pass<THREADS>(p, sh, n1, n2, qo, dof, (uint32_t)pass, nrows, pass ? ck + (uint64_t)(pass - 1) * S : nullptr,
                          nullptr, blk, P, y_base, false, unused);
    if (g == 0) {
      while (wk.alive() && wk.y > y_base) {
        wk.emit();
        wk.step(wk.x >= 1 ? blk[(uint64_t)(wk.x - 1) * P + (wk.y - 1 - y_base)] : 0u);
      }
    }
    __syncthreads();
  }
  if (g == 0) {
    // what is left runs along the border: column 0 / row 0 cells have no continuing parent
    while (wk.alive()) {
      wk.emit();
      wk.step(0u);
    }
    if (wk.run_len) *--wk.out = (wk.run_len << 2) | wk.run_op;
    p.cigar_len[id] = wk.nruns;
  }
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
  uint8_t* tb = keep_tb ? p.tb + p.tb_off[k] : nullptr;

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
        if (tb) tb[(uint64_t)(i - 1) * n2 + (j - 1)] = (uint8_t)mv;
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
