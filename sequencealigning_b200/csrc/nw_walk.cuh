// nw_walk.cuh -- traceback ("first printed alignment") over the packed 4-bit matrix, and the
// small scan that turns per-pair CIGAR lengths into pool offsets.
//
// The reference's traceback (/root/reference/src/needleman_wunsch_affine.rs:242-329) is a
// LIFO depth-first enumeration of every co-optimal path.  The FIRST path it prints is the
// greedy walk that always follows the LAST-pushed parent:
//   end cell   : D if D == max, else M if M == max, else I            (:251-280 push I,M,D)
//   from M     : the state of (x-1,y-1) with priority D > I > M       (:120-153 push M,I,D)
//   from I     : M if opening ties/wins, else I                       (:108-119 push I,M)
//   from D     : M if opening ties/wins, else D                       (:96-107  push D,M)
// Reaching x == 0 with y > 0 (or y == 0 with x > 0) means the path runs into the boundary
// chain D[0][y] / I[x][0], where the reference indexes seq2[0-1] / seq1[0-1] and panics
// (:299/:303) -> SA_REF_PANIC_EARLY, no alignment.
//
// Why the "panic bonus" of the first fill cannot corrupt these bits for the pairs that keep
// them: the bonus (+1 on boundary-chain cells, scores doubled) only ever changes the outcome
// of a comparison between two candidates of EQUAL score, in favour of the one whose value
// descends from a boundary-chain cell.  Suppose a tie bit read by this walk were polluted.
// The walk only reads bits of cells on a co-optimal path (it starts at an optimal end state
// and every step follows a tie parent), and the polluted comparison's tainted candidate is
// then itself a tie parent of that cell, so a co-optimal path runs through a boundary-chain
// cell and the bonus propagates, along maxima, to the end cell.  Contrapositive: an end cell
// WITHOUT the bonus means every bit on every co-optimal path is clean.  Pairs WITH the bonus
// are re-filled without it (phase 1) before they are walked.
//
// One thread walks one pair; lanes of a warp walk neighbouring pairs of the same tile, whose
// paths stay close to each other, so the 8-byte traceback words they read share sectors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nw_affine_s16.cuh"

namespace sa {

#ifndef SA_STATUS_CODES
#define SA_STATUS_CODES
enum : uint8_t { kOk = 0, kRefPanic = 1, kRefNoConv = 2, kNotImpl = 3, kRefPanicEarly = 4, kRefNoOutput = 5 };
#endif

struct WalkParams {
  const uint32_t* __restrict__ q_len;
  const uint32_t* __restrict__ d_len;
  const uint32_t* __restrict__ pair_ids;  // launch index -> pair id (rerun list) or nullptr
  uint32_t pair_base;
  uint32_t n_launch_pairs;
  const uint32_t* __restrict__ n_launch_dev;  // if non-null, overrides n_launch_pairs (device-side count)
  const uint2* __restrict__ tb;
  uint64_t tb_tile_stride;
  uint32_t tb_rows;
  uint32_t ng;           // pair-of-pairs per tile (32 / G)
  uint32_t k, k_inv, w;  // columns per strip, ceil(2^32 / k), traceback words per strip row
  const uint32_t* __restrict__ end;  // per launch index
  int32_t match, open, ext;
  // outputs, indexed by pair id
  int32_t* __restrict__ score;
  uint8_t* __restrict__ status;
  uint32_t* __restrict__ cigar_len;
  const uint64_t* __restrict__ cigar_off;
  uint32_t* __restrict__ pool;
  uint64_t pool_cap;
  int phase;  // 0: first fill (bonus on)  1: clean refill of queued pairs
  // The count pass already sees every run: it parks up to kTmpRuns of them per pair (back to
  // front, like the pool) so that the pool can be filled by a plain gather after the scan;
  // only pairs with more runs are walked a second time.
  uint32_t* __restrict__ tmp_runs;  // [segment pairs][kTmpRuns], indexed by (id - tmp_base)
  uint32_t tmp_base;
  // decoding of the fill's end word: 2*score (+ taint bit) = H' - bias + diag2*(n1+n2)
  // (affine: diag2 = 2*ext; linear: diag2 = match)
  int32_t bias, diag2;
  // Look-ahead of the walk, in steps (0 = off): every step also asks L2 for the word the walk reaches
  // after pf more diagonal steps.  A step is one dependent read of a word nobody else has touched (a
  // segment's traceback is GBs, far beyond L2), so without it a walk costs DRAM latency x path length;
  // reads are mostly diagonal runs, so the guess is right except just after a gap.
  uint32_t pf;
};

constexpr uint32_t kTmpRuns = 48;  // (a 150-250 bp read pair at 5 % has ~8-14 runs; beyond 48 the second walk does the work)

// Position of a walk inside its tile's traceback words.  The walks move one row and / or one column per step, so
// the word index is kept incrementally (two predicated subtracts per step) instead of being rebuilt from (x, y)
// with a division and three 64-bit multiplies in front of every dependent read.
struct TbCursor {
  const uint32_t* tile;  // the tile's words as 32-bit halves (pair A / pair B), offset by the lane group and the half
  uint32_t off;          // word 0 of (strip of column y-1, row x-1): ((s * tb_rows + (x - 1)) * w) * ng
  uint32_t c;            // column y-1 inside its strip
  uint32_t row_step;     // w * ng: one row up
  uint32_t strip_step;   // tb_rows * w * ng: one strip to the left
  uint32_t k, ng;
  uint32_t pf, pf_rows;  // look-ahead in steps (0xffffffff = off for the test below) and in words

  __device__ __forceinline__ void init(const WalkParams& p, uint64_t tile_base, uint32_t grp, uint32_t half, uint32_t x, uint32_t y) {
    tile = reinterpret_cast<const uint32_t*>(p.tb + tile_base + grp) + half;
    k = p.k;
    ng = p.ng;
    row_step = p.w * p.ng;
    strip_step = p.tb_rows * row_step;
    const uint32_t steps = min(p.pf, p.k);  // the look-ahead crosses at most one strip edge
    pf = steps ? steps : 0xffffffffu;
    pf_rows = steps * row_step;
    const uint32_t s = __umulhi(y - 1, p.k_inv);  // exact for every y < 2^32 / k
    c = (y - 1) - s * p.k;
    off = (s * p.tb_rows + (x - 1)) * row_step;
  }
  // one step to (x - dx, y - dy); leaving the matrix (x or y = 0) wraps harmlessly, nothing is read there
  __device__ __forceinline__ void move(bool dx, bool dy) {
    if (dx) off -= row_step;
    if (dy) {
      if (c == 0) {
        c = k;
        off -= strip_step;
      }
      --c;
    }
  }
  // the same with the step held in registers: row_dec = dx ? row_step : 0, dy = 0 / 1 (constant along a run)
  __device__ __forceinline__ void move_run(uint32_t row_dec, uint32_t dy) {
    off -= row_dec;
    c -= dy;
    if ((int32_t)c < 0) {
      c += k;
      off -= strip_step;
    }
  }
  __device__ __forceinline__ uint32_t nibble() const {
    return (__ldg(tile + 2 * (off + (c >> 3) * ng)) >> (4 * (c & 7))) & 15u;
  }
  // ask L2 for the word the walk reaches after pf more diagonal steps from (x, y)
  __device__ __forceinline__ void prefetch(uint32_t x, uint32_t y) const {
    if (min(x, y) <= pf) return;
    uint32_t o = off - pf_rows, cp = c - pf;
    if ((int32_t)cp < 0) {
      cp += k;
      o -= strip_step;
    }
    asm volatile("prefetch.global.L2 [%0];" ::"l"(tile + 2 * (o + (cp >> 3) * ng)));
  }
};

// Next-state tables of the walk, 2 bits per nibble value (0 = M, 1 = I, 2 = D):
//   from M: D if bit1, else I if bit0, else M      (:120-153 push M,I,D)
//   from I: M if bit2 (opening ties/wins) else I   (:108-119)
//   from D: M if bit3 else D                       (:96-107)
__host__ __device__ constexpr uint32_t walk_lut(int st) {
  uint32_t v = 0;
  for (uint32_t nb = 0; nb < 16; ++nb) {
    uint32_t nx = 0;
    if (st == 0) nx = (nb & 2u) ? 2u : ((nb & 1u) ? 1u : 0u);
    if (st == 1) nx = (nb & 4u) ? 0u : 1u;
    if (st == 2) nx = (nb & 8u) ? 0u : 2u;
    v |= nx << (2 * nb);
  }
  return v;
}
constexpr uint32_t kLutM = walk_lut(0), kLutI = walk_lut(1), kLutD = walk_lut(2);

// MODE 0: classify + count runs (writes score/status/cigar_len, queues tainted pairs)
// MODE 1: write runs into the pool (needs cigar_off)
template <int MODE>
__global__ void __launch_bounds__(128) nw_affine_walk(const WalkParams p) {
  const uint32_t n_launch = p.n_launch_dev ? *p.n_launch_dev : p.n_launch_pairs;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_launch) return;
  const uint32_t id = p.pair_ids ? p.pair_ids[i] : p.pair_base + i;
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint32_t ppt = 2 * p.ng;
  const uint32_t tile = i / ppt, grp = (i % ppt) >> 1, half = i & 1;
  const uint64_t tile_base = (uint64_t)tile * p.tb_tile_stride;

  if (n1 == 0 || n2 == 0) {
    if (MODE == 0 && p.phase == 0) {
      if (n1 == 0 && n2 == 0) {  // M[0][0] = 0 is popped at (0,0): one empty alignment (:283-286)
        p.score[id] = 0;
        p.status[id] = kOk;
      } else {  // the end cell IS a boundary-chain cell (:183-216): expanding it panics
        const int32_t n = (int32_t)(n1 + n2);
        p.score[id] = p.open + (n + 1) * p.ext;
        p.status[id] = kRefPanicEarly;
      }
      p.cigar_len[id] = 0;
    }
    return;
  }

  const uint32_t w = p.end[i];
  uint32_t st = (w >> 16) & 3u;
  if (MODE == 0) {
    const int32_t t2 = (int32_t)(w & 0xffffu) - p.bias + p.diag2 * (int32_t)(n1 + n2);
    const int32_t taint = t2 & 1;
    if (p.phase == 0) {
      p.score[id] = (t2 - taint) / 2;
      // some co-optimal path starts with a gap: the reference panics somewhere.  The fill queued the pair for the
      // clean refill (nw_affine_s16.cuh), whose walk (phase 1) owns its status, length and runs -- it may run
      // beside this kernel, so nothing else is written here.
      if (taint) return;
    }
  } else {
    // write pass: only pairs whose runs did not fit the temp slot; skip pairs that have no
    // alignment, and (phase 0) pairs owned by the refill
    if (p.cigar_len[id] <= kTmpRuns) return;
    if (p.phase == 0 && p.status[id] != kOk) return;
  }

  uint32_t x = n2, y = n1;
  uint32_t nruns = 0, run_op = 3, run_len = 0;
  uint64_t wpos = 0;
  if (MODE == 1) wpos = p.cigar_off[id] + p.cigar_len[id];  // runs are produced last-to-first
  uint32_t* tmp = (MODE == 0 && p.tmp_runs) ? p.tmp_runs + (uint64_t)(id - p.tmp_base) * kTmpRuns : nullptr;
  TbCursor cur;
  cur.init(p, tile_base, grp, half, x, y);  // (the end cell's own nibble is never read: its state comes from the end word)
  // What a step does is fixed along a run (the state), so the per-state quantities -- the move, the next-state table --
  // are set where the run changes and the common path is: count, move, test for the border, read, look up.
  //   M: emit a diagonal column, go to (x-1,y-1), next state = its best state, priority D > I > M
  //   I: seq1[y-1] against '-', go to (x,y-1), next state M if opening ties/wins else I
  //   D: '-' against seq2[x-1], go to (x-1,y), next state M if opening ties/wins else D
  uint32_t dx = 0, dy = 0, row_dec = 0, lut = 0;
  for (;;) {  // (x > 0 and y > 0 here)
    if (st != run_op) {
      if (MODE == 1 && run_len) {
        --wpos;
        if (wpos < p.pool_cap) p.pool[wpos] = (run_len << 2) | run_op;
      }
      if (MODE == 0 && run_len && tmp && nruns <= kTmpRuns) tmp[kTmpRuns - nruns] = (run_len << 2) | run_op;
      run_op = st;
      run_len = 0;
      ++nruns;
      dx = st != 1u;
      dy = st != 2u;
      row_dec = dx ? cur.row_step : 0u;
      lut = st == 0 ? kLutM : (st == 1 ? kLutI : kLutD);
    }
    ++run_len;
    x -= dx;
    y -= dy;
    cur.move_run(row_dec, dy);
    if (min(x, y) == 0) break;
    const uint32_t nb = cur.nibble();
    cur.prefetch(x, y);
    st = (lut >> (2 * nb)) & 3u;
  }
  const bool complete = (x == 0 && y == 0);
  if (MODE == 0) {
    if (complete && run_len && tmp && nruns <= kTmpRuns) tmp[kTmpRuns - nruns] = (run_len << 2) | run_op;
    if (p.phase == 0) {
      // untainted pairs cannot run into the boundary chain (that path would carry the bonus)
      p.status[id] = complete ? kOk : kRefPanicEarly;
      p.cigar_len[id] = complete ? nruns : 0;
    } else {
      p.status[id] = complete ? kRefPanic : kRefPanicEarly;
      p.cigar_len[id] = complete ? nruns : 0;
    }
  } else if (complete && run_len) {
    --wpos;
    if (wpos < p.pool_cap) p.pool[wpos] = (run_len << 2) | run_op;
  }
}

// Linear ("pseudo-affine") NW walk: the first hit of get_next (needleman_wunsch.rs:205-254)
// follows the first stored move of every cell, in push order Down, Right, Diag (:92-100).
// The fill kernel ran with rows = seq1 (query) and columns = seq2 (db), like the reference.
//   nibble bit0 = down >= right, bit1 = max(down,right) >= diag
//   Down  (seq1[i-1] over '-') -> SA_OP_I      Right ('-' over seq2[j-1]) -> SA_OP_D
// Row 0 holds [Right] and column 0 [Down] (:44-65), so the walk runs along the border to (0,0);
// nothing can panic in global mode.
template <int MODE>
__global__ void __launch_bounds__(128) nw_linear_walk(const WalkParams p) {
  const uint32_t i_launch = blockIdx.x * blockDim.x + threadIdx.x;
  if (i_launch >= p.n_launch_pairs) return;
  const uint32_t id = p.pair_ids ? p.pair_ids[i_launch] : p.pair_base + i_launch;
  const uint32_t n1 = p.q_len[id], n2 = p.d_len[id];
  const uint32_t ppt = 2 * p.ng;
  const uint32_t tile = i_launch / ppt, grp = (i_launch % ppt) >> 1, half = i_launch & 1;
  const uint64_t tile_base = (uint64_t)tile * p.tb_tile_stride;
  if (MODE == 0) {
    int32_t score;
    if (n1 == 0 || n2 == 0) {  // border cells (:44-65); scores[0][0] gets both increments
      score = (n1 == 0 && n2 == 0) ? 2 * p.open : p.open + (int32_t)(n1 + n2) * p.ext;
    } else {
      const int32_t t2 = (int32_t)(p.end[i_launch] & 0xffffu) - p.bias + p.diag2 * (int32_t)(n1 + n2);
      score = t2 / 2;
    }
    p.score[id] = score;
    p.status[id] = kOk;
  } else if (p.cigar_len[id] <= kTmpRuns) {
    return;
  }
  uint32_t i = n1, j = n2;  // rows walk seq1, columns walk seq2
  uint32_t nruns = 0, run_op = 3, run_len = 0;
  uint64_t wpos = 0;
  if (MODE == 1) wpos = p.cigar_off[id] + p.cigar_len[id];
  uint32_t* tmp = (MODE == 0 && p.tmp_runs) ? p.tmp_runs + (uint64_t)(id - p.tmp_base) * kTmpRuns : nullptr;
  TbCursor cur;
  if (i > 0 && j > 0) cur.init(p, tile_base, grp, half, i, j);
  while (i > 0 || j > 0) {
    uint32_t op;
    if (i == 0) {
      op = 2;
      --j;
    } else if (j == 0) {
      op = 1;
      --i;
    } else {
      const uint32_t nb = cur.nibble();
      cur.prefetch(i, j);
      if ((nb & 3u) == 3u) {
        op = 1;
        --i;
      } else if (nb & 2u) {
        op = 2;
        --j;
      } else {
        op = 0;
        --i;
        --j;
      }
      cur.move(op != 2u, op != 1u);
    }
    if (op != run_op) {
      if (MODE == 1 && run_len) {
        --wpos;
        if (wpos < p.pool_cap) p.pool[wpos] = (run_len << 2) | run_op;
      }
      if (MODE == 0 && run_len && tmp && nruns <= kTmpRuns) tmp[kTmpRuns - nruns] = (run_len << 2) | run_op;
      run_op = op;
      run_len = 0;
      ++nruns;
    }
    ++run_len;
  }
  if (MODE == 0) {
    if (run_len && tmp && nruns <= kTmpRuns) tmp[kTmpRuns - nruns] = (run_len << 2) | run_op;
    p.cigar_len[id] = nruns;
  } else if (run_len) {
    --wpos;
    if (wpos < p.pool_cap) p.pool[wpos] = (run_len << 2) | run_op;
  }
}

// Pool fill for the common case: the count pass left pair p's runs at the END of its temp slot
// (tmp[kTmpRuns - len .. kTmpRuns)); copy them to pool[cigar_off[p] ..].
static __global__ void __launch_bounds__(256) cigar_gather(const uint32_t* __restrict__ tmp_runs,
                                                    const uint32_t* __restrict__ cigar_len,
                                                    const uint64_t* __restrict__ cigar_off,
                                                    uint32_t* __restrict__ pool, uint64_t pool_cap,
                                                    uint32_t base, uint32_t n) {
  // 8 lanes per pair: each lane copies runs k, k+8, ..
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 3, sub = t & 7;
  if (i >= n) return;
  const uint32_t len = cigar_len[base + i];
  if (len == 0 || len > kTmpRuns) return;
  const uint64_t off = cigar_off[base + i];
  const uint32_t* src = tmp_runs + (uint64_t)i * kTmpRuns + (kTmpRuns - len);
  for (uint32_t k = sub; k < len; k += 8)
    if (off + k < pool_cap) pool[off + k] = src[k];
}

// ---- exclusive scan of cigar_len -> cigar_off (three small kernels, no library) -------------
constexpr int kScanBlock = 1024;

static __global__ void __launch_bounds__(kScanBlock) scan_block_sums(const uint32_t* __restrict__ len,
                                                              uint64_t* __restrict__ block_sums,
                                                              uint32_t n) {
  __shared__ uint64_t warp_sums[32];
  const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
  uint64_t v = i < n ? len[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint64_t s = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
  }
}

// single block: exclusive scan of block sums in place, starting from *carry; updates *carry
static __global__ void __launch_bounds__(kScanBlock) scan_block_offsets(uint64_t* __restrict__ block_sums,
                                                                 uint32_t n_blocks,
                                                                 uint64_t* __restrict__ carry) {
  __shared__ uint64_t sh[kScanBlock];
  __shared__ uint64_t running;
  if (threadIdx.x == 0) running = *carry;
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += kScanBlock) {
    const uint32_t i = base + threadIdx.x;
    const uint64_t v = i < n_blocks ? block_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
      const uint64_t a = threadIdx.x >= (uint32_t)o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += a;
      __syncthreads();
    }
    if (i < n_blocks) block_sums[i] = running + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) running += sh[kScanBlock - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *carry = running;
}

static __global__ void __launch_bounds__(kScanBlock) scan_apply(const uint32_t* __restrict__ len,
                                                         const uint64_t* __restrict__ block_sums,
                                                         uint64_t* __restrict__ off, uint32_t n) {
  __shared__ uint64_t warp_sums[32];
  const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
  const uint64_t v = i < n ? len[i] : 0;
  uint64_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t a = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= (uint32_t)o) incl += a;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint64_t w = warp_sums[threadIdx.x];
    uint64_t s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t a = __shfl_up_sync(0xffffffffu, s, o);
      if (threadIdx.x >= (uint32_t)o) s += a;
    }
    warp_sums[threadIdx.x] = s - w;  // exclusive
  }
  __syncthreads();
  if (i < n) off[i] = block_sums[blockIdx.x] + warp_sums[threadIdx.x >> 5] + incl - v;
}

}  // namespace sa
