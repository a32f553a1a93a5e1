// nw_order.cuh -- device-side ordering of a segment's pairs by their number of DP rows.
//
// The 8-64 pairs of a warp tile all run to the tile's longest pair (nw_affine_s16.cuh), so read sets whose
// lengths differ by a few residues (indels) pay ~2-5 % of padded rows.  Ordering the pairs by rows removes
// that.  Done on the host it costs more than it saves (a counting sort of 128 Ki pairs is ~0.4 ms on the
// critical path of a 1 ms segment); done here it is one tiny launch on the segment's own stream and the host
// never sees the permutation (results are indexed by pair id).
//
// The sort is WINDOWED and STABLE: each block orders its own 1024 consecutive pairs, equal keys keep their
// order.  Tiles are then made of pairs that are neighbours in memory (a global sort with atomics scatters
// them: the sequence loads and the walks' result stores lose their locality and the step gets slower, measured),
// and a window of 1024 reads still holds tens of pairs per row count, so almost every tile is uniform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sa {

constexpr uint32_t kOrderWindow = 1024;

// order[w * 1024 + pos] = id of the pair that comes pos-th in window w of the segment (launch index -> pair id).
// Dynamic shared memory: nbins words.
__global__ void __launch_bounds__(kOrderWindow) order_window(const uint32_t* __restrict__ rows, uint32_t pair_base, uint32_t n,
                                                             uint32_t nbins, uint32_t* __restrict__ order) {
  extern __shared__ uint32_t cursor[];  // per key: first the count, then the next free position in the window
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t i = blockIdx.x * kOrderWindow + t;
  const bool valid = i < n;
  const uint32_t key = valid ? min(rows[pair_base + i], nbins - 1) : nbins - 1;
  for (uint32_t k = t; k < nbins; k += kOrderWindow) cursor[k] = 0;
  if (t == 0) carry = 0;
  __syncthreads();
  if (valid) atomicAdd(&cursor[key], 1u);
  __syncthreads();
  // exclusive scan of the counts, 1024 bins per trip
  for (uint32_t base = 0; base < nbins; base += kOrderWindow) {
    const uint32_t k = base + t;
    const uint32_t v = k < nbins ? cursor[k] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += u;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= (uint32_t)o) w += u;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    if (k < nbins) cursor[k] = carry + (warp ? warp_sums[warp - 1] : 0u) + incl - v;
    __syncthreads();
    if (t == 0) carry += warp_sums[31];
    __syncthreads();
  }
  // stable placement: the warps take their turns in order; inside a warp the lanes with one key rank by lane
  const uint32_t same = __match_any_sync(0xffffffffu, key);
  const uint32_t rank = __popc(same & ((1u << lane) - 1u));
  for (uint32_t w = 0; w < kOrderWindow / 32; ++w) {
    if (warp == w && valid) {
      const uint32_t pos = cursor[key] + rank;
      order[blockIdx.x * kOrderWindow + pos] = pair_base + i;
    }
    __syncthreads();
    if (warp == w && valid && rank == 0) cursor[key] += __popc(same);  // (invalid lanes carry the last key: they sit at the end)
    __syncthreads();
  }
}

}  // namespace sa
