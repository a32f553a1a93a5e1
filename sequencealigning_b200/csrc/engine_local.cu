// engine_local.cu -- host side of the linear aligner's LOCAL mode (nw_local.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine_internal.h"
#include "engine_util.h"
#include "nw_walk.cuh"
#include "nw_local.cuh"

namespace sa_host {

// ---------------------------------------------------------------------------------------------
// Linear NW, LOCAL mode (needleman_wunsch.rs:88-89, :107-111, :256-272): nw_local.cuh, one warp
// per pair.  Pairs go through in chunks that bound the CIGAR staging; per chunk: fill + argmax +
// walk in one kernel, scan of the CIGAR lengths (offsets continue across chunks), gather.
// ---------------------------------------------------------------------------------------------
template <int K>
sa_status_t launch_local(sa_engine* e, const sa::LocalParams& lp, uint32_t blocks, size_t smem, cudaStream_t sx) {
  auto kern = sa::nw_linear_local_kernel<K>;
  size_t& configured = e->smem_configured[(const void*)kern];
  if (smem > configured) {
    CUDA_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)std::min(e->smem_optin, std::max<size_t>(smem, 48 * 1024))));
    configured = std::max<size_t>(smem, 48 * 1024);
  }
  kern<<<blocks, 32 * sa::kLocalWarps, smem, sx>>>(lp);
  CUDA_TRY(e, cudaGetLastError());
  e->timing.kernel_launches++;
  return SA_OK;
}

sa_status_t run_linear_local(sa_engine* e, DeviceBatch& db, uint64_t n, const uint32_t* h_q_len,
                             const uint32_t* h_d_len, const sa_scheme_t& sc, bool want_cigar,
                             const sa_batch_t* in, sa_result_t* out, uint64_t* used_out) {
  sa_status_t st;
  *used_out = 0;
  cudaStream_t sx = e->stream;
  if (in) {
    uint64_t max_end = 0;
    for (uint64_t p = 0; p < n; ++p) {
      max_end = std::max(max_end, std::max(view_end(in->q_off[p], h_q_len[p]), view_end(in->d_off[p], h_d_len[p])));
      e->timing.cells += (uint64_t)h_q_len[p] * h_d_len[p];
    }
    if (max_end > (in->packing ? in->residues_len * 4 : in->residues_len))
      return fail(e, SA_E_ARG, "a pair reaches past residues_len");
    CUDA_TRY(e, cudaMemcpyAsync(db.residues, in->residues, in->residues_len, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(db.q_off, in->q_off, n * 8, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(db.d_off, in->d_off, n * 8, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(db.q_len, in->q_len, n * 4, cudaMemcpyHostToDevice, sx));
    CUDA_TRY(e, cudaMemcpyAsync(db.d_len, in->d_len, n * 4, cudaMemcpyHostToDevice, sx));
    e->timing.h2d_bytes += in->residues_len + n * 24;
  }
  if ((st = ensure(e, e->misc, 256)) != SA_OK) return st;
  uint32_t* d_next = (uint32_t*)e->misc.p + 8;
  CUDA_TRY(e, cudaMemsetAsync(db.carry, 0, 16, sx));
  if (db.pool_base) {  // offsets continue from the slice's start (multi-device calls)
    memcpy(e->h_count + 12, &db.pool_base, 8);
    CUDA_TRY(e, cudaMemcpyAsync(db.carry, e->h_count + 12, 8, cudaMemcpyHostToDevice, sx));
  }
  CUDA_TRY(e, cudaEventRecord(e->ev_t0, sx));
  const size_t budget = e->tb_budget ? e->tb_budget : (size_t)8 << 30;  // traceback scratch of the resident warps
  const uint64_t stage_cap = (uint64_t)1 << 29;                          // CIGAR staging words per chunk (2 GB)
  std::vector<uint64_t> runs_end;
  for (uint64_t base = 0; base < n;) {
    // chunk: as many pairs as the staging holds; shape maxima pick the kernel form
    uint32_t n1max = 0, n2max = 0;
    uint64_t words = 0, cnt = 0;
    runs_end.clear();
    while (base + cnt < n && cnt < (1u << 22)) {
      const uint64_t p = base + cnt;
      const uint64_t w = (uint64_t)h_q_len[p] + h_d_len[p];
      if (cnt && words + w > stage_cap) break;
      words += w;
      runs_end.push_back(words);
      n1max = std::max(n1max, h_q_len[p]);
      n2max = std::max(n2max, h_d_len[p]);
      ++cnt;
    }
    const int K = n2max <= 160 ? 5 : (n2max <= 256 ? 8 : 16);
    const uint32_t wbytes = K == 16 ? 4 : 2;
    const uint64_t ns = ((uint64_t)n2max + K - 1) / K;
    const uint64_t tb_words = std::max<uint64_t>((uint64_t)n1max * ns, 1);
    // the matrix of a pair stays in shared memory when four warps' worth leaves >= 2 blocks per SM
    const bool in_smem = want_cigar && tb_words * wbytes * sa::kLocalWarps <= 96 * 1024;
    const size_t smem = in_smem ? (size_t)tb_words * wbytes * sa::kLocalWarps : 0;
    uint64_t warps = std::min<uint64_t>(cnt, (uint64_t)e->sm_count * 16);
    if (in_smem) warps = std::min<uint64_t>(warps, (uint64_t)e->sm_count * sa::kLocalWarps * std::max<uint64_t>(1, (224 * 1024) / (smem + 1024)));
    const uint64_t tb_stride = in_smem || !want_cigar ? 16 : ((tb_words * wbytes + 15) & ~(uint64_t)15);
    bool omit = false;
    if (!in_smem && want_cigar) {
      if (tb_stride > budget) omit = true;  // not even one pair's matrix fits: score and end cell only
      else warps = std::min<uint64_t>(warps, std::max<uint64_t>(1, budget / tb_stride));
    }
    const uint32_t blocks = (uint32_t)((warps + sa::kLocalWarps - 1) / sa::kLocalWarps);
    const uint64_t gwarps = (uint64_t)blocks * sa::kLocalWarps;
    const bool tb_on = want_cigar && !omit;
    if ((st = ensure(e, e->tb2, (size_t)(gwarps * tb_stride))) != SA_OK) return st;
    if ((st = ensure(e, e->par_rows, (size_t)(gwarps * ((uint64_t)n1max + 2) * 8))) != SA_OK) return st;
    if (tb_on) {
      if ((st = ensure(e, e->par_bytes, (size_t)(words * 4 + 256))) != SA_OK) return st;
      if ((st = ensure(e, e->par_in, (size_t)(cnt * 8))) != SA_OK) return st;
      CUDA_TRY(e, cudaMemcpyAsync(e->par_in.p, runs_end.data(), cnt * 8, cudaMemcpyHostToDevice, sx));
    }
    sa::LocalParams lp{};
    lp.residues = db.residues;
    lp.q_off = db.q_off;
    lp.q_len = db.q_len;
    lp.d_off = db.d_off;
    lp.d_len = db.d_len;
    lp.packing = db.packing;
    lp.pair_base = (uint32_t)base;
    lp.n_launch_pairs = (uint32_t)cnt;
    lp.match = sc.match;
    lp.mismatch = sc.mismatch;
    lp.open = sc.gap_open;
    lp.ext = sc.gap_ext;
    lp.next_pair = d_next;
    lp.smem_words = in_smem ? (uint32_t)tb_words : 0;
    lp.tb = (uint8_t*)e->tb2.p;
    lp.tb_stride = tb_stride;
    lp.bnd = (int2*)e->par_rows.p;
    lp.bnd_stride = (uint64_t)n1max + 2;
    lp.runs = tb_on ? (uint32_t*)e->par_bytes.p : nullptr;
    lp.runs_end = tb_on ? (const uint64_t*)e->par_in.p : nullptr;
    lp.score = db.score;
    lp.status = db.status;
    lp.cigar_len = db.cigar_len;
    lp.end1 = db.end1;
    lp.end2 = db.end2;
    lp.omit_flag = (want_cigar && omit) ? (uint32_t)SA_ALIGNMENT_OMITTED : 0u;
    CUDA_TRY(e, cudaMemsetAsync(d_next, 0, 4, sx));
    if (K == 5) st = launch_local<5>(e, lp, blocks, smem, sx);
    else if (K == 8) st = launch_local<8>(e, lp, blocks, smem, sx);
    else st = launch_local<16>(e, lp, blocks, smem, sx);
    if (st != SA_OK) return st;
    const uint32_t cn = (uint32_t)cnt;
    const uint32_t sb = (cn + sa::kScanBlock - 1) / sa::kScanBlock;
    if ((st = ensure(e, e->block_sums, (size_t)sb * 8)) != SA_OK) return st;
    sa::scan_block_sums<<<sb, sa::kScanBlock, 0, sx>>>(db.cigar_len + base, (uint64_t*)e->block_sums.p, cn);
    sa::scan_block_offsets<<<1, sa::kScanBlock, 0, sx>>>((uint64_t*)e->block_sums.p, sb, db.carry);
    sa::scan_apply<<<sb, sa::kScanBlock, 0, sx>>>(db.cigar_len + base, (const uint64_t*)e->block_sums.p, db.cigar_off + base, cn);
    e->timing.kernel_launches += 3;
    if (tb_on) {
      sa::local_runs_to_pool<<<(cn + 3) / 4, 128, 0, sx>>>((uint32_t)base, cn, (const uint32_t*)e->par_bytes.p, (const uint64_t*)e->par_in.p,
                                                           db.cigar_len, db.cigar_off, db.pool, db.pool_cap);
      e->timing.kernel_launches++;
    }
    CUDA_TRY(e, cudaGetLastError());
    if (base + cnt < n) CUDA_TRY(e, cudaStreamSynchronize(sx));  // runs_end (a host vector) and the staging are reused
    base += cnt;
  }
  CUDA_TRY(e, cudaEventRecord(e->ev_t1, sx));
  CUDA_TRY(e, cudaMemcpyAsync(e->h_count + 2, db.carry, 8, cudaMemcpyDeviceToHost, sx));
  if (out) {
    auto cp = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
      if (!dst || !bytes) return cudaSuccess;
      e->timing.d2h_bytes += bytes;
      return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, sx);
    };
    CUDA_TRY(e, cp(out->score, db.score, n * 4));
    CUDA_TRY(e, cp(out->status, db.status, n));
    CUDA_TRY(e, cp(out->cigar_len, db.cigar_len, n * 4));
    CUDA_TRY(e, cp(out->cigar_off, db.cigar_off, n * 8));
    CUDA_TRY(e, cp(out->end1, db.end1, n * 4));
    CUDA_TRY(e, cp(out->end2, db.end2, n * 4));
  }
  CUDA_TRY(e, cudaStreamSynchronize(sx));
  memcpy(used_out, e->h_count + 2, 8);
  return SA_OK;
}

}  // namespace sa_host
