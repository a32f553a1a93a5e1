// engine_internal.h -- what the translation units of libsa_engine.so share: the engine object and
// the single-device entry points (engine.cu) that the multi-device front (multi.cu) and the C ABI
// (api.cu) call.  Nothing here is part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/sa_engine.h"

namespace sa_host {
struct MultiFront;  // multi.cu: worker threads and shard bookkeeping of a multi-device engine
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// device views the kernels work on (either engine-owned staging or a resident batch)
struct DeviceBatch {
  uint8_t* residues = nullptr;
  uint64_t *q_off = nullptr, *d_off = nullptr;
  uint32_t *q_len = nullptr, *d_len = nullptr;
  int32_t* score = nullptr;
  uint8_t* status = nullptr;
  uint32_t* cigar_len = nullptr;
  uint64_t* cigar_off = nullptr;
  uint32_t *end1 = nullptr, *end2 = nullptr;  // local mode: start cell of the traceback (else unused)
  uint32_t* pool = nullptr;
  uint64_t pool_cap = 0;
  uint64_t* carry = nullptr;
  uint32_t packing = 0;
  // Multi-device calls give every device a slice [pool_base, pool_cap) of ONE caller-side pool:
  // cigar_off is absolute (the scan starts at pool_base) and `pool` is the device buffer shifted
  // down by pool_base words, so pool[cigar_off] is the right word on both sides.
  uint64_t pool_base = 0;
};

struct sa_resident {
  uint64_t n_pairs = 0;
  uint64_t residues_len = 0;
  DeviceBatch d;
  std::vector<uint32_t> h_q_len, h_d_len;
  uint64_t cells = 0;
  uint64_t used = 0;  // CIGAR words of the last alignment
  bool aligned = false;
  bool want_cigar = false;
  bool local = false;  // the last alignment ran in local mode (end cells live on the device)
  // segment plan of the last alignment (shapes do not change while the batch is resident):
  // opaque here, owned through the deleter
  void* plan = nullptr;
  void (*plan_free)(void*) = nullptr;
  uint64_t plan_key = 0;
  size_t plan_budget = 0;
  ~sa_resident() {
    if (plan && plan_free) plan_free(plan);
  }
};

struct sa_engine {
  int device = 0;
  cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_carry[2] = {nullptr, nullptr}, ev_t0 = nullptr, ev_t1 = nullptr;
  std::string err;
  // Two segments are in flight on the compute stream (the fill of segment i+1 is queued before
  // the host reads segment i's refill count), so per-segment scratch is double-buffered.
  struct LitBufs {  // scratch of the literal long-pair kernels (nw_general.cuh)
    DevBuf ids, meta, tb, rows, info, runs;
  };
  struct Slot {
    DevBuf tb, end, rerun_ids, tmp_runs, order;
    LitBufs lit;
    DevBuf f_meta, f_edges, f_tb, f_runs;  // tiled long pairs (nw_long.cuh)
    cudaStream_t stream = nullptr;  // stage A of alternating segments runs on its own stream, so
                                    // the next fill overlaps the tail of the previous one
    cudaStream_t fill_stream = nullptr;  // LOW priority: only the fill kernels.  The walks, scans and
                                         // copies of other segments then get SMs as fill CTAs retire,
                                         // instead of queueing behind a whole fill
    cudaEvent_t ev_count = nullptr, ev_f0 = nullptr, ev_f1 = nullptr, ev_bdone = nullptr, ev_w0 = nullptr;
    cudaEvent_t ev_l0 = nullptr, ev_l1 = nullptr, ev_l2 = nullptr;  // tiled long pairs: forward start / end, traceback end
  } slot[2];
  // scratch (grow-only)
  DevBuf tb2, end2, misc, block_sums, wfa_scratch, par_bytes, par_rows, par_in;
  // staging for sa_align_batch (grow-only)
  DevBuf b_res, b_qoff, b_doff, b_qlen, b_dlen, b_score, b_status, b_clen, b_coff, b_pool, b_carry, b_end1, b_end2;
  uint32_t* h_count = nullptr;  // pinned
  sa_timing_t timing = {};
  int sm_count = 0;
  size_t smem_optin = 0;
  int force_g = 0, force_k = 0;
  bool long_ckpt_always = false;  // SA_LONG_CKPT: checkpointed traceback for every long pair (tests)
  bool long_literal = false;      // SA_LONG_LITERAL: affine long pairs through the literal kernel only (tests)
  uint32_t long_s = 0, long_r = 0;  // SA_LONG_S / SA_LONG_R: tile shape of the tiled long-pair path (0 = auto)
  uint32_t long_cell = 2;           // SA_LONG_CELL: instruction mix of the score-only long-pair cell (nw_long.cuh long_cells)
  uint32_t long_minb = 4;           // SA_LONG_MINB: register allocation of nw_long_fwd (4 or 5 CTAs per SM)
  LitBufs fb_lit;                 // literal-kernel scratch for pairs the tiled path hands over
  uint32_t ormask = 0x00;
  uint32_t fill_minb = 0;           // SA_FILL_MINB: 16 = the 128-register build of the K = 19 fill forms, 1 = the full one, 0 = per form
  int scan_threads = 1;             // SA_SCAN_THREADS: host threads for the scan of a streamed call's first segments (seg_scan.h);
                                    // measured: 4 threads LOSE 3 % end to end (8.15 -> 8.40 ms: starting them costs more than the scan saves)
  uint64_t scan_mt_pairs = 1u << 18;  // ... segments that start below this pair index
  uint32_t walk_pf = 6;  // look-ahead of the traceback walks in steps (SA_WALK_PF; 0 = off)
  size_t tb_budget = 0;
  size_t budget_cached = 0;
  uint32_t seg_pairs = 131072;  // measured (1 M x 150 bp): 512 Ki 2672 / 256 Ki 2691 / 128 Ki 2696 GCUPS resident, e2e 1915 / 2126 / 2302
  std::map<const void*, size_t> smem_configured;  // kernel -> opted-in dynamic smem ON THIS DEVICE
  uint32_t seg_head = 65536;  // SA_SEG_HEAD: first segment of a call that streams from the host (sizes double from here)
  int sort_mode = 0;  // 0 auto, 1 always, 2 never (SA_SORT)
  bool seg_pairs_forced = false;
  // A multi-device engine (sa_engine_create_multi) owns no device itself: it shards a call over
  // its children (one single-device engine and one worker thread each) and gathers the results.
  std::vector<sa_engine*> children;
  sa_host::MultiFront* front = nullptr;
};


namespace sa_host {

sa_status_t fail(sa_engine* e, sa_status_t st, const char* fmt, ...);

// ---- single-device engine (engine.cu) ----------------------------------------------------------
sa_status_t sd_create(int device_id, sa_engine** out);
sa_status_t sd_destroy(sa_engine* e);
sa_status_t sd_synchronize(sa_engine* e);
// sa_align_batch on one device.  pool_base: first word of this device's slice of the caller's
// CIGAR pool (0 for a single-device call); res->cigar is the pool's start, res->cigar_capacity
// the slice's END, cigar_off and res->cigar_used come back absolute.
sa_status_t sd_align_batch(sa_engine* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                           const sa_batch_t* b, sa_result_t* res, uint64_t pool_base);
sa_status_t sd_batch_upload(sa_engine* e, const sa_batch_t* b, sa_resident_t** out);
sa_status_t sd_batch_free(sa_engine* e, sa_resident_t* r);
sa_status_t sd_align_resident(sa_engine* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                              sa_resident_t* r, int want_cigar);
sa_status_t sd_resident_download(sa_engine* e, sa_resident_t* r, sa_result_t* res);
sa_status_t sd_count_cooptimal(sa_engine* e, const sa_scheme_t* scheme, const sa_batch_t* b, int64_t* counts);
int64_t sd_all_alignments(sa_engine* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                          const sa_scheme_t* scheme, uint64_t max_alignments, char* buf, size_t cap,
                          uint64_t* n_printed, int32_t* panicked);

int64_t sd_linear_all_hits(sa_engine* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, int local,
                           const sa_scheme_t* scheme, uint64_t max_hits, char* buf, size_t cap, uint64_t* n_printed);
int64_t sd_wfa_stdout(sa_engine* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, char* buf, size_t cap,
                      int32_t* status_out);

// ---- multi-device front (multi.cu) ---------------------------------------------------------------
sa_status_t md_create(const int* device_ids, int n_devices, sa_engine** out);
void md_destroy(sa_engine* e);
sa_status_t md_align_batch(sa_engine* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                           const sa_batch_t* b, sa_result_t* res);
sa_status_t md_last_shards(const sa_engine* e, sa_shard_info_t* out, int cap, int* n_out);
sa_status_t plan_shards(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs, int n_parts,
                        uint64_t* begin, int32_t* part, int* contiguous);

}  // namespace sa_host
