// engine.cu -- C ABI (include/sa_engine.h) over the CUDA kernels.  sm_100a only, no CPU path.
//
// Replaces the reference's per-pair dispatch loop (/root/reference/src/main.rs:61-79) with
// batched launches: the pair list is cut into chunks whose packed traceback matrices fit in
// the device scratch, each chunk is filled (nw_affine_s16.cuh), walked (nw_walk.cuh) and its
// CIGARs are gathered into one pool in pair order.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/sa_engine.h"
#include "nw_affine_s16.cuh"
#include "nw_walk.cuh"

namespace {

constexpr int kK = 8;  // columns per strip (8 cells x 4 bits = one 32-bit traceback word)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct sa_resident {
  uint64_t n_pairs = 0;
  uint64_t residues_len = 0;
  uint8_t* residues = nullptr;
  uint64_t *q_off = nullptr, *d_off = nullptr;
  uint32_t *q_len = nullptr, *d_len = nullptr;
  // results (device)
  int32_t* score = nullptr;
  uint8_t* status = nullptr;
  uint32_t* cigar_len = nullptr;
  uint64_t* cigar_off = nullptr;
  uint32_t* pool = nullptr;
  uint64_t pool_cap = 0;
  uint64_t* carry = nullptr;  // total words used (device)
  // host-side shape summary
  uint32_t n1max = 0, n2max = 0;
  uint64_t cells = 0;
  std::vector<uint32_t> h_q_len, h_d_len;
  bool aligned = false;
  bool want_cigar = false;
};

struct sa_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  std::string err;
  DevBuf tb, tb2, end, end2, rerun_ids, misc, block_sums;
  uint32_t* h_count = nullptr;  // pinned
  sa_timing_t timing = {};
  int sm_count = 0;
  size_t smem_optin = 0;
  int force_g = 0;
  uint32_t ormask = 0x0F;
  size_t tb_budget = 0;
};

namespace {

sa_status_t fail(sa_engine* e, sa_status_t st, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return st;
}

#define CUDA_TRY(e, call)                                                                   \
  do {                                                                                      \
    cudaError_t err__ = (call);                                                             \
    if (err__ != cudaSuccess)                                                               \
      return fail(e, SA_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

sa_status_t ensure(sa_engine* e, DevBuf& b, size_t bytes) {
  if (b.cap >= bytes) return SA_OK;
  if (b.p) CUDA_TRY(e, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  cudaError_t err = cudaMalloc(&b.p, bytes);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(e, SA_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
  }
  b.cap = bytes;
  return SA_OK;
}

struct Geometry {
  int G = 4;
  uint32_t ng = 8, ppt = 16;
  uint32_t nstrips_pad = 0, n1pad = 0, tb_rows = 0;
  uint64_t tile_stride = 0;  // uint2 per tile
  uint32_t d_halfs = 0, q_halfs = 0;
  size_t smem_bytes = 0;
};

Geometry make_geometry(int G, uint32_t n1max, uint32_t n2max) {
  Geometry g;
  g.G = G;
  g.ng = 32 / G;
  g.ppt = 2 * g.ng;
  const uint32_t nstrips = (n1max + kK - 1) / kK;
  const uint32_t npass = (nstrips + G - 1) / G;
  g.nstrips_pad = std::max(1u, npass * G);
  g.n1pad = g.nstrips_pad * kK;
  g.tb_rows = std::max(1u, n2max);
  g.tile_stride = (uint64_t)g.nstrips_pad * g.tb_rows * g.ng;
  g.d_halfs = g.tb_rows * g.ng;  // one u16 (two residues) per row per pair-of-pairs
  g.q_halfs = g.n1pad * g.ng;
  g.smem_bytes = (size_t)g.tb_rows * g.ng * 8 + (size_t)(g.d_halfs + g.q_halfs) * 2;
  g.smem_bytes = (g.smem_bytes + 15) & ~(size_t)15;
  return g;
}

// Pick lanes-per-pair-of-pairs: least padded work among the configurations that leave at
// least 8 resident warps per SM; ties go to the smaller G (fewer shuffles, fewer ramp steps).
int choose_g(const sa_engine* e, uint32_t n1max, uint32_t n2max) {
  if (e->force_g) return e->force_g;
  int best = 0;
  double best_cost = 1e300;
  for (int G : {1, 2, 4, 8, 16, 32}) {
    const Geometry g = make_geometry(G, n1max, n2max);
    if (g.smem_bytes > e->smem_optin) continue;
    const double warps = std::min(32.0, std::floor(227.0 * 1024 / (double)(g.smem_bytes + 1024)));
    if (warps < 1) continue;
    const double npass = g.nstrips_pad / G;
    // steps per tile / pairs per tile, discounted when occupancy cannot cover latencies
    double cost = npass * (n2max + G - 1) / (double)g.ppt;
    const double occ = std::min(1.0, warps / 8.0);
    cost /= (0.55 + 0.45 * occ);
    if (cost < best_cost - 1e-12) {
      best_cost = cost;
      best = G;
    }
  }
  return best;
}

template <int G, uint32_t ORMASK>
sa_status_t launch_fill_m(sa_engine* e, const sa::AffineS16Params& p, const Geometry& g,
                          uint32_t n_tiles) {
  auto kern = sa::nw_affine_fill_s16<kK, G, ORMASK>;
  CUDA_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)g.smem_bytes));
  kern<<<n_tiles, 32, g.smem_bytes, e->stream>>>(p);
  CUDA_TRY(e, cudaGetLastError());
  e->timing.kernel_launches++;
  return SA_OK;
}

template <int G>
sa_status_t launch_fill(sa_engine* e, const sa::AffineS16Params& p, const Geometry& g,
                        uint32_t n_tiles) {
  switch (e->ormask) {
    case 0x00: return launch_fill_m<G, 0x00>(e, p, g, n_tiles);
    case 0x0F: return launch_fill_m<G, 0x0F>(e, p, g, n_tiles);
    case 0x1F: return launch_fill_m<G, 0x1F>(e, p, g, n_tiles);
    case 0x3F: return launch_fill_m<G, 0x3F>(e, p, g, n_tiles);
    case 0x7F: return launch_fill_m<G, 0x7F>(e, p, g, n_tiles);
    case 0xFF: return launch_fill_m<G, 0xFF>(e, p, g, n_tiles);
  }
  return fail(e, SA_E_ARG, "SA_ORMASK 0x%x has no instantiation", e->ormask);
}

sa_status_t launch_fill_g(sa_engine* e, const sa::AffineS16Params& p, const Geometry& g,
                          uint32_t n_tiles) {
  switch (g.G) {
    case 1: return launch_fill<1>(e, p, g, n_tiles);
    case 2: return launch_fill<2>(e, p, g, n_tiles);
    case 4: return launch_fill<4>(e, p, g, n_tiles);
    case 8: return launch_fill<8>(e, p, g, n_tiles);
    case 16: return launch_fill<16>(e, p, g, n_tiles);
    case 32: return launch_fill<32>(e, p, g, n_tiles);
  }
  return fail(e, SA_E_ARG, "bad G %d", g.G);
}

uint32_t pack2(uint32_t v) { return v | (v << 16); }

}  // namespace

extern "C" {

int sa_abi_version(void) { return SA_ABI_VERSION; }

const char* sa_last_error(const sa_engine_t* e) { return e ? e->err.c_str() : "null engine"; }

sa_status_t sa_engine_create(int device_id, sa_engine_t** out) {
  if (!out) return SA_E_ARG;
  *out = nullptr;
  sa_engine* e = new (std::nothrow) sa_engine();
  if (!e) return SA_E_NOMEM;
  *out = e;  // returned even on failure so the caller can read sa_last_error
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0)
    return fail(e, SA_E_CUDA, "no CUDA device: %s (this engine has no CPU fallback)",
                cudaGetErrorString(err));
  if (device_id < 0 || device_id >= n) return fail(e, SA_E_ARG, "device %d of %d", device_id, n);
  e->device = device_id;
  CUDA_TRY(e, cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CUDA_TRY(e, cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10)
    return fail(e, SA_E_CUDA, "device %s is sm_%d%d; this build is sm_100a only", prop.name,
                prop.major, prop.minor);
  e->sm_count = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  CUDA_TRY(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) CUDA_TRY(e, cudaEventCreate(&ev));
  CUDA_TRY(e, cudaMallocHost((void**)&e->h_count, 64));
  if (const char* s = getenv("SA_FORCE_G")) e->force_g = atoi(s);
  if (const char* s = getenv("SA_ORMASK")) e->ormask = (uint32_t)strtoul(s, nullptr, 0);
  if (const char* s = getenv("SA_TB_BUDGET_MB")) e->tb_budget = (size_t)atoll(s) << 20;
  return SA_OK;
}

sa_status_t sa_engine_destroy(sa_engine_t* e) {
  if (!e) return SA_OK;
  if (e->stream) {
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (DevBuf* b : {&e->tb, &e->tb2, &e->end, &e->end2, &e->rerun_ids, &e->misc, &e->block_sums})
      if (b->p) cudaFree(b->p);
    for (auto& ev : e->ev)
      if (ev) cudaEventDestroy(ev);
    if (e->h_count) cudaFreeHost(e->h_count);
    cudaStreamDestroy(e->stream);
  }
  delete e;
  return SA_OK;
}

void* sa_engine_stream(sa_engine_t* e) { return e ? (void*)e->stream : nullptr; }

sa_status_t sa_engine_synchronize(sa_engine_t* e) {
  if (!e) return SA_E_ARG;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return SA_OK;
}

sa_status_t sa_last_timing(const sa_engine_t* e, sa_timing_t* out) {
  if (!e || !out) return SA_E_ARG;
  *out = e->timing;
  return SA_OK;
}

void* sa_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void sa_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

sa_status_t sa_batch_free(sa_engine_t* e, sa_resident_t* r) {
  if (!r) return SA_OK;
  if (e) {
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
  }
  for (void* p : {(void*)r->residues, (void*)r->q_off, (void*)r->d_off, (void*)r->q_len,
                  (void*)r->d_len, (void*)r->score, (void*)r->status, (void*)r->cigar_len,
                  (void*)r->cigar_off, (void*)r->pool, (void*)r->carry})
    if (p) cudaFree(p);
  delete r;
  return SA_OK;
}

sa_status_t sa_batch_upload(sa_engine_t* e, const sa_batch_t* b, sa_resident_t** out) {
  if (!e || !b || !out) return SA_E_ARG;
  *out = nullptr;
  if (b->packing != 0) return fail(e, SA_E_UNSUPPORTED, "packing %u not supported in ABI v1", b->packing);
  if (b->n_pairs >= (1ull << 31)) return fail(e, SA_E_ARG, "n_pairs %llu too large", (unsigned long long)b->n_pairs);
  if (b->n_pairs && (!b->q_off || !b->q_len || !b->d_off || !b->d_len))
    return fail(e, SA_E_ARG, "null offset/length array");
  CUDA_TRY(e, cudaSetDevice(e->device));
  sa_resident* r = new (std::nothrow) sa_resident();
  if (!r) return SA_E_NOMEM;
  const uint64_t n = b->n_pairs;
  r->n_pairs = n;
  r->residues_len = b->residues_len;
  r->h_q_len.assign(b->q_len, b->q_len + n);
  r->h_d_len.assign(b->d_len, b->d_len + n);
  for (uint64_t i = 0; i < n; ++i) {
    if (b->q_off[i] + b->q_len[i] > b->residues_len || b->d_off[i] + b->d_len[i] > b->residues_len) {
      delete r;
      return fail(e, SA_E_ARG, "pair %llu reaches past residues_len", (unsigned long long)i);
    }
    r->n1max = std::max(r->n1max, b->q_len[i]);
    r->n2max = std::max(r->n2max, b->d_len[i]);
    r->cells += (uint64_t)b->q_len[i] * b->d_len[i];
  }
  const size_t n1 = std::max<uint64_t>(n, 1);
  auto alloc = [&](void** p, size_t bytes) { return cudaMalloc(p, std::max<size_t>(bytes, 16)); };
  cudaError_t err = cudaSuccess;
  if (err == cudaSuccess) err = alloc((void**)&r->residues, b->residues_len);
  if (err == cudaSuccess) err = alloc((void**)&r->q_off, n1 * 8);
  if (err == cudaSuccess) err = alloc((void**)&r->d_off, n1 * 8);
  if (err == cudaSuccess) err = alloc((void**)&r->q_len, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->d_len, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->score, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->status, n1);
  if (err == cudaSuccess) err = alloc((void**)&r->cigar_len, n1 * 4);
  if (err == cudaSuccess) err = alloc((void**)&r->cigar_off, n1 * 8);
  if (err == cudaSuccess) err = alloc((void**)&r->carry, 16);
  if (err != cudaSuccess) {
    cudaGetLastError();
    sa_batch_free(e, r);
    return fail(e, SA_E_NOMEM, "device allocation for the batch failed: %s", cudaGetErrorString(err));
  }
  cudaEventRecord(e->ev[0], e->stream);
  if (b->residues_len)
    cudaMemcpyAsync(r->residues, b->residues, b->residues_len, cudaMemcpyHostToDevice, e->stream);
  if (n) {
    cudaMemcpyAsync(r->q_off, b->q_off, n * 8, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(r->d_off, b->d_off, n * 8, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(r->q_len, b->q_len, n * 4, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(r->d_len, b->d_len, n * 4, cudaMemcpyHostToDevice, e->stream);
  }
  cudaEventRecord(e->ev[1], e->stream);
  cudaError_t last = cudaGetLastError();
  if (last != cudaSuccess) {
    sa_batch_free(e, r);
    return fail(e, SA_E_CUDA, "upload failed: %s", cudaGetErrorString(last));
  }
  e->timing = sa_timing_t{};
  e->timing.h2d_bytes = b->residues_len + n * 24;
  *out = r;
  return SA_OK;
}

sa_status_t sa_align_resident(sa_engine_t* e, sa_algo_t algo, sa_mode_t mode,
                              const sa_scheme_t* scheme, sa_resident_t* r, int want_cigar) {
  if (!e || !r) return SA_E_ARG;
  CUDA_TRY(e, cudaSetDevice(e->device));
  const uint64_t n = r->n_pairs;
  r->want_cigar = want_cigar != 0;
  e->timing.cells = r->cells;
  e->timing.pairs_rerun = 0;
  e->timing.kernel_launches = 0;
  e->timing.walk_ms = 0;
  if (n == 0) {
    r->aligned = true;
    return SA_OK;
  }
  if (mode != SA_MODE_GLOBAL) {
    // nw_affine:433-434, wfa.rs:26: every pair returns Err("not implemented")
    if (algo == SA_ALGO_NW_AFFINE || algo == SA_ALGO_WFA) {
      CUDA_TRY(e, cudaMemsetAsync(r->status, SA_NOT_IMPLEMENTED, n, e->stream));
      CUDA_TRY(e, cudaMemsetAsync(r->score, 0, n * 4, e->stream));
      CUDA_TRY(e, cudaMemsetAsync(r->cigar_len, 0, n * 4, e->stream));
      CUDA_TRY(e, cudaMemsetAsync(r->cigar_off, 0, n * 8, e->stream));
      CUDA_TRY(e, cudaMemsetAsync(r->carry, 0, 16, e->stream));
      r->aligned = true;
      return SA_OK;
    }
    return fail(e, SA_E_UNSUPPORTED, "mode %d for algo %d is not built yet", (int)mode, (int)algo);
  }
  if (algo != SA_ALGO_NW_AFFINE)
    return fail(e, SA_E_UNSUPPORTED, "algo %d is not built yet", (int)algo);

  sa_scheme_t sc = {5, -4, -8, -6};  // nw_affine.rs:15-20
  if (scheme) sc = *scheme;
  if (!(sc.match > sc.mismatch) || sc.gap_open > 0 || sc.gap_ext > 0 || sc.match < 0)
    return fail(e, SA_E_UNSUPPORTED, "scheme (%d,%d,%d,%d) outside the packed kernel's domain",
                sc.match, sc.mismatch, sc.gap_open, sc.gap_ext);
  const int pen = 2 * (sc.match - sc.mismatch), openp = -2 * sc.gap_open,
            extp = sc.match - 2 * sc.gap_ext;
  if (pen > 128 || extp <= 0)
    return fail(e, SA_E_UNSUPPORTED, "scheme magnitudes exceed the packed kernel's range");

  const int G = choose_g(e, r->n1max, r->n2max);
  if (!G) return fail(e, SA_E_UNSUPPORTED, "pair shape %u x %u needs more shared memory than one SM has", r->n1max, r->n2max);
  const Geometry g = make_geometry(G, r->n1max, r->n2max);
  const uint32_t need = sa::s16_min_value_bound(sc.match, sc.mismatch, sc.gap_open, sc.gap_ext, g.n1pad, r->n2max);
  if (need + 64 > sa::kBias)
    return fail(e, SA_E_UNSUPPORTED, "pair shape %u x %u exceeds the 16-bit packed range (long-pair kernel not built yet)", r->n1max, r->n2max);

  // ---- chunking: traceback scratch for the main fill and for the clean refill ------------
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(e, cudaMemGetInfo(&free_b, &total_b));
  size_t budget = e->tb_budget ? e->tb_budget : (size_t)((double)(free_b + e->tb.cap + e->tb2.cap) * 0.70);
  const size_t tile_bytes = (size_t)g.tile_stride * 8;
  const uint64_t n_tiles_all = (n + g.ppt - 1) / g.ppt;
  // main region gets 4/5 of the budget, refill region 1/5 (refills are a few % of pairs)
  uint64_t tiles_main = std::max<uint64_t>(1, std::min<uint64_t>(n_tiles_all, budget * 4 / 5 / tile_bytes));
  uint64_t tiles_re = std::max<uint64_t>(1, std::min<uint64_t>(tiles_main, std::max<uint64_t>(budget / 5 / tile_bytes, 1)));
  tiles_re = std::min<uint64_t>(tiles_re, std::max<uint64_t>(1, (tiles_main + 3) / 4));
  const uint64_t chunk_pairs = tiles_main * g.ppt;
  sa_status_t st;
  if ((st = ensure(e, e->tb, tiles_main * tile_bytes)) != SA_OK) return st;
  if ((st = ensure(e, e->tb2, tiles_re * tile_bytes)) != SA_OK) return st;
  if ((st = ensure(e, e->end, chunk_pairs * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->end2, tiles_re * g.ppt * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->rerun_ids, chunk_pairs * 4)) != SA_OK) return st;
  if ((st = ensure(e, e->misc, 256)) != SA_OK) return st;
  const uint32_t scan_blocks_max = (uint32_t)((chunk_pairs + sa::kScanBlock - 1) / sa::kScanBlock);
  if ((st = ensure(e, e->block_sums, (size_t)scan_blocks_max * 8)) != SA_OK) return st;
  uint32_t* d_rerun_count = (uint32_t*)e->misc.p;

  if (want_cigar && !r->pool) {
    // worst case is n1+n2 runs per pair; size for the common case and grow on demand
    r->pool_cap = std::max<uint64_t>(1024, n * 24);
    cudaError_t err = cudaMalloc((void**)&r->pool, r->pool_cap * 4);
    if (err != cudaSuccess) {
      cudaGetLastError();
      r->pool = nullptr;
      return fail(e, SA_E_NOMEM, "cigar pool allocation failed");
    }
  }

  sa::AffineS16Params fp{};
  fp.residues = r->residues;
  fp.q_off = r->q_off;
  fp.q_len = r->q_len;
  fp.d_off = r->d_off;
  fp.d_len = r->d_len;
  fp.tb_tile_stride = g.tile_stride;
  fp.tb_rows = g.tb_rows;
  fp.smem_bnd_rows = g.tb_rows;
  fp.smem_d_halfs = g.d_halfs;
  fp.pen2 = pack2((uint32_t)pen);
  fp.open2 = pack2((uint32_t)openp);
  fp.ext2 = pack2((uint32_t)extp);
  fp.origin = pack2(sa::kBias);
  const uint32_t row0_clean = sa::kBias - (uint32_t)(openp + (-2 * sc.gap_ext));

  sa::WalkParams wp{};
  wp.q_len = r->q_len;
  wp.d_len = r->d_len;
  wp.tb_tile_stride = g.tile_stride;
  wp.tb_rows = g.tb_rows;
  wp.ng = g.ng;
  wp.match = sc.match;
  wp.open = sc.gap_open;
  wp.ext = sc.gap_ext;
  wp.score = r->score;
  wp.status = r->status;
  wp.cigar_len = r->cigar_len;
  wp.cigar_off = r->cigar_off;
  wp.pool = r->pool;
  wp.pool_cap = r->pool_cap;
  wp.rerun_ids = (uint32_t*)e->rerun_ids.p;
  wp.rerun_count = d_rerun_count;

  CUDA_TRY(e, cudaMemsetAsync(r->carry, 0, 16, e->stream));
  CUDA_TRY(e, cudaEventRecord(e->ev[2], e->stream));

  for (uint64_t base = 0; base < n; base += chunk_pairs) {
    const uint32_t cn = (uint32_t)std::min<uint64_t>(chunk_pairs, n - base);
    const uint32_t ctiles = (cn + g.ppt - 1) / g.ppt;
    // 1. fill with the panic bonus on
    fp.pair_ids = nullptr;
    fp.pair_base = (uint32_t)base;
    fp.n_launch_pairs = cn;
    fp.tb = (uint2*)e->tb.p;
    fp.end = (uint32_t*)e->end.p;
    fp.row0 = pack2(row0_clean + 1);
    CUDA_TRY(e, cudaMemsetAsync(d_rerun_count, 0, 4, e->stream));
    CUDA_TRY(e, cudaEventRecord(e->ev[6], e->stream));
    if ((st = launch_fill_g(e, fp, g, ctiles)) != SA_OK) return st;
    CUDA_TRY(e, cudaEventRecord(e->ev[7], e->stream));
    // 2. classify + count
    wp.pair_ids = nullptr;
    wp.pair_base = (uint32_t)base;
    wp.n_launch_pairs = cn;
    wp.n_launch_dev = nullptr;
    wp.tb = (const uint2*)e->tb.p;
    wp.end = (const uint32_t*)e->end.p;
    wp.phase = 0;
    sa::nw_affine_walk<0><<<(cn + 127) / 128, 128, 0, e->stream>>>(wp);
    CUDA_TRY(e, cudaGetLastError());
    e->timing.kernel_launches++;
    // 3. clean refill of the pairs whose end cell carries the bonus
    CUDA_TRY(e, cudaMemcpyAsync(e->h_count, d_rerun_count, 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    const uint32_t n_re = *e->h_count;
    e->timing.pairs_rerun += n_re;
    {
      float fms = 0;
      if (cudaEventElapsedTime(&fms, e->ev[6], e->ev[7]) == cudaSuccess) e->timing.walk_ms += fms;
    }
    const uint32_t re_chunk = (uint32_t)(tiles_re * g.ppt);
    struct ReLaunch { uint32_t off, cnt; };
    std::vector<ReLaunch> re_launches;
    // The refill region may be smaller than the queue: refill, count and (after the scan)
    // write in slices.  Writing needs the offsets of the whole chunk, so when more than one
    // slice is needed the count pass runs per slice first and the fills are repeated for the
    // write pass.
    for (uint32_t off = 0; off < n_re; off += re_chunk)
      re_launches.push_back({off, std::min(re_chunk, n_re - off)});
    auto refill = [&](const ReLaunch& rl, bool do_fill) -> sa_status_t {
      fp.pair_ids = (const uint32_t*)e->rerun_ids.p + rl.off;
      fp.pair_base = 0;
      fp.n_launch_pairs = rl.cnt;
      fp.tb = (uint2*)e->tb2.p;
      fp.end = (uint32_t*)e->end2.p;
      fp.row0 = pack2(row0_clean);
      if (do_fill) {
        sa_status_t s2 = launch_fill_g(e, fp, g, (rl.cnt + g.ppt - 1) / g.ppt);
        if (s2 != SA_OK) return s2;
      }
      wp.pair_ids = fp.pair_ids;
      wp.pair_base = 0;
      wp.n_launch_pairs = rl.cnt;
      wp.tb = (const uint2*)e->tb2.p;
      wp.end = (const uint32_t*)e->end2.p;
      wp.phase = 1;
      return SA_OK;
    };
    for (const ReLaunch& rl : re_launches) {
      if ((st = refill(rl, true)) != SA_OK) return st;
      sa::nw_affine_walk<0><<<(rl.cnt + 127) / 128, 128, 0, e->stream>>>(wp);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches++;
    }
    // 4. offsets for this chunk (continuing from the previous chunk's total)
    const uint32_t sb = (cn + sa::kScanBlock - 1) / sa::kScanBlock;
    sa::scan_block_sums<<<sb, sa::kScanBlock, 0, e->stream>>>(r->cigar_len + base, (uint64_t*)e->block_sums.p, cn);
    sa::scan_block_offsets<<<1, sa::kScanBlock, 0, e->stream>>>((uint64_t*)e->block_sums.p, sb, r->carry);
    sa::scan_apply<<<sb, sa::kScanBlock, 0, e->stream>>>(r->cigar_len + base, (const uint64_t*)e->block_sums.p, r->cigar_off + base, cn);
    CUDA_TRY(e, cudaGetLastError());
    e->timing.kernel_launches += 3;
    if (want_cigar) {
      // grow the pool if this chunk does not fit
      CUDA_TRY(e, cudaMemcpyAsync(e->h_count + 2, r->carry, 8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      uint64_t used;
      memcpy(&used, e->h_count + 2, 8);
      if (used > r->pool_cap) {
        const uint64_t remaining_pairs = n - (base + cn);
        const uint64_t new_cap = used + remaining_pairs * 24 + 1024;
        uint32_t* np = nullptr;
        cudaError_t err = cudaMalloc((void**)&np, new_cap * 4);
        if (err != cudaSuccess) {
          cudaGetLastError();
          return fail(e, SA_E_NOMEM, "cigar pool growth to %llu words failed", (unsigned long long)new_cap);
        }
        // only earlier chunks' words exist so far
        uint64_t prev_used = 0;
        if (base) {
          CUDA_TRY(e, cudaMemcpy(&prev_used, r->cigar_off + base, 8, cudaMemcpyDeviceToHost));
          CUDA_TRY(e, cudaMemcpyAsync(np, r->pool, prev_used * 4, cudaMemcpyDeviceToDevice, e->stream));
          CUDA_TRY(e, cudaStreamSynchronize(e->stream));
        }
        cudaFree(r->pool);
        r->pool = np;
        r->pool_cap = new_cap;
      }
      wp.pool = r->pool;
      wp.pool_cap = r->pool_cap;
      // 5. write pass: main region, then each refill slice
      wp.pair_ids = nullptr;
      wp.pair_base = (uint32_t)base;
      wp.n_launch_pairs = cn;
      wp.tb = (const uint2*)e->tb.p;
      wp.end = (const uint32_t*)e->end.p;
      wp.phase = 0;
      sa::nw_affine_walk<1><<<(cn + 127) / 128, 128, 0, e->stream>>>(wp);
      CUDA_TRY(e, cudaGetLastError());
      e->timing.kernel_launches++;
      for (const ReLaunch& rl : re_launches) {
        // with a single slice the refill region still holds its traceback matrix
        if ((st = refill(rl, re_launches.size() > 1)) != SA_OK) return st;
        sa::nw_affine_walk<1><<<(rl.cnt + 127) / 128, 128, 0, e->stream>>>(wp);
        CUDA_TRY(e, cudaGetLastError());
        e->timing.kernel_launches++;
      }
    }
  }
  CUDA_TRY(e, cudaEventRecord(e->ev[3], e->stream));
  r->aligned = true;
  return SA_OK;
}

sa_status_t sa_resident_download(sa_engine_t* e, sa_resident_t* r, sa_result_t* res) {
  if (!e || !r || !res) return SA_E_ARG;
  if (!r->aligned) return fail(e, SA_E_ARG, "sa_align_resident has not run on this batch");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const uint64_t n = r->n_pairs;
  uint64_t used = 0;
  CUDA_TRY(e, cudaEventRecord(e->ev[4], e->stream));
  if (n) {
    CUDA_TRY(e, cudaMemcpyAsync(&used, r->carry, 8, cudaMemcpyDeviceToHost, e->stream));
    if (res->score) CUDA_TRY(e, cudaMemcpyAsync(res->score, r->score, n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (res->status) CUDA_TRY(e, cudaMemcpyAsync(res->status, r->status, n, cudaMemcpyDeviceToHost, e->stream));
    if (res->cigar_len) CUDA_TRY(e, cudaMemcpyAsync(res->cigar_len, r->cigar_len, n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (res->cigar_off) CUDA_TRY(e, cudaMemcpyAsync(res->cigar_off, r->cigar_off, n * 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  }
  res->cigar_used = used;
  uint64_t d2h = n * 17 + 8;
  sa_status_t rc = SA_OK;
  if (res->cigar && r->want_cigar && used) {
    if (used > res->cigar_capacity) {
      rc = fail(e, SA_E_CIGAR_CAPACITY, "cigar pool needs %llu words, capacity is %llu",
                (unsigned long long)used, (unsigned long long)res->cigar_capacity);
    } else {
      CUDA_TRY(e, cudaMemcpyAsync(res->cigar, r->pool, used * 4, cudaMemcpyDeviceToHost, e->stream));
      d2h += used * 4;
    }
  }
  CUDA_TRY(e, cudaEventRecord(e->ev[5], e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]) == cudaSuccess) e->timing.fill_ms = ms;
  if (cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]) == cudaSuccess) e->timing.d2h_ms = ms;
  if (cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]) == cudaSuccess) e->timing.h2d_ms = ms;
  cudaGetLastError();
  e->timing.d2h_bytes = d2h;
  e->timing.total_ms = e->timing.h2d_ms + e->timing.fill_ms + e->timing.d2h_ms;
  return rc;
}

sa_status_t sa_align_batch(sa_engine_t* e, sa_algo_t algo, sa_mode_t mode,
                           const sa_scheme_t* scheme, const sa_batch_t* batch,
                           sa_result_t* result) {
  if (!e || !batch || !result) return SA_E_ARG;
  sa_resident_t* r = nullptr;
  sa_status_t st = sa_batch_upload(e, batch, &r);
  if (st != SA_OK) return st;
  st = sa_align_resident(e, algo, mode, scheme, r, result->cigar != nullptr && result->cigar_capacity > 0);
  if (st == SA_OK) st = sa_resident_download(e, r, result);
  sa_batch_free(e, r);
  return st;
}

sa_status_t sa_partition_lpt(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs,
                             int n_parts, int32_t* part) {
  if (n_parts < 1 || (n_pairs && (!q_len || !d_len || !part))) return SA_E_ARG;
  // Greedy LPT on n1*n2.  Equal-cost pairs are dealt in index order, so the result is
  // deterministic and, for uniform batches, contiguous-cyclic.
  std::vector<uint64_t> order(n_pairs);
  for (uint64_t i = 0; i < n_pairs; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
    return (uint64_t)q_len[a] * d_len[a] > (uint64_t)q_len[b] * d_len[b];
  });
  std::vector<uint64_t> load(n_parts, 0);
  for (uint64_t k = 0; k < n_pairs; ++k) {
    const uint64_t i = order[k];
    int best = 0;
    for (int p = 1; p < n_parts; ++p)
      if (load[p] < load[best]) best = p;
    part[i] = best;
    load[best] += (uint64_t)q_len[i] * d_len[i] + 1;
  }
  return SA_OK;
}

}  // extern "C"
