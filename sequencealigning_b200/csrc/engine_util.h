// engine_util.h -- small host helpers shared by the translation units of the engine (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "engine_internal.h"
#include "seg_scan.h"  // view_end / view_in_bounds, the host's pass over a segment

#define CUDA_TRY(e, call)                                                                   \
  do {                                                                                      \
    cudaError_t err__ = (call);                                                             \
    if (err__ != cudaSuccess)                                                               \
      return fail(e, SA_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

namespace sa_host {

// grow-only device buffer
inline sa_status_t ensure(sa_engine* e, DevBuf& b, size_t bytes) {
  if (b.cap >= bytes && b.p) return SA_OK;
  if (b.p) {
    CUDA_TRY(e, cudaDeviceSynchronize());
    CUDA_TRY(e, cudaFree(b.p));
  }
  b.p = nullptr;
  b.cap = 0;
  bytes = std::max<size_t>(bytes, 256);
  cudaError_t err = cudaMalloc(&b.p, bytes);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(e, SA_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
  }
  b.cap = bytes;
  return SA_OK;
}


// ---- the other host paths (their own translation units) -------------------------------------------------
// WFA, literal and standard mode (engine_wfa.cu)
sa_status_t run_wfa(sa_engine* e, DeviceBatch& db, uint64_t n, const uint32_t* h_q_len, const uint32_t* h_d_len,
                    const sa_scheme_t* scheme, bool literal, const sa_batch_t* in, sa_result_t* out);
// linear NW, local mode (engine_local.cu)
sa_status_t run_linear_local(sa_engine* e, DeviceBatch& db, uint64_t n, const uint32_t* h_q_len, const uint32_t* h_d_len,
                             const sa_scheme_t& sc, bool want_cigar, const sa_batch_t* in, sa_result_t* out, uint64_t* used_out);

}  // namespace sa_host
