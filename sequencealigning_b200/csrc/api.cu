// api.cu -- the C ABI of include/sa_engine.h: argument checks, dispatch to the single-device engine
// (engine.cu) or the multi-device front (multi.cu), and the exception fence: nothing C++ crosses
// the boundary; a failed host allocation comes back as SA_E_NOMEM.
#include <exception>
#include <new>

#include "engine_internal.h"

using namespace sa_host;

namespace {

template <class F>
sa_status_t fenced(sa_engine* e, F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(e, SA_E_NOMEM, "host allocation failed");
  } catch (const std::exception& ex) {
    return fail(e, SA_E_NOMEM, "host error: %s", ex.what());
  }
}

inline bool is_multi(const sa_engine* e) { return e && e->front != nullptr; }

}  // namespace

extern "C" {

int sa_abi_version(void) { return SA_ABI_VERSION; }

const char* sa_last_error(const sa_engine_t* e) { return e ? e->err.c_str() : "null engine"; }

sa_status_t sa_engine_create(int device_id, sa_engine_t** out) {
  return fenced(nullptr, [&] { return sd_create(device_id, out); });
}

sa_status_t sa_engine_create_multi(const int* device_ids, int n_devices, sa_engine_t** out) {
  return fenced(nullptr, [&] { return md_create(device_ids, n_devices, out); });
}

int sa_engine_device_count(const sa_engine_t* e) {
  if (!e) return 0;
  return is_multi(e) ? (int)e->children.size() : 1;
}

sa_status_t sa_engine_destroy(sa_engine_t* e) {
  if (!e) return SA_OK;
  if (is_multi(e) || !e->children.empty()) {
    md_destroy(e);
    delete e;
    return SA_OK;
  }
  return sd_destroy(e);
}

void* sa_engine_stream(sa_engine_t* e) {
  if (!e) return nullptr;
  return is_multi(e) ? (e->children.empty() ? nullptr : (void*)e->children[0]->stream) : (void*)e->stream;
}

sa_status_t sa_engine_synchronize(sa_engine_t* e) {
  if (!e) return SA_E_ARG;
  if (!is_multi(e)) return sd_synchronize(e);
  for (sa_engine* c : e->children) {
    const sa_status_t st = sd_synchronize(c);
    if (st != SA_OK) return fail(e, st, "device %d: %s", c->device, c->err.c_str());
  }
  return SA_OK;
}

sa_status_t sa_last_timing(const sa_engine_t* e, sa_timing_t* out) {
  if (!e || !out) return SA_E_ARG;
  *out = e->timing;
  return SA_OK;
}

sa_status_t sa_last_shards(const sa_engine_t* e, sa_shard_info_t* out, int cap, int* n_out) {
  if (!e) return SA_E_ARG;
  if (is_multi(e)) return md_last_shards(e, out, cap, n_out);
  if (n_out) *n_out = 1;
  if (out && cap >= 1) {
    *out = sa_shard_info_t{};
    out->device = e->device;
    out->contiguous = 1;
    out->cells = e->timing.cells;
    out->h2d_bytes = e->timing.h2d_bytes;
    out->d2h_bytes = e->timing.d2h_bytes;
    out->kernel_launches = e->timing.kernel_launches;
    out->device_ms = e->timing.kernels_ms;
  }
  return SA_OK;
}

sa_status_t sa_align_batch(sa_engine_t* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                           const sa_batch_t* batch, sa_result_t* result) {
  if (!e) return SA_E_ARG;
  return fenced(e, [&] {
    return is_multi(e) ? md_align_batch(e, algo, mode, scheme, batch, result)
                       : sd_align_batch(e, algo, mode, scheme, batch, result, 0);
  });
}

// The device-resident entry points are single-device (a benchmark/rescoring aid): on a multi-device
// engine they say so instead of silently using one GPU.
sa_status_t sa_batch_upload(sa_engine_t* e, const sa_batch_t* batch, sa_resident_t** out) {
  if (!e) return SA_E_ARG;
  if (is_multi(e)) return fail(e, SA_E_UNSUPPORTED, "resident batches are single-device; use sa_align_batch");
  return fenced(e, [&] { return sd_batch_upload(e, batch, out); });
}

sa_status_t sa_batch_free(sa_engine_t* e, sa_resident_t* r) {
  if (e && is_multi(e)) return fail(e, SA_E_UNSUPPORTED, "resident batches are single-device");
  return sd_batch_free(e, r);
}

sa_status_t sa_align_resident(sa_engine_t* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                              sa_resident_t* r, int want_cigar) {
  if (!e) return SA_E_ARG;
  if (is_multi(e)) return fail(e, SA_E_UNSUPPORTED, "resident batches are single-device");
  return fenced(e, [&] { return sd_align_resident(e, algo, mode, scheme, r, want_cigar); });
}

sa_status_t sa_resident_download(sa_engine_t* e, sa_resident_t* r, sa_result_t* result) {
  if (!e) return SA_E_ARG;
  if (is_multi(e)) return fail(e, SA_E_UNSUPPORTED, "resident batches are single-device");
  return fenced(e, [&] { return sd_resident_download(e, r, result); });
}

sa_status_t sa_affine_count_cooptimal(sa_engine_t* e, const sa_scheme_t* scheme, const sa_batch_t* batch,
                                      int64_t* counts) {
  if (!e) return SA_E_ARG;
  sa_engine* t = is_multi(e) ? e->children[0] : e;  // a side API: one device is enough
  const sa_status_t st = fenced(t, [&] { return sd_count_cooptimal(t, scheme, batch, counts); });
  if (st != SA_OK && t != e) e->err = t->err;
  return st;
}

int64_t sa_affine_all_alignments(sa_engine_t* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                                 const sa_scheme_t* scheme, uint64_t max_alignments, char* buf, size_t cap,
                                 uint64_t* n_printed, int32_t* panicked) {
  if (!e) return SA_E_ARG;
  sa_engine* t = is_multi(e) ? e->children[0] : e;
  int64_t r;
  try {
    r = sd_all_alignments(t, seq1, n1, seq2, n2, scheme, max_alignments, buf, cap, n_printed, panicked);
  } catch (const std::exception& ex) {
    r = fail(t, SA_E_NOMEM, "host allocation failed: %s", ex.what());
  }
  if (r < 0 && t != e) e->err = t->err;
  return r;
}

int64_t sa_linear_all_hits(sa_engine_t* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2, int local,
                           const sa_scheme_t* scheme, uint64_t max_hits, char* buf, size_t cap, uint64_t* n_printed) {
  if (!e) return SA_E_ARG;
  sa_engine* t = is_multi(e) ? e->children[0] : e;
  int64_t r;
  try {
    r = sd_linear_all_hits(t, seq1, n1, seq2, n2, local, scheme, max_hits, buf, cap, n_printed);
  } catch (const std::exception& ex) {
    r = fail(t, SA_E_NOMEM, "host allocation failed: %s", ex.what());
  }
  if (r < 0 && t != e) e->err = t->err;
  return r;
}

int64_t sa_wfa_reference_stdout(sa_engine_t* e, const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                                char* buf, size_t cap, int32_t* status) {
  if (!e) return SA_E_ARG;
  sa_engine* t = is_multi(e) ? e->children[0] : e;
  int64_t r;
  try {
    r = sd_wfa_stdout(t, seq1, n1, seq2, n2, buf, cap, status);
  } catch (const std::exception& ex) {
    r = fail(t, SA_E_NOMEM, "host allocation failed: %s", ex.what());
  }
  if (r < 0 && t != e) e->err = t->err;
  return r;
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

sa_status_t sa_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return SA_E_ARG;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    return SA_E_CUDA;
  }
  return SA_OK;
}

sa_status_t sa_host_unregister(void* p) {
  if (!p) return SA_E_ARG;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return SA_E_CUDA;
  }
  return SA_OK;
}

}  // extern "C"
