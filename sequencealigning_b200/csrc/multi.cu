// multi.cu -- one engine over several GPUs: shard ONE pair list, run a worker + a single-device
// engine per GPU, put the results back in input order.  Host code only.
//
// The reference iterates one list of (query, db) pairs (/root/reference/src/main.rs:61-62) with
// no state shared between iterations, so the list shards freely: there is NO collective and no
// device-to-device traffic; every GPU streams its own pairs over its own PCIe link.
//
// Plan (sa_plan_shards): pairs weigh n1*n2 + 1.  When the cell-balanced CONTIGUOUS split leaves
// every part within 2 % of the mean -- any batch of many reads -- a shard is a range of the
// caller's arrays: nothing is gathered, every device DMAs straight from / to the caller's
// (pinned) buffers at the range's offset, and the CIGAR pool is cut into one region per device.
// Otherwise (few, uneven pairs) the shards are the greedy-LPT index sets of sa_partition_lpt,
// gathered into per-device arrays and scattered back.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <exception>
#include <mutex>
#include <new>
#include <thread>

#include "engine_internal.h"

namespace sa_host {

namespace {

// One long-lived thread per device: keeps the device's context warm and the call free of thread
// start-up; a job is a closure, wait() returns when it has run.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, idle = true, stop = false;

  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(mu);
      for (;;) {
        cv.wait(lk, [&] { return has_job || stop; });
        if (stop) return;
        std::function<void()> j = std::move(job);
        has_job = false;
        lk.unlock();
        j();
        lk.lock();
        idle = true;
        cv.notify_all();
      }
    });
  }
  void submit(std::function<void()> j) {
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(j);
    has_job = true;
    idle = false;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return idle; });
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

inline uint64_t weight(uint32_t a, uint32_t b) { return (uint64_t)a * b + 1; }

// Contiguous split balanced on the weights: block sums (in parallel when `par` runs closures
// concurrently), then boundaries by a walk over the blocks and inside the boundary block.
template <class Par>
void plan_contiguous(const uint32_t* q_len, const uint32_t* d_len, uint64_t n, int parts, uint64_t* begin,
                     uint64_t* part_weight, Par&& par) {
  const uint64_t nblk = std::max<uint64_t>(1, std::min<uint64_t>(n, (uint64_t)parts * 256));
  const uint64_t per = (n + nblk - 1) / std::max<uint64_t>(nblk, 1);
  std::vector<uint64_t> blk(nblk + 1, 0);
  // Shards of >= 512 Ki pairs: a block's sum is estimated from 1/16 of its pairs (one cache line of each length
  // array out of sixteen, at a per-block pseudo-random offset, so that a periodic list does not alias with it).  The pass over 2 x 4 bytes per
  // pair is host-memory bound (~1 ms per 10 M pairs), which is 3 % of an 8-device call; an estimate that is off by
  // a few blocks' worth of noise moves a boundary by a few thousand pairs out of a million.
  const bool sampled = per >= 1024 && n / (uint64_t)parts >= (1u << 19);
  par(nblk, [&](uint64_t b) {
    const uint64_t lo = std::min(n, b * per), hi = std::min(n, lo + per);
    uint64_t s = 0;
    if (sampled && hi - lo == per) {
      uint64_t cnt = 0;
      for (uint64_t at = lo + (b * 0x9E3779B97F4A7C15ull >> 33) % 240; at + 16 <= hi; at += 256, cnt += 16)
        for (uint64_t i = at; i < at + 16; ++i) s += weight(q_len[i], d_len[i]);
      s = (uint64_t)((unsigned __int128)s * per / cnt);
    } else {
      for (uint64_t i = lo; i < hi; ++i) s += weight(q_len[i], d_len[i]);
    }
    blk[b] = s;
  });
  uint64_t total = 0;
  for (uint64_t b = 0; b < nblk; ++b) total += blk[b];
  begin[0] = 0;
  uint64_t b = 0, acc = 0;  // acc = weight of the pairs before block b
  for (int k = 1; k < parts; ++k) {
    // first pair index at which the prefix weight reaches k/parts of the total
    const uint64_t want = (uint64_t)((unsigned __int128)total * (unsigned)k / (unsigned)parts);
    while (b < nblk && acc + blk[b] <= want) acc += blk[b++];
    uint64_t i = std::min(n, b * per), run = acc;
    const uint64_t hi = std::min(n, i + per);
    while (i < hi && run + weight(q_len[i], d_len[i]) <= want) run += weight(q_len[i], d_len[i]), ++i;
    begin[k] = std::max(begin[k - 1], i);
  }
  begin[parts] = n;
  if (part_weight) {
    for (int k = 0; k < parts; ++k) part_weight[k] = 0;
    // (weights per part: whole blocks from the sums, boundary blocks pair by pair)
    for (int k = 0; k < parts; ++k) {
      uint64_t i = begin[k];
      const uint64_t e = begin[k + 1];
      while (i < e) {
        const uint64_t bi = per ? i / per : 0;
        const uint64_t blo = bi * per, bhi = std::min(n, blo + per);
        if (i == blo && bhi <= e) {
          part_weight[k] += blk[bi];
          i = bhi;
        } else {
          const uint64_t stop = std::min(e, bhi);
          for (; i < stop; ++i) part_weight[k] += weight(q_len[i], d_len[i]);
        }
      }
    }
  }
}

void lpt(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs, int n_parts, int32_t* part) {
  // Greedy LPT on n1*n2.  Equal-cost pairs are dealt in index order, so the result is
  // deterministic and, for uniform batches, cyclic.
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
    load[best] += weight(q_len[i], d_len[i]);
  }
}

bool balanced(const uint64_t* part_weight, int parts) {
  uint64_t total = 0, worst = 0;
  for (int k = 0; k < parts; ++k) {
    total += part_weight[k];
    worst = std::max(worst, part_weight[k]);
  }
  // worst <= 1.02 * mean
  return (unsigned __int128)worst * parts * 100 <= (unsigned __int128)total * 102;
}

}  // namespace

struct MultiFront {
  std::vector<Worker*> workers;
  std::vector<sa_shard_info_t> shards;
};

sa_status_t plan_shards(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs, int n_parts,
                        uint64_t* begin, int32_t* part, int* contiguous) {
  if (n_parts < 1 || !begin || (n_pairs && (!q_len || !d_len))) return SA_E_ARG;
  std::vector<uint64_t> w(n_parts);
  plan_contiguous(q_len, d_len, n_pairs, n_parts, begin, w.data(), [](uint64_t nb, auto&& f) {
    for (uint64_t b = 0; b < nb; ++b) f(b);
  });
  const bool ok = n_parts == 1 || balanced(w.data(), n_parts);
  if (contiguous) *contiguous = ok ? 1 : 0;
  if (!ok && part) lpt(q_len, d_len, n_pairs, n_parts, part);
  if (ok && part)
    for (int k = 0; k < n_parts; ++k)
      for (uint64_t i = begin[k]; i < begin[k + 1]; ++i) part[i] = k;
  return SA_OK;
}

sa_status_t md_create(const int* device_ids, int n_devices, sa_engine** out) {
  if (!out) return SA_E_ARG;
  *out = nullptr;
  sa_engine* e = new (std::nothrow) sa_engine();
  if (!e) return SA_E_NOMEM;
  *out = e;  // returned even on failure so the caller can read sa_last_error
  if (!device_ids || n_devices < 1 || n_devices > 64) return fail(e, SA_E_ARG, "n_devices %d", n_devices);
  e->front = new MultiFront();
  e->device = device_ids[0];
  for (int k = 0; k < n_devices; ++k) {
    sa_engine* c = nullptr;
    const sa_status_t st = sd_create(device_ids[k], &c);
    if (st != SA_OK) {
      fail(e, st, "device %d: %s", device_ids[k], c ? c->err.c_str() : "allocation failed");
      if (c) sd_destroy(c);
      return st;
    }
    e->children.push_back(c);
    Worker* w = new Worker();
    w->start();
    e->front->workers.push_back(w);
  }
  e->sm_count = e->children[0]->sm_count;
  e->smem_optin = e->children[0]->smem_optin;
  return SA_OK;
}

void md_destroy(sa_engine* e) {
  if (!e) return;
  if (e->front) {
    for (Worker* w : e->front->workers) {
      w->shutdown();
      delete w;
    }
    delete e->front;
    e->front = nullptr;
  }
  for (sa_engine* c : e->children) sd_destroy(c);
  e->children.clear();
}

sa_status_t md_last_shards(const sa_engine* e, sa_shard_info_t* out, int cap, int* n_out) {
  if (!e) return SA_E_ARG;
  const int n = e->front ? (int)e->front->shards.size() : 0;
  if (n_out) *n_out = n;
  for (int k = 0; k < n && k < cap && out; ++k) out[k] = e->front->shards[k];
  return SA_OK;
}

sa_status_t md_align_batch(sa_engine* e, sa_algo_t algo, sa_mode_t mode, const sa_scheme_t* scheme,
                           const sa_batch_t* b, sa_result_t* res) {
  if (!e || !b || !res || !e->front) return SA_E_ARG;
  const int N = (int)e->children.size();
  if (b->packing > 1) return fail(e, SA_E_ARG, "packing %u (0 = bytes, 1 = 2-bit)", b->packing);
  if (b->n_pairs >= (1ull << 31)) return fail(e, SA_E_ARG, "n_pairs %llu too large", (unsigned long long)b->n_pairs);
  if (b->n_pairs && (!b->q_off || !b->q_len || !b->d_off || !b->d_len || (!b->residues && b->residues_len)))
    return fail(e, SA_E_ARG, "null input array");
  const uint64_t n = b->n_pairs;
  const auto t_call = std::chrono::steady_clock::now();
  e->timing = sa_timing_t{};
  e->err.clear();
  res->cigar_used = 0;
  MultiFront& mf = *e->front;
  mf.shards.assign(N, sa_shard_info_t{});
  for (int k = 0; k < N; ++k) mf.shards[k].device = e->children[k]->device;
  const bool is_wfa = algo == SA_ALGO_WFA || algo == SA_ALGO_WFA_STANDARD;
  const bool local_ok = mode == SA_MODE_LOCAL && algo == SA_ALGO_NW_LINEAR;  // needleman_wunsch.rs:88-89
  const bool want_cigar = !is_wfa && (mode == SA_MODE_GLOBAL || local_ok) && res->cigar != nullptr && res->cigar_capacity > 0;

  // ---- plan --------------------------------------------------------------------------------
  std::vector<uint64_t> begin(N + 1, 0), pw(N, 0);
  auto on_workers = [&](uint64_t nb, auto&& f) {
    // The planning pass of a 10 M pair list is worth spreading: block b on thread b % T.  The device workers
    // take the first N shares; with few devices extra threads take the rest, so that the pass costs the
    // same (~1 ms per 10 M pairs on 8 threads, less on 16) whatever the number of devices.
    if (n < (1u << 18)) {
      for (uint64_t bl = 0; bl < nb; ++bl) f(bl);
      return;
    }
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int T = std::max(N, std::min(16, hw));
    auto share = [&, T](int k) {
      for (uint64_t bl = (uint64_t)k; bl < nb; bl += (uint64_t)T) f(bl);
    };
    for (int k = 0; k < N; ++k) mf.workers[k]->submit([&, k] { share(k); });
    std::vector<std::thread> extra;
    for (int k = N; k < T; ++k) extra.emplace_back([&, k] { share(k); });
    for (auto& t : extra) t.join();
    for (int k = 0; k < N; ++k) mf.workers[k]->wait();
  };
  if (N == 1) {  // nothing to balance: no pass over the lengths
    begin[1] = n;
    pw[0] = n + 1;
  } else {
    plan_contiguous(b->q_len, b->d_len, n, N, begin.data(), pw.data(), on_workers);
  }
  const bool contiguous = N == 1 || n == 0 || balanced(pw.data(), N);
  const bool trace = getenv("SA_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "[sa trace] multi: planned %llu pairs over %d device(s) in %.3f ms (%s)\n", (unsigned long long)n, N,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(), contiguous ? "contiguous" : "LPT");

  std::vector<sa_status_t> rc(N, SA_OK);
  std::vector<uint64_t> used(N, 0);
  auto finish_shard = [&](int k, const std::chrono::steady_clock::time_point& t0) {
    sa_engine* c = e->children[k];
    sa_shard_info_t& si = mf.shards[k];
    si.h2d_bytes = c->timing.h2d_bytes;
    si.d2h_bytes = c->timing.d2h_bytes;
    si.kernel_launches = c->timing.kernel_launches;
    si.device_ms = c->timing.kernels_ms;
    if (c->timing.cells) si.cells = c->timing.cells;  // exact (the plan's weights of a large list are sampled estimates)
    si.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };

  if (contiguous) {
    // ---- ranges of the caller's arrays; the pool in one region per device --------------------
    std::vector<uint64_t> lo(N + 1, 0);
    if (want_cigar) {
      uint64_t total = 0;
      for (int k = 0; k < N; ++k) total += pw[k];
      uint64_t acc = 0;
      for (int k = 0; k < N; ++k) {
        lo[k] = total ? (uint64_t)((unsigned __int128)res->cigar_capacity * acc / total) : 0;
        acc += pw[k];
      }
      lo[N] = res->cigar_capacity;
      for (int k = 0; k < N; ++k)
        if (begin[k + 1] > begin[k] && lo[k + 1] <= lo[k]) {  // a shard without any room: ask for a sane pool
          res->cigar_used = std::max<uint64_t>(2 * res->cigar_capacity, 32 * n + 1024);
          return fail(e, SA_E_CIGAR_CAPACITY, "cigar pool of %llu words is too small to share among %d devices",
                      (unsigned long long)res->cigar_capacity, N);
        }
    }
    for (int k = 0; k < N; ++k) {
      sa_shard_info_t& si = mf.shards[k];
      si.contiguous = 1;
      si.first_pair = begin[k];
      si.pairs = begin[k + 1] - begin[k];
      si.cells = pw[k] - si.pairs;
      mf.workers[k]->submit([&, k] {
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t p0 = begin[k], cnt = begin[k + 1] - begin[k];
        sa_batch_t sb = *b;
        sb.q_off = b->q_off + p0;
        sb.q_len = b->q_len + p0;
        sb.d_off = b->d_off + p0;
        sb.d_len = b->d_len + p0;
        sb.n_pairs = cnt;
        sa_result_t sr = *res;
        sr.score = res->score ? res->score + p0 : nullptr;
        sr.status = res->status ? res->status + p0 : nullptr;
        sr.cigar_off = res->cigar_off ? res->cigar_off + p0 : nullptr;
        sr.cigar_len = res->cigar_len ? res->cigar_len + p0 : nullptr;
        sr.end1 = res->end1 ? res->end1 + p0 : nullptr;
        sr.end2 = res->end2 ? res->end2 + p0 : nullptr;
        sr.cigar = want_cigar ? res->cigar : nullptr;
        sr.cigar_capacity = want_cigar ? lo[k + 1] : 0;
        sr.cigar_used = 0;
        try {
          rc[k] = cnt ? sd_align_batch(e->children[k], algo, mode, scheme, &sb, &sr, want_cigar ? lo[k] : 0) : SA_OK;
        } catch (const std::exception& ex) {
          rc[k] = fail(e->children[k], SA_E_NOMEM, "host allocation failed: %s", ex.what());
        }
        used[k] = cnt && want_cigar ? sr.cigar_used : lo[k];
        if (cnt) finish_shard(k, t0);
      });
    }
    for (int k = 0; k < N; ++k) mf.workers[k]->wait();
    sa_status_t first = SA_OK;
    bool overflow = false;
    for (int k = 0; k < N; ++k) {
      if (rc[k] == SA_E_CIGAR_CAPACITY) overflow = true;
      else if (rc[k] != SA_OK && first == SA_OK) first = fail(e, rc[k], "device %d: %s", e->children[k]->device, e->children[k]->err.c_str());
    }
    if (first != SA_OK) return first;
    if (overflow) {
      // the capacity with which every region would have held its device's words
      uint64_t total = 0, need_cap = 0;
      for (int k = 0; k < N; ++k) total += pw[k];
      for (int k = 0; k < N; ++k) {
        const uint64_t need = used[k] > lo[k] ? used[k] - lo[k] : 0;
        if (pw[k]) need_cap = std::max<uint64_t>(need_cap, (uint64_t)((unsigned __int128)need * total / pw[k]) + 2 * N + 16);
      }
      res->cigar_used = need_cap;
      return fail(e, SA_E_CIGAR_CAPACITY, "a device's region of the cigar pool overflowed; capacity %llu fits",
                  (unsigned long long)need_cap);
    }
    uint64_t end = 0;
    for (int k = 0; k < N; ++k)
      if (begin[k + 1] > begin[k]) end = std::max(end, used[k]);
    res->cigar_used = want_cigar ? end : 0;
  } else {
    // ---- LPT index sets: gather inputs, scatter results ---------------------------------------
    std::vector<int32_t> part(n);
    lpt(b->q_len, b->d_len, n, N, part.data());
    struct Shard {
      std::vector<uint32_t> idx, q_len, d_len, clen, pool, end1, end2;
      std::vector<uint64_t> q_off, d_off, coff;
      std::vector<int32_t> score;
      std::vector<uint8_t> status;
      uint64_t pool_used = 0;
    };
    std::vector<Shard> sh(N);
    for (uint64_t i = 0; i < n; ++i) sh[part[i]].idx.push_back((uint32_t)i);
    uint64_t total_len = 0;
    for (uint64_t i = 0; i < n; ++i) total_len += (uint64_t)b->q_len[i] + b->d_len[i];
    for (int k = 0; k < N; ++k) {
      sa_shard_info_t& si = mf.shards[k];
      si.contiguous = 0;
      si.first_pair = sh[k].idx.empty() ? 0 : sh[k].idx[0];
      si.pairs = sh[k].idx.size();
      mf.workers[k]->submit([&, k] {
        const auto t0 = std::chrono::steady_clock::now();
        Shard& s = sh[k];
        const size_t cnt = s.idx.size();
        if (!cnt) return;
        try {
          s.q_off.resize(cnt); s.d_off.resize(cnt); s.q_len.resize(cnt); s.d_len.resize(cnt);
          s.score.resize(cnt); s.status.resize(cnt); s.coff.resize(cnt); s.clen.resize(cnt);
          if (res->end1) s.end1.resize(cnt);
          if (res->end2) s.end2.resize(cnt);
          uint64_t cells = 0, len = 0;
          for (size_t t = 0; t < cnt; ++t) {
            const uint32_t i = s.idx[t];
            s.q_off[t] = b->q_off[i]; s.q_len[t] = b->q_len[i];
            s.d_off[t] = b->d_off[i]; s.d_len[t] = b->d_len[i];
            cells += (uint64_t)b->q_len[i] * b->d_len[i];
            len += (uint64_t)b->q_len[i] + b->d_len[i];
          }
          mf.shards[k].cells = cells;
          sa_batch_t sb = *b;
          sb.q_off = s.q_off.data(); sb.q_len = s.q_len.data();
          sb.d_off = s.d_off.data(); sb.d_len = s.d_len.data();
          sb.n_pairs = cnt;
          // local pool: this shard's share of the caller's capacity by sequence length, with slack;
          // a second attempt gets exactly what the first one asked for
          uint64_t cap = want_cigar ? (uint64_t)((unsigned __int128)res->cigar_capacity * len / std::max<uint64_t>(total_len, 1)) + 4096 : 0;
          for (int attempt = 0; attempt < 2; ++attempt) {
            if (want_cigar) s.pool.resize(cap);
            sa_result_t sr{s.score.data(), s.status.data(), s.coff.data(), s.clen.data(), want_cigar ? s.pool.data() : nullptr, cap, 0,
                           res->end1 ? s.end1.data() : nullptr, res->end2 ? s.end2.data() : nullptr};
            rc[k] = sd_align_batch(e->children[k], algo, mode, scheme, &sb, &sr, 0);
            s.pool_used = sr.cigar_used;
            if (rc[k] != SA_E_CIGAR_CAPACITY) break;
            cap = sr.cigar_used + 16;
          }
        } catch (const std::exception& ex) {
          rc[k] = fail(e->children[k], SA_E_NOMEM, "host allocation failed: %s", ex.what());
        }
        finish_shard(k, t0);
      });
    }
    for (int k = 0; k < N; ++k) mf.workers[k]->wait();
    for (int k = 0; k < N; ++k)
      if (rc[k] != SA_OK) return fail(e, rc[k], "device %d: %s", e->children[k]->device, e->children[k]->err.c_str());
    // per-pair results back to input order; offsets = exclusive scan of the lengths
    std::vector<uint32_t> clen(n, 0);
    for (int k = 0; k < N; ++k)
      for (size_t t = 0; t < sh[k].idx.size(); ++t) {
        const uint32_t i = sh[k].idx[t];
        if (res->score) res->score[i] = sh[k].score[t];
        if (res->status) res->status[i] = sh[k].status[t];
        if (res->end1) res->end1[i] = sh[k].end1[t];
        if (res->end2) res->end2[i] = sh[k].end2[t];
        clen[i] = sh[k].clen[t];
      }
    std::vector<uint64_t> coff(n, 0);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
      coff[i] = total;
      total += clen[i];
    }
    if (res->cigar_len) memcpy(res->cigar_len, clen.data(), n * 4);
    if (res->cigar_off) memcpy(res->cigar_off, coff.data(), n * 8);
    res->cigar_used = total;
    if (want_cigar) {
      if (total > res->cigar_capacity)
        return fail(e, SA_E_CIGAR_CAPACITY, "cigar pool needs %llu words, capacity is %llu", (unsigned long long)total,
                    (unsigned long long)res->cigar_capacity);
      for (int k = 0; k < N; ++k)
        mf.workers[k]->submit([&, k] {
          const Shard& s = sh[k];
          for (size_t t = 0; t < s.idx.size(); ++t)
            if (s.clen[t]) memcpy(res->cigar + coff[s.idx[t]], s.pool.data() + s.coff[t], (size_t)s.clen[t] * 4);
        });
      for (int k = 0; k < N; ++k) mf.workers[k]->wait();
    }
  }
  // ---- the call's timing: sums over devices, the slowest device's time -------------------------
  for (int k = 0; k < N; ++k) {
    const sa_timing_t& t = e->children[k]->timing;
    if (!mf.shards[k].pairs) continue;
    e->timing.cells += t.cells;
    e->timing.kernel_launches += t.kernel_launches;
    e->timing.h2d_bytes += t.h2d_bytes;
    e->timing.d2h_bytes += t.d2h_bytes;
    e->timing.pairs_rerun += t.pairs_rerun;
    e->timing.pairs_fallback += t.pairs_fallback;
    e->timing.wfa_cells += t.wfa_cells;
    e->timing.wfa_extended += t.wfa_extended;
    e->timing.kernels_ms = std::max(e->timing.kernels_ms, t.kernels_ms);
    e->timing.fill_ms = std::max(e->timing.fill_ms, t.fill_ms);
    e->timing.long_fwd_ms = std::max(e->timing.long_fwd_ms, t.long_fwd_ms);
    e->timing.long_back_ms = std::max(e->timing.long_back_ms, t.long_back_ms);
  }
  e->timing.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
  if (trace) fprintf(stderr, "[sa trace] multi: call done after %.3f ms\n", e->timing.wall_ms);
  return SA_OK;
}

}  // namespace sa_host

extern "C" {

sa_status_t sa_partition_lpt(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs, int n_parts,
                             int32_t* part) {
  if (n_parts < 1 || (n_pairs && (!q_len || !d_len || !part))) return SA_E_ARG;
  try {
    sa_host::lpt(q_len, d_len, n_pairs, n_parts, part);
  } catch (const std::bad_alloc&) {
    return SA_E_NOMEM;
  }
  return SA_OK;
}

sa_status_t sa_plan_shards(const uint32_t* q_len, const uint32_t* d_len, uint64_t n_pairs, int n_parts,
                           uint64_t* begin, int32_t* part, int* contiguous) {
  try {
    return sa_host::plan_shards(q_len, d_len, n_pairs, n_parts, begin, part, contiguous);
  } catch (const std::bad_alloc&) {
    return SA_E_NOMEM;
  }
}

}  // extern "C"
