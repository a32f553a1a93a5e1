// sa_align -- command line with the reference's flags (/root/reference/src/parse.rs:8-34) over
// the C ABI.  It replaces the `for d in db { for q in query { .. } }` loop of
// /root/reference/src/main.rs:61-79 by ONE batched call and prints, per pair and in the same
// db-major order, the text the reference prints for the FIRST alignment
// (needleman_wunsch_affine.rs:283-286, Display :390-411).
//
//   sa_align -q <query.fa> -d <db.fa> [-o PATH] [-v] [-m global|local|semi-global]
//            [-a a-star|needleman-wunsch|needleman-wunsch-linear|wfa|wfa-standard] [--strict]
//
// Differences from the reference binary, all deliberate and printed on stderr when they apply:
//   * only the first co-optimal alignment of a pair is printed (the reference prints all);
//   * where the reference process would panic (nw_affine:299/:303, wfa.rs:577/:603) this tool
//     reports it on stderr and goes on; with --strict it exits with status 101 like a Rust panic;
//   * a-star (the reference's default algorithm) is not part of the GPU path.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "sa_engine.h"

// A record as views into the parser's output buffer (names are copied: they are printed with %s).
struct SeqView {
  const char* p = nullptr;
  size_t n = 0;
  const char* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
};
struct Rec {
  std::string name;
  SeqView seq;
  uint64_t seq_off = 0;  // offset of the sequence in the file's output buffer (= in its 2-bit image, in residues)
};

// One FASTA file through the fused parser + packer (sa_parse_fasta_packed): `out` receives names and
// sequences, `packed` (pinned, the engine streams from it) their 2-bit image.
struct Fasta {
  std::vector<uint8_t> out;
  std::vector<Rec> recs;
  uint64_t out_len = 0;
  bool all_acgt = true;
};

// Page-aligned plain memory: nothing here touches CUDA, so parsing runs while the engine (and the CUDA
// context, ~1 s) comes up on another thread; the large buffers are page-locked afterwards in place
// (sa_host_register), which is what lets sa_align_batch stream from / to them asynchronously.
static void* host_alloc(size_t bytes) {
  const size_t rounded = (std::max<size_t>(bytes, 1) + 4095) & ~(size_t)4095;
  return aligned_alloc(4096, rounded);
}

// One render call per text: the snprintf-style entry points write into a reused buffer and say how much they
// needed; only a text that did not fit is rendered a second time.  Returns false when the call failed.
template <class F>
static bool emit_text(std::string& buf, F&& call) {
  if (buf.size() < (1u << 16)) buf.resize(1u << 16);
  int64_t need = call(&buf[0], buf.size());
  if (need < 0) return false;
  if ((size_t)need + 1 > buf.size()) {
    buf.resize((size_t)need + 1);
    need = call(&buf[0], buf.size());
  }
  fwrite(buf.data(), 1, (size_t)need, stdout);
  return true;
}

static size_t file_size(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  size_t size = 0;
  if (f) {
    fseek(f, 0, SEEK_END);
    size = (size_t)ftell(f);
    fclose(f);
  }
  return size;
}

static bool load_fasta(const char* what, const std::string& path, Fasta& fa, uint8_t* packed, size_t packed_cap) {
  const size_t size = file_size(path);
  fa.out.resize(size + 1);
  std::vector<uint8_t> err(4096);  // the first rejected bytes; the parser counts the rest
  // one index row per '>' of the file (a scan at memory speed; the parser reads the page cache afterwards)
  size_t n_gt = 0;
  if (FILE* f = fopen(path.c_str(), "rb")) {
    std::vector<char> chunk(1 << 22);
    size_t got;
    while ((got = fread(chunk.data(), 1, chunk.size(), f)) > 0)
      for (const char* c = chunk.data(), *e = c + got; (c = (const char*)memchr(c, '>', (size_t)(e - c))) != nullptr; ++c) ++n_gt;
    fclose(f);
  }
  std::vector<uint64_t> index(4 * (n_gt + 2));
  size_t nerr = 0;
  int acgt = 1;
  const int64_t n = sa_parse_fasta_packed(path.c_str(), fa.out.data(), fa.out.size(), index.data(), index.size() / 4, err.data(), err.size(),
                                          &nerr, packed, packed_cap, &acgt, &fa.out_len);
  if (n < 0) {  // main.rs:24-28 / :44-48
    fprintf(stderr, "%s fasta could not be opened: invalid input parameter\naborting\n", what);
    return false;
  }
  fa.all_acgt = acgt != 0;
  if (nerr) {  // main.rs:29-35: continue with the partial records
    std::string chars;
    for (size_t k = 0; k < nerr && k < err.size(); ++k) {
      if (k) chars += ", ";
      chars += "'";
      chars += (char)err[k];
      chars += "'";
    }
    fprintf(stderr, "Invalid character '[%s]' detected in %s fasta; continuing by ignoring it\n", chars.c_str(),
            strcmp(what, "DB") == 0 ? "db" : "query");
  }
  fa.recs.resize((size_t)n);
  for (int64_t r = 0; r < n; ++r) {
    Rec& rec = fa.recs[(size_t)r];
    rec.name.assign((const char*)fa.out.data() + index[4 * r], index[4 * r + 1]);
    rec.seq = SeqView{(const char*)fa.out.data() + index[4 * r + 2], (size_t)index[4 * r + 3]};
    rec.seq_off = index[4 * r + 2];
  }
  return true;
}

static void usage() {
  fprintf(stderr,
          "Usage: sa_align --query-file <QUERY_FILE> --db-file <DB_FILE> [OPTIONS]\n\n"
          "Options:\n  -q, --query-file <QUERY_FILE>  Path to query sequence\n"
          "  -d, --db-file <DB_FILE>        path to db sequence\n"
          "  -o, --out-path <OUT_PATH>      out path [default: ./results]\n"
          "  -v, --verbose                  verbose\n"
          "  -m, --mode <MODE>              modus [default: global] [possible values: global, local, semi-global]\n"
          "  -a, --algo <ALGO>              algo [default: needleman-wunsch] [possible values: a-star, needleman-wunsch,\n"
          "                                 needleman-wunsch-linear, wfa, wfa-standard]\n"
          "      --strict                   exit 101 where the reference would panic\n"
          "      --all                      needleman-wunsch(-linear): print EVERY co-optimal alignment / hit, like the reference\n"
          "      --timing                   wall time of parse+pack / align / print on stderr\n"
          "      --no-output                align without printing the alignments\n"
          "      --pageable                 leave the input buffers pageable (default: page-locked in place for >= 64 Ki pairs)\n"
          "      --device <N>               CUDA device [default: 0]\n"
          "      --devices <LIST>           several CUDA devices, e.g. 0,1,2,3: the pair list is sharded over them\n"
          "  -h, --help                     Print help\n  -V, --version                  Print version\n");
}

// Rust's `{:#?}` of a Duration, close enough for the line the reference prints per pair
static std::string duration_debug(double seconds) {
  char buf[64];
  if (seconds >= 1.0) snprintf(buf, sizeof(buf), "%.9gs", seconds);
  else if (seconds >= 1e-3) snprintf(buf, sizeof(buf), "%.6gms", seconds * 1e3);
  else if (seconds >= 1e-6) snprintf(buf, sizeof(buf), "%.3f\xC2\xB5s", seconds * 1e6);
  else snprintf(buf, sizeof(buf), "%.0fns", seconds * 1e9);
  return buf;
}

int main(int argc, char** argv) {
  std::string qpath, dpath, mode = "global", algo = "needleman-wunsch";
  bool verbose = false, strict = false, all = false, algo_given = false, timing = false, quiet = false, pageable = false;
  std::vector<int> devices;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&](const char* s, const char* l) -> const char* {
      if (a == s || a == l) {
        if (i + 1 >= argc) { usage(); exit(2); }
        return argv[++i];
      }
      const std::string pre = std::string(l) + "=";
      if (a.rfind(pre, 0) == 0) return argv[i] + pre.size();
      return nullptr;
    };
    if (const char* v = val("-q", "--query-file")) qpath = v;
    else if (const char* v = val("-d", "--db-file")) dpath = v;
    else if (const char* v = val("-o", "--out-path")) (void)v;  // parsed and unused, as in the reference
    else if (const char* v = val("-m", "--mode")) mode = v;
    else if (const char* v = val("-a", "--algo")) { algo = v; algo_given = true; }
    else if (const char* v = val("--device", "--device")) devices.assign(1, atoi(v));
    else if (const char* v = val("--devices", "--devices")) {  // 0,1,2,3: one engine over several GPUs
      devices.clear();
      for (const char* c = v; *c;) {
        char* end = nullptr;
        devices.push_back((int)strtol(c, &end, 10));
        if (end == c) { fprintf(stderr, "error: invalid value '%s' for '--devices <LIST>'\n", v); return 2; }
        c = *end == ',' ? end + 1 : end;
      }
    }
    else if (a == "-v" || a == "--verbose") verbose = true;
    else if (a == "--strict") strict = true;
    else if (a == "--all") all = true;
    else if (a == "--timing") timing = true;   // phase times on stderr
    else if (a == "--pageable") pageable = true;  // do not page-lock the inputs
    else if (a == "--no-output") quiet = true;  // align, but print nothing (throughput measurements)
    else if (a == "-h" || a == "--help") { usage(); return 0; }
    else if (a == "-V" || a == "--version") { printf("sa_align 0.1.0 (ABI %d)\n", sa_abi_version()); return 0; }
    else { fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str()); usage(); return 2; }
  }
  if (qpath.empty() || dpath.empty()) { usage(); return 2; }
  if (devices.empty()) devices.push_back(0);
  if (!algo_given)  // parse.rs:36-42: the reference's default is a-star, which is not part of the GPU path
    fprintf(stderr, "note: no -a given; the reference defaults to a-star, this tool to needleman-wunsch\n");
  sa_mode_t m;
  if (mode == "global") m = SA_MODE_GLOBAL;
  else if (mode == "local") m = SA_MODE_LOCAL;
  else if (mode == "semi-global") m = SA_MODE_SEMIGLOBAL;
  else { fprintf(stderr, "error: invalid value '%s' for '--mode <MODE>'\n", mode.c_str()); return 2; }
  sa_algo_t al;
  if (algo == "needleman-wunsch") al = SA_ALGO_NW_AFFINE;
  else if (algo == "needleman-wunsch-linear") al = SA_ALGO_NW_LINEAR;
  else if (algo == "wfa") al = SA_ALGO_WFA;
  else if (algo == "wfa-standard") al = SA_ALGO_WFA_STANDARD;
  else if (algo == "a-star") { fprintf(stderr, "a-star is not part of the GPU path (see DESIGN.md); use the reference binary\n"); return 2; }
  else { fprintf(stderr, "error: invalid value '%s' for '--algo <ALGO>'\n", algo.c_str()); return 2; }

  // Both files through the fused parser + packer into ONE pinned 2-bit image: the query file's bytes at
  // residue offset 0, the db file's behind them at a byte boundary (main.rs:22-59 loads the db first).
  // Records with an 'N' cannot be 2-bit coded: then the byte images are concatenated (pinned) instead.
  const auto t_start = std::chrono::steady_clock::now();
  // the engine (CUDA context, streams, events) comes up while the FASTA files are parsed and packed
  sa_engine_t* eng = nullptr;
  sa_status_t created = SA_OK;
  std::thread engine_thread([&] {
    created = devices.size() == 1 ? sa_engine_create(devices[0], &eng)
                                  : sa_engine_create_multi(devices.data(), (int)devices.size(), &eng);
  });
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } joiner{engine_thread};
  const size_t sq = file_size(qpath), sd = file_size(dpath);
  const size_t q_bytes = (sq + 3) / 4 + 4, d_bytes = (sd + 3) / 4 + 4;
  uint8_t* packed = (uint8_t*)host_alloc(q_bytes + d_bytes);
  if (!packed) { fprintf(stderr, "host allocation failed\n"); return 1; }
  Fasta fdb, fq;
  if (!load_fasta("DB", dpath, fdb, packed + q_bytes, d_bytes)) return 0;       // the reference returns from main, status 0
  if (!load_fasta("Query", qpath, fq, packed, q_bytes)) return 0;
  const std::vector<Rec>&db = fdb.recs, &query = fq.recs;
  const bool two_bit = fdb.all_acgt && fq.all_acgt;
  uint8_t* bytes = nullptr;
  if (!two_bit) {
    bytes = (uint8_t*)host_alloc(fq.out_len + fdb.out_len + 1);
    if (!bytes) { fprintf(stderr, "host allocation failed\n"); return 1; }
    memcpy(bytes, fq.out.data(), fq.out_len);
    memcpy(bytes + fq.out_len, fdb.out.data(), fdb.out_len);
  }
  const uint64_t d_base = two_bit ? 4 * (uint64_t)q_bytes : fq.out_len;  // where the db file's residues start

  // pairs in the db-major order of main.rs:61-62, offsets and lengths in pinned memory
  const size_t nq = query.size(), nd = db.size(), n = nq * nd;
  uint64_t* q_off = (uint64_t*)host_alloc((n + 1) * 8);
  uint64_t* d_off = (uint64_t*)host_alloc((n + 1) * 8);
  uint32_t* q_len = (uint32_t*)host_alloc((n + 1) * 4);
  uint32_t* d_len = (uint32_t*)host_alloc((n + 1) * 4);
  if (!q_off || !d_off || !q_len || !d_len) { fprintf(stderr, "host allocation failed\n"); return 1; }
  for (size_t d = 0, p = 0; d < nd; ++d)
    for (size_t q = 0; q < nq; ++q, ++p) {
      q_off[p] = query[q].seq_off; q_len[p] = (uint32_t)query[q].seq.size();
      d_off[p] = d_base + db[d].seq_off; d_len[p] = (uint32_t)db[d].seq.size();
    }
  const auto t_parsed = std::chrono::steady_clock::now();
  engine_thread.join();
  if (created != SA_OK) {
    fprintf(stderr, "sa_engine_create: %s\n", sa_last_error(eng));
    sa_engine_destroy(eng);
    return 1;
  }
  sa_batch_t batch{two_bit ? packed : bytes, two_bit ? (uint64_t)(q_bytes + d_bytes) : fq.out_len + fdb.out_len, q_off, q_len, d_off, d_len, n,
                   two_bit ? 1u : 0u};
  const auto t_engine = std::chrono::steady_clock::now();
  // page-lock the inputs in place (large batches only: locking costs about as much as one staged copy)
  const bool pin = !pageable && n >= 65536;
  if (pin) {
    sa_host_register(two_bit ? (void*)packed : (void*)bytes, (size_t)batch.residues_len);
    sa_host_register(q_off, (n + 1) * 8);
    sa_host_register(d_off, (n + 1) * 8);
    sa_host_register(q_len, (n + 1) * 4);
    sa_host_register(d_len, (n + 1) * 4);
  }
  // results (pageable: a one-shot process would spend longer locking ~100 bytes per pair than the staged copy takes)
  int32_t* score = (int32_t*)host_alloc((n + 1) * 4);
  uint8_t* status = (uint8_t*)host_alloc(n + 1);
  uint64_t* coff = (uint64_t*)host_alloc((n + 1) * 8);
  uint32_t* clen = (uint32_t*)host_alloc((n + 1) * 4);
  uint32_t* end1 = (uint32_t*)host_alloc((n + 1) * 4);
  uint32_t* end2 = (uint32_t*)host_alloc((n + 1) * 4);
  uint64_t pool_cap = 24 * (uint64_t)n + 4096;
  uint32_t* pool = (uint32_t*)host_alloc(pool_cap * 4);
  if (!score || !status || !coff || !clen || !end1 || !end2 || !pool) { fprintf(stderr, "host allocation failed\n"); return 1; }
  sa_result_t res{score, status, coff, clen, pool, pool_cap, 0, end1, end2};
  const auto t0 = std::chrono::steady_clock::now();
  sa_status_t rc = sa_align_batch(eng, al, m, nullptr, &batch, &res);
  if (rc == SA_E_CIGAR_CAPACITY) {
    pool_cap = res.cigar_used + 16;  // (the first pool is left to process exit)
    pool = (uint32_t*)host_alloc(pool_cap * 4);
    if (!pool) { fprintf(stderr, "host allocation failed\n"); return 1; }
    res.cigar = pool;
    res.cigar_capacity = pool_cap;
    rc = sa_align_batch(eng, al, m, nullptr, &batch, &res);
  }
  const double per_pair = n ? std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (double)n : 0;
  if (rc != SA_OK) {
    fprintf(stderr, "sa_align_batch: %s\n", sa_last_error(eng));
    sa_engine_destroy(eng);
    return 1;
  }
  const auto t_aligned = std::chrono::steady_clock::now();
  auto report_timing = [&]() {
    if (!timing) return;
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    sa_timing_t tm{};
    sa_last_timing(eng, &tm);
    const auto t_end = std::chrono::steady_clock::now();
    fprintf(stderr,
            "timing: %zu pairs, %.3e cells | parse+pack %.1f ms (%.2f GB/s, %s; engine start-up runs beside it) | "
            "wait for the engine %.1f ms | page-lock + buffers %.1f ms | sa_align_batch %.1f ms "
            "(kernels %.1f ms, %.1f GCUPS end to end) | print %.1f ms | total %.1f ms\n",
            n, (double)tm.cells, ms(t_start, t_parsed), (double)(sq + sd) / 1e6 / std::max(ms(t_start, t_parsed), 1e-6),
            two_bit ? "2-bit" : "bytes", ms(t_parsed, t_engine), ms(t_engine, t0), ms(t0, t_aligned), tm.kernels_ms,
            (double)tm.cells / 1e6 / std::max(ms(t0, t_aligned), 1e-6), ms(t_aligned, t_end), ms(t_start, t_end));
  };
  if (quiet) {
    report_timing();
    sa_engine_destroy(eng);
    return 0;
  }
  int exit_code = 0;
  // SA_ALIGNMENT_OMITTED is a flag on top of the status: report it, then treat the status as usual
  std::vector<uint8_t> omitted(n, 0);
  for (size_t p = 0; p < n; ++p) {
    omitted[p] = (status[p] & SA_ALIGNMENT_OMITTED) ? 1 : 0;
    status[p] &= 0x7f;
  }
  std::string text, all_text;  // reused render buffers
  static char out_buf[1 << 20];
  setvbuf(stdout, out_buf, _IOFBF, sizeof(out_buf));
  for (size_t d = 0, p = 0; d < nd && !exit_code; ++d)
    for (size_t q = 0; q < nq; ++q, ++p) {
      const Rec &Q = query[q], &D = db[d];
      if (omitted[p])
        fprintf(stderr, "%s vs %s: the alignment was not materialised (traceback exceeds the scratch budget); score %d and status are exact\n",
                Q.name.c_str(), D.name.c_str(), score[p]);
      if (status[p] == SA_NOT_IMPLEMENTED) {  // main.rs:68-74 + errors.rs:11-12
        fprintf(stderr, "An error occured during alignment of %s and %s\nError in alignment: not implemented\n",
                Q.name.c_str(), D.name.c_str());
        continue;
      }
      if (al == SA_ALGO_WFA) {
        // wfa_align's complete stdout (wfa.rs:23-42) from a traced run of the literal
        // kernel: the `lo: .., hi: ..` lines, and after convergence the score, the `huhu` block, rec_tr's
        // lines and the empty Alignment.  Where the reference dies or loops: what it had printed by then.
        int32_t st = 0;
        if (!emit_text(text, [&](char* b, size_t c) {
              return sa_wfa_reference_stdout(eng, (const uint8_t*)Q.seq.data(), (uint32_t)Q.seq.size(), (const uint8_t*)D.seq.data(),
                                             (uint32_t)D.seq.size(), b, c, &st);
            })) {
          fprintf(stderr, "sa_wfa_reference_stdout: %s\n", sa_last_error(eng));
          exit_code = 1;
          break;
        }
        if (st != SA_OK) {
          fprintf(stderr, "%s vs %s: the reference %s here\n", Q.name.c_str(), D.name.c_str(),
                  st == SA_REF_PANIC ? "panics in trim (wfa.rs:577/603)" : "never converges (wfa.rs:189)");
          if (strict && st == SA_REF_PANIC) { exit_code = 101; break; }
        }
        continue;
      }
      if (al == SA_ALGO_WFA_STANDARD) {  // an extension (the reference has no such mode): the optimal gap-affine cost
        printf("%s vs %s: gap-affine cost %d\n", Q.name.c_str(), D.name.c_str(), score[p]);
        continue;
      }
      if (al == SA_ALGO_NW_LINEAR) {
        // needleman_wunsch.rs:193-201, then the FIRST hit of backtrace (:106-116, :205-213)
        if (verbose) printf("search finished after %s\n", duration_debug(per_pair).c_str());
        printf("Alignment between sequences %s and %s found\n", Q.name.c_str(), D.name.c_str());
        if (all) {  // every hit of every start cell, as the reference prints them (:106-116, :205-254)
          uint64_t n_hits = 0;
          if (all_text.size() < (1u << 20)) all_text.resize(1u << 20);
          if (!emit_text(all_text, [&](char* b, size_t c) {
                return sa_linear_all_hits(eng, (const uint8_t*)Q.seq.data(), (uint32_t)Q.seq.size(), (const uint8_t*)D.seq.data(),
                                          (uint32_t)D.seq.size(), m == SA_MODE_LOCAL, nullptr, ~0ull, b, c, &n_hits);
              })) {
            fprintf(stderr, "sa_linear_all_hits: %s\n", sa_last_error(eng));
            exit_code = 1;
            break;
          }
          continue;
        }
        if (!omitted[p]) {
          if (!emit_text(text, [&](char* b, size_t c) {
                return sa_render_linear_hit((const uint8_t*)Q.seq.data(), (uint32_t)Q.seq.size(), (const uint8_t*)D.seq.data(),
                                            (uint32_t)D.seq.size(), pool + coff[p], clen[p], end1[p], end2[p], b, c);
              })) {
            fprintf(stderr, "sa_render_linear_hit: the CIGAR does not fit the pair\n");
            exit_code = 1;
            break;
          }
        }
        continue;
      }
      const bool has_alignment = !omitted[p] && (clen[p] > 0 || (Q.seq.empty() && D.seq.empty()));
      if (all && al == SA_ALGO_NW_AFFINE) {
        // the reference's full output for the pair: every co-optimal alignment in DFS order,
        // up to the point where it would panic
        uint64_t n_printed = 0;
        int32_t panicked = 0;
        // one enumeration per pair: only a text beyond the reused buffer (1 MB to start with) is produced twice
        if (all_text.size() < (1u << 20)) all_text.resize(1u << 20);
        if (!emit_text(all_text, [&](char* b, size_t c) {
              return sa_affine_all_alignments(eng, (const uint8_t*)Q.seq.data(), (uint32_t)Q.seq.size(), (const uint8_t*)D.seq.data(),
                                              (uint32_t)D.seq.size(), nullptr, ~0ull, b, c, &n_printed, &panicked);
            })) {
          fprintf(stderr, "sa_affine_all_alignments: %s\n", sa_last_error(eng));
          exit_code = 1;
          break;
        }
        if (panicked) {
          fprintf(stderr, "%s vs %s: the reference panics here (index out of bounds, needleman_wunsch_affine.rs:299/303) after %llu alignment(s)\n",
                  Q.name.c_str(), D.name.c_str(), (unsigned long long)n_printed);
          if (strict) { exit_code = 101; break; }
        }
        printf("%s\n", duration_debug(per_pair).c_str());
        continue;
      }
      if (has_alignment)
        emit_text(text, [&](char* b, size_t c) {
          return sa_render_affine((const uint8_t*)Q.seq.data(), (uint32_t)Q.seq.size(), (const uint8_t*)D.seq.data(), (uint32_t)D.seq.size(),
                                  pool + coff[p], clen[p], b, c);
        });
      if (verbose) printf("score: %d\n", score[p]);
      if (status[p] == SA_REF_PANIC || status[p] == SA_REF_PANIC_EARLY) {
        fprintf(stderr, "%s vs %s: the reference panics here (index out of bounds, needleman_wunsch_affine.rs:299/303)%s\n",
                Q.name.c_str(), D.name.c_str(), status[p] == SA_REF_PANIC ? " after printing" : " before printing anything");
        if (strict) { exit_code = 101; break; }
      }
      if (al == SA_ALGO_NW_AFFINE) printf("%s\n", duration_debug(per_pair).c_str());  // nw_affine:431
    }
  fflush(stdout);
  report_timing();
  sa_engine_destroy(eng);
  return exit_code;
}
