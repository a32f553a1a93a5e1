// host_io.cu -- host-side pieces of the drop-in boundary (no device code).
//
//   sa_parse_fasta      parse_fasta of /root/reference/src/parse.rs:54-99, same quirks
//   sa_render_affine    the text the reference prints for one alignment
//                       (needleman_wunsch_affine.rs:283-286 + Display :390-411)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <exception>
#include <memory>
#include <new>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/sa_engine.h"

namespace {

// Path::extension() semantics used by has_extension (parse.rs:101-106)
bool has_ext(const char* path, const char* ext) {
  const char* base = strrchr(path, '/');
  base = base ? base + 1 : path;
  const char* dot = strrchr(base, '.');
  if (!dot || dot == base) return false;
  return strcmp(dot + 1, ext) == 0;
}

inline bool allowed(uint8_t c) {  // ALLOWED_CHARS parse.rs:52
  return c == 'A' || c == 'G' || c == 'C' || c == 'T' || c == 'N';
}

struct RecordSpan {
  uint64_t name_off, name_len, seq_off, seq_len;
};

// ---- 2-bit packer ---------------------------------------------------------------------------------
// Two residues per table look-up: tab[b0 | b1 << 8] = code(b0) | code(b1) << 2, or 0xFF when either
// byte is not A/C/G/T (A=0 C=1 G=2 T=3).  64 KB, built once; a look-up per two bytes does the
// validation and the coding together.
struct PackTable {
  uint8_t strict[65536];   // 0xFF for any other byte
  uint8_t lenient[65536];  // other bytes code as 0 (names, 'N': the caller knows whether it matters)
  PackTable() {
    int code[256];
    for (int i = 0; i < 256; ++i) code[i] = -1;
    code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3;
    for (int b1 = 0; b1 < 256; ++b1)
      for (int b0 = 0; b0 < 256; ++b0) {
        const int c0 = code[b0], c1 = code[b1];
        strict[b0 | (b1 << 8)] = (c0 < 0 || c1 < 0) ? 0xFF : (uint8_t)(c0 | (c1 << 2));
        lenient[b0 | (b1 << 8)] = (uint8_t)((c0 < 0 ? 0 : c0) | ((c1 < 0 ? 0 : c1) << 2));
      }
  }
};
const PackTable& pack_table() {
  static const PackTable t;
  return t;
}

// packs src[0, n) to whole bytes at dst (n a multiple of 4 except for the last call of a buffer);
// STRICT: returns false at the first byte that is not A/C/G/T
template <bool STRICT>
bool pack_run(const uint8_t* src, uint64_t n, uint8_t* dst) {
  const uint8_t* tab = STRICT ? pack_table().strict : pack_table().lenient;
  uint64_t k = 0;
  for (; k + 8 <= n; k += 8) {
    uint16_t w[4];
    memcpy(w, src + k, 8);
    const uint8_t a = tab[w[0]], b = tab[w[1]], c = tab[w[2]], d = tab[w[3]];
    if (STRICT && ((a | b | c | d) & 0xF0)) return false;  // (valid entries are 0 .. 15)
    dst[k >> 2] = (uint8_t)(a | (b << 4));
    dst[(k >> 2) + 1] = (uint8_t)(c | (d << 4));
  }
  for (; k < n; k += 4) {
    uint8_t v = 0;
    for (uint64_t j = 0; j < 4 && k + j < n; ++j) {
      const uint8_t t = tab[src[k + j] | ('A' << 8)];
      if (STRICT && t == 0xFF) return false;
      v |= (uint8_t)((t & 3) << (2 * j));
    }
    dst[k >> 2] = v;
  }
  return true;
}

// src[0, n) -> dst bytes [0, ceil(n / 4)), over `nt` threads on 4-residue boundaries
template <bool STRICT>
bool pack_parallel(const uint8_t* src, uint64_t n, uint8_t* dst, unsigned nt) {
  if (n < ((uint64_t)1 << 20) || nt <= 1) return pack_run<STRICT>(src, n, dst);
  std::vector<uint8_t> ok(nt, 1);
  std::vector<std::thread> th;
  const uint64_t per = ((n / nt) + 3) & ~(uint64_t)3;
  for (unsigned k = 0; k < nt; ++k)
    th.emplace_back([&, k] {
      const uint64_t lo = std::min(n, per * k), hi = (k + 1 == nt) ? n : std::min(n, per * (k + 1));
      if (lo < hi) ok[k] = pack_run<STRICT>(src + lo, hi - lo, dst + (lo >> 2)) ? 1 : 0;
    });
  for (auto& t : th) t.join();
  for (unsigned k = 0; k < nt; ++k)
    if (!ok[k]) return false;
  return true;
}

unsigned host_threads() {
  unsigned nt = std::thread::hardware_concurrency();
  if (const char* s = getenv("SA_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(s));
  return std::max(1u, std::min(nt, 32u));
}

int64_t parse_impl(const char* path, uint8_t* out, size_t out_cap, uint64_t* index, size_t index_cap, uint8_t* err_chars,
                   size_t err_cap, size_t* n_err, uint8_t* packed, size_t packed_cap, int* all_acgt, uint64_t* out_len) {
  if (n_err) *n_err = 0;
  if (all_acgt) *all_acgt = 1;
  if (out_len) *out_len = 0;
  if (!path) return SA_E_ARG;
  // parse.rs:55-60: anything but .fa/.fasta/.fna is FastaError(InvalidInput)
  if (!(has_ext(path, "fa") || has_ext(path, "fasta") || has_ext(path, "fna"))) return SA_E_ARG;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return SA_E_ARG;  // parse.rs:62 `read(path)?`
  unsigned nt = host_threads();
  // The whole file in one (uninitialised) buffer.  A regular file is read by all threads at once
  // (pread of disjoint ranges); anything else (pipe, /dev/stdin) by a plain read loop.
  std::unique_ptr<uint8_t[]> storage;
  size_t size = 0;
  {
    struct stat sb;
    const bool regular = fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0;
    size_t cap = regular ? (size_t)sb.st_size : ((size_t)1 << 20);
    storage.reset(new (std::nothrow) uint8_t[cap + 1]);
    if (!storage) {
      close(fd);
      return SA_E_NOMEM;
    }
    if (regular) {
      const unsigned rt = cap >= ((size_t)8 << 20) ? nt : 1;
      std::vector<size_t> got(rt, 0);
      auto rd = [&](unsigned k) {
        size_t lo = cap / rt * k, hi = (k + 1 == rt) ? cap : cap / rt * (k + 1);
        while (lo < hi) {
          const ssize_t n = pread(fd, storage.get() + lo, hi - lo, (off_t)lo);
          if (n <= 0) break;
          lo += (size_t)n;
          got[k] += (size_t)n;
        }
      };
      if (rt == 1) {
        rd(0);
      } else {
        std::vector<std::thread> th;
        for (unsigned k = 0; k < rt; ++k) th.emplace_back(rd, k);
        for (auto& t : th) t.join();
      }
      // (a file truncated while being read ends at the first short range)
      for (unsigned k = 0; k < rt; ++k) {
        const size_t want = ((k + 1 == rt) ? cap : cap / rt * (k + 1)) - cap / rt * k;
        size += got[k];
        if (got[k] < want) break;
      }
    } else {
      for (;;) {
        if (size == cap) {
          std::unique_ptr<uint8_t[]> bigger(new (std::nothrow) uint8_t[cap * 2 + 1]);
          if (!bigger) {
            close(fd);
            return SA_E_NOMEM;
          }
          memcpy(bigger.get(), storage.get(), size);
          storage.swap(bigger);
          cap *= 2;
        }
        const ssize_t n = read(fd, storage.get() + size, cap - size);
        if (n <= 0) break;
        size += (size_t)n;
      }
    }
  }
  close(fd);
  uint8_t* const data = storage.get();
  const bool trace = getenv("SA_TRACE") != nullptr;
  const auto t_read = std::chrono::steady_clock::now();

  // The state machine of parse.rs:66-89, restated over byte classes: '>' ANYWHERE starts a
  // record (also inside a header line); a header runs to the next '\n'; elsewhere '\n' is
  // skipped, a byte outside ACGTN is reported, an allowed byte is appended -- to the default
  // record before the first '>', which parse.rs:91 then removes.
  //
  // Parallel form: the file is cut at '>' bytes into one chunk per thread; a chunk is parsed IN
  // PLACE (the output of a chunk is never longer than its input), then the compacted chunks are
  // copied to `out` at their prefix offsets.  Chunk 0 starts at byte 0 (in the default record).
  if (size < ((size_t)1 << 20)) nt = 1;
  std::vector<size_t> start(nt + 1, size);
  start[0] = 0;
  for (unsigned k = 1; k < nt; ++k) {
    const size_t from = std::max(start[k - 1], size / nt * k);
    const void* gt = from < size ? memchr(data + from, '>', size - from) : nullptr;
    start[k] = gt ? (size_t)((const uint8_t*)gt - data) : size;
  }
  struct Chunk {
    size_t bytes = 0;                 // compacted output bytes, at data[start .. start + bytes)
    std::vector<RecordSpan> recs;     // offsets relative to the chunk's start
    std::vector<uint8_t> errs;        // first err_cap offending bytes, in order
    size_t nerr = 0;
    bool has_n = false;               // a sequence holds an allowed byte that is not A/C/G/T ('N')
  };
  std::vector<Chunk> chunks(nt);
  auto work = [&](unsigned k) {
    Chunk& ch = chunks[k];
    uint8_t* const base = data + start[k];
    const uint8_t* p = base;
    const uint8_t* const end = data + start[k + 1];
    uint8_t* w = base;
    bool have_rec = false;  // chunk 0 begins inside the default record
    RecordSpan r{0, 0, 0, 0};
    while (p < end) {
      if (*p == '>') {
        if (have_rec) ch.recs.push_back(r);
        have_rec = true;
        // header: '>' and everything up to the next '\n' or '>' (a '>' starts the next record)
        const uint8_t* q = p + 1;
        while (q < end && *q != '\n' && *q != '>') ++q;
        r.name_off = (uint64_t)(w - base);
        r.name_len = (uint64_t)(q - p);
        memmove(w, p, (size_t)(q - p));
        w += q - p;
        r.seq_off = (uint64_t)(w - base);
        r.seq_len = 0;
        p = (q < end && *q == '\n') ? q + 1 : q;
        continue;
      }
      // sequence bytes up to the next '>' : runs of allowed bytes are copied in one piece
      const uint8_t* q = p;
      while (q < end && allowed(*q)) ++q;
      if (q > p) {
        if (have_rec) {
          if (!ch.has_n && memchr(p, 'N', (size_t)(q - p))) ch.has_n = true;
          memmove(w, p, (size_t)(q - p));
          w += q - p;
          r.seq_len += (uint64_t)(q - p);
        }
        p = q;
        continue;
      }
      if (*p != '\n') {
        if (ch.errs.size() < err_cap) ch.errs.push_back(*p);
        ++ch.nerr;
      }
      ++p;
    }
    if (have_rec) ch.recs.push_back(r);
    ch.bytes = (size_t)(w - base);
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; ++k) th.emplace_back(work, k);
    for (auto& t : th) t.join();
  }
  const auto t_parse = std::chrono::steady_clock::now();
  size_t cur = 0, nerr = 0;
  int64_t nrec = 0;
  struct Copy {
    uint8_t* dst;
    const uint8_t* src;
    size_t n;
  };
  std::vector<Copy> copies;
  for (unsigned k = 0; k < nt; ++k) {
    const Chunk& ch = chunks[k];
    if (out && cur < out_cap) copies.push_back({out + cur, data + start[k], std::min(ch.bytes, out_cap - cur)});
    for (const RecordSpan& r : ch.recs) {
      if (index && (size_t)nrec < index_cap) {
        index[4 * nrec + 0] = r.name_off + cur;
        index[4 * nrec + 1] = r.name_len;
        index[4 * nrec + 2] = r.seq_off + cur;
        index[4 * nrec + 3] = r.seq_len;
      }
      ++nrec;
    }
    for (uint8_t c : ch.errs) {
      if (err_chars && nerr < err_cap) err_chars[nerr] = c;
      ++nerr;
    }
    nerr += ch.nerr - ch.errs.size();
    cur += ch.bytes;
  }
  if (copies.size() <= 1) {
    for (const Copy& c : copies) memcpy(c.dst, c.src, c.n);
  } else {
    std::vector<std::thread> th;
    for (const Copy& c : copies) th.emplace_back([c] { memcpy(c.dst, c.src, c.n); });
    for (auto& t : th) t.join();
  }
  if (n_err) *n_err = nerr;
  if (out_len) *out_len = cur;
  const auto t_gather = std::chrono::steady_clock::now();
  // The packer, fused into the parse: the 2-bit codes of the WHOLE output buffer (names and 'N'
  // code as 0), so the record offsets of `index` address both formats.
  if (packed && out) {
    const uint64_t have = std::min<uint64_t>(cur, out_cap);
    if ((have + 3) / 4 > packed_cap) return SA_E_ARG;
    pack_parallel<false>(out, have, packed, nt);
  }
  if (all_acgt)
    for (const Chunk& ch : chunks)
      if (ch.has_n) *all_acgt = 0;
  if (trace)
    fprintf(stderr, "[sa trace] parse_fasta %zu bytes, %u threads: parse %.1f ms, gather %.1f ms, pack %.1f ms\n", size, nt,
            std::chrono::duration<double, std::milli>(t_parse - t_read).count(),
            std::chrono::duration<double, std::milli>(t_gather - t_parse).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_gather).count());
  return nrec;
}

}  // namespace

extern "C" {

int64_t sa_parse_fasta(const char* path, uint8_t* out, size_t out_cap, uint64_t* index, size_t index_cap,
                       uint8_t* err_chars, size_t err_cap, size_t* n_err) {
  try {
    return parse_impl(path, out, out_cap, index, index_cap, err_chars, err_cap, n_err, nullptr, 0, nullptr, nullptr);
  } catch (const std::exception&) {
    return SA_E_NOMEM;
  }
}

int64_t sa_parse_fasta_packed(const char* path, uint8_t* out, size_t out_cap, uint64_t* index, size_t index_cap,
                              uint8_t* err_chars, size_t err_cap, size_t* n_err, uint8_t* packed, size_t packed_cap,
                              int* all_acgt, uint64_t* out_len) {
  try {
    return parse_impl(path, out, out_cap, index, index_cap, err_chars, err_cap, n_err, packed, packed_cap, all_acgt, out_len);
  } catch (const std::exception&) {
    return SA_E_NOMEM;
  }
}

sa_status_t sa_pack_2bit_mt(const uint8_t* src, uint64_t n, uint8_t* dst, int n_threads) {
  if (n && (!src || !dst)) return SA_E_ARG;
  try {
    const unsigned nt = n_threads > 0 ? (unsigned)n_threads : host_threads();
    return pack_parallel<true>(src, n, dst, nt) ? SA_OK : SA_E_ARG;
  } catch (const std::exception&) {
    return SA_E_NOMEM;
  }
}


// 2-bit packer for sa_batch_t.packing = 1: appends n residues (A, C, G, T only) to `dst`
// starting at residue index dst_pos.  Returns SA_OK, or SA_E_ARG at the first other byte
// (e.g. 'N': such inputs stay in the byte format, where raw bytes are compared).
sa_status_t sa_pack_2bit(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_pos) {
  static const int8_t code[256] = {
#define X -1
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
      X, 0, X, 1, X, X, X, 2, X, X, X, X, X, X, X, X, X, X, X, X, 3, X, X, X, X, X, X, X, X, X, X, X,
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
      X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X
#undef X
  };
  if (n && (!src || !dst)) return SA_E_ARG;
  uint64_t k = 0;
  auto put_one = [&](uint64_t i) -> bool {
    const int c = code[src[i]];
    if (c < 0) return false;
    const uint64_t pos = dst_pos + i;
    uint8_t& b = dst[pos >> 2];
    const int sh = 2 * (int)(pos & 3);
    b = (uint8_t)((b & ~(3 << sh)) | (c << sh));
    return true;
  };
  // head: up to the next output byte boundary
  for (; k < n && ((dst_pos + k) & 3); ++k)
    if (!put_one(k)) return SA_E_ARG;
  // body: four residues -> one byte
  uint8_t* out = dst + ((dst_pos + k) >> 2);
  for (; k + 4 <= n; k += 4) {
    const int c0 = code[src[k]], c1 = code[src[k + 1]], c2 = code[src[k + 2]], c3 = code[src[k + 3]];
    if ((c0 | c1 | c2 | c3) < 0) return SA_E_ARG;
    *out++ = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
  }
  for (; k < n; ++k)
    if (!put_one(k)) return SA_E_ARG;
  return SA_OK;
}

// Renders "alignment found\n\nseq1: ..\n      ..\nseq2: ..\n" for one CIGAR.
// Returns the number of bytes needed (snprintf style).
int64_t sa_render_affine(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                         const uint32_t* cigar, uint32_t cigar_len, char* buf, size_t cap) {
  std::string r1, r2;
  uint32_t x = 0, y = 0;  // consumed residues of seq2 / seq1
  for (uint32_t k = 0; k < cigar_len; ++k) {
    const uint32_t op = cigar[k] & 3u, len = cigar[k] >> 2;
    for (uint32_t t = 0; t < len; ++t) {
      if (op == SA_OP_M) {
        if (y >= n1 || x >= n2) return SA_E_ARG;
        r1.push_back((char)seq1[y++]);
        r2.push_back((char)seq2[x++]);
      } else if (op == SA_OP_I) {
        if (y >= n1) return SA_E_ARG;
        r1.push_back((char)seq1[y++]);
        r2.push_back('-');
      } else if (op == SA_OP_D) {
        if (x >= n2) return SA_E_ARG;
        r1.push_back('-');
        r2.push_back((char)seq2[x++]);
      } else {
        return SA_E_ARG;
      }
    }
  }
  std::string bars(r1.size(), ' ');
  for (size_t k = 0; k < r1.size(); ++k)
    if (r1[k] == r2[k]) bars[k] = '|';
  std::string s = "alignment found\n\nseq1: " + r1 + "\n      " + bars + "\nseq2: " + r2 + "\n";
  if (buf && cap) {
    const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size();
}

// The text the reference's linear aligner prints for one hit: println!("\nHit: {}\n", hit)
// (needleman_wunsch.rs:207/:211) with Display for Hit (:155-178).  The alignment ends at cell
// (end1, end2) and its CIGAR runs backwards from there (global mode: end = (n1, n2); local mode:
// the start cell sa_result_t.end1/end2 name).  "start in seq1/seq2" are the coordinates of the last
// cell the recursion left before it printed (:215-216: max(i,1)-1, max(j,1)-1), i.e. one step
// after the alignment's first column; an empty hit prints 0 / 0 (Hit::default()).
int64_t sa_render_linear_hit(const uint8_t* seq1, uint32_t n1, const uint8_t* seq2, uint32_t n2,
                             const uint32_t* cigar, uint32_t cigar_len, uint32_t end1, uint32_t end2,
                             char* buf, size_t cap) {
  if (end1 > n1 || end2 > n2) return SA_E_ARG;
  uint64_t c1 = 0, c2 = 0;  // residues the alignment consumes
  for (uint32_t k = 0; k < cigar_len; ++k) {
    const uint32_t op = cigar[k] & 3u, len = cigar[k] >> 2;
    if (op == SA_OP_M) { c1 += len; c2 += len; }
    else if (op == SA_OP_I) c1 += len;
    else if (op == SA_OP_D) c2 += len;
    else return SA_E_ARG;
  }
  if (c1 > end1 || c2 > end2) return SA_E_ARG;
  uint32_t i = (uint32_t)(end1 - c1), j = (uint32_t)(end2 - c2);  // the cell the hit was printed at
  uint32_t s1 = 0, s2 = 0;
  if (cigar_len) {  // the cell one move before it
    const uint32_t op = cigar[0] & 3u;
    const uint32_t pi = i + (op != SA_OP_D ? 1u : 0u), pj = j + (op != SA_OP_I ? 1u : 0u);
    s1 = (pi > 1 ? pi : 1) - 1;
    s2 = (pj > 1 ? pj : 1) - 1;
  }
  std::string r1, r2;
  for (uint32_t k = 0; k < cigar_len; ++k) {
    const uint32_t op = cigar[k] & 3u, len = cigar[k] >> 2;
    for (uint32_t t = 0; t < len; ++t) {
      r1.push_back(op == SA_OP_D ? '-' : (char)seq1[i++]);
      r2.push_back(op == SA_OP_I ? '-' : (char)seq2[j++]);
    }
  }
  std::string bars(r1.size(), ' ');
  for (size_t k = 0; k < r1.size(); ++k)
    if (r1[k] == r2[k]) bars[k] = '|';
  std::string s = "\nHit: \nseq1: " + r1 + "\n      " + bars + "\nseq2: " + r2 + "\nstart in seq1: " + std::to_string(s1) +
                  "\nstart in seq2: " + std::to_string(s2) + "\n\n\n\n";
  if (buf && cap) {
    const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size();
}

}  // extern "C"
