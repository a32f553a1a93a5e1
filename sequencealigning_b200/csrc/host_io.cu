// host_io.cu -- host-side pieces of the drop-in boundary (no device code).
//
//   sa_parse_fasta      parse_fasta of /root/reference/src/parse.rs:54-99, same quirks
//   sa_render_affine    the text the reference prints for one alignment
//                       (needleman_wunsch_affine.rs:283-286 + Display :390-411)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

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

}  // namespace

extern "C" {

int64_t sa_parse_fasta(const char* path, uint8_t* out, size_t out_cap, uint64_t* index,
                       size_t index_cap, uint8_t* err_chars, size_t err_cap, size_t* n_err) {
  if (n_err) *n_err = 0;
  if (!path) return SA_E_ARG;
  // parse.rs:55-60: anything but .fa/.fasta/.fna is FastaError(InvalidInput)
  if (!(has_ext(path, "fa") || has_ext(path, "fasta") || has_ext(path, "fna"))) return SA_E_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return SA_E_ARG;  // parse.rs:62 `read(path)?`
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t got;
  while ((got = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
  fclose(f);

  // The state machine of parse.rs:66-89.  A record's name is complete before its first
  // residue arrives, so records are appended to `out` as name bytes then sequence bytes.
  size_t cur = 0, nerr = 0;
  int64_t nrec = -1;  // -1 while inside the default record that parse.rs:91 removes
  RecordSpan r{0, 0, 0, 0};
  bool in_name = false;
  auto flush = [&]() {
    if (nrec >= 0 && (size_t)nrec < index_cap && index) {
      index[4 * nrec + 0] = r.name_off;
      index[4 * nrec + 1] = r.name_len;
      index[4 * nrec + 2] = r.seq_off;
      index[4 * nrec + 3] = r.seq_len;
    }
  };
  auto put = [&](uint8_t c) {
    if (cur < out_cap && out) out[cur] = c;
    ++cur;
  };
  for (uint8_t c : buf) {
    if (c == '>') {
      flush();
      ++nrec;
      if (nrec == 0) cur = 0;
      r = RecordSpan{cur, 1, cur + 1, 0};
      put(c);
      in_name = true;
      continue;
    }
    if (in_name) {
      if (c == '\n') {
        in_name = false;
        continue;
      }
      put(c);
      ++r.name_len;
      r.seq_off = cur;
    } else if (c == '\n') {
      continue;
    } else if (!allowed(c)) {
      if (err_chars && nerr < err_cap) err_chars[nerr] = c;
      ++nerr;
    } else if (nrec >= 0) {
      put(c);
      ++r.seq_len;
    }
  }
  flush();
  if (n_err) *n_err = nerr;
  return nrec + 1;
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

}  // extern "C"
