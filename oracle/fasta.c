/*
 * fasta.c -- ORACLE (test infrastructure, never linked into the product).
 *
 * Literal restatement of parse_fasta, /root/reference/src/parse.rs:52-106.
 * Pinned by the reference's own tests parse.rs:166-251 (tests/test_oracle_fasta.py).
 */
#include "sa_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* has_extension :101-106 -- Path::extension(): text after the last '.' of the file name,
 * None if there is no '.', or the name starts with '.' and has no other '.'. */
static int has_ext(const char* path, const char* ext) {
  const char* base = strrchr(path, '/');
  base = base ? base + 1 : path;
  const char* dot = strrchr(base, '.');
  if (!dot || dot == base) return 0;
  return strcmp(dot + 1, ext) == 0;
}

static int allowed(uint8_t c) { /* ALLOWED_CHARS :52 */
  return c == 'A' || c == 'G' || c == 'C' || c == 'T' || c == 'N';
}

int64_t sao_parse_fasta_mem(const uint8_t* contents, size_t n, uint8_t* out, size_t out_cap,
                            uint64_t* index, size_t index_cap, uint8_t* err_chars,
                            size_t err_cap, size_t* n_err) {
  /* Records are laid out in `out` as name bytes followed by seq bytes.  Because a record's
   * name is complete before its first seq byte arrives (:76-88) a single append cursor works. */
  size_t cur = 0, nerr = 0;
  int64_t nrec = -1; /* -1: still inside the default record that :91 removes */
  uint64_t name_off = 0, name_len = 0, seq_off = 0, seq_len = 0;
  int in_name = 0;
  for (size_t k = 0; k < n; ++k) {
    const uint8_t c = contents[k];
    if (c == '>') { /* :67-75 */
      if (nrec >= 0 && (size_t)nrec < index_cap) {
        index[4 * nrec + 0] = name_off;
        index[4 * nrec + 1] = name_len;
        index[4 * nrec + 2] = seq_off;
        index[4 * nrec + 3] = seq_len;
      }
      ++nrec;
      if (nrec == 0) cur = 0; /* drop whatever the default record had collected */
      name_off = cur;
      if (cur < out_cap) out[cur] = c;
      ++cur;
      name_len = 1;
      seq_off = cur;
      seq_len = 0;
      in_name = 1;
      continue;
    }
    if (in_name) { /* :76-81 */
      if (c == '\n') {
        in_name = 0;
        seq_off = cur;
        continue;
      }
      if (cur < out_cap) out[cur] = c;
      ++cur;
      ++name_len;
      seq_off = cur;
    } else if (c == '\n') { /* :82-83 */
      continue;
    } else if (!allowed(c)) { /* :84-85 */
      if (nerr < err_cap && err_chars) err_chars[nerr] = c;
      ++nerr;
    } else { /* :86-88 */
      if (nrec >= 0) {
        if (cur < out_cap) out[cur] = c;
        ++cur;
        ++seq_len;
      }
    }
  }
  /* :90-91 push the last record, remove the default one */
  if (nrec >= 0 && (size_t)nrec < index_cap) {
    index[4 * nrec + 0] = name_off;
    index[4 * nrec + 1] = name_len;
    index[4 * nrec + 2] = seq_off;
    index[4 * nrec + 3] = seq_len;
  }
  if (n_err) *n_err = nerr;
  return nrec + 1;
}

int64_t sao_parse_fasta_path(const char* path, uint8_t* out, size_t out_cap, uint64_t* index,
                             size_t index_cap, uint8_t* err_chars, size_t err_cap,
                             size_t* n_err) {
  if (!(has_ext(path, "fa") || has_ext(path, "fasta") || has_ext(path, "fna"))) return -1; /* :55-60 */
  FILE* f = fopen(path, "rb");
  if (!f) return -1; /* read(path)? :62 */
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* buf = (uint8_t*)malloc(sz > 0 ? (size_t)sz : 1);
  if (!buf) {
    fclose(f);
    return -1;
  }
  size_t got = fread(buf, 1, (size_t)sz, f);
  fclose(f);
  int64_t r = sao_parse_fasta_mem(buf, got, out, out_cap, index, index_cap, err_chars, err_cap, n_err);
  free(buf);
  return r;
}
