#include "db_format.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <queue>

namespace prib {

// ------------------------------------------------------------------------------------------------
// FASTA
// ------------------------------------------------------------------------------------------------
static void strip_eol(std::string &line) {
  // fastafile_reader.cpp:58-67: "\r\n" then a single trailing '\r' or '\n'
  if (line.size() >= 2 && line.compare(line.size() - 2, 2, "\r\n") == 0) line.erase(line.size() - 2, 2);
  if (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
}

bool read_fasta(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs,
                std::string &err) {
  std::ifstream fp(path.c_str(), std::ios::in);
  if (!fp) {
    err = "Error: can't open input_file: " + path + ".";
    return false;
  }
  std::string line;
  bool have = false;
  while (std::getline(fp, line)) {
    if (!have || (!line.empty() && line[0] == '>')) {
      // the first line is always taken as a header (fastafile_reader.cpp:53-55)
      names.push_back(line.empty() ? std::string() : line.substr(1));
      seqs.emplace_back();
      have = true;
    } else {
      strip_eol(line);
      seqs.back() += line;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// encoding
// ------------------------------------------------------------------------------------------------
bool encode_reversed(const std::vector<std::string> &seqs, size_t first, size_t count, int repeat_flag,
                     std::vector<uint8_t> &out) {
  if (repeat_flag < 0 || repeat_flag > 2) return false;
  uint8_t table[256];
  std::memset(table, 1, sizeof(table));
  table[(int)'A'] = 2;
  table[(int)'C'] = 3;
  table[(int)'G'] = 4;
  table[(int)'T'] = 5;
  table[(int)'U'] = 5;
  if (repeat_flag != 0) {
    const uint8_t base = repeat_flag == 1 ? 6 : 2;
    table[(int)'a'] = base;
    table[(int)'c'] = base + 1;
    table[(int)'g'] = base + 2;
    table[(int)'t'] = base + 3;
    table[(int)'u'] = base + 3;
  }
  for (size_t k = first; k < first + count; k++) {
    const std::string &s = seqs[k];
    for (size_t j = s.size(); j-- > 0;) out.push_back(table[(unsigned char)s[j]]);
    out.push_back(0);
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// suffix array: prefix doubling, each round one stable two-key radix sort on (rank[i], rank[i+k])
// ------------------------------------------------------------------------------------------------
void build_suffix_array(const uint8_t *text, int n, std::vector<int32_t> &sa) {
  sa.resize((size_t)std::max(n, 0));
  if (n <= 0) return;
  std::vector<int32_t> rank(n), tmp(n), sa2(n);
  std::vector<int32_t> cnt((size_t)std::max(n, 256) + 2);
  // round 0: by first byte
  std::fill(cnt.begin(), cnt.begin() + 257, 0);
  for (int i = 0; i < n; i++) cnt[text[i] + 1]++;
  for (int c = 0; c < 256; c++) cnt[c + 1] += cnt[c];
  for (int i = 0; i < n; i++) sa[cnt[text[i]]++] = i;
  int classes = 0;
  rank[sa[0]] = 0;
  for (int i = 1; i < n; i++) {
    if (text[sa[i]] != text[sa[i - 1]]) classes++;
    rank[sa[i]] = classes;
  }
  classes++;
  for (long long k = 1; classes < n && k < n; k <<= 1) {
    // sort by second key: suffixes whose second half is past the end come first (shorter = smaller),
    // then the others in the order of the current SA shifted left by k
    int m = 0;
    for (int i = n - (int)std::min<long long>(k, n); i < n; i++) sa2[m++] = i;
    for (int i = 0; i < n; i++)
      if (sa[i] >= k) sa2[m++] = sa[i] - (int32_t)k;
    // stable counting sort by first key
    std::fill(cnt.begin(), cnt.begin() + classes + 1, 0);
    for (int i = 0; i < n; i++) cnt[rank[i] + 1]++;
    for (int c = 0; c < classes; c++) cnt[c + 1] += cnt[c];
    for (int i = 0; i < n; i++) sa[cnt[rank[sa2[i]]]++] = sa2[i];
    // new ranks
    tmp[sa[0]] = 0;
    int cls = 0;
    for (int i = 1; i < n; i++) {
      const int a = sa[i - 1], b = sa[i];
      const int ra2 = a + k < n ? rank[a + k] : -1, rb2 = b + k < n ? rank[b + k] : -1;
      if (rank[a] != rank[b] || ra2 != rb2) cls++;
      tmp[b] = cls;
    }
    rank.swap(tmp);
    classes = cls + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// k-mer interval hash
// ------------------------------------------------------------------------------------------------
// One narrowing step: from the interval [*start, *end] of suffixes sharing a prefix of length `offset`,
// to those whose next character is c.  Behaviour follows DbConstruction::Search line by line, including
// the start bump for a suffix that ends exactly at `offset` while the local copy keeps the old start.
static void narrow(const std::vector<uint8_t> &text, const std::vector<int32_t> &sa, int32_t *start, int32_t *end,
                   uint8_t c, int offset) {
  int32_t s = *start, e = *end;
  const size_t n = text.size();
  // (the reference reads out of bounds when a suffix is shorter than `offset`; that cannot happen for
  //  intervals of ACGU k-mers because the text ends with a sentinel: treated as a 0 byte here)
  auto ch = [&](int32_t k) -> uint8_t {
    const size_t pos = (size_t)sa[k] + offset;
    return pos < n ? text[pos] : (uint8_t)0;
  };
  if ((size_t)s < sa.size() && (size_t)((unsigned)(sa[s] + offset)) >= n) ++(*start);
  if (s > e) {
    *start = 1;
    *end = 0;
    return;
  }
  if (s == e) {
    if (ch(s) != c) {
      *start = 1;
      *end = 0;
    }
    return;
  }
  if (ch(s) != c) {
    while (s < e - 1) {
      const int32_t m = (s + e) / 2;
      if (ch(m) < c) s = m;
      else e = m;
    }
    if (ch(e) != c) {
      *start = 1;
      *end = 0;
      return;
    }
    *start = e;
    s = e;
    e = *end;
  }
  if (ch(e) != c) {
    while (s < e - 1) {
      const int32_t m = (s + e) / 2;
      if (ch(m) > c) e = m;
      else s = m;
    }
    if (ch(s) != c) {
      *start = 1;
      *end = 0;
      return;
    }
    *end = s;
  }
}

void build_kmer_hash(const std::vector<uint8_t> &text, const std::vector<int32_t> &sa, int hash_size,
                     std::vector<std::vector<int32_t>> &start_hash, std::vector<std::vector<int32_t>> &end_hash) {
  start_hash.assign((size_t)std::max(hash_size, 0), {});
  end_hash.assign((size_t)std::max(hash_size, 0), {});
  for (int i = 0; i < hash_size; i++) {
    const int count = (int)std::pow(4, i + 1);
    start_hash[i].reserve(count);
    end_hash[i].reserve(count);
    for (int j = 0; j < count; j++) {
      const uint8_t c = (uint8_t)((j % 4) + 2);
      int32_t s, e;
      if (i == 0) {
        s = 0;
        e = (int32_t)sa.size() - 1;
      } else {
        s = start_hash[i - 1][j / 4];
        e = end_hash[i - 1][j / 4];
      }
      narrow(text, sa, &s, &e, c, i);
      start_hash[i].push_back(s);
      end_hash[i].push_back(e);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// writers
// ------------------------------------------------------------------------------------------------
static bool put(std::FILE *f, const void *p, size_t bytes) { return bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes; }

bool write_bas(const std::string &db, const DbParams &p, std::string &err) {
  std::FILE *f = std::fopen((db + ".bas").c_str(), "wb");
  if (!f) {
    err = "Error: can't open " + db + ".bas";
    return false;
  }
  const int32_t v[4] = {p.hash_size, p.repeat_flag, p.maximal_span, p.min_accessible_length};
  const bool ok = put(f, v, sizeof(v));
  std::fclose(f);
  if (!ok) err = "Error: short write on " + db + ".bas";
  return ok;
}

bool write_nam(const std::string &db, const std::vector<std::string> &names, std::string &err) {
  std::FILE *f = std::fopen((db + ".nam").c_str(), "w");
  if (!f) {
    err = "Error: can't open name file";
    return false;
  }
  bool ok = true;
  for (const std::string &n : names) ok = ok && put(f, n.data(), n.size()) && put(f, "\n", 1);
  std::fclose(f);
  if (!ok) err = "Error: short write on " + db + ".nam";
  return ok;
}

bool write_acc(const std::string &db, const std::vector<std::string> &seqs, const float *image,
               const std::vector<int64_t> &acc_off, const std::vector<int64_t> &cond_off, int delta, std::string &err) {
  std::FILE *f = std::fopen((db + ".acc").c_str(), "wb");
  if (!f) {
    err = "Error: can't open " + db + ".acc";
    return false;
  }
  bool ok = true;
  for (size_t k = 0; k < seqs.size() && ok; k++) {
    // one record per sequence, raccess.cpp:447-481: n1, acc[0..n1), L, cond[0..L) (cond[0..delta) are zeros)
    const int32_t L = (int32_t)seqs[k].size(), n1 = L - delta + 1;
    ok = put(f, &n1, 4) && put(f, image + acc_off[k], 4 * (size_t)n1) && put(f, &L, 4) &&
         put(f, image + cond_off[k], 4 * (size_t)L);
  }
  std::fclose(f);
  if (!ok) err = "Error: short write on " + db + ".acc";
  return ok;
}

bool write_seq_ind(const std::string &db, const std::vector<std::string> &seqs, const DbParams &p, std::string &err,
                   SaBuilder sa_builder) {
  std::FILE *fs = std::fopen((db + ".seq").c_str(), "wb");
  std::FILE *fi = std::fopen((db + ".ind").c_str(), "wb");
  if (!fs || !fi) {
    if (fs) std::fclose(fs);
    if (fi) std::fclose(fi);
    err = "Error: can't open " + db + ".seq/.ind";
    return false;
  }
  bool ok = true;
  const size_t n = seqs.size(), chunk = (size_t)std::max(p.chunk_size, 1);
  // pages of `chunk` sequences, db_construction.cpp:116-144
  for (size_t first = 0; first < n && ok; first += chunk) {
    const size_t count = std::min(chunk, n - first);
    std::vector<uint8_t> text;
    std::vector<int32_t> lens(count);
    for (size_t k = 0; k < count; k++) lens[k] = (int32_t)seqs[first + k].size();
    if (!encode_reversed(seqs, first, count, p.repeat_flag, text)) {
      err = "Error: -r option must be 0, 1, or 2";
      ok = false;
      break;
    }
    std::vector<int32_t> sa;
    if (sa_builder) {
      if (!sa_builder(text.data(), (int)text.size(), sa, err)) {
        ok = false;
        break;
      }
    } else {
      build_suffix_array(text.data(), (int)text.size(), sa);
    }
    std::vector<std::vector<int32_t>> sh, eh;
    build_kmer_hash(text, sa, p.hash_size, sh, eh);
    const int32_t nseq = (int32_t)count, nbytes = (int32_t)text.size();
    ok = put(fs, &nseq, 4) && put(fs, lens.data(), 4 * count) && put(fs, &nbytes, 4) && put(fs, text.data(), text.size());
    ok = ok && put(fi, &nbytes, 4) && put(fi, sa.data(), 4 * sa.size());
    for (auto &v : sh) ok = ok && put(fi, v.data(), 4 * v.size());
    for (auto &v : eh) ok = ok && put(fi, v.data(), 4 * v.size());
  }
  std::fclose(fs);
  std::fclose(fi);
  if (!ok && err.empty()) err = "Error: short write on " + db + ".seq/.ind";
  return ok;
}

// ------------------------------------------------------------------------------------------------
// partition
// ------------------------------------------------------------------------------------------------
void lpt_partition(const std::vector<std::string> &seqs, int parts, std::vector<std::vector<int>> &part) {
  part.assign((size_t)std::max(parts, 1), {});
  std::vector<int> order(seqs.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seqs[a].size() > seqs[b].size(); });
  typedef std::pair<long long, int> Load;  // (assigned nucleotides, device)
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int d = 0; d < (int)part.size(); d++) heap.push(Load(0, d));
  for (int idx : order) {
    Load l = heap.top();
    heap.pop();
    part[l.second].push_back(idx);
    l.first += (long long)seqs[idx].size();
    heap.push(l);
  }
}

}  // namespace prib
