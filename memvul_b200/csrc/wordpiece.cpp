// libmemvul_tok.so -- batched ASCII fast path of BERT's BasicTokenizer + WordPiece (include/memvul_tok.h).
// Host-only C++17, no dependencies.  Two tries over the vocabulary (word-initial pieces and "##" continuation
// pieces) stored in one open-addressed (node, byte) -> node table make a piece lookup O(length).
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/memvul_tok.h"

namespace {

thread_local char g_err[256] = "";

struct Trie {
  // nodes are numbered from 0 (root of the word-initial trie) and 1 (root of the continuation trie)
  std::vector<int32_t> token_of;            // node -> token id or -1
  std::vector<uint64_t> keys;               // open addressing: key = node * 256 + byte + 1 (0 = empty)
  std::vector<int32_t> vals;
  uint64_t mask = 0;

  static uint64_t hash(uint64_t k) { k *= 0x9E3779B97F4A7C15ull; return k ^ (k >> 29); }
  int32_t child(int32_t node, unsigned char c) const {
    const uint64_t key = (static_cast<uint64_t>(node) << 8) + c + 1;
    for (uint64_t h = hash(key) & mask;; h = (h + 1) & mask) {
      if (keys[h] == key) return vals[h];
      if (keys[h] == 0) return -1;
    }
  }
  void put(int32_t node, unsigned char c, int32_t child_node) {
    const uint64_t key = (static_cast<uint64_t>(node) << 8) + c + 1;
    for (uint64_t h = hash(key) & mask;; h = (h + 1) & mask) {
      if (keys[h] == 0) { keys[h] = key; vals[h] = child_node; return; }
    }
  }
};

struct Tok {
  Trie trie;
  int lowercase = 1;
  int32_t unk = -1, cls = -1, sep = -1;
  std::unordered_map<std::string, int32_t> vocab;
};

inline bool is_punct(unsigned char c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }
inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
inline bool is_dropped(unsigned char c) { return (c < 0x20 && !is_space(c)) || c == 0x7f; }   // Cc except \t \n \r

bool build(Tok* t, const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { snprintf(g_err, sizeof(g_err), "cannot open vocabulary %s", path); return false; }
  std::vector<std::string> toks;
  std::string line;
  int ch;
  while ((ch = fgetc(f)) != EOF) {
    if (ch == '\n') { if (!line.empty() && line.back() == '\r') line.pop_back(); toks.push_back(line); line.clear(); }
    else line.push_back(static_cast<char>(ch));
  }
  if (!line.empty()) toks.push_back(line);
  fclose(f);
  size_t edges = 2;
  for (const auto& s : toks) edges += s.size();
  uint64_t cap = 1;
  while (cap < 2 * edges + 16) cap <<= 1;
  t->trie.keys.assign(cap, 0);
  t->trie.vals.assign(cap, -1);
  t->trie.mask = cap - 1;
  t->trie.token_of.assign(2, -1);
  for (size_t id = 0; id < toks.size(); ++id) {
    const std::string& s = toks[id];
    if (s.empty()) continue;
    t->vocab.emplace(s, static_cast<int32_t>(id));       // first occurrence wins, like a dict built in file order would not: HF keeps the LAST; vocab files have no duplicates
    bool ascii = true;
    for (unsigned char c : s) if (c >= 0x80) { ascii = false; break; }
    if (!ascii) continue;                                 // cannot match an ASCII text
    int32_t node = 0;
    size_t start = 0;
    if (s.size() > 2 && s[0] == '#' && s[1] == '#') { node = 1; start = 2; }
    for (size_t i = start; i < s.size(); ++i) {
      const unsigned char c = static_cast<unsigned char>(s[i]);
      int32_t nx = t->trie.child(node, c);
      if (nx < 0) {
        nx = static_cast<int32_t>(t->trie.token_of.size());
        t->trie.token_of.push_back(-1);
        t->trie.put(node, c, nx);
      }
      node = nx;
    }
    if (node > 1 && t->trie.token_of[node] < 0) t->trie.token_of[node] = static_cast<int32_t>(id);
  }
  auto find = [&](const char* s) { auto it = t->vocab.find(s); return it == t->vocab.end() ? -1 : it->second; };
  t->unk = find("[UNK]"); t->cls = find("[CLS]"); t->sep = find("[SEP]");
  if (t->unk < 0 || t->cls < 0 || t->sep < 0) { snprintf(g_err, sizeof(g_err), "vocabulary lacks [UNK]/[CLS]/[SEP]"); return false; }
  return true;
}

// word [w, w+len) (already lower-cased, no whitespace / punctuation inside unless it IS a single punctuation char)
inline void wordpiece(const Tok* t, const unsigned char* w, int len, int64_t* out, int& n, int cap) {
  if (n >= cap) return;
  if (len > 100) { out[n++] = t->unk; return; }
  const int n0 = n;
  int start = 0;
  while (start < len) {
    int32_t node = start == 0 ? 0 : 1, best = -1;
    int best_end = start;
    for (int i = start; i < len; ++i) {
      node = t->trie.child(node, w[i]);
      if (node < 0) break;
      const int32_t tk = t->trie.token_of[node];
      if (tk >= 0) { best = tk; best_end = i + 1; }
    }
    if (best < 0) { n = n0; out[n++] = t->unk; return; }   // the whole word becomes [UNK]
    if (n < cap) out[n++] = best;
    else { /* truncated: later pieces cannot change earlier ones, but an unmatched remainder would turn the word into
              [UNK]; keep scanning without storing */ }
    start = best_end;
  }
}

// returns false if the text needs the full-Unicode path
bool encode_one(const Tok* t, const unsigned char* s, int64_t len, int add_special, int max_length, int64_t* row, int32_t* out_len) {
  for (int64_t i = 0; i < len; ++i) {
    if (s[i] >= 0x80) return false;
    if (s[i] == '[' && i + 4 < len && (!memcmp(s + i, "[UNK]", 5) || !memcmp(s + i, "[SEP]", 5) || !memcmp(s + i, "[PAD]", 5) ||
                                      !memcmp(s + i, "[CLS]", 5) || (i + 5 < len && !memcmp(s + i, "[MASK]", 6))))
      return false;
  }
  const int cap = add_special ? std::max(max_length - 2, 0) : max_length;    // pieces kept
  // a word that straddles the truncation point may still become [UNK] as a whole, so pieces are produced into a
  // buffer with head room and cut afterwards
  std::vector<int64_t> buf(static_cast<size_t>(cap) + 128);
  int n = 0;
  unsigned char word[104];
  int wl = 0;
  bool too_long = false;
  auto flush = [&]() {
    if (wl == 0 && !too_long) return;
    if (n < cap) {
      if (too_long) buf[n++] = t->unk;
      else wordpiece(t, word, wl, buf.data(), n, cap + 120);
    }
    wl = 0; too_long = false;
  };
  for (int64_t i = 0; i < len && n < cap; ++i) {
    unsigned char c = s[i];
    if (is_dropped(c)) continue;
    if (is_space(c)) { flush(); continue; }
    if (is_punct(c)) { flush(); word[0] = c; wl = 1; flush(); continue; }
    if (t->lowercase && c >= 'A' && c <= 'Z') c = static_cast<unsigned char>(c + 32);
    if (wl < 101) word[wl++] = c; else too_long = true;
    if (wl == 101) { too_long = true; }
  }
  flush();
  if (n > cap) n = cap;
  int k = 0;
  if (add_special && max_length >= 1) row[k++] = t->cls;
  for (int i = 0; i < n && k < max_length; ++i) row[k++] = buf[i];
  if (add_special && k < max_length) row[k++] = t->sep;
  *out_len = k;
  return true;
}

}  // namespace

extern "C" {

const char* memvul_tok_last_error(void) { return g_err; }

void* memvul_tok_create(const char* vocab_path, int lowercase) {
  if (!vocab_path) { snprintf(g_err, sizeof(g_err), "null vocabulary path"); return nullptr; }
  Tok* t = new Tok();
  t->lowercase = lowercase;
  if (!build(t, vocab_path)) { delete t; return nullptr; }
  return t;
}

void memvul_tok_destroy(void* tok) { delete static_cast<Tok*>(tok); }

int32_t memvul_tok_token_to_id(const void* tok, const char* token) {
  if (!tok || !token) return -1;
  const Tok* t = static_cast<const Tok*>(tok);
  auto it = t->vocab.find(token);
  return it == t->vocab.end() ? -1 : it->second;
}

int memvul_tok_encode_batch(const void* tok, const char* data, const int64_t* offsets, int n, int add_special,
                            int max_length, int64_t* out_ids, int32_t* out_lens, uint8_t* status, int n_threads) {
  if (!tok || !data || !offsets || n < 0 || max_length <= 0 || !out_ids || !out_lens || !status) {
    snprintf(g_err, sizeof(g_err), "encode_batch: invalid argument");
    return -1;
  }
  const Tok* t = static_cast<const Tok*>(tok);
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  n_threads = std::max(1, std::min(n_threads, std::max(1, n / 8)));
  std::atomic<int> next{0}, fallback{0};
  auto work = [&]() {
    for (;;) {
      const int i0 = next.fetch_add(16);
      if (i0 >= n) return;
      for (int i = i0; i < std::min(n, i0 + 16); ++i) {
        int64_t* row = out_ids + static_cast<int64_t>(i) * max_length;
        out_lens[i] = 0;
        const bool ok = encode_one(t, reinterpret_cast<const unsigned char*>(data) + offsets[i], offsets[i + 1] - offsets[i],
                                   add_special, max_length, row, &out_lens[i]);
        status[i] = ok ? 0 : 1;
        if (!ok) fallback.fetch_add(1);
      }
    }
  };
  std::vector<std::thread> th;
  for (int k = 1; k < n_threads; ++k) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  return fallback.load();
}

}  // extern "C"
