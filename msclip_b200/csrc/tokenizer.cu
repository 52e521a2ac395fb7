// Native byte-pair-encoding tokenizer: host code, no GPU work.  Restates
// lib/dataset/languages/simple_tokenizer.py of the reference (SimpleTokenizer.__init__ :64-82, bpe :84-123, encode :125-131,
// tokenize :150-166) for whole batches of prompts on all host cores:
//   whitespace_clean (re.sub(r'\s+', ' ', text).strip(), :59-62) -> str.lower() -> findall of the CLIP pattern
//   <|startoftext|> | <|endoftext|> | 's | 't | 're | 've | 'm | 'll | 'd | \p{L}+ | \p{N} | [^\s\p{L}\p{N}]+   (IGNORECASE)
//   -> UTF-8 bytes mapped to the reversible byte alphabet (bytes_to_unicode :21-41) -> greedy lowest-rank pair merges
//   -> vocabulary ids -> [SOT] + ids + [EOT], truncated to the context length, zero padded.
// `basic_clean` (ftfy.fix_text + html.unescape, :53-56 - third-party text repair) stays in the Python wrapper.
// The Unicode behaviour of the reference comes from two libraries (Python's str methods, the regex module); the tables in
// unicode_tables.inc are read off those libraries code point by code point (tools/gen_unicode_tables.py), including the
// Final_Sigma rule of str.lower() and regex's simple case folding of U+017F in the contraction alternatives.
// The merges file (bpe_simple_vocab_16e6.txt.gz) is the reference's own data and is NOT shipped: the caller passes its text.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/msclip_b200.h"
#include "common.cuh"

namespace msclip {
namespace {

#include "unicode_tables.inc"

bool in_ranges(const uint32_t (*r)[2], int n, uint32_t c) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) / 2;
    if (c < r[mid][0]) hi = mid - 1;
    else if (c > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
bool is_letter(uint32_t c) { return c < 128 ? ((c | 32) >= 'a' && (c | 32) <= 'z') : in_ranges(kLetterRanges, kLetterRangesCount, c); }
bool is_number(uint32_t c) { return c < 128 ? (c >= '0' && c <= '9') : in_ranges(kNumberRanges, kNumberRangesCount, c); }
bool is_space(uint32_t c) { return in_ranges(kRegexSpace, kRegexSpaceCount, c); }
// [^\s\p{L}\p{N}] under IGNORECASE: NOT simply the complement of the three classes (see tools/gen_unicode_tables.py)
bool is_other(uint32_t c) { return in_ranges(kOtherRanges, kOtherRangesCount, c); }
bool is_py_space(uint32_t c) { return is_space(c) || (c >= 0x1c && c <= 0x1f); }  // str.isspace(): what str.strip() removes

uint32_t lower_simple(uint32_t c) {
  if (c < 128) return (c >= 'A' && c <= 'Z') ? c + 32 : c;
  int lo = 0, hi = kLowerMapCount - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) / 2;
    if (c < kLowerMap[mid][0]) hi = mid - 1;
    else if (c > kLowerMap[mid][0]) lo = mid + 1;
    else return kLowerMap[mid][1];
  }
  return c;
}

// UTF-8 -> code points.  The wrapper hands over what Python's str.encode('utf-8') produced, i.e. valid UTF-8.
void decode_utf8(const char* s, size_t n, std::vector<uint32_t>& out) {
  out.clear();
  size_t i = 0;
  while (i < n) {
    const uint8_t b = static_cast<uint8_t>(s[i]);
    uint32_t c;
    int len;
    if (b < 0x80) { c = b; len = 1; }
    else if ((b >> 5) == 6) { c = b & 31; len = 2; }
    else if ((b >> 4) == 14) { c = b & 15; len = 3; }
    else { c = b & 7; len = 4; }
    for (int k = 1; k < len && i + k < n; ++k) c = (c << 6) | (static_cast<uint8_t>(s[i + k]) & 63);
    out.push_back(c);
    i += len;
  }
}
void append_utf8(uint32_t c, std::string& out) {
  if (c < 0x80) out.push_back(static_cast<char>(c));
  else if (c < 0x800) { out.push_back(static_cast<char>(0xC0 | (c >> 6))); out.push_back(static_cast<char>(0x80 | (c & 63))); }
  else if (c < 0x10000) {
    out.push_back(static_cast<char>(0xE0 | (c >> 12)));
    out.push_back(static_cast<char>(0x80 | ((c >> 6) & 63)));
    out.push_back(static_cast<char>(0x80 | (c & 63)));
  } else {
    out.push_back(static_cast<char>(0xF0 | (c >> 18)));
    out.push_back(static_cast<char>(0x80 | ((c >> 12) & 63)));
    out.push_back(static_cast<char>(0x80 | ((c >> 6) & 63)));
    out.push_back(static_cast<char>(0x80 | (c & 63)));
  }
}

// str.lower(): simple mappings, U+0130 -> "i" + U+0307, and the Final_Sigma rule for U+03A3 (CPython handle_capital_sigma)
void lower_text(const std::vector<uint32_t>& in, std::vector<uint32_t>& out) {
  out.clear();
  const long n = static_cast<long>(in.size());
  for (long i = 0; i < n; ++i) {
    const uint32_t c = in[i];
    if (c == 0x130) {
      out.push_back('i');
      out.push_back(0x307);
    } else if (c == 0x3A3) {
      long j = i - 1;
      while (j >= 0 && in_ranges(kCaseIgnorable, kCaseIgnorableCount, in[j])) --j;
      bool final_sigma = j >= 0 && in_ranges(kCasedNI, kCasedNICount, in[j]);
      if (final_sigma) {
        j = i + 1;
        while (j < n && in_ranges(kCaseIgnorable, kCaseIgnorableCount, in[j])) ++j;
        final_sigma = j == n || !in_ranges(kCasedNI, kCasedNICount, in[j]);
      }
      out.push_back(final_sigma ? 0x3C2 : 0x3C3);
    } else {
      out.push_back(lower_simple(c));
    }
  }
}

struct Tokenizer {
  std::string byte_sym[256];                        // byte -> its symbol of the reversible alphabet, UTF-8 encoded
  std::unordered_map<std::string, int> encoder;     // symbol (sequence) -> id
  std::unordered_map<std::string, int> ranks;       // "first\x01second" -> merge rank
  int sot = 0, eot = 0, vocab = 0;
  static constexpr int kShards = 64;
  std::unordered_map<std::string, std::vector<int>> cache[kShards];  // word (byte symbols) -> ids
  std::mutex lock[kShards];
};

// bytes_to_unicode (simple_tokenizer.py:21-41): printable bytes map to themselves, the others to U+0100 + n
void build_byte_alphabet(Tokenizer& t, std::vector<int>& order) {
  std::vector<int> bs;
  for (int b = '!'; b <= '~'; ++b) bs.push_back(b);
  for (int b = 0xA1; b <= 0xAC; ++b) bs.push_back(b);
  for (int b = 0xAE; b <= 0xFF; ++b) bs.push_back(b);
  std::vector<uint32_t> cs(bs.begin(), bs.end());
  int n = 0;
  for (int b = 0; b < 256; ++b)
    if (std::find(bs.begin(), bs.end(), b) == bs.end()) {
      bs.push_back(b);
      cs.push_back(256 + n++);
    }
  for (size_t i = 0; i < bs.size(); ++i) {
    t.byte_sym[bs[i]].clear();
    append_utf8(cs[i], t.byte_sym[bs[i]]);
  }
  order = bs;
}

// word = symbols of one pre-token (last one carries "</w>") -> ids after the greedy merges of SimpleTokenizer.bpe
void bpe_word(const Tokenizer& t, std::vector<std::string>& word, std::vector<int>& ids) {
  while (word.size() > 1) {
    int best = INT32_MAX;
    size_t best_i = 0;
    for (size_t i = 0; i + 1 < word.size(); ++i) {
      auto it = t.ranks.find(word[i] + '\x01' + word[i + 1]);
      if (it != t.ranks.end() && it->second < best) {
        best = it->second;
        best_i = i;
      }
    }
    if (best == INT32_MAX) break;
    // merge EVERY occurrence of the best pair, left to right (the reference's while loop over word.index)
    const std::string first = word[best_i], second = word[best_i + 1];
    std::vector<std::string> merged;
    merged.reserve(word.size());
    for (size_t i = 0; i < word.size();) {
      if (i + 1 < word.size() && word[i] == first && word[i + 1] == second) {
        merged.push_back(first + second);
        i += 2;
      } else {
        merged.push_back(word[i]);
        i += 1;
      }
    }
    word.swap(merged);
  }
  ids.clear();
  for (const std::string& sym : word) {
    auto it = t.encoder.find(sym);
    ids.push_back(it == t.encoder.end() ? 0 : it->second);  // every reachable symbol is in the vocabulary by construction
  }
}

void encode_pretoken(Tokenizer& t, const std::string& utf8, std::vector<int>& out) {
  if (utf8 == "<|startoftext|>") { out.push_back(t.sot); return; }   // pre-seeded cache entries of the reference (:80)
  if (utf8 == "<|endoftext|>") { out.push_back(t.eot); return; }
  std::string key;
  for (unsigned char b : utf8) key += t.byte_sym[b];
  const size_t shard = std::hash<std::string>()(key) % Tokenizer::kShards;
  {
    std::lock_guard<std::mutex> g(t.lock[shard]);
    auto it = t.cache[shard].find(key);
    if (it != t.cache[shard].end()) {
      out.insert(out.end(), it->second.begin(), it->second.end());
      return;
    }
  }
  std::vector<std::string> word;
  for (unsigned char b : utf8) word.push_back(t.byte_sym[b]);
  word.back() += "</w>";
  std::vector<int> ids;
  bpe_word(t, word, ids);
  out.insert(out.end(), ids.begin(), ids.end());
  std::lock_guard<std::mutex> g(t.lock[shard]);
  t.cache[shard].emplace(std::move(key), std::move(ids));
}

bool starts_with(const std::vector<uint32_t>& s, size_t i, const char* lit) {
  for (size_t k = 0; lit[k]; ++k)
    if (i + k >= s.size() || s[i + k] != static_cast<uint32_t>(static_cast<unsigned char>(lit[k]))) return false;
  return true;
}
// length of a contraction alternative at i ('s 't 're 've 'm 'll 'd, IGNORECASE with simple case folding: U+017F counts as s)
size_t contraction_len(const std::vector<uint32_t>& s, size_t i) {
  if (s[i] != '\'' || i + 1 >= s.size()) return 0;
  auto fold = [](uint32_t c) -> uint32_t { return c == 0x17F ? 's' : (c == 0x212A ? 'k' : lower_simple(c)); };
  const uint32_t a = fold(s[i + 1]);
  const uint32_t b = i + 2 < s.size() ? fold(s[i + 2]) : 0;
  if (a == 's' || a == 't') return 2;
  if (a == 'r' && b == 'e') return 3;
  if (a == 'v' && b == 'e') return 3;
  if (a == 'm') return 2;
  if (a == 'l' && b == 'l') return 3;
  if (a == 'd') return 2;
  return 0;
}

void encode_text(Tokenizer& t, const char* text, size_t len, std::vector<int>& ids) {
  std::vector<uint32_t> raw, clean, s;
  decode_utf8(text, len, raw);
  // whitespace_clean: runs of \s -> one space, then str.strip()
  for (size_t i = 0; i < raw.size();) {
    if (is_space(raw[i])) {
      while (i < raw.size() && is_space(raw[i])) ++i;
      clean.push_back(' ');
    } else {
      clean.push_back(raw[i++]);
    }
  }
  size_t b = 0, e = clean.size();
  while (b < e && is_py_space(clean[b])) ++b;
  while (e > b && is_py_space(clean[e - 1])) --e;
  clean.assign(clean.begin() + b, clean.begin() + e);
  lower_text(clean, s);
  ids.clear();
  std::string tok;
  size_t i = 0;
  while (i < s.size()) {
    size_t n = 0;
    if (starts_with(s, i, "<|startoftext|>")) n = 15;
    else if (starts_with(s, i, "<|endoftext|>")) n = 13;
    else if ((n = contraction_len(s, i)) != 0) {}
    else if (is_letter(s[i])) { while (i + n < s.size() && is_letter(s[i + n])) ++n; }
    else if (is_number(s[i])) n = 1;
    else if (is_other(s[i])) { while (i + n < s.size() && is_other(s[i + n])) ++n; }
    else { ++i; continue; }   // whitespace (and the few case-folding oddities) match no alternative
    tok.clear();
    for (size_t k = 0; k < n; ++k) append_utf8(s[i + k], tok);
    encode_pretoken(t, tok, ids);
    i += n;
  }
}

}  // namespace
}  // namespace msclip

using namespace msclip;

extern "C" {

// merges_text: the decompressed text of the reference's bpe_simple_vocab_16e6.txt.gz (first line = header)
int msclip_tokenizer_create(const char* merges_text, int64_t length, void** out) {
  MSCLIP_REQUIRE(merges_text != nullptr && length > 0 && out != nullptr, "msclip_tokenizer_create: bad arguments");
  Tokenizer* t = new Tokenizer();
  std::vector<int> order;
  build_byte_alphabet(*t, order);
  int id = 0;
  for (int b : order) t->encoder[t->byte_sym[b]] = id++;
  for (int b : order) t->encoder[t->byte_sym[b] + "</w>"] = id++;
  // merges = text.split('\n')[1 : 49152 - 256 - 2 + 1]; each line = two symbols separated by whitespace (:68-70)
  const int64_t max_merges = 49152 - 256 - 2;
  int64_t pos = 0, line = 0, rank = 0;
  while (pos < length && rank < max_merges) {
    int64_t end = pos;
    while (end < length && merges_text[end] != '\n') ++end;
    if (line >= 1) {
      std::string ln(merges_text + pos, merges_text + end);
      const size_t sp = ln.find(' ');
      if (sp == std::string::npos || sp == 0 || sp + 1 >= ln.size()) {
        delete t;
        MSCLIP_REQUIRE(false, "msclip_tokenizer_create: malformed merges line " + std::to_string(line));
      }
      const std::string a = ln.substr(0, sp), b = ln.substr(sp + 1);
      t->ranks[a + '\x01' + b] = static_cast<int>(rank++);
      t->encoder[a + b] = id++;
    }
    pos = end + 1;
    ++line;
  }
  if (rank != max_merges) {
    delete t;
    MSCLIP_REQUIRE(false, "msclip_tokenizer_create: the merges file holds fewer than 48894 merges");
  }
  t->sot = id++;
  t->eot = id++;
  t->vocab = id;
  *out = t;
  return 0;
}

int msclip_tokenizer_destroy(void* tok) {
  delete static_cast<Tokenizer*>(tok);
  return 0;
}

int msclip_tokenizer_info(void* tok, int* vocab_size, int* sot, int* eot) {
  MSCLIP_REQUIRE(tok != nullptr, "msclip_tokenizer_info: null tokenizer");
  const Tokenizer* t = static_cast<const Tokenizer*>(tok);
  if (vocab_size) *vocab_size = t->vocab;
  if (sot) *sot = t->sot;
  if (eot) *eot = t->eot;
  return 0;
}

// ids of one cleaned text (SimpleTokenizer.encode without basic_clean): returns the count, writes at most `capacity` ids
int64_t msclip_tokenizer_encode(void* tok, const char* text, int64_t length, int32_t* ids_out, int64_t capacity) {
  if (tok == nullptr || text == nullptr || length < 0) return -1;
  std::vector<int> ids;
  encode_text(*static_cast<Tokenizer*>(tok), text, static_cast<size_t>(length), ids);
  for (int64_t i = 0; i < static_cast<int64_t>(ids.size()) && i < capacity; ++i) ids_out[i] = ids[i];
  return static_cast<int64_t>(ids.size());
}

// SimpleTokenizer.tokenize (:150-166) for n texts (UTF-8, text i = texts + offsets[i], offsets[n] = total length):
// out [n, context_length] int64 = [SOT] + ids + [EOT] truncated to context_length, zero padded; all host cores
int msclip_tokenizer_tokenize(void* tok, const char* texts, const int64_t* offsets, int n, int context_length, int64_t* out,
                              int threads) {
  MSCLIP_REQUIRE(tok != nullptr && texts != nullptr && offsets != nullptr && out != nullptr && n >= 0 && context_length >= 1,
                 "msclip_tokenizer_tokenize: bad arguments");
  Tokenizer& t = *static_cast<Tokenizer*>(tok);
  int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (nt > n) nt = n > 0 ? n : 1;
  auto work = [&](int tid) {
    std::vector<int> ids;
    for (int i = tid; i < n; i += nt) {
      encode_text(t, texts + offsets[i], static_cast<size_t>(offsets[i + 1] - offsets[i]), ids);
      int64_t* row = out + static_cast<int64_t>(i) * context_length;
      int k = 0;
      row[k++] = t.sot;
      for (size_t j = 0; j < ids.size() && k < context_length; ++j) row[k++] = ids[j];
      if (k < context_length) row[k++] = t.eot;
      for (; k < context_length; ++k) row[k] = 0;
    }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int tid = 0; tid < nt; ++tid) pool.emplace_back(work, tid);
    for (auto& th : pool) th.join();
  }
  return 0;
}

}  // extern "C"
