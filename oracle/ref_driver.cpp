// oracle/ref_driver.cpp — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern-"C" wrapper that is compiled TOGETHER WITH the unmodified reference
// sources where they lie (/root/reference/youtokentome/cpp/{bpe,utils,utf8}.cpp) by
// oracle/Makefile into oracle/_ref/libyttm_ref_{det,prod}.so.  No reference source is
// copied into this repository; this file only calls the reference's public C++ API
// (bpe.h:19 train_bpe, bpe.h:22-82 BaseEncoder) plus learn_bpe_from_string, which the
// reference's own stress test also declares by hand (tests/unit_tests/stress_test.h:8-13).
//
//   det  build: -DDETERMINISTIC_QUEUE  -> parity oracle (thread-invariant tie-breaks)
//   prod build: shipped flags          -> timing baseline ("cpu_baseline.kind = reference")
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "bpe.h"
#include "utf8.h"
#include "utils.h"

namespace vkcom {
Status learn_bpe_from_string(std::string &text_utf8, int n_tokens, const std::string &output_file,
                             BpeConfig bpe_config, BPEState *bpe_state);
}

namespace {
void set_err(char *err, int errlen, const std::string &m) {
  if (err && errlen > 0) {
    std::strncpy(err, m.c_str(), errlen - 1);
    err[errlen - 1] = 0;
  }
}
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct RefEncoder {
  vkcom::BaseEncoder *enc = nullptr;
  std::vector<std::vector<int>> ids;
  std::vector<std::vector<std::string>> pieces;
};
}  // namespace

extern "C" {

// train_bpe (bpe.cpp:1368): reads the file, trains, writes the model.  Returns Status.code.
int ref_train_file(const char *input_path, const char *model_path, int vocab_size, double coverage, int n_threads,
                   int pad_id, int unk_id, int bos_id, int eos_id, double *seconds, char *err, int errlen) {
  vkcom::BpeConfig cfg(coverage, n_threads, vkcom::SpecialTokens(pad_id, unk_id, bos_id, eos_id));
  double t0 = now_s();
  vkcom::Status st = vkcom::train_bpe(input_path, model_path, vocab_size, cfg);
  if (seconds) *seconds = now_s() - t0;
  if (!st.ok()) set_err(err, errlen, st.message);
  return st.code;
}

// learn_bpe_from_string (bpe.cpp:859) on an in-memory corpus (the entry the stress test uses).
// n_threads must already be in [1,8] (check_config, bpe.cpp:1345-1348, is not run on this path).
int ref_train_memory(const char *text, uint64_t n, const char *model_path, int vocab_size, double coverage,
                     int n_threads, int pad_id, int unk_id, int bos_id, int eos_id, double *seconds, char *err,
                     int errlen) {
  std::string data(text, text + n);
  vkcom::BpeConfig cfg(coverage, n_threads, vkcom::SpecialTokens(pad_id, unk_id, bos_id, eos_id));
  vkcom::BPEState state;
  double t0 = now_s();
  vkcom::Status st = vkcom::learn_bpe_from_string(data, vocab_size, model_path, cfg, &state);
  if (seconds) *seconds = now_s() - t0;
  if (!st.ok()) set_err(err, errlen, st.message);
  return st.code;
}

void *ref_encoder_new(const char *model_path, int n_threads, char *err, int errlen) {
  vkcom::Status st;
  auto *h = new RefEncoder();
  h->enc = new vkcom::BaseEncoder(std::string(model_path), n_threads, &st);
  if (!st.ok()) {
    set_err(err, errlen, st.message);
    delete h->enc;
    delete h;
    return nullptr;
  }
  return h;
}

void ref_encoder_free(void *hv) {
  auto *h = static_cast<RefEncoder *>(hv);
  if (!h) return;
  delete h->enc;
  delete h;
}

int ref_vocab_size(void *hv) { return static_cast<RefEncoder *>(hv)->enc->vocab_size(); }

// encode_as_ids (bpe.cpp:1740).  Sentences are bytes[offsets[i], offsets[i+1]).  Only the
// encode_as_ids call itself is timed (the std::vector<std::string> marshalling is excluded,
// as SURVEY.md §8d specifies).  Results stay in the handle; fetch with ref_result_*.
int ref_encode_ids(void *hv, const char *bytes, const uint64_t *offsets, uint64_t n_sent, int bos, int eos,
                   int reverse, double dropout, double *seconds, uint64_t *total_ids, char *err, int errlen) {
  auto *h = static_cast<RefEncoder *>(hv);
  std::vector<std::string> s(n_sent);
  for (uint64_t i = 0; i < n_sent; i++) s[i].assign(bytes + offsets[i], bytes + offsets[i + 1]);
  h->ids.clear();
  double t0 = now_s();
  vkcom::Status st = h->enc->encode_as_ids(s, &h->ids, bos != 0, eos != 0, reverse != 0, dropout);
  if (seconds) *seconds = now_s() - t0;
  if (!st.ok()) {
    set_err(err, errlen, st.message);
    return st.code;
  }
  uint64_t tot = 0;
  for (auto &v : h->ids) tot += v.size();
  if (total_ids) *total_ids = tot;
  return 0;
}

// Copies the last encode_as_ids result: out_offsets has n_sent+1 entries.
void ref_result_ids(void *hv, int32_t *out_ids, uint64_t *out_offsets) {
  auto *h = static_cast<RefEncoder *>(hv);
  uint64_t pos = 0;
  for (size_t i = 0; i < h->ids.size(); i++) {
    out_offsets[i] = pos;
    for (int v : h->ids[i]) out_ids[pos++] = v;
  }
  out_offsets[h->ids.size()] = pos;
}

// encode_as_subwords (bpe.cpp:1757); pieces of all sentences joined by '\x01', sentences by '\n'.
// Returns needed length; copies min(len, cap) bytes.
int64_t ref_encode_subwords(void *hv, const char *bytes, const uint64_t *offsets, uint64_t n_sent, int bos, int eos,
                            int reverse, char *out, int64_t cap) {
  auto *h = static_cast<RefEncoder *>(hv);
  std::vector<std::string> s(n_sent);
  for (uint64_t i = 0; i < n_sent; i++) s[i].assign(bytes + offsets[i], bytes + offsets[i + 1]);
  h->pieces.clear();
  vkcom::Status st = h->enc->encode_as_subwords(s, &h->pieces, bos != 0, eos != 0, reverse != 0, 0.0);
  if (!st.ok()) return -1;
  std::string joined;
  for (auto &sent : h->pieces) {
    for (size_t j = 0; j < sent.size(); j++) {
      if (j) joined.push_back('\x01');
      joined += sent[j];
    }
    joined.push_back('\n');
  }
  int64_t need = (int64_t)joined.size();
  if (out && cap > 0) std::memcpy(out, joined.data(), (size_t)std::min<int64_t>(need, cap));
  return need;
}

// decode (bpe.cpp:1843) of one id sequence; returns needed length.
int64_t ref_decode_ids(void *hv, const int32_t *ids, uint64_t n, char *out, int64_t cap) {
  auto *h = static_cast<RefEncoder *>(hv);
  std::vector<int> v(ids, ids + n);
  std::string sent;
  vkcom::Status st = h->enc->decode(v, &sent, nullptr);
  if (!st.ok()) return -1;
  int64_t need = (int64_t)sent.size();
  if (out && cap > 0) std::memcpy(out, sent.data(), (size_t)std::min<int64_t>(need, cap));
  return need;
}

// The reference's own container: fill a flat_hash_map<uint32_t,uint32_t> the way compute_alphabet_helper
// does (operator[] per key, bpe.cpp:340-353), copy it the way `*bpe_state = {char2id, ...}` does
// (bpe.cpp:1289) and list the copy's iteration order, i.e. the order BPEState::dump writes (utils.cpp:57).
// Pins the product's slot-order replay (yttm_api_dump_order) on arbitrary key sets.
int ref_char2id_order(const uint32_t *keys, uint64_t n, uint32_t *out) {
  vkcom::flat_hash_map<uint32_t, uint32_t> filled;
  for (uint64_t i = 0; i < n; i++) filled[keys[i]] = (uint32_t)i;
  vkcom::BPEState state = {filled, {}, vkcom::SpecialTokens()};
  uint64_t j = 0;
  for (auto kv : state.char2id) out[j++] = kv.first;
  return j == n ? 0 : 1;
}

int ref_is_deterministic_queue() {
#ifdef DETERMINISTIC_QUEUE
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"
