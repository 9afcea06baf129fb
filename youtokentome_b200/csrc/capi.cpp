// capi.cpp — flat C entry points over the host C++ surface (bpe_b200.h) for the Python package
// (youtokentome_b200/_lib.py binds them with ctypes; it plays the role of the reference's
// Cython module youtokentome/cpp/yttm.pyx).  Strings are UTF-8; errors come back as
// (non-zero return, message in the handle / thread-local buffer).
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/bpe_b200.h"
#include "../../include/yttm_b200_api.h"

using namespace vkcom;

namespace {
thread_local std::string g_err;
struct Handle {
  BaseEncoder *enc = nullptr;
  std::string err;
  std::vector<int32_t> ids;
  std::vector<uint64_t> offs;
  std::string text;  // last string result
};
int fail(Handle *h, const std::string &m) { (h ? h->err : g_err) = m; return 1; }
}  // namespace

extern "C" {

const char *yttm_api_last_error(void *hv) { return hv ? static_cast<Handle *>(hv)->err.c_str() : g_err.c_str(); }

int yttm_api_train(const char *data_path, const char *model_path, int vocab_size, double coverage, int n_threads,
                   int pad_id, int unk_id, int bos_id, int eos_id) {
  BpeConfig cfg(coverage, n_threads, SpecialTokens(pad_id, unk_id, bos_id, eos_id));
  Status st = train_bpe(data_path, model_path, vocab_size, cfg);
  if (!st.ok()) return fail(nullptr, st.message);
  return 0;
}

int yttm_api_train_memory(const char *text, uint64_t n, const char *model_path, int vocab_size, double coverage,
                          int pad_id, int unk_id, int bos_id, int eos_id) {
  BpeConfig cfg(coverage, 1, SpecialTokens(pad_id, unk_id, bos_id, eos_id));
  std::string data(text, text + n);
  BPEState state;
  Status st = learn_bpe_from_string(data, vocab_size, model_path ? model_path : "", cfg, &state);
  if (!st.ok()) return fail(nullptr, st.message);
  return 0;
}

void yttm_api_release_training_cache(void) { release_training_cache(); }

int yttm_api_train_report(double *out, int n) {
  const TrainReport &r = last_train_report();
  double v[] = {(double)r.n_bytes, (double)r.data_len, (double)r.n_words, (double)r.n_unique, (double)r.n_tokens,
                (double)r.n_pairs, (double)r.n_merges, r.read_s, r.h2d_ms, r.char_hist_ms, r.word_count_ms,
                r.tokenise_ms, r.pair_hist_ms, r.merge_loop_ms, r.total_s, (double)r.launches};
  int m = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < m; i++) out[i] = v[i];
  return m;
}

void *yttm_api_open(const char *model_path, int n_threads) {
  Status st;
  auto *h = new Handle();
  h->enc = new BaseEncoder(std::string(model_path), n_threads, &st);
  if (!st.ok()) {
    g_err = st.message;
    delete h->enc;
    delete h;
    return nullptr;
  }
  return h;
}

void yttm_api_close(void *hv) {
  auto *h = static_cast<Handle *>(hv);
  if (!h) return;
  delete h->enc;
  delete h;
}

int yttm_api_vocab_size(void *hv) { return static_cast<Handle *>(hv)->enc->vocab_size(); }

void yttm_api_set_dropout_seed(void *hv, uint64_t seed) { static_cast<Handle *>(hv)->enc->set_dropout_seed(seed); }

int yttm_api_encode_ids(void *hv, const char *bytes, const uint64_t *offsets, uint64_t n_sent, int bos, int eos,
                        int reverse, double dropout, uint64_t *total_ids) {
  auto *h = static_cast<Handle *>(hv);
  Status st = h->enc->encode_packed(bytes, offsets, n_sent, &h->ids, &h->offs, bos != 0, eos != 0, reverse != 0,
                                    dropout);
  if (!st.ok()) return fail(h, st.message);
  *total_ids = h->ids.size();
  return 0;
}

void yttm_api_result_ids(void *hv, int32_t *ids, uint64_t *offsets) {
  auto *h = static_cast<Handle *>(hv);
  if (!h->ids.empty()) std::memcpy(ids, h->ids.data(), h->ids.size() * 4);
  std::memcpy(offsets, h->offs.data(), h->offs.size() * 8);
}

// pieces joined by '\x01', sentences terminated by '\n'; returns the byte length (fetch with
// yttm_api_result_text) or -1.
int64_t yttm_api_encode_subwords(void *hv, const char *bytes, const uint64_t *offsets, uint64_t n_sent, int bos,
                                 int eos, int reverse, double dropout) {
  auto *h = static_cast<Handle *>(hv);
  std::vector<std::string> s(n_sent);
  for (uint64_t i = 0; i < n_sent; i++) s[i].assign(bytes + offsets[i], bytes + offsets[i + 1]);
  std::vector<std::vector<std::string>> out;
  Status st = h->enc->encode_as_subwords(s, &out, bos != 0, eos != 0, reverse != 0, dropout);
  if (!st.ok()) { fail(h, st.message); return -1; }
  h->text.clear();
  for (auto &sent : out) {
    for (size_t j = 0; j < sent.size(); j++) {
      if (j) h->text.push_back('\x01');
      h->text += sent[j];
    }
    h->text.push_back('\n');
  }
  return (int64_t)h->text.size();
}

void yttm_api_result_text(void *hv, char *out) {
  auto *h = static_cast<Handle *>(hv);
  std::memcpy(out, h->text.data(), h->text.size());
}

int64_t yttm_api_decode(void *hv, const int32_t *ids, const uint64_t *offsets, uint64_t n_sent,
                        const int32_t *ignore, uint64_t n_ignore) {
  auto *h = static_cast<Handle *>(hv);
  std::unordered_set<int> ign(ignore, ignore + n_ignore);
  h->text.clear();
  for (uint64_t i = 0; i < n_sent; i++) {
    std::vector<int> v(ids + offsets[i], ids + offsets[i + 1]);
    std::string sent;
    Status st = h->enc->decode(v, &sent, &ign);
    if (!st.ok()) { fail(h, st.message); return -1; }
    h->text += sent;
    h->text.push_back('\n');
  }
  return (int64_t)h->text.size();
}

int64_t yttm_api_id_to_subword(void *hv, int id) {
  auto *h = static_cast<Handle *>(hv);
  Status st = h->enc->id_to_subword(id, &h->text);
  if (!st.ok()) { fail(h, st.message); return -1; }
  return (int64_t)h->text.size();
}

int yttm_api_subword_to_id(void *hv, const char *subword) {
  return static_cast<Handle *>(hv)->enc->subword_to_id(subword);
}

int64_t yttm_api_vocab(void *hv) {  // '\n'-joined is ambiguous (a piece may be "\n"): use '\x01'
  auto *h = static_cast<Handle *>(hv);
  h->text.clear();
  auto v = h->enc->vocabulary();
  for (size_t i = 0; i < v.size(); i++) {
    if (i) h->text.push_back('\x01');
    h->text += v[i];
  }
  return (int64_t)h->text.size();
}

int yttm_api_encode_cli(void *hv, const char *output_type, int stream, int bos, int eos, int reverse, double dropout) {
  auto *h = static_cast<Handle *>(hv);
  Status st = h->enc->encode_cli(output_type, stream != 0, bos != 0, eos != 0, reverse != 0, dropout);
  if (!st.ok()) return fail(h, st.message);
  return 0;
}
int yttm_api_decode_cli(void *hv, const int32_t *ignore, uint64_t n_ignore) {
  auto *h = static_cast<Handle *>(hv);
  std::unordered_set<int> ign(ignore, ignore + n_ignore);
  Status st = h->enc->decode_cli(&ign);
  if (!st.ok()) return fail(h, st.message);
  return 0;
}
void yttm_api_vocab_cli(void *hv, int verbose) { static_cast<Handle *>(hv)->enc->vocab_cli(verbose != 0); }

int yttm_api_dump_order(const uint32_t *filled, uint64_t n, uint32_t *out) {
  const std::vector<uint32_t> o = reference_dump_order(std::vector<uint32_t>(filled, filled + n));
  if (o.size() != n) return 1;  // duplicate keys in `filled`
  std::copy(o.begin(), o.end(), out);
  return 0;
}

int yttm_api_redump(const char *in_path, const char *out_path) {
  BPEState st;
  if (!st.load(in_path).ok()) return 1;
  st.dump(out_path);
  return 0;
}

// raw handles for bench.py / tests that drive the device ABI of yttm_b200.h directly
void *yttm_api_device_context(void *hv) { return static_cast<Handle *>(hv)->enc->device_context(); }
void *yttm_api_device_encoder(void *hv) { return static_cast<Handle *>(hv)->enc->device_encoder(); }

}  // extern "C"
