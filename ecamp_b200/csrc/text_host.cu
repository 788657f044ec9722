// Host side of the report branch in native code (no CUDA in this file): entity-centred context masking and the template
// re-weighting of `ContextBertDataset` (ECAMP/Pre-training/module/pretrain_datasets.py:60-110 and :141-184), the per-token
// Python loops that bound the reference's loader.  The decisions and their ORDER are those of the reference; the random
// numbers are handed in pre-drawn (`random.random()` values in draw order), and the number the function will consume is a
// function of the tokens alone (text_mask_draw_count), so the caller draws exactly that many and Python's generator is
// left at the same position as after the reference code.  Pinned bit-exactly by tests/golden/text_masking.json.
#include "kernels.cuh"

#include <vector>

namespace ecamp {
namespace {
constexpr long long T_PAD = 0, T_MASK = 3, T_PERIOD = 16;        // mimic_wordpiece.json ids (pretrain_datasets.py:72-90)
constexpr long long TEMPLATE1[5] = {219, 149, 152, 422, 158};    // "there is no evidence of"   (:23)
constexpr long long TEMPLATE2[3] = {219, 149, 152};              // "there is no"               (:24)

inline bool flag(const uint8_t* table, int vocab, long long id) { return id >= 0 && id < vocab && table[id] != 0; }
inline bool contains(const std::vector<int>& v, int x) {
  for (int y : v)
    if (y == x) return true;
  return false;
}
}  // namespace

// draws of loop 2 (one per visited non-continuation position before the first [PAD]) + loop 3 (one per entity position)
int text_mask_draw_count(const long long* ids, int T, const uint8_t* is_sub, const uint8_t* is_entity, int vocab) {
  int n = 0;
  for (int i = 1; i < T - 1; ++i) {
    const long long cur = ids[i];
    if (cur == T_PAD) break;
    if (flag(is_sub, vocab, cur)) continue;
    n += flag(is_entity, vocab, cur) ? 2 : 1;
  }
  return n;
}

int text_context_mask(const long long* ids, int T, const uint8_t* is_sub, const uint8_t* is_entity, int vocab,
                      const double* draws, int n_draws, long long* masked, int* mask_pos, int* n_mask_pos) {
  ECAMP_REQUIRE(ids && is_sub && is_entity && masked && mask_pos && n_mask_pos && T >= 2, "text_context_mask: bad argument");
  ECAMP_REQUIRE(n_draws == 0 || draws, "text_context_mask: null draws");
  for (int i = 0; i < T; ++i) masked[i] = ids[i];
  std::vector<int> entity_pos, mpos;
  bool entity_exist = false;
  for (int i = 1; i < T - 1; ++i)
    if (flag(is_entity, vocab, masked[i])) { entity_exist = true; break; }
  int used = 0;
  for (int i = 1; i < T - 1; ++i) {
    const long long cur = masked[i];
    if (cur == T_PAD) break;
    if (flag(is_sub, vocab, cur)) {
      if (masked[i - 1] == T_MASK) masked[i] = T_MASK;  // a continuation piece follows its (masked) head
      continue;
    }
    if (flag(is_entity, vocab, cur)) {
      entity_pos.push_back(i);
      for (int j = 1; j < 3; ++j) {
        if (i - j <= 0) break;
        if (ids[i - j] != T_PERIOD && !contains(mpos, i - j)) mpos.push_back(i - j);
      }
    }
    ECAMP_REQUIRE(used < n_draws, "text_context_mask: ran out of random draws (%d given)", n_draws);
    const double prob = draws[used++];
    if (!entity_exist) {
      if (prob < 0.75) masked[i] = T_MASK;
    } else if (prob < 0.7 && !contains(entity_pos, i) && !contains(mpos, i)) {
      masked[i] = T_MASK;
    }
  }
  for (int i = 1; i < T - 1; ++i) {  // mask entity on 75 % prob
    if (contains(entity_pos, i)) {
      ECAMP_REQUIRE(used < n_draws, "text_context_mask: ran out of random draws (%d given)", n_draws);
      if (draws[used++] < 0.75) masked[i] = T_MASK;
    }
  }
  ECAMP_REQUIRE(used == n_draws, "text_context_mask: %d draws given, %d consumed", n_draws, used);
  for (size_t k = 0; k < mpos.size(); ++k) mask_pos[k] = mpos[k];  // at most 2 per entity position: < T entries
  *n_mask_pos = (int)mpos.size();
  return 0;
}

int text_template_weights(const long long* ids, int n_ids, const int* mask_pos, int n_mask_pos, int max_len, float* w) {
  ECAMP_REQUIRE(ids && w && n_ids >= 0 && max_len >= n_ids && (n_mask_pos == 0 || mask_pos), "text_template_weights: bad argument");
  for (int i = 0; i < max_len; ++i) w[i] = 1.0f;
  std::vector<uint8_t> dim((size_t)max_len, 0);
  int diminish_cnt = 0, i = 0;
  auto match = [&](const long long* tpl, int len) {
    for (int k = 0; k < len; ++k)
      if (ids[i + k] != tpl[k]) return false;
    return true;
  };
  while (i < n_ids - 4) {
    if (match(TEMPLATE1, 5)) {
      for (int k = 0; k < 5; ++k) { w[i + k] = 0.05f; dim[i + k] = 1; }
      diminish_cnt += 5; i += 5;
    } else if (match(TEMPLATE2, 3)) {
      for (int k = 0; k < 3; ++k) { w[i + k] = 0.05f; dim[i + k] = 1; }
      diminish_cnt += 3; i += 3;
    } else {
      ++i;
    }
  }
  int len_dm = 0;
  for (int k = 0; k < n_mask_pos; ++k) {
    ECAMP_REQUIRE(mask_pos[k] >= 0 && mask_pos[k] < max_len, "text_template_weights: mask position out of range");
    len_dm += dim[mask_pos[k]];
  }
  if (n_mask_pos > 0 && diminish_cnt > 0) {  // the removed weight goes to the entity-context positions (double, then fp32)
    const double expand = (0.95 * (double)(diminish_cnt - len_dm) + (double)n_mask_pos) / ((double)n_mask_pos - 0.95 * (double)len_dm);
    const float e = (float)expand;
    for (int k = 0; k < n_mask_pos; ++k) w[mask_pos[k]] = w[mask_pos[k]] * e;
  } else if (diminish_cnt > 0) {             // ... or is spread over the whole report
    const double expand = (double)max_len / ((double)max_len - 0.95 * (double)diminish_cnt);
    const float e = (float)expand;
    for (int k = 0; k < max_len; ++k) w[k] = w[k] * e;
  }
  return 0;
}

}  // namespace ecamp
