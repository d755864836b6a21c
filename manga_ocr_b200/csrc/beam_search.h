// Beam-search bookkeeping on the host: the per-step selection of
// transformers' GenerationMixin._beam_search (generation/utils.py:3076-3370 and its helpers
// _get_top_k_continuations :2945, _get_running_beams_for_next_iteration :2999, _update_finished_beams
// :3021, _check_early_stop_heuristic :2876, _beam_search_has_unfinished_sequences :2923), the
// NoRepeatNGramLogitsProcessor (generation/logits_process.py:1012-1136) and the MaxLength / EosToken
// stopping criteria, restated for one EOS id and a one-token prompt ([CLS]).
//
// The device hands over, per running beam, its 2*num_beams best continuations (log-probability after
// the n-gram ban, token id); everything else - accumulated scores, the top-2B merge over the beams of a
// crop, running / finished beam sets, length penalty, early stopping - is O(beams^2) work per crop and
// stays here, in float32 exactly as the reference computes it.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace mocr {

struct BeamSearch {
  int n = 0, beams = 0, keep = 0, max_length = 0, ngram = 0, early = 0;   // early: 0 False, 1 True, 2 "never"
  float length_penalty = 1.f;
  int eos = 3, pad = 0, prompt_len = 1, cur_len = 1;
  bool unfinished = true;
  std::vector<int32_t> running;        // [n][beams][max_length]
  std::vector<float> running_score;    // [n][beams]
  std::vector<int32_t> finished;       // [n][beams][max_length]
  std::vector<float> finished_score;   // [n][beams]
  std::vector<int32_t> finished_len;   // [n][beams] generated tokens (for the output length)
  std::vector<uint8_t> is_finished;    // [n][beams]
  std::vector<uint8_t> unsatisfied;    // [n] early-stop heuristic still allows an improvement

  void init(int n_, int beams_, int max_length_, int ngram_, float lp, int early_, int start_token) {
    n = n_; beams = beams_; keep = 2 * beams_; max_length = max_length_; ngram = ngram_; length_penalty = lp; early = early_;
    cur_len = prompt_len;
    unfinished = max_length > prompt_len;
    // (:3163 `output_fill_value = pad_token_id or eos_token_id[0]`: a pad id of 0 is falsy, so the reference fills with EOS)
    const int32_t fill = pad != 0 ? pad : eos;
    running.assign(static_cast<size_t>(n) * beams * max_length, fill);
    for (int i = 0; i < n * beams; ++i) running[static_cast<size_t>(i) * max_length] = start_token;
    finished = running;
    running_score.assign(static_cast<size_t>(n) * beams, -1e9f);
    for (int b = 0; b < n; ++b) running_score[static_cast<size_t>(b) * beams] = 0.f;
    finished_score.assign(static_cast<size_t>(n) * beams, -1e9f);
    finished_len.assign(static_cast<size_t>(n) * beams, 0);
    is_finished.assign(static_cast<size_t>(n) * beams, 0);
    unsatisfied.assign(n, 1);
  }

  // Tokens the next step may not produce for running row `row` (= crop * beams + beam): every w such that the
  // n-gram (last ngram-1 tokens, w) already occurs in the row's sequence (prompt included).
  int banned(int row, int32_t* out, int cap) const {
    if (ngram <= 0 || cur_len + 1 < ngram) return 0;
    const int32_t* s = &running[static_cast<size_t>(row) * max_length];
    const int m = ngram - 1;
    int cnt = 0;
    for (int i = 0; i + ngram <= cur_len; ++i) {
      bool same = true;
      for (int k = 0; k < m && same; ++k) same = s[i + k] == s[cur_len - m + k];
      if (same) {
        const int32_t w = s[i + m];
        bool dup = false;
        for (int k = 0; k < cnt && k < cap && !dup; ++k) dup = out[k] == w;
        if (!dup) {
          if (cnt < cap) out[cnt] = w;
          ++cnt;
        }
      }
    }
    return cnt;
  }

  // One step.  cand_lp / cand_tok: [n*beams][keep], each row sorted by descending log-probability.
  // next_tok / parent: [n*beams] the token every running row consumes next and the row whose cache it continues.
  // Returns whether the search goes on.
  bool step(const float* cand_lp, const int32_t* cand_tok, int32_t* next_tok, int32_t* parent) {
    const int K = keep;
    bool all_hit = true;
    std::vector<int32_t> new_running(static_cast<size_t>(beams) * max_length), merged_seq(static_cast<size_t>(beams + K) * max_length);
    std::vector<float> acc(static_cast<size_t>(beams) * K);
    std::vector<int> order(static_cast<size_t>(beams) * K), top(K), sel(beams);
    for (int b = 0; b < n; ++b) {
      // accumulated log-probabilities of the beams*K candidates; top-K of them (torch.topk: descending, stable on ties by index)
      for (int j = 0; j < beams; ++j)
        for (int k = 0; k < K; ++k) acc[j * K + k] = cand_lp[(static_cast<size_t>(b) * beams + j) * K + k] + running_score[static_cast<size_t>(b) * beams + j];
      // the candidates of a beam are its K best tokens: order by (value desc, flattened vocabulary index asc) as topk over [beams*vocab]
      for (int i = 0; i < beams * K; ++i) order[i] = i;
      std::stable_sort(order.begin(), order.begin() + beams * K, [&](int x, int y) {
        if (acc[x] != acc[y]) return acc[x] > acc[y];
        const int bx = x / K, by = y / K;
        if (bx != by) return bx < by;
        return cand_tok[(static_cast<size_t>(b) * beams + bx) * K + x % K] < cand_tok[(static_cast<size_t>(b) * beams + by) * K + y % K];
      });
      std::vector<float> top_lp(K), run_lp(K), fin_lp(K);
      std::vector<int> top_beam(K), top_tok(K);
      std::vector<uint8_t> hit(K);
      for (int k = 0; k < K; ++k) {
        const int c = order[k];
        top_lp[k] = acc[c];
        top_beam[k] = c / K;
        top_tok[k] = cand_tok[(static_cast<size_t>(b) * beams + c / K) * K + c % K];
        hit[k] = (cur_len + 1 >= max_length) || top_tok[k] == eos;
        all_hit = all_hit && hit[k];
        run_lp[k] = top_lp[k] + (hit[k] ? 1.0f : 0.0f) * -1.0e9f;
      }
      // ---- running beams of the next iteration: the best `beams` candidates that did not stop
      for (int k = 0; k < K; ++k) top[k] = k;
      std::stable_sort(top.begin(), top.end(), [&](int x, int y) { return run_lp[x] > run_lp[y]; });
      for (int j = 0; j < beams; ++j) sel[j] = top[j];
      // ---- finished beams: candidates among the first `beams` that stopped, length-penalised, merged with the previous set
      const float denom = static_cast<float>(pow(static_cast<double>(cur_len + 1 - prompt_len), static_cast<double>(length_penalty)));
      bool full = early == 1;
      for (int j = 0; j < beams; ++j) full = full && is_finished[static_cast<size_t>(b) * beams + j];
      for (int k = 0; k < K; ++k) {
        const bool did = hit[k] && k < beams;
        float v = top_lp[k] / denom;
        v += (full ? 1.0f : 0.0f) * -1.0e9f;
        v += (unsatisfied[b] ? 0.0f : 1.0f) * -1.0e9f;
        v += (did ? 0.0f : 1.0f) * -1.0e9f;
        fin_lp[k] = v;
      }
      std::vector<float> mscore(beams + K);
      std::vector<uint8_t> mfin(beams + K);
      std::vector<int> mlen(beams + K), midx(beams + K);
      for (int j = 0; j < beams; ++j) {
        mscore[j] = finished_score[static_cast<size_t>(b) * beams + j];
        mfin[j] = is_finished[static_cast<size_t>(b) * beams + j];
        mlen[j] = finished_len[static_cast<size_t>(b) * beams + j];
        std::copy_n(&finished[(static_cast<size_t>(b) * beams + j) * max_length], max_length, &merged_seq[static_cast<size_t>(j) * max_length]);
      }
      for (int k = 0; k < K; ++k) {
        mscore[beams + k] = fin_lp[k];
        mfin[beams + k] = hit[k] && k < beams;
        mlen[beams + k] = cur_len + 1 - prompt_len;
        int32_t* dst = &merged_seq[static_cast<size_t>(beams + k) * max_length];
        std::copy_n(&running[(static_cast<size_t>(b) * beams + top_beam[k]) * max_length], max_length, dst);
        dst[cur_len] = top_tok[k];
      }
      for (int i = 0; i < beams + K; ++i) midx[i] = i;
      std::stable_sort(midx.begin(), midx.end(), [&](int x, int y) { return mscore[x] > mscore[y]; });
      std::vector<int32_t> fin_new(static_cast<size_t>(beams) * max_length);
      std::vector<float> fs(beams);
      std::vector<uint8_t> ff(beams);
      std::vector<int> fl(beams);
      for (int j = 0; j < beams; ++j) {
        const int i = midx[j];
        std::copy_n(&merged_seq[static_cast<size_t>(i) * max_length], max_length, &fin_new[static_cast<size_t>(j) * max_length]);
        fs[j] = mscore[i];
        ff[j] = mfin[i];
        fl[j] = mlen[i];
      }
      for (int j = 0; j < beams; ++j) {
        std::copy_n(&fin_new[static_cast<size_t>(j) * max_length], max_length, &finished[(static_cast<size_t>(b) * beams + j) * max_length]);
        finished_score[static_cast<size_t>(b) * beams + j] = fs[j];
        is_finished[static_cast<size_t>(b) * beams + j] = ff[j];
        finished_len[static_cast<size_t>(b) * beams + j] = fl[j];
      }
      // ---- commit the running beams
      for (int j = 0; j < beams; ++j) {
        const int k = sel[j];
        int32_t* dst = &new_running[static_cast<size_t>(j) * max_length];
        std::copy_n(&running[(static_cast<size_t>(b) * beams + top_beam[k]) * max_length], max_length, dst);
        dst[cur_len] = top_tok[k];
        next_tok[static_cast<size_t>(b) * beams + j] = top_tok[k];
        parent[static_cast<size_t>(b) * beams + j] = b * beams + top_beam[k];
      }
      for (int j = 0; j < beams; ++j) {
        std::copy_n(&new_running[static_cast<size_t>(j) * max_length], max_length, &running[(static_cast<size_t>(b) * beams + j) * max_length]);
        running_score[static_cast<size_t>(b) * beams + j] = run_lp[sel[j]];
      }
    }
    ++cur_len;
    // ---- can the open beams still beat the finished ones?  (per crop; :2876-2921)
    bool any_unsat = false, all_fin = true;
    for (int b = 0; b < n; ++b) {
      const int best_len = (early == 2 && length_penalty > 0.0f) ? max_length - prompt_len : cur_len - prompt_len;
      const float best_possible = running_score[static_cast<size_t>(b) * beams] /
                                  static_cast<float>(pow(static_cast<double>(best_len), static_cast<double>(length_penalty)));
      float mn = finished_score[static_cast<size_t>(b) * beams];
      for (int j = 1; j < beams; ++j) mn = std::min(mn, finished_score[static_cast<size_t>(b) * beams + j]);
      bool any = false;
      for (int j = 0; j < beams; ++j) {
        const float worst = is_finished[static_cast<size_t>(b) * beams + j] ? mn : -1.0e9f;
        any = any || best_possible > worst;
        all_fin = all_fin && is_finished[static_cast<size_t>(b) * beams + j];
      }
      unsatisfied[b] = unsatisfied[b] && any;
      any_unsat = any_unsat || unsatisfied[b];
    }
    unfinished = any_unsat && !(all_fin && early == 1) && !all_hit;
    return unfinished;
  }

  // Best finished sequence of every crop (filled as the reference fills it, see init); lens = prompt + generated tokens.
  void result(int32_t* ids, int32_t* lens, float* scores) const {
    for (int b = 0; b < n; ++b) {
      std::copy_n(&finished[static_cast<size_t>(b) * beams * max_length], max_length, ids + static_cast<size_t>(b) * max_length);
      if (lens) lens[b] = prompt_len + finished_len[static_cast<size_t>(b) * beams];
      if (scores) scores[b] = finished_score[static_cast<size_t>(b) * beams];
    }
  }
};

}  // namespace mocr
