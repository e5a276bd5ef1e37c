// beam.cpp -- the two beam-search CTC decoders of EasyOCR (readtext(decoder='beamsearch' | 'wordbeamsearch')), host side.
//   easyocr/utils.py : BeamEntry, BeamState.sort / norm / wordsearch, fast_simplify_label, ctcBeamSearch,
//                      CTCLabelConverter.decode_beamsearch / decode_wordbeamsearch           
// Upstream runs these in Python over the float32 probability matrix it copied back from the device; so do we (the matrix
// comes from k_row_probs).  Every beam probability is a float64 product / sum of float32 inputs in upstream's order of
// operations, beams live in an insertion-ordered table and are ranked with a stable descending sort -- Python's dict and
// sorted(reverse=True) -- so ties resolve the same way.  No language model is ever applied upstream (prText == 1).
#include <algorithm>
#include <cstdint>
#include <map>
#include <set>
#include <vector>

#include "engine.h"

namespace bbocr {

namespace {

struct Beam {
    std::vector<int32_t> lab;
    double total = 0, non_blank = 0, blank = 0;
};

struct BeamTable {                      // Python dict: lookup by labeling, iteration in insertion order
    std::vector<Beam> rows;
    std::map<std::vector<int32_t>, int> index;
    Beam& at(const std::vector<int32_t>& lab) {
        auto it = index.find(lab);
        if (it != index.end()) return rows[it->second];
        index.emplace(lab, (int)rows.size());
        rows.emplace_back();
        rows.back().lab = lab;
        return rows.back();
    }
    // BeamState.sort(): indices by descending prTotal * prText (prText == 1.0), stable
    std::vector<int> ranked() const {
        std::vector<int> o(rows.size());
        for (size_t i = 0; i < o.size(); ++i) o[i] = (int)i;
        std::stable_sort(o.begin(), o.end(), [&](int a, int b) { return rows[a].total * 1.0 > rows[b].total * 1.0; });
        return o;
    }
};

// utils.fast_simplify_label(labeling, c, blankIdx = 0)
void extend_label(const std::vector<int32_t>& lab, int32_t c, std::vector<int32_t>& out) {
    out = lab;
    const bool has = !lab.empty();
    if (has && c == 0 && lab.back() != 0) { out.push_back(c); return; }
    if (has && c != 0 && lab.back() == 0) {
        if (lab[lab.size() - 2] == c) out.push_back(c);         // blank between equal characters stays
        else out.back() = c;                                   // blank between different characters goes
        return;
    }
    if (has && c == 0 && lab.back() == 0) return;               // consecutive blanks
    if (!has && c == 0) return;                                 // leading blank
    out.push_back(c);
}

// characters of a labeling: drop blanks and repeats (the `res` loop at the end of ctcBeamSearch / wordsearch)
void label_text(const std::vector<int32_t>& lab, std::vector<int32_t>& text) {
    text.clear();
    for (size_t i = 0; i < lab.size(); ++i)
        if (lab[i] != 0 && !(i > 0 && lab[i - 1] == lab[i])) text.push_back(lab[i]);
}

}  // namespace

// utils.ctcBeamSearch(mat, classes, ignore_idx = [0], lm = None, beamWidth, dict_list).  mat: T rows of C float32
// probabilities, row t at mat + rows[t] * C.  dict == nullptr or empty: the most probable labeling; otherwise
// BeamState.wordsearch(maxCandidate = 20): the first of the 20 best labelings whose text is a dictionary word, else the best.
void ctc_beam_search(const float* mat, const int* rows, int T, int C, int beam_width, const std::set<std::vector<int32_t>>* dict,
                     std::vector<int32_t>& text) {
    BeamTable last;
    {
        Beam& e = last.at({});
        e.blank = 1.0;
        e.total = 1.0;
    }
    const float thr = (float)(0.5 / (double)C);                 // np.where(mat[t, :] >= 0.5 / maxC): compared in float32
    std::vector<int32_t> cand, nl;
    for (int t = 0; t < T; ++t) {
        const float* m = mat + (size_t)rows[t] * C;
        cand.clear();
        for (int c = 0; c < C; ++c)
            if (m[c] >= thr) cand.push_back(c);
        BeamTable curr;
        std::vector<int> order = last.ranked();
        if ((int)order.size() > beam_width) order.resize(std::max(beam_width, 0));
        for (int bi : order) {
            const Beam prev = last.rows[bi];                    // by value: `curr` grows below, `last` does not, but keep it simple
            double pr_non_blank = 0.0;
            if (!prev.lab.empty()) pr_non_blank = prev.non_blank * (double)m[prev.lab.back()];
            const double pr_blank = prev.total * (double)m[0];
            {
                Beam& e = curr.at(prev.lab);
                e.non_blank += pr_non_blank;
                e.blank += pr_blank;
                e.total += pr_blank + pr_non_blank;
            }
            for (int32_t c : cand) {
                extend_label(prev.lab, c, nl);
                double p;
                if (!prev.lab.empty() && prev.lab.back() == c) p = (double)m[c] * prev.blank;
                else p = (double)m[c] * prev.total;
                Beam& e = curr.at(nl);
                e.non_blank += p;
                e.total += p;
            }
        }
        last = std::move(curr);
    }
    std::vector<int> order = last.ranked();
    text.clear();
    if (order.empty()) return;
    if (!dict || dict->empty()) {
        label_text(last.rows[order[0]].lab, text);
        return;
    }
    std::vector<int32_t> cur;
    for (size_t j = 0; j < order.size() && j < 20; ++j) {
        label_text(last.rows[order[j]].lab, cur);
        if (j == 0) text = cur;
        if (dict->count(cur)) { text = cur; break; }
    }
}

// CTCLabelConverter.decode_beamsearch / decode_wordbeamsearch for ONE crop: probs = T x C float32.
//   decoder 1: beam search over the whole sequence.
//   decoder 2: the arg-max path is cut at its space symbols; every run of non-space steps is searched on its own against
//              the dictionary (empty dictionary = plain beam search per word) and the words are joined by one space.
void decode_beam(const float* probs, int T, int C, int decoder, int beam_width, int space_idx,
                 const std::set<std::vector<int32_t>>* dict, std::vector<int32_t>& text) {
    std::vector<int> rows;
    text.clear();
    if (decoder == 1) {
        rows.resize(T);
        for (int t = 0; t < T; ++t) rows[t] = t;
        ctc_beam_search(probs, rows.data(), T, C, beam_width, nullptr, text);
        return;
    }
    std::vector<int32_t> word;
    bool first = true;
    int t = 0;
    while (t < T) {
        // np.argmax: first maximum
        auto argmax = [&](int tt) {
            const float* m = probs + (size_t)tt * C;
            int b = 0;
            for (int c = 1; c < C; ++c)
                if (m[c] > m[b]) b = c;
            return b;
        };
        if (argmax(t) == space_idx) { ++t; continue; }
        rows.clear();
        while (t < T && argmax(t) != space_idx) rows.push_back(t++);
        ctc_beam_search(probs, rows.data(), (int)rows.size(), C, beam_width, dict, word);
        if (!first) text.push_back(space_idx);
        text.insert(text.end(), word.begin(), word.end());
        first = false;
    }
}

}  // namespace bbocr
