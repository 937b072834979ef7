#include "tokenizer.h"

#include <algorithm>
#include <climits>

namespace gabby {
namespace inference {

namespace {
void AppendUtf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) {
        out.push_back(static_cast<char>(cp));
    } else if (cp < 0x800) {
        out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else {
        out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    }
}

// split a UTF-8 string into code-point substrings
std::vector<std::string> Utf8Chars(const std::string& s) {
    std::vector<std::string> out;
    for (size_t i = 0; i < s.size();) {
        const unsigned char c = static_cast<unsigned char>(s[i]);
        const size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : 4;
        out.push_back(s.substr(i, n));
        i += n;
    }
    return out;
}

#include "unicode_tables.inc"

template <size_t N>
bool InRanges(const uint32_t (&r)[N][2], uint32_t cp) {
    size_t lo = 0, hi = N;
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (cp < r[mid][0]) hi = mid;
        else if (cp > r[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
bool IsLetter(uint32_t c) { return c < 0x80 ? ((c | 0x20) >= 'a' && (c | 0x20) <= 'z') : InRanges(kUnicodeLetter, c); }
bool IsNumber(uint32_t c) { return c < 0x80 ? (c >= '0' && c <= '9') : InRanges(kUnicodeNumber, c); }
bool IsSpace(uint32_t c) { return c < 0x80 ? (c == ' ' || (c >= 0x9 && c <= 0xD)) : InRanges(kUnicodeSpace, c); }
bool IsNewline(uint32_t c) { return c == '\r' || c == '\n'; }

// UTF-8 -> code points with the byte offset of each (malformed bytes become one code point each, so that every
// input byte is covered and the byte-level BPE below still sees it)
void DecodeUtf8(std::string_view s, std::vector<uint32_t>& cps, std::vector<size_t>& offs) {
    for (size_t i = 0; i < s.size();) {
        const unsigned char c = static_cast<unsigned char>(s[i]);
        size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
        if (i + n > s.size()) n = 1;
        uint32_t cp = c;
        if (n > 1) {
            cp = c & (0xFF >> (n + 1));
            for (size_t k = 1; k < n; k++) {
                const unsigned char cc = static_cast<unsigned char>(s[i + k]);
                if ((cc & 0xC0) != 0x80) { n = 1; cp = c; break; }
                cp = (cp << 6) | (cc & 0x3F);
            }
        }
        if (n == 1 && c >= 0x80) cp = 0xFFFD;   // stray byte: not a letter, number or space
        cps.push_back(cp);
        offs.push_back(i);
        i += n;
    }
    offs.push_back(s.size());
}

// The Llama-3 pre-tokenizer (tokenizer.json: pre_tokenizer.pretokenizers[0], Split / Isolated):
//   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N}{1,3} | ' '?[^\s\p{L}\p{N}]+[\r\n]* | \s*[\r\n]+ | \s+(?!\S) | \s+
// matched leftmost-first, alternatives in this order, each greedy with backtracking -- written out by hand over code points.
std::vector<std::string> PreTokenize(std::string_view s) {
    std::vector<uint32_t> c;
    std::vector<size_t> off;
    DecodeUtf8(s, c, off);
    const size_t n = c.size();
    std::vector<std::string> out;
    auto lower = [](uint32_t x) { return x >= 'A' && x <= 'Z' ? x + 32 : x; };
    size_t i = 0;
    while (i < n) {
        size_t end = 0;
        // 1. contractions, case-insensitive
        if (c[i] == '\'' && i + 1 < n) {
            const uint32_t a = lower(c[i + 1]), b = i + 2 < n ? lower(c[i + 2]) : 0;
            if (a == 's' || a == 't') end = i + 2;
            else if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e')) end = i + 3;
            else if (a == 'm') end = i + 2;
            else if (a == 'l' && b == 'l') end = i + 3;
            else if (a == 'd') end = i + 2;
        }
        // 2. an optional non-letter/number/newline character, then letters
        if (!end) {
            size_t j = i;
            if (!IsNewline(c[j]) && !IsLetter(c[j]) && !IsNumber(c[j]) && j + 1 < n && IsLetter(c[j + 1])) j++;
            if (IsLetter(c[j])) {
                while (j < n && IsLetter(c[j])) j++;
                end = j;
            }
        }
        // 3. one to three numbers
        if (!end && IsNumber(c[i])) {
            size_t j = i;
            while (j < n && j < i + 3 && IsNumber(c[j])) j++;
            end = j;
        }
        // 4. an optional space, then characters that are neither space, letter nor number, then newlines
        if (!end) {
            size_t j = i;
            auto other = [&](size_t k) { return k < n && !IsSpace(c[k]) && !IsLetter(c[k]) && !IsNumber(c[k]); };
            if (c[j] == ' ' && other(j + 1)) j++;
            if (other(j)) {
                while (other(j)) j++;
                while (j < n && IsNewline(c[j])) j++;
                end = j;
            }
        }
        if (!end) {   // c[i] is whitespace from here on
            size_t run = i;
            while (run < n && IsSpace(c[run])) run++;
            // 5. whitespace up to and including its last newline
            size_t last_nl = 0;
            bool have_nl = false;
            for (size_t k = i; k < run; k++)
                if (IsNewline(c[k])) last_nl = k, have_nl = true;
            if (have_nl) end = last_nl + 1;
            // 6. whitespace not followed by a non-space: everything at the end of the text, else all but the last one
            else if (run == n) end = run;
            else if (run - i >= 2) end = run - 1;
            // 7. whitespace
            else end = run;
        }
        out.emplace_back(s.substr(off[i], off[end] - off[i]));
        i = end;
    }
    return out;
}
}  // namespace

Tokenizer::Tokenizer(json::ValuePtr, json::ValuePtr, json::ValuePtr tokens) {
    // GPT-2 byte <-> printable-unicode table
    std::vector<int> bs;
    for (int b = '!'; b <= '~'; b++) bs.push_back(b);
    for (int b = 0xA1; b <= 0xAC; b++) bs.push_back(b);
    for (int b = 0xAE; b <= 0xFF; b++) bs.push_back(b);
    std::vector<int> cs = bs;
    int extra = 0;
    for (int b = 0; b < 256; b++) {
        if (std::find(bs.begin(), bs.end(), b) == bs.end()) {
            bs.push_back(b);
            cs.push_back(256 + extra++);
        }
    }
    for (size_t i = 0; i < bs.size(); i++) {
        std::string u;
        AppendUtf8(u, static_cast<uint32_t>(cs[i]));
        byte_to_unicode_[bs[i]] = u;
        unicode_to_byte_[u] = static_cast<uint8_t>(bs[i]);
    }
    if (!tokens || !tokens->is(json::Type::OBJ)) return;
    if (tokens->contains("model") && tokens->at("model").is(json::Type::OBJ)) {
        const json::Value& model = tokens->at("model");
        if (model.contains("vocab") && model.at("vocab").is(json::Type::OBJ)) {
            for (const auto& [tok, id] : model.at("vocab").members()) {
                const int i = static_cast<int>(id->as_number());
                vocab_[tok] = i;
                if (static_cast<int>(id_to_token_.size()) <= i) id_to_token_.resize(i + 1);
                id_to_token_[i] = tok;
            }
        }
        if (model.contains("ignore_merges") && model.at("ignore_merges").is(json::Type::BOOL)) ignore_merges_ = model.at("ignore_merges").as_boolean();
        if (model.contains("merges") && model.at("merges").is(json::Type::ARRAY)) {
            int rank = 0;
            for (const auto& m : model.at("merges").as_array()) {
                std::string key;
                if (m->is(json::Type::STR)) key = m->as_string();                       // "a b"
                else if (m->is(json::Type::ARRAY) && m->size() == 2) key = (*m)[0].as_string() + " " + (*m)[1].as_string();
                if (!key.empty()) merge_rank_[key] = rank++;
            }
        }
    }
    if (tokens->contains("added_tokens") && tokens->at("added_tokens").is(json::Type::ARRAY)) {
        for (const auto& t : tokens->at("added_tokens").as_array()) {
            const int id = static_cast<int>(t->at("id").as_number());
            specials_[t->at("content").as_string()] = id;
            if (static_cast<int>(id_to_token_.size()) <= id) id_to_token_.resize(id + 1);
        }
    }
}

int Tokenizer::special(const std::string& name) const {
    auto it = specials_.find(name);
    return it == specials_.end() ? -1 : it->second;
}

std::vector<int> Tokenizer::BpeWord(const std::string& mapped) const {
    std::vector<std::string> parts = Utf8Chars(mapped);
    while (parts.size() > 1) {
        int best = INT_MAX;
        size_t at = 0;
        for (size_t i = 0; i + 1 < parts.size(); i++) {
            auto it = merge_rank_.find(parts[i] + " " + parts[i + 1]);
            if (it != merge_rank_.end() && it->second < best) best = it->second, at = i;
        }
        if (best == INT_MAX) break;
        parts[at] += parts[at + 1];
        parts.erase(parts.begin() + at + 1);
    }
    std::vector<int> ids;
    for (const auto& p : parts) {
        auto it = vocab_.find(p);
        if (it != vocab_.end()) {
            ids.push_back(it->second);
        } else {  // unknown piece: fall back to its single-character tokens
            for (const auto& ch : Utf8Chars(p)) {
                auto jt = vocab_.find(ch);
                if (jt != vocab_.end()) ids.push_back(jt->second);
            }
        }
    }
    return ids;
}

std::vector<int> Tokenizer::Tokenize(const std::string_view input) {
    std::vector<int> out;
    if (input.empty()) return out;   // the reference's only pinned behaviour (tokenizer_test.cc:9-25)
    if (vocab_.empty()) {
        for (unsigned char c : input) out.push_back(c);   // byte fallback (no vocabulary in tokenizer.json)
        return out;
    }
    auto encode_text = [&](std::string_view text) {
        for (const std::string& word : PreTokenize(text)) {
            std::string mapped;
            for (unsigned char c : word) mapped += byte_to_unicode_[c];
            if (ignore_merges_) {   // Llama-3's tokenizer.json: a pre-token that is itself in the vocabulary is one token
                auto it = vocab_.find(mapped);
                if (it != vocab_.end()) {
                    out.push_back(it->second);
                    continue;
                }
            }
            const std::vector<int> ids = BpeWord(mapped);
            out.insert(out.end(), ids.begin(), ids.end());
        }
    };
    // added (special) tokens are cut out of the raw text first: leftmost occurrence, longest content on ties
    size_t pos = 0;
    while (pos < input.size()) {
        size_t best_at = std::string_view::npos, best_len = 0;
        int best_id = -1;
        for (const auto& [content, id] : specials_) {
            if (content.empty()) continue;
            const size_t at = input.find(content, pos);
            if (at == std::string_view::npos) continue;
            if (at < best_at || (at == best_at && content.size() > best_len)) best_at = at, best_len = content.size(), best_id = id;
        }
        if (best_id < 0) break;
        if (best_at > pos) encode_text(input.substr(pos, best_at - pos));
        out.push_back(best_id);
        pos = best_at + best_len;
    }
    if (pos < input.size()) encode_text(input.substr(pos));
    return out;
}

std::string Tokenizer::Detokenize(const std::vector<int32_t>& ids) const {
    std::string out;
    if (vocab_.empty()) {
        for (int32_t id : ids)
            if (id >= 0 && id < 256) out.push_back(static_cast<char>(id));
        return out;
    }
    for (int32_t id : ids) {
        if (id < 0 || id >= static_cast<int>(id_to_token_.size())) continue;
        for (const auto& ch : Utf8Chars(id_to_token_[id])) {
            auto it = unicode_to_byte_.find(ch);
            if (it != unicode_to_byte_.end()) out.push_back(static_cast<char>(it->second));
        }
    }
    return out;
}

std::vector<int32_t> Tokenizer::ChatPrompt(const std::string& system, const std::string& user) {
    std::vector<int32_t> ids;
    auto push_special = [&](const char* name) {
        const int id = special(name);
        if (id >= 0) ids.push_back(id);
    };
    auto push_text = [&](const std::string& s) {
        for (int t : Tokenize(s)) ids.push_back(t);
    };
    auto turn = [&](const char* role, const std::string& content) {
        push_special("<|start_header_id|>");
        push_text(role);
        push_special("<|end_header_id|>");
        push_text("\n\n");
        push_text(content);
        push_special("<|eot_id|>");
    };
    push_special("<|begin_of_text|>");
    turn("system", system);
    turn("user", user);
    push_special("<|start_header_id|>");
    push_text("assistant");
    push_special("<|end_header_id|>");
    push_text("\n\n");
    return ids;
}

}  // namespace inference
}  // namespace gabby
