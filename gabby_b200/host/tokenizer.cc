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

enum CharClass { kSpace, kLetter, kDigit, kOther };
CharClass Classify(unsigned char c) {
    if (c == ' ' || c == '\t' || c == '\n' || c == '\r') return kSpace;
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80) return kLetter;  // non-ASCII bytes group with letters
    if (c >= '0' && c <= '9') return kDigit;
    return kOther;
}

// Approximation of the Llama-3 pre-tokenizer regex: an optional single leading space glued to a run of
// letters / a run of punctuation, digits in groups of <= 3, whitespace runs kept together.
std::vector<std::string> PreTokenize(std::string_view s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        size_t start = i;
        unsigned char c = static_cast<unsigned char>(s[i]);
        if (c == ' ' && i + 1 < s.size() && Classify(static_cast<unsigned char>(s[i + 1])) != kSpace) {
            i++;  // the space travels with the following word
            c = static_cast<unsigned char>(s[i]);
        }
        const CharClass cls = Classify(c);
        if (cls == kDigit) {
            size_t n = 0;
            while (i < s.size() && Classify(static_cast<unsigned char>(s[i])) == kDigit && n < 3) i++, n++;
        } else if (cls == kSpace) {
            while (i < s.size() && Classify(static_cast<unsigned char>(s[i])) == kSpace) i++;
            // leave the last space for the next word when one follows
            if (i < s.size() && i - start > 1 && s[i - 1] == ' ') i--;
        } else {
            while (i < s.size() && Classify(static_cast<unsigned char>(s[i])) == cls) i++;
        }
        out.emplace_back(s.substr(start, i - start));
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
    for (const std::string& word : PreTokenize(input)) {
        std::string mapped;
        for (unsigned char c : word) mapped += byte_to_unicode_[c];
        const std::vector<int> ids = BpeWord(mapped);
        out.insert(out.end(), ids.begin(), ids.end());
    }
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
