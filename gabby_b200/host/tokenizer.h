// tokenizer.h -- mirrors gabby's Tokenizer (/root/reference/src/inference/tokenizer.h:12-24: ctor
// from special_tokens_map / tokenizer_config / tokenizer.json, Tokenize(string_view) -> vector<int>),
// whose reference body returns {} (tokenizer.cc:6-8). This one implements byte-level BPE from
// tokenizer.json (model.vocab + model.merges, added_tokens as specials) with a byte fallback when the
// file carries no vocabulary (synthetic test directories), plus the Llama-3 chat template and the
// inverse mapping. The pre-tokenizer is the Llama-3 split regex (contractions, [^\r\n\p{L}\p{N}]?\p{L}+, \p{N}{1,3},
// punctuation runs, the whitespace alternations) over Unicode general categories (unicode_tables.inc); added tokens are
// cut out of the text before it. SURVEY.md section 8(f) row 1.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "json.h"

namespace gabby {
namespace inference {

class Tokenizer {
public:
    Tokenizer(json::ValuePtr special_tokens_map, json::ValuePtr tokenizer_config, json::ValuePtr tokens);
    virtual ~Tokenizer() = default;

    virtual std::vector<int> Tokenize(const std::string_view input);   // no specials added
    std::string Detokenize(const std::vector<int32_t>& ids) const;     // specials are skipped
    // <|begin_of_text|><|start_header_id|>system<|end_header_id|>\n\n{system}<|eot_id|> ... assistant header
    std::vector<int32_t> ChatPrompt(const std::string& system, const std::string& user);

    bool has_vocab() const { return !vocab_.empty(); }
    int special(const std::string& name) const;   // -1 when absent

private:
    std::vector<int> BpeWord(const std::string& mapped) const;

    std::unordered_map<std::string, int> vocab_;          // byte-mapped token string -> id
    std::vector<std::string> id_to_token_;
    std::unordered_map<std::string, int> merge_rank_;     // "a b" -> rank
    std::map<std::string, int> specials_;
    bool ignore_merges_ = false;                          // model.ignore_merges (true in Llama-3's tokenizer.json)
    std::string byte_to_unicode_[256];
    std::unordered_map<std::string, uint8_t> unicode_to_byte_;
};

}  // namespace inference
}  // namespace gabby
