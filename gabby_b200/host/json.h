// json.h -- small JSON reader for config.json, the safetensors header and tokenizer.json.
// Interface shaped after the reference's json::Value tree (/root/reference/src/json/json.h:
// as_object().at(k), as_array()[i], as_number().get(), *as_string()) but value-typed, with real
// string escapes (the reference drops the backslash: parser.cc:112-121, which breaks
// tokenizer.json vocab entries) and 64-bit sizes (reference: int, parser.h:36,71).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace gabby {
namespace json {

class JSONError : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};

enum class Type { NUM, BOOL, STR, ARRAY, OBJ, NIL };

class Value;
using ValuePtr = std::shared_ptr<Value>;

class Value {
public:
    Type type() const { return type_; }
    bool is(Type t) const { return type_ == t; }

    double as_number() const { Want(Type::NUM); return num_; }
    int64_t as_int() const { Want(Type::NUM); return static_cast<int64_t>(num_); }
    bool as_boolean() const { Want(Type::BOOL); return bool_; }
    const std::string& as_string() const { Want(Type::STR); return str_; }
    const std::vector<ValuePtr>& as_array() const { Want(Type::ARRAY); return arr_; }
    // insertion-ordered members (tokenizer merges / vocab order matters to nobody, but keep it stable)
    const std::vector<std::pair<std::string, ValuePtr>>& members() const { Want(Type::OBJ); return obj_; }

    bool contains(const std::string& key) const;
    // throws JSONError("key not present in object: k") like the reference's ObjectValue::at
    const Value& at(const std::string& key) const;
    const Value& operator[](size_t i) const;
    size_t size() const;

    // typed getters with defaults
    double number_or(const std::string& key, double dflt) const;
    bool boolean_or(const std::string& key, bool dflt) const;

    static ValuePtr MakeNil();
    static ValuePtr MakeBool(bool b);
    static ValuePtr MakeNumber(double d);
    static ValuePtr MakeString(std::string s);
    static ValuePtr MakeArray(std::vector<ValuePtr> v);
    static ValuePtr MakeObject(std::vector<std::pair<std::string, ValuePtr>> m);

private:
    void Want(Type t) const;
    Type type_ = Type::NIL;
    double num_ = 0;
    bool bool_ = false;
    std::string str_;
    std::vector<ValuePtr> arr_;
    std::vector<std::pair<std::string, ValuePtr>> obj_;
    std::map<std::string, size_t> index_;
};

ValuePtr Parse(std::string_view text);
ValuePtr ParseFile(const std::string& path);

}  // namespace json
}  // namespace gabby
