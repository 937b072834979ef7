#include "json.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gabby {
namespace json {

namespace {
const char* TypeName(Type t) {
    switch (t) {
        case Type::NUM: return "number";
        case Type::BOOL: return "boolean";
        case Type::STR: return "string";
        case Type::ARRAY: return "array";
        case Type::OBJ: return "object";
        default: return "null";
    }
}

void AppendUtf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) {
        out.push_back(static_cast<char>(cp));
    } else if (cp < 0x800) {
        out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
        out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else {
        out.push_back(static_cast<char>(0xF0 | (cp >> 18)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    }
}

class Reader {
public:
    explicit Reader(std::string_view s) : s_(s) {}

    ValuePtr Document() {
        ValuePtr v = ParseValue(0);
        SkipWs();
        if (pos_ != s_.size()) Fail("trailing characters after the JSON value");
        return v;
    }

private:
    [[noreturn]] void Fail(const std::string& what) const {
        throw JSONError("json: " + what + " at offset " + std::to_string(pos_));
    }
    void SkipWs() {
        while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\n' || s_[pos_] == '\t' || s_[pos_] == '\r')) pos_++;
    }
    char Peek() {
        if (pos_ >= s_.size()) Fail("unexpected end of input");
        return s_[pos_];
    }
    void Expect(char c) {
        if (Peek() != c) Fail(std::string("expected '") + c + "'");
        pos_++;
    }
    bool Literal(const char* lit) {
        const size_t n = std::strlen(lit);
        if (s_.compare(pos_, n, lit) == 0) {
            pos_ += n;
            return true;
        }
        return false;
    }
    uint32_t Hex4() {
        if (pos_ + 4 > s_.size()) Fail("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            const char c = s_[pos_++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else Fail("bad hex digit in \\u escape");
        }
        return v;
    }
    std::string ParseString() {
        Expect('"');
        std::string out;
        for (;;) {
            if (pos_ >= s_.size()) Fail("unterminated string");
            const char c = s_[pos_++];
            if (c == '"') return out;
            if (c != '\\') {
                out.push_back(c);
                continue;
            }
            if (pos_ >= s_.size()) Fail("unterminated escape");
            const char e = s_[pos_++];
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp = Hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && s_.compare(pos_, 2, "\\u") == 0) {
                        pos_ += 2;
                        const uint32_t lo = Hex4();
                        if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else Fail("unpaired surrogate");
                    }
                    AppendUtf8(out, cp);
                    break;
                }
                default: Fail("unknown escape");
            }
        }
    }
    ValuePtr ParseNumber() {
        const size_t start = pos_;
        if (Peek() == '-') pos_++;
        while (pos_ < s_.size() && ((s_[pos_] >= '0' && s_[pos_] <= '9') || s_[pos_] == '.' || s_[pos_] == 'e' ||
                                    s_[pos_] == 'E' || s_[pos_] == '+' || s_[pos_] == '-'))
            pos_++;
        const std::string tok(s_.substr(start, pos_ - start));
        char* end = nullptr;
        errno = 0;
        const double d = std::strtod(tok.c_str(), &end);
        if (tok.empty() || end != tok.c_str() + tok.size()) Fail("malformed number '" + tok + "'");
        return Value::MakeNumber(d);
    }
    ValuePtr ParseValue(int depth) {
        if (depth > 256) Fail("nesting too deep");
        SkipWs();
        const char c = Peek();
        if (c == '{') {
            pos_++;
            std::vector<std::pair<std::string, ValuePtr>> members;
            SkipWs();
            if (Peek() == '}') {
                pos_++;
                return Value::MakeObject(std::move(members));
            }
            for (;;) {
                SkipWs();
                std::string key = ParseString();
                SkipWs();
                Expect(':');
                members.emplace_back(std::move(key), ParseValue(depth + 1));
                SkipWs();
                if (Peek() == ',') {
                    pos_++;
                    continue;
                }
                Expect('}');
                return Value::MakeObject(std::move(members));
            }
        }
        if (c == '[') {
            pos_++;
            std::vector<ValuePtr> items;
            SkipWs();
            if (Peek() == ']') {
                pos_++;
                return Value::MakeArray(std::move(items));
            }
            for (;;) {
                items.push_back(ParseValue(depth + 1));
                SkipWs();
                if (Peek() == ',') {
                    pos_++;
                    continue;
                }
                Expect(']');
                return Value::MakeArray(std::move(items));
            }
        }
        if (c == '"') return Value::MakeString(ParseString());
        if (Literal("true")) return Value::MakeBool(true);
        if (Literal("false")) return Value::MakeBool(false);
        if (Literal("null")) return Value::MakeNil();
        if (c == '-' || (c >= '0' && c <= '9')) return ParseNumber();
        Fail(std::string("unexpected character '") + c + "'");
    }

    std::string_view s_;
    size_t pos_ = 0;
};
}  // namespace

void Value::Want(Type t) const {
    if (type_ != t) throw JSONError(std::string("json: wanted ") + TypeName(t) + ", got " + TypeName(type_));
}

bool Value::contains(const std::string& key) const {
    Want(Type::OBJ);
    return index_.count(key) != 0;
}

const Value& Value::at(const std::string& key) const {
    Want(Type::OBJ);
    auto it = index_.find(key);
    if (it == index_.end()) throw JSONError("key not present in object: " + key);
    return *obj_[it->second].second;
}

const Value& Value::operator[](size_t i) const {
    Want(Type::ARRAY);
    if (i >= arr_.size()) throw std::out_of_range("out of range: " + std::to_string(i));
    return *arr_[i];
}

size_t Value::size() const {
    if (type_ == Type::ARRAY) return arr_.size();
    Want(Type::OBJ);
    return obj_.size();
}

double Value::number_or(const std::string& key, double dflt) const {
    return contains(key) && at(key).is(Type::NUM) ? at(key).as_number() : dflt;
}
bool Value::boolean_or(const std::string& key, bool dflt) const {
    return contains(key) && at(key).is(Type::BOOL) ? at(key).as_boolean() : dflt;
}

ValuePtr Value::MakeNil() { return std::make_shared<Value>(); }
ValuePtr Value::MakeBool(bool b) {
    auto v = std::make_shared<Value>();
    v->type_ = Type::BOOL;
    v->bool_ = b;
    return v;
}
ValuePtr Value::MakeNumber(double d) {
    auto v = std::make_shared<Value>();
    v->type_ = Type::NUM;
    v->num_ = d;
    return v;
}
ValuePtr Value::MakeString(std::string s) {
    auto v = std::make_shared<Value>();
    v->type_ = Type::STR;
    v->str_ = std::move(s);
    return v;
}
ValuePtr Value::MakeArray(std::vector<ValuePtr> items) {
    auto v = std::make_shared<Value>();
    v->type_ = Type::ARRAY;
    v->arr_ = std::move(items);
    return v;
}
ValuePtr Value::MakeObject(std::vector<std::pair<std::string, ValuePtr>> m) {
    auto v = std::make_shared<Value>();
    v->type_ = Type::OBJ;
    v->obj_ = std::move(m);
    for (size_t i = 0; i < v->obj_.size(); i++) v->index_[v->obj_[i].first] = i;  // last duplicate wins
    return v;
}

ValuePtr Parse(std::string_view text) { return Reader(text).Document(); }

ValuePtr ParseFile(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error(path + ": " + std::strerror(errno));
    std::string data;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
    std::fclose(f);
    try {
        return Parse(data);
    } catch (const JSONError& e) {
        throw JSONError(path + ": " + e.what());
    }
}

}  // namespace json
}  // namespace gabby
