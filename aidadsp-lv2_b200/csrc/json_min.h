// json_min.h -- small recursive-descent JSON reader (RFC 8259) for model files.
//
// The reference parses model files with the vendored nlohmann::json
// (rt-neural-generic.cpp:970-975).  The engine only needs read access to a handful of fields
// (SURVEY.md Appendix B), so it carries its own ~200-line parser instead of a dependency.
// Numbers are held as double and narrowed on access, which is what `.get<float>()` does there.
#pragma once

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace aidax {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object };
    Kind kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;

    bool is_null() const { return kind == Null; }
    bool is_number() const { return kind == Number; }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }

    // Missing keys read as null (the reference uses the inserting operator[], same observable result).
    const JsonValue& operator[](const char* key) const {
        static const JsonValue null_value;
        if (kind == Null) return null_value;
        if (kind != Object) throw std::runtime_error(std::string("json: cannot use key '") + key + "' on a non-object");
        for (const auto& kv : obj)
            if (kv.first == key) return kv.second;
        return null_value;
    }
    const JsonValue& at(const char* key) const {
        if (kind != Object) throw std::runtime_error(std::string("json: key '") + key + "' requested on a non-object");
        for (const auto& kv : obj)
            if (kv.first == key) return kv.second;
        throw std::runtime_error(std::string("json: key '") + key + "' not found");
    }
    const JsonValue& at(size_t i) const {
        if (kind != Array) throw std::runtime_error("json: index requested on a non-array");
        if (i >= arr.size()) throw std::runtime_error("json: array index out of range");
        return arr[i];
    }
    const JsonValue& back() const {
        if (kind != Array || arr.empty()) throw std::runtime_error("json: back() on an empty or non-array value");
        return arr.back();
    }
    size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
    double as_double() const {
        if (kind == Number) return num;
        if (kind == Bool) return b ? 1.0 : 0.0;
        throw std::runtime_error("json: value is not a number");
    }
    int as_int() const { return (int)as_double(); }
    float as_float() const { return (float)as_double(); }
    const std::string& as_string() const {
        if (kind != String) throw std::runtime_error("json: value is not a string");
        return str;
    }
};

class JsonParser {
public:
    JsonParser(const char* begin, const char* end) : p_(begin), end_(end) {}

    JsonValue parse_document() {
        JsonValue v = parse_value(0);
        skip_ws();
        if (p_ != end_) fail("trailing characters after document");
        return v;
    }

private:
    const char* p_;
    const char* end_;

    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json parse error: ") + what); }

    void skip_ws() {
        while (p_ != end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
    }
    bool consume(const char* lit) {
        const size_t n = std::strlen(lit);
        if ((size_t)(end_ - p_) >= n && std::memcmp(p_, lit, n) == 0) { p_ += n; return true; }
        return false;
    }

    JsonValue parse_value(int depth) {
        if (depth > 64) fail("nesting too deep");
        skip_ws();
        if (p_ == end_) fail("unexpected end of input");
        JsonValue v;
        switch (*p_) {
        case '{': {
            ++p_;
            v.kind = JsonValue::Object;
            skip_ws();
            if (p_ != end_ && *p_ == '}') { ++p_; return v; }
            for (;;) {
                skip_ws();
                if (p_ == end_ || *p_ != '"') fail("expected string key");
                std::string key = parse_string();
                skip_ws();
                if (p_ == end_ || *p_ != ':') fail("expected ':'");
                ++p_;
                v.obj.emplace_back(std::move(key), parse_value(depth + 1));
                skip_ws();
                if (p_ == end_) fail("unterminated object");
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; return v; }
                fail("expected ',' or '}'");
            }
        }
        case '[': {
            ++p_;
            v.kind = JsonValue::Array;
            skip_ws();
            if (p_ != end_ && *p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(parse_value(depth + 1));
                skip_ws();
                if (p_ == end_) fail("unterminated array");
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == ']') { ++p_; return v; }
                fail("expected ',' or ']'");
            }
        }
        case '"':
            v.kind = JsonValue::String;
            v.str = parse_string();
            return v;
        case 't':
            if (!consume("true")) fail("bad literal");
            v.kind = JsonValue::Bool; v.b = true; return v;
        case 'f':
            if (!consume("false")) fail("bad literal");
            v.kind = JsonValue::Bool; v.b = false; return v;
        case 'n':
            if (!consume("null")) fail("bad literal");
            return v;
        default:
            return parse_number();
        }
    }

    JsonValue parse_number() {
        // strtod accepts a superset of JSON numbers (NaN, Infinity, hex); model exporters
        // (numpy/json.dump) can emit NaN/Infinity, which nlohmann rejects -- so do we.
        const char* s = p_;
        if (p_ != end_ && (*p_ == '-')) ++p_;
        if (p_ == end_ || !(*p_ >= '0' && *p_ <= '9')) fail("invalid number");
        while (p_ != end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) ++p_;
        std::string tmp(s, p_);
        char* endp = nullptr;
        const double d = std::strtod(tmp.c_str(), &endp);
        if (endp != tmp.c_str() + tmp.size()) fail("invalid number");
        JsonValue v;
        v.kind = JsonValue::Number;
        v.num = d;
        return v;
    }

    static void append_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    unsigned parse_hex4() {
        if (end_ - p_ < 4) fail("bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string parse_string() {
        ++p_; // opening quote
        std::string out;
        for (;;) {
            if (p_ == end_) fail("unterminated string");
            const char c = *p_++;
            if (c == '"') return out;
            if (c != '\\') { out.push_back(c); continue; }
            if (p_ == end_) fail("unterminated escape");
            const char e = *p_++;
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
                unsigned cp = parse_hex4();
                if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                    p_ += 2;
                    const unsigned lo = parse_hex4();
                    if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                append_utf8(out, cp);
                break;
            }
            default: fail("unknown escape");
            }
        }
    }
};

inline JsonValue json_parse(const std::string& text) {
    JsonParser p(text.data(), text.data() + text.size());
    return p.parse_document();
}

} // namespace aidax
