// Minimal JSON DOM reader for the SNN model format (replaces picojson in core/src/ic2/modelparser.cpp).
// Model files carry weights as JSON number arrays hundreds of MB long, so arrays whose elements are all numbers
// are stored as one contiguous std::vector<double> instead of a vector of nodes.
#pragma once

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace snn {
namespace json {

struct Value;
typedef std::shared_ptr<Value> ValuePtr;

struct Value {
    enum Type { Null, Bool, Number, String, Array, NumArray, Object } type = Null;
    bool b     = false;
    double num = 0.0;
    std::string str;
    std::vector<ValuePtr> arr;              // Array (mixed / nested)
    std::vector<double> nums;               // NumArray (all numbers)
    std::vector<std::pair<std::string, ValuePtr>> obj; // Object (insertion order kept)

    bool isNumber() const { return type == Number; }
    bool isString() const { return type == String; }
    bool isObject() const { return type == Object; }
    bool isArray() const { return type == Array || type == NumArray; }
    size_t size() const { return type == NumArray ? nums.size() : arr.size(); }
    bool has(const std::string& key) const {
        for (auto& kv : obj)
            if (kv.first == key) return true;
        return false;
    }
    const Value& at(const std::string& key) const {
        for (auto& kv : obj)
            if (kv.first == key) return *kv.second;
        throw std::runtime_error("JSON: missing key '" + key + "'");
    }
    const Value* find(const std::string& key) const {
        for (auto& kv : obj)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    double asNumber() const {
        if (type != Number) throw std::runtime_error("JSON: value is not a number");
        return num;
    }
    const std::string& asString() const {
        if (type != String) throw std::runtime_error("JSON: value is not a string");
        return str;
    }
    // element i of an array as a number (works for both array flavours)
    double numAt(size_t i) const {
        if (type == NumArray) return nums.at(i);
        if (type == Array) return arr.at(i)->asNumber();
        throw std::runtime_error("JSON: value is not an array");
    }
    const Value& elemAt(size_t i) const {
        if (type != Array) throw std::runtime_error("JSON: value is not a nested array");
        return *arr.at(i);
    }
};

class Parser {
public:
    Parser(const char* begin, const char* end): p(begin), e(end) {}
    ValuePtr parse() {
        ValuePtr v = value();
        ws();
        if (p != e) fail("trailing characters");
        return v;
    }

private:
    const char* p;
    const char* e;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("JSON parse error: ") + what); }
    void ws() {
        while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    ValuePtr value() {
        ws();
        if (p >= e) fail("unexpected end");
        switch (*p) {
        case '{': return object();
        case '[': return array();
        case '"': {
            auto v  = std::make_shared<Value>();
            v->type = Value::String;
            v->str  = string();
            return v;
        }
        case 't':
        case 'f':
        case 'n': return literal();
        default: {
            auto v  = std::make_shared<Value>();
            v->type = Value::Number;
            v->num  = number();
            return v;
        }
        }
    }
    double number() {
        char* endp = nullptr;
        double d   = strtod(p, &endp);
        if (endp == p) {
            // picojson/Python may emit NaN / Infinity literals
            if (e - p >= 3 && !strncmp(p, "NaN", 3)) {
                p += 3;
                return NAN;
            }
            if (e - p >= 8 && !strncmp(p, "Infinity", 8)) {
                p += 8;
                return INFINITY;
            }
            if (e - p >= 9 && !strncmp(p, "-Infinity", 9)) {
                p += 9;
                return -INFINITY;
            }
            fail("bad number");
        }
        p = endp;
        return d;
    }
    std::string string() {
        std::string s;
        ++p; // opening quote
        while (p < e && *p != '"') {
            if (*p == '\\') {
                ++p;
                if (p >= e) fail("bad escape");
                switch (*p) {
                case 'n': s += '\n'; break;
                case 't': s += '\t'; break;
                case 'r': s += '\r'; break;
                case 'b': s += '\b'; break;
                case 'f': s += '\f'; break;
                case 'u': { // keep BMP code points as UTF-8
                    if (e - p < 5) fail("bad \\u escape");
                    unsigned cp = (unsigned) strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    p += 4;
                    if (cp < 0x80)
                        s += (char) cp;
                    else if (cp < 0x800) {
                        s += (char) (0xC0 | (cp >> 6));
                        s += (char) (0x80 | (cp & 0x3F));
                    } else {
                        s += (char) (0xE0 | (cp >> 12));
                        s += (char) (0x80 | ((cp >> 6) & 0x3F));
                        s += (char) (0x80 | (cp & 0x3F));
                    }
                    break;
                }
                default: s += *p; break;
                }
                ++p;
            } else {
                s += *p++;
            }
        }
        if (p >= e) fail("unterminated string");
        ++p;
        return s;
    }
    ValuePtr literal() {
        auto v = std::make_shared<Value>();
        if (e - p >= 4 && !strncmp(p, "true", 4)) {
            v->type = Value::Bool, v->b = true, p += 4;
        } else if (e - p >= 5 && !strncmp(p, "false", 5)) {
            v->type = Value::Bool, v->b = false, p += 5;
        } else if (e - p >= 4 && !strncmp(p, "null", 4)) {
            v->type = Value::Null, p += 4;
        } else
            fail("bad literal");
        return v;
    }
    ValuePtr array() {
        auto v = std::make_shared<Value>();
        ++p;
        ws();
        if (p < e && *p == ']') {
            ++p;
            v->type = Value::NumArray;
            return v;
        }
        // fast path: a run of numbers
        bool numeric = true;
        {
            ws();
            char c  = *p;
            numeric = (c == '-' || (c >= '0' && c <= '9') || c == 'N' || c == 'I');
        }
        if (numeric) {
            v->type = Value::NumArray;
            for (;;) {
                ws();
                char c = p < e ? *p : 0;
                if (!(c == '-' || (c >= '0' && c <= '9') || c == 'N' || c == 'I')) {
                    // mixed array after all: convert what we have and continue on the slow path
                    v->type = Value::Array;
                    for (double d : v->nums) {
                        auto n  = std::make_shared<Value>();
                        n->type = Value::Number, n->num = d;
                        v->arr.push_back(n);
                    }
                    v->nums.clear();
                    break;
                }
                v->nums.push_back(number());
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == ']') {
                    ++p;
                    return v;
                }
                fail("expected , or ] in array");
            }
        } else {
            v->type = Value::Array;
        }
        for (;;) {
            v->arr.push_back(value());
            ws();
            if (p < e && *p == ',') {
                ++p;
                continue;
            }
            if (p < e && *p == ']') {
                ++p;
                return v;
            }
            fail("expected , or ] in array");
        }
    }
    ValuePtr object() {
        auto v  = std::make_shared<Value>();
        v->type = Value::Object;
        ++p;
        ws();
        if (p < e && *p == '}') {
            ++p;
            return v;
        }
        for (;;) {
            ws();
            if (p >= e || *p != '"') fail("expected key string");
            std::string key = string();
            ws();
            if (p >= e || *p != ':') fail("expected :");
            ++p;
            v->obj.emplace_back(std::move(key), value());
            ws();
            if (p < e && *p == ',') {
                ++p;
                continue;
            }
            if (p < e && *p == '}') {
                ++p;
                return v;
            }
            fail("expected , or } in object");
        }
    }
};

inline ValuePtr parse(const std::string& text) { return Parser(text.data(), text.data() + text.size()).parse(); }

} // namespace json
} // namespace snn
