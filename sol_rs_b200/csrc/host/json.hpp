// json.hpp — minimal JSON reader for glTF documents (objects, arrays, strings, numbers, bools, null).
// Stands in for the serde_json layer underneath the `gltf` crate the reference uses (Cargo.lock:518-519).
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace sol {
namespace json {

struct Value;
using Array = std::vector<Value>;
using Object = std::map<std::string, Value>;

struct Value {
    enum Type { Null, Bool, Number, String, Arr, Obj } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::shared_ptr<Array> arr;
    std::shared_ptr<Object> obj;

    bool is_null() const { return type == Null; }
    bool has(const std::string &k) const { return type == Obj && obj->count(k) != 0; }
    const Value &operator[](const std::string &k) const {
        static const Value none;
        if (type != Obj) return none;
        auto it = obj->find(k);
        return it == obj->end() ? none : it->second;
    }
    const Value &operator[](size_t i) const {
        if (type != Arr || i >= arr->size()) throw std::runtime_error("json: array index out of range");
        return (*arr)[i];
    }
    size_t size() const { return type == Arr ? arr->size() : (type == Obj ? obj->size() : 0); }
    double number(double dflt) const { return type == Number ? num : dflt; }
    long integer(long dflt) const { return type == Number ? (long)num : dflt; }
    const std::string &string() const { return str; }
};

class Parser {
public:
    explicit Parser(const std::string &s) : s_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string &s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const char *m) const { throw std::runtime_error(std::string("json: ") + m + " at byte " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) p_++; }
    char peek() { ws(); if (p_ >= s_.size()) fail("unexpected end"); return s_[p_]; }
    void expect(char c) { if (peek() != c) fail("unexpected character"); p_++; }
    Value value() {
        const char c = peek();
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') { Value v; v.type = Value::String; v.str = string(); return v; }
        if (c == 't' || c == 'f' || c == 'n') return literal();
        return number();
    }
    Value literal() {
        Value v;
        if (s_.compare(p_, 4, "true") == 0) { v.type = Value::Bool; v.b = true; p_ += 4; }
        else if (s_.compare(p_, 5, "false") == 0) { v.type = Value::Bool; v.b = false; p_ += 5; }
        else if (s_.compare(p_, 4, "null") == 0) { p_ += 4; }
        else fail("bad literal");
        return v;
    }
    Value number() {
        const char *begin = s_.c_str() + p_;
        char *end = nullptr;
        const double d = std::strtod(begin, &end);
        if (end == begin) fail("bad number");
        p_ += (size_t)(end - begin);
        Value v; v.type = Value::Number; v.num = d;
        return v;
    }
    std::string string() {
        expect('"');
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\') {
                if (p_ >= s_.size()) fail("bad escape");
                const char e = s_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (p_ + 4 > s_.size()) fail("bad \\u escape");
                        const unsigned cp = (unsigned)std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
                        p_ += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xc0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3f)); }
                        else { out += (char)(0xe0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3f)); out += (char)(0x80 | (cp & 0x3f)); }
                        break;
                    }
                    default: out += e;
                }
            } else out += c;
        }
        if (p_ >= s_.size()) fail("unterminated string");
        p_++;
        return out;
    }
    Value array() {
        expect('[');
        Value v; v.type = Value::Arr; v.arr = std::make_shared<Array>();
        if (peek() == ']') { p_++; return v; }
        for (;;) {
            v.arr->push_back(value());
            const char c = peek();
            p_++;
            if (c == ']') break;
            if (c != ',') fail("expected , or ]");
        }
        return v;
    }
    Value object() {
        expect('{');
        Value v; v.type = Value::Obj; v.obj = std::make_shared<Object>();
        if (peek() == '}') { p_++; return v; }
        for (;;) {
            if (peek() != '"') fail("expected key");
            std::string k = string();
            expect(':');
            (*v.obj)[k] = value();
            const char c = peek();
            p_++;
            if (c == '}') break;
            if (c != ',') fail("expected , or }");
        }
        return v;
    }
};

inline Value parse(const std::string &text) { return Parser(text).parse(); }

}  // namespace json
}  // namespace sol
