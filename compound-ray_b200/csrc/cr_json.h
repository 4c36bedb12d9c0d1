// cr_json.h -- minimal JSON DOM (RFC 8259 subset sufficient for glTF 2.0 ASCII files).
// Replaces the vendored nlohmann/tinygltf JSON layer the reference uses through
// tinygltf::TinyGLTF::LoadASCIIFromFile (libEyeRenderer3/MulticamScene.cpp:533-538).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cr {

class Json {
public:
    enum Type { Null, Bool, Number, String, Array, Object };
    Type type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // insertion order preserved

    bool isNull() const { return type == Null; }
    bool isBool() const { return type == Bool; }
    bool isNumber() const { return type == Number; }
    bool isString() const { return type == String; }
    bool isArray() const { return type == Array; }
    bool isObject() const { return type == Object; }

    const Json* find(const char* key) const
    {
        if (type != Object) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char* key) const { return find(key) != nullptr; }
    const Json& operator[](const char* key) const
    {
        static const Json nil;
        const Json* j = find(key);
        return j ? *j : nil;
    }
    const Json& operator[](size_t i) const
    {
        static const Json nil;
        return (type == Array && i < arr.size()) ? arr[i] : nil;
    }
    const Json& operator[](int i) const { return (*this)[static_cast<size_t>(i < 0 ? ~size_t(0) : static_cast<size_t>(i))]; }
    size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }
    double number(double dflt = 0.0) const { return type == Number ? num : dflt; }
    int integer(int dflt = 0) const      // numbers outside int's range (or NaN) count as absent: the cast would be undefined
    { return (type == Number && num >= -2147483648.0 && num <= 2147483647.0) ? static_cast<int>(num) : dflt; }
    const std::string& string() const { static const std::string e; return type == String ? str : e; }

    static Json parse(const std::string& text)
    {
        Parser p{text.data(), text.data() + text.size()};
        p.skipWs();
        Json j = p.value(0);
        p.skipWs();
        if (p.cur != p.end) p.fail("trailing characters");
        return j;
    }

private:
    struct Parser {
        const char* cur;
        const char* end;
        [[noreturn]] void fail(const char* what) const
        { throw std::runtime_error(std::string("JSON parse error: ") + what); }
        void skipWs()
        { while (cur < end && (*cur == ' ' || *cur == '\t' || *cur == '\n' || *cur == '\r')) ++cur; }
        bool consume(const char* lit)
        {
            size_t n = strlen(lit);
            if (static_cast<size_t>(end - cur) >= n && memcmp(cur, lit, n) == 0) { cur += n; return true; }
            return false;
        }
        static void appendUtf8(std::string& s, uint32_t cp)
        {
            if (cp < 0x80) s += static_cast<char>(cp);
            else if (cp < 0x800) { s += static_cast<char>(0xC0 | (cp >> 6)); s += static_cast<char>(0x80 | (cp & 0x3F)); }
            else if (cp < 0x10000) {
                s += static_cast<char>(0xE0 | (cp >> 12)); s += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
                s += static_cast<char>(0x80 | (cp & 0x3F));
            } else {
                s += static_cast<char>(0xF0 | (cp >> 18)); s += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
                s += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); s += static_cast<char>(0x80 | (cp & 0x3F));
            }
        }
        uint32_t hex4()
        {
            if (end - cur < 4) fail("bad \\u escape");
            uint32_t v = 0;
            for (int i = 0; i < 4; i++) {
                char c = *cur++;
                v <<= 4;
                if (c >= '0' && c <= '9') v |= static_cast<uint32_t>(c - '0');
                else if (c >= 'a' && c <= 'f') v |= static_cast<uint32_t>(c - 'a' + 10);
                else if (c >= 'A' && c <= 'F') v |= static_cast<uint32_t>(c - 'A' + 10);
                else fail("bad hex digit");
            }
            return v;
        }
        std::string stringLit()
        {
            if (cur >= end || *cur != '"') fail("expected string");
            ++cur;
            std::string s;
            // fast path: scan to the closing quote, copying runs without escapes (base64 buffers are MBs)
            for (;;) {
                const char* run = cur;
                while (cur < end && *cur != '"' && *cur != '\\') ++cur;
                s.append(run, static_cast<size_t>(cur - run));
                if (cur >= end) fail("unterminated string");
                if (*cur == '"') { ++cur; return s; }
                ++cur;   // backslash
                if (cur >= end) fail("bad escape");
                char c = *cur++;
                switch (c) {
                    case '"': s += '"'; break;
                    case '\\': s += '\\'; break;
                    case '/': s += '/'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'n': s += '\n'; break;
                    case 'r': s += '\r'; break;
                    case 't': s += '\t'; break;
                    case 'u': {
                        uint32_t cp = hex4();
                        if (cp >= 0xD800 && cp <= 0xDBFF && end - cur >= 6 && cur[0] == '\\' && cur[1] == 'u') {
                            cur += 2;
                            uint32_t lo = hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        appendUtf8(s, cp);
                        break;
                    }
                    default: fail("unknown escape");
                }
            }
        }
        Json value(int depth)
        {
            if (depth > 256) fail("nesting too deep");
            skipWs();
            if (cur >= end) fail("unexpected end");
            Json j;
            char c = *cur;
            if (c == '{') {
                ++cur;
                j.type = Object;
                skipWs();
                if (cur < end && *cur == '}') { ++cur; return j; }
                for (;;) {
                    skipWs();
                    std::string key = stringLit();
                    skipWs();
                    if (cur >= end || *cur != ':') fail("expected ':'");
                    ++cur;
                    j.obj.emplace_back(std::move(key), value(depth + 1));
                    skipWs();
                    if (cur < end && *cur == ',') { ++cur; continue; }
                    if (cur < end && *cur == '}') { ++cur; break; }
                    fail("expected ',' or '}'");
                }
            } else if (c == '[') {
                ++cur;
                j.type = Array;
                skipWs();
                if (cur < end && *cur == ']') { ++cur; return j; }
                for (;;) {
                    j.arr.push_back(value(depth + 1));
                    skipWs();
                    if (cur < end && *cur == ',') { ++cur; continue; }
                    if (cur < end && *cur == ']') { ++cur; break; }
                    fail("expected ',' or ']'");
                }
            } else if (c == '"') {
                j.type = String;
                j.str = stringLit();
            } else if (consume("true")) { j.type = Bool; j.b = true; }
            else if (consume("false")) { j.type = Bool; j.b = false; }
            else if (consume("null")) { j.type = Null; }
            else {
                // number: strtod on a bounded copy
                const char* s = cur;
                while (cur < end && (strchr("+-0123456789.eE", *cur) != nullptr)) ++cur;
                if (s == cur) fail("unexpected character");
                std::string tmp(s, static_cast<size_t>(cur - s));
                char* ep = nullptr;
                j.num = strtod(tmp.c_str(), &ep);
                if (ep == tmp.c_str()) fail("bad number");
                j.type = Number;
            }
            return j;
        }
    };
};

}  // namespace cr
